"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI, against the oracle (bit-exact
for codes / ids / table bits; stated tolerances for floating-point aggregates), the committed golden
vectors and the reference's own linscan .so.  Nothing here reads /root/reference."""
import glob
import os

import numpy as np
import pytest

from util import make_problem, make_scan_problem

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def gpu(lsq):
    assert lsq.device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    lsq.init(0)
    return lsq


# ---------------------------------------------------------------- tables (a3, a4) and costs (a5)
@pytest.mark.parametrize("kind", ["sift", "gauss"])
@pytest.mark.parametrize("n,d,m", [(300, 128, 8), (77, 32, 7), (130, 20, 3), (1, 128, 16), (513, 6, 2), (1000, 100, 5)])
def test_unaries_bit_exact(gpu, oracle, n, d, m, kind):
    # "sift" data sums exactly in fp32 (any order passes); "gauss" exercises the sequential-k FMA chain
    X, C, _ = make_problem(100 + n, n, d, m, kind=kind)
    assert np.array_equal(gpu.get_unaries(X, C), oracle.get_unaries(X, C))


@pytest.mark.parametrize("d,m", [(128, 8), (32, 4), (10, 3), (64, 16)])
def test_binaries_bit_exact(gpu, oracle, d, m):
    _, C, _ = make_problem(200 + d, 4, d, m, kind="gauss")
    G, cbi = gpu.get_binaries(C)
    Go, cbo = oracle.get_binaries(C)
    assert np.array_equal(G, Go)
    assert np.array_equal(cbi, cbo + 1)  # reference cbi is 1-based (utils.jl:138)


@pytest.mark.parametrize("n,d,m,kind", [(1000, 128, 8, "sift"), (257, 100, 7, "gauss"), (64, 7, 16, "gauss")])
def test_veccost_qerror_reconstruct(gpu, oracle, n, d, m, kind):
    X, C, B = make_problem(300 + n, n, d, m, kind=kind)
    B0 = (B - 1).astype(np.int16)
    assert np.array_equal(gpu.veccost(X, B, C), oracle.veccost(X, B0, C))
    assert np.array_equal(gpu.reconstruct(B, C), oracle.reconstruct(B0, C))
    q, qo = gpu.qerror(X, B, C), oracle.qerror(X, B0, C)
    assert abs(q - qo) <= 1e-6 * abs(qo)  # float64 sum of identical float32 costs, rounded once


def test_quantize_norms(gpu, oracle):
    X, C, B = make_problem(17, 500, 32, 7)
    rng = np.random.default_rng(0)
    rec = oracle.reconstruct((B - 1).astype(np.int16), C)
    cb = np.sort(rng.choice((rec ** 2).sum(1), 256)).astype(np.float32)
    assert np.array_equal(gpu.quantize_norms(B, C, cb), oracle.quantize_norms((B - 1).astype(np.int16), C, cb) + 1)


# ---------------------------------------------------------------- encoding_icm (a1, a2)
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "icm_*.npz"))))
def test_encoding_icm_golden(gpu, path):
    g = np.load(path)
    X, C, B = make_problem(int(g["seed"]), int(g["n"]), int(g["d"]), int(g["m"]))
    for it in range(int(g["iters"])):
        B = gpu.encoding_icm(X, B, C, int(g["niter"]), bool(g["randord"]), int(g["npert"]),
                             seed=int(g["seed"]), ils_iter=it)
        assert np.array_equal(B, g["codes"][it])
        assert np.array_equal(gpu.veccost(X, B, C), g["cost"][it])


@pytest.mark.parametrize("n,d,m,niter,npert,randord,kind", [
    (2000, 128, 8, 4, 4, True, "sift"),      # demo hyper-parameters (demo_lsq.jl:34-37), m = 8
    (1500, 128, 7, 4, 4, True, "sift"),      # the demos' m = 7 (+1 norm byte)
    (600, 128, 16, 4, 4, True, "sift"),      # 128-bit codes
    (700, 64, 15, 2, 6, False, "gauss"),
    (333, 96, 5, 3, 0, True, "gauss"),       # npert = 0: plain ICM
    (100, 40, 1, 2, 1, True, "gauss"),       # a single codebook: no pair terms at all
    (1, 128, 8, 4, 4, True, "sift"),         # a single vector
])
def test_encoding_icm_bit_exact(gpu, oracle, n, d, m, niter, npert, randord, kind):
    X, C, B = make_problem(400 + n + m, n, d, m, kind=kind)
    Bo, _ = oracle.encoding_icm(X, (B - 1).astype(np.int16), C, niter, randord, npert, seed=42, ils_iter=3, g0=7)
    Bg = gpu.encoding_icm(X, B, C, niter, randord, npert, seed=42, ils_iter=3, g0=7)
    assert np.array_equal(Bg, Bo + 1)


def test_encoding_icm_explicit_schedule(gpu, oracle):
    """The schedule is an input: any permutation / perturbation the caller draws must give the
    oracle's codes (this is how a Julia RNG stream would be replayed)."""
    n, d, m = 800, 128, 8
    X, C, B = make_problem(500, n, d, m)
    rng = np.random.default_rng(5)
    to_look = rng.permutation(m).astype(np.int32)
    slots = np.sort(np.stack([rng.permutation(m)[:3] for _ in range(n)]), axis=1).astype(np.uint8)
    vals = rng.integers(0, 256, size=(n, 3)).astype(np.int16)
    Bo, _ = oracle.encoding_icm_sched(X, (B - 1).astype(np.int16), C, 3, to_look, slots, vals)
    Bg = gpu.encoding_icm_sched(X, B, C, 3, to_look, slots, vals)
    assert np.array_equal(Bg, Bo + 1)


def test_encoding_icm_empty_and_errors(gpu):
    X, C, B = make_problem(1, 0, 16, 4)
    assert gpu.encoding_icm(X, B, C, 2, True, 2).shape == (0, 4)
    X, C, B = make_problem(1, 10, 16, 4)
    bad = B.copy()
    bad[3, 1] = 257
    with pytest.raises(gpu.LsqError, match="1-based"):
        gpu.encoding_icm(X, bad, C, 2, True, 2)
    bad[3, 1] = 0
    with pytest.raises(gpu.LsqError, match="1-based"):
        gpu.encoding_icm(X, bad, C, 2, True, 2)


def test_encoding_icm_properties_config1(gpu, oracle):
    """BASELINE config 1 shape (10 K SIFT-like vectors, m = 8): size-independent properties, plus the
    bit-exact check against the oracle on the same inputs."""
    n, d, m = 10000, 128, 8
    X, C, B = make_problem(600, n, d, m)
    prev = gpu.veccost(X, B, C)
    cur = B
    for it in range(3):
        nxt = gpu.encoding_icm(X, cur, C, 4, True, 4, seed=1, ils_iter=it)
        cost = gpu.veccost(X, nxt, C)
        assert nxt.min() >= 1 and nxt.max() <= 256
        assert np.all(cost <= prev)
        changed = np.any(nxt != cur, axis=1)
        assert np.all(cost[changed] < prev[changed])
        cur, prev = nxt, cost
    Bo = (B - 1).astype(np.int16)
    for it in range(3):
        Bo, _ = oracle.encoding_icm(X, Bo, C, 4, True, 4, seed=1, ils_iter=it, nworkers=oracle.num_threads())
    assert np.array_equal(cur, Bo + 1)
    # ICM fixed point: with npert = 0 a converged code does not move, and never gets worse
    fix = cur
    for it in range(6):
        fix = gpu.encoding_icm(X, fix, C, 4, False, 0, seed=1, ils_iter=10 + it)
    again = gpu.encoding_icm(X, fix, C, 4, False, 0, seed=1, ils_iter=99)
    assert np.array_equal(again, fix)
    # sharding invariance: any split gives the same codes (perturbations keyed by global index)
    parts = gpu.splitarray(n, 3)
    sharded = np.concatenate([gpu.encoding_icm(X[lo:hi], B[lo:hi], C, 4, True, 4, seed=1, ils_iter=0, g0=lo)
                              for lo, hi in parts])
    whole = gpu.encoding_icm(X, B, C, 4, True, 4, seed=1, ils_iter=0)
    assert np.array_equal(sharded, whole)


# ---------------------------------------------------------------- encode_icm_cuda (a6)
def test_encode_icm_cuda_matches_looped_encoding_icm(gpu, oracle):
    n, d, m = 3000, 128, 8
    X, C, B = make_problem(700, n, d, m)
    Bs, objs = gpu.encode_icm_cuda(X, B, C, [2, 5], 4, 4, True, 3, seed=11)
    Bo, objo = oracle.encode_icm_ils(X, (B - 1).astype(np.int16), C, [2, 5], 4, 4, True, seed=11,
                                     nworkers=oracle.num_threads())
    assert np.array_equal(Bs[0], Bo[0] + 1) and np.array_equal(Bs[1], Bo[1] + 1)
    assert np.allclose(objs, objo, rtol=1e-6, atol=0)
    assert objs[1] <= objs[0]
    # nsplits only bounds memory: 1 split and 4 splits give the same codes
    Bs1, _ = gpu.encode_icm_cuda(X, B, C, [5], 4, 4, True, 1, seed=11)
    Bs4, _ = gpu.encode_icm_cuda(X, B, C, [5], 4, 4, True, 4, seed=11)
    assert np.array_equal(Bs1[0], Bs[1]) and np.array_equal(Bs4[0], Bs[1])
    # and equals looping the single-iteration call, as train_lsq does (LSQ.jl:45-48)
    cur = B
    for i in range(5):
        cur = gpu.encoding_icm(X, cur, C, 4, True, 4, seed=11, ils_iter=i)
    assert np.array_equal(cur, Bs[1])


# ---------------------------------------------------------------- linscan (a8, a9)
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "linscan_*.npz"))))
def test_linscan_golden_reference_vectors(gpu, path):
    g = np.load(path)
    n, nq, d, m, nn = (int(g[k]) for k in ("n", "nq", "d", "m", "nn"))
    codes, queries, codebooks, norms = make_scan_problem(int(g["seed"]), n, nq, d, m)
    if "pq" in os.path.basename(path):
        centers = codebooks[:, : d // m].reshape(m, 256, d // m).copy()
        dists, ids = gpu.linscan_pq(codes, queries, centers, 8 * m, nn)
        ids = ids.astype(np.int64) - 1   # the wrapper adds 1 like Linscan.jl:25
    else:
        dists, ids = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), norms, np.eye(d, dtype=np.float32), nn)
    assert np.array_equal(ids.astype(np.int64), g["ids"])
    assert np.array_equal(dists, g["dists"])


def _ref_or_oracle_lsq(oracle, *a):
    return oracle.ref_linscan_lsq(*a) if oracle.ref_available() else oracle.linscan_lsq(*a)


def _ref_or_oracle_pq(oracle, *a):
    return oracle.ref_linscan_pq(*a) if oracle.ref_available() else oracle.linscan_pq(*a)


@pytest.mark.parametrize("n,nq,d,m,nn", [
    (200000, 64, 128, 8, 1000),    # sampled-threshold path, config-5 shape scaled down
    (120000, 40, 128, 16, 1000),
    (100000, 33, 128, 7, 100),
    (90000, 50, 64, 9, 200),       # 24-query tiles (6 of 8 lanes per group)
    (90000, 37, 64, 12, 200),      # 16-query tiles, groups of 4 lanes
    (70000, 29, 32, 15, 100),      # 14-query tiles, two queries per lane
    (60000, 65, 32, 3, 50),        # 32-query tiles
    (50000, 9, 64, 8, 10000),      # large k -> exhaustive path + radix select
    (5000, 70, 32, 4, 1),
    (300, 5, 16, 2, 300),          # nn == n
    (40000, 21, 30, 8, 64),        # d % 4 != 0: scalar LUT build
    (40000, 3, 128, 1, 10),        # a single codebook
])
def test_linscan_lsq_exact(gpu, oracle, n, nq, d, m, nn):
    codes, queries, codebooks, norms = make_scan_problem(800 + m + nn, n, nq, d, m)
    dr, ir = _ref_or_oracle_lsq(oracle, codes, queries, codebooks, norms, nn)
    dg, ig = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), norms, np.eye(d, dtype=np.float32), nn)
    assert np.array_equal(ig, ir)
    assert np.array_equal(dg, dr)


@pytest.mark.parametrize("n,nq,d,m,nn", [(150000, 50, 128, 8, 1000), (40000, 20, 64, 16, 200), (2000, 3, 32, 8, 50),
                                         (30000, 9, 32, 16, 100)])   # subdim = 2: scalar LUT build
def test_linscan_pq_exact(gpu, oracle, n, nq, d, m, nn):
    codes, queries, codebooks, _ = make_scan_problem(900 + m, n, nq, d, m)
    centers = codebooks[:, : d // m].reshape(m, 256, d // m).copy()
    dr, ir = _ref_or_oracle_pq(oracle, codes, queries, centers, nn)
    dg, ig = gpu.linscan_pq(codes, queries, centers, 8 * m, nn)
    assert np.array_equal(ig.astype(np.int64) - 1, ir.astype(np.int64))
    assert np.array_equal(dg, dr)


def test_linscan_ties_and_adversarial_order(gpu, oracle):
    """Duplicates (ties -> lower id first) and a base set sorted so that every near neighbour sits at
    the END (a strided sample under-represents them: exercises the checked-threshold fallback)."""
    n, nq, d, m, nn = 60000, 12, 32, 8, 500
    codes, queries, codebooks, norms = make_scan_problem(1000, n, nq, d, m)
    codes[n // 2:] = codes[: n // 2]
    norms[n // 2:] = norms[: n // 2]
    dr, ir = _ref_or_oracle_lsq(oracle, codes, queries, codebooks, norms, nn)
    dg, ig = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), norms, np.eye(d, dtype=np.float32), nn)
    assert np.array_equal(ig, ir) and np.array_equal(dg, dr)
    norms2 = norms.copy()
    norms2[-600:] -= 1e6   # the last 600 vectors are by far the closest for every query
    dr, ir = _ref_or_oracle_lsq(oracle, codes, queries, codebooks, norms2, nn)
    dg, ig = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), norms2, np.eye(d, dtype=np.float32), nn)
    assert np.array_equal(ig, ir) and np.array_equal(dg, dr)


def test_linscan_reference_symbols_drop_in(gpu, oracle):
    """Call the two reference-named void symbols exactly as Linscan.jl's ccall does."""
    import ctypes as ct
    L = gpu.lib()
    codes, queries, codebooks, norms = make_scan_problem(1100, 30000, 8, 64, 8)
    nn = 100
    dists = np.zeros((8, nn), np.float32)
    idx = np.zeros((8, nn), np.int32)
    P = lambda a: a.ctypes.data_as(ct.c_void_p)
    L.linscan_aqd_query_extra_byte(P(dists), P(idx), P(codes), P(queries), P(codebooks), P(norms), 8, 30000, 8, 256, 64, nn)
    dr, ir = _ref_or_oracle_lsq(oracle, codes, queries, codebooks, norms, nn)
    assert np.array_equal(idx, ir) and np.array_equal(dists, dr)
    centers = codebooks[:, :8].reshape(8, 256, 8).copy()
    res = np.zeros((8, nn), np.uint32)
    L.linscan_aqd_query(P(dists), P(res), P(codes), P(centers), P(queries), 30000, ct.c_uint32(8), 64, nn, 8, 64, 8)
    dr, ir = _ref_or_oracle_pq(oracle, codes, queries, centers, nn)
    assert np.array_equal(res, ir) and np.array_equal(dists, dr)


# ---------------------------------------------------------------- update_codebooks (a7)
@pytest.mark.parametrize("n,d,m", [(20000, 32, 4), (6000, 128, 8)])
def test_update_codebooks(gpu, oracle, n, d, m):
    """Parity unpinned at this boundary (IterativeSolvers.jl is unvendored/unversioned): criterion is
    qerror(X, B, C_gpu) <= qerror(X, B, C_exact) * (1 + 1e-5), C_exact = float64 pseudo-inverse."""
    from oracle import codebook_update as cu
    X, C, B = make_problem(1200 + m, n, d, m)
    B[:, 0] = np.where(B[:, 0] == 17, 18, B[:, 0])  # make code 17 of codebook 0 unused
    B0 = (B - 1).astype(np.int16)
    Cg = gpu.update_codebooks(X, B, 256)
    Ce = cu.update_codebooks_exact(X, B0, 256)
    qg, qe = oracle.qerror(X, B0, Cg), oracle.qerror(X, B0, Ce)
    assert qg <= qe * (1 + 1e-5)
    assert qg <= oracle.qerror(X, B0, C)            # never worse than the codebooks we started from
    assert np.all(Cg[0, 16] == 0)                   # unused code -> zero codeword (min-norm solution)
    assert np.allclose(Cg, Ce, rtol=0, atol=2e-3 * np.abs(Ce).max())  # same min-norm gauge


def test_train_lsq_alternation_monotone(gpu):
    """The caller's loop (LSQ.jl:57-66): update_codebooks <-> encoding_icm never increases qerror."""
    X, C, B = make_problem(1300, 5000, 64, 4)
    C = gpu.update_codebooks(X, B, 256)
    objs = []
    for it in range(3):
        objs.append(gpu.qerror(X, B, C))
        C = gpu.update_codebooks(X, B, 256)
        assert gpu.qerror(X, B, C) <= objs[-1] * (1 + 1e-6)
        for i in range(2):
            B = gpu.encoding_icm(X, B, C, 4, True, 2, seed=3, ils_iter=2 * it + i)
    objs.append(gpu.qerror(X, B, C))
    assert all(b <= a * (1 + 1e-6) for a, b in zip(objs, objs[1:]))


# ---------------------------------------------------------------- both ICM kernels, forced
@pytest.fixture(params=["warp", "slice"])
def icm_kernel(request, monkeypatch):
    """LSQ_B200_ICM_KERNEL forces the warp-per-vector kernel or the shared-memory-slice kernel."""
    monkeypatch.setenv("LSQ_B200_ICM_KERNEL", request.param)
    return request.param


@pytest.mark.parametrize("n,d,m,niter,npert,randord", [
    (3000, 128, 8, 4, 4, True), (2500, 128, 7, 4, 4, True), (1111, 64, 4, 3, 2, False), (700, 32, 2, 2, 1, True),
    (5, 128, 8, 4, 4, True), (149, 16, 8, 1, 8, True),
])
def test_both_kernels_bit_exact_single_iteration(gpu, oracle, icm_kernel, n, d, m, niter, npert, randord):
    X, C, B = make_problem(1500 + n + m, n, d, m)
    Bo, _ = oracle.encoding_icm(X, (B - 1).astype(np.int16), C, niter, randord, npert, seed=21, ils_iter=2, g0=11)
    Bg = gpu.encoding_icm(X, B, C, niter, randord, npert, seed=21, ils_iter=2, g0=11)
    assert np.array_equal(Bg, Bo + 1)


def test_both_kernels_bit_exact_ils_with_snapshots(gpu, oracle, icm_kernel):
    n, d, m = 4000, 128, 8
    X, C, B = make_problem(1600, n, d, m)
    Bs, objs = gpu.encode_icm_cuda(X, B, C, [1, 3, 6], 4, 4, True, 2, seed=5, g0=3)
    Bo, objo = oracle.encode_icm_ils(X, (B - 1).astype(np.int16), C, [1, 3, 6], 4, 4, True, seed=5, g0=3,
                                     nworkers=oracle.num_threads())
    for r in range(3):
        assert np.array_equal(Bs[r], Bo[r] + 1)
    assert np.allclose(objs, objo, rtol=1e-6, atol=0)


def test_both_kernels_explicit_schedule(gpu, oracle, icm_kernel):
    n, d, m = 1500, 64, 8
    X, C, B = make_problem(1700, n, d, m)
    rng = np.random.default_rng(6)
    to_look = rng.permutation(m).astype(np.int32)
    slots = np.sort(np.stack([rng.permutation(m)[:4] for _ in range(n)]), axis=1).astype(np.uint8)
    vals = rng.integers(0, 256, size=(n, 4)).astype(np.int16)
    Bo, _ = oracle.encoding_icm_sched(X, (B - 1).astype(np.int16), C, 4, to_look, slots, vals)
    assert np.array_equal(gpu.encoding_icm_sched(X, B, C, 4, to_look, slots, vals), Bo + 1)


def test_slice_kernel_large_n_matches_warp_kernel(gpu, monkeypatch):
    """At a large size (every CTA of the slice kernel owns > 1000 vectors) both kernels give identical codes."""
    n, d, m = 200000, 128, 8
    X, C, B = make_problem(1800, n, d, m)
    monkeypatch.setenv("LSQ_B200_ICM_KERNEL", "warp")
    Bw, ow = gpu.encode_icm_cuda(X, B, C, [4], 4, 4, True, 1, seed=2)
    monkeypatch.setenv("LSQ_B200_ICM_KERNEL", "slice")
    Bs, os_ = gpu.encode_icm_cuda(X, B, C, [4], 4, 4, True, 1, seed=2)
    assert np.array_equal(Bw[0], Bs[0]) and ow[0] == os_[0]


# ---------------------------------------------------------------- tensor-core unaries (fast mode)
@pytest.mark.parametrize("kernel", ["pipelined", "serial"])
@pytest.mark.parametrize("n,d,m", [(3000, 128, 8), (1000, 64, 7), (129, 128, 16), (5, 8, 2), (70000, 128, 8), (64, 40, 3)])
def test_unaries_tensor_core_tolerance(gpu, oracle, monkeypatch, n, d, m, kernel):
    """tcgen05 3xTF32 build of the unary tables: NOT bit-exact by construction; |dU| <= 4e-6 * max|U|.
    Both kernels: the warp-specialised pipeline (codebook operand in TMEM, 3 shared-memory stages, two
    accumulators; n = 70000 makes every CTA walk many tiles and wrap all its barriers) and the serial one."""
    import ctypes as ct
    import torch
    monkeypatch.setenv("LSQ_B200_UNARY_TC", kernel)
    X, C, _ = make_problem(1900 + n, n, d, m, kind="gauss")
    X *= 37.0
    Xd, Cd = torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda()
    U = torch.full((m, n, 256), float("nan"), dtype=torch.float32, device="cuda")
    L = gpu.lib()
    P = lambda t: ct.c_void_p(t.data_ptr())
    rc = L.lsq_dev_build_unaries_tc(P(Xd), d, ct.c_int64(n), P(Cd), m, P(U), ct.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, L.lsq_last_error()
    torch.cuda.synchronize()
    Uo = oracle.get_unaries(X, C)
    err = np.abs(U.cpu().numpy() - Uo).max() / np.abs(Uo).max()
    assert err <= 4e-6, err


def test_tensor_core_unaries_keep_quantisation_error(gpu, oracle, monkeypatch):
    """Fast mode criterion (BASELINE north_star): quantisation error within 1e-5 relative of the parity path."""
    n, d, m = 20000, 128, 8
    X, C, B = make_problem(2000, n, d, m, kind="gauss")
    X *= 25.0
    C *= 25.0
    Bs0, o0 = gpu.encode_icm_cuda(X, B, C, [4], 4, 4, True, 1, seed=3)
    monkeypatch.setenv("LSQ_B200_UNARY", "tc")
    Bs1, o1 = gpu.encode_icm_cuda(X, B, C, [4], 4, 4, True, 1, seed=3)
    assert abs(o1[0] - o0[0]) <= 1e-5 * o0[0]
    assert np.mean(np.any(Bs0[0] != Bs1[0], axis=1)) < 1e-3


# ---------------------------------------------------------------- device-resident training loop
def test_train_lsq_sharded_single_gpu_matches_host_loop(gpu):
    """parallel.train_lsq_sharded (world size 1: X / codes / tables resident, no re-uploads) reproduces the
    caller-side loop of LSQ.jl:57-66 driven through the host API."""
    import torch
    n, d, m = 12000, 64, 8
    X, C, B = make_problem(2100, n, d, m)
    Xd = torch.from_numpy(X).cuda()
    cd = torch.from_numpy((B - 1).astype(np.uint8)).cuda()
    C1, codes, obj = gpu.parallel.train_lsq_sharded(Xd, cd, torch.from_numpy(C).cuda(), 2, 3, 4, True, 4, seed=9)
    Bc, Cc = B.copy(), C
    for it in range(2):
        Cc = gpu.update_codebooks(X, Bc, 256)
        for i in range(3):
            Bc = gpu.encoding_icm(X, Bc, Cc, 4, True, 4, seed=9, ils_iter=3 * it + i)
    assert np.array_equal(codes.cpu().numpy().astype(np.int16) + 1, Bc)
    q = gpu.qerror(X, Bc, Cc)
    assert abs(obj[-1] - q) <= 1e-5 * q
    assert obj[-1] <= obj[0]


# ---------------------------------------------------------------- §8(f): train_lsq, norm codebook, eval_recall
@pytest.mark.parametrize("use_R", [False, True])
def test_train_lsq_equals_public_calls(gpu, use_R):
    """lsq_train_lsq (everything resident on the GPU) == the caller-side loop of LSQ.jl:31-66 driven through
    the public host calls with ils_iter = 0, 1, 2, ... — codes, codebooks and objective bit for bit."""
    n, d, m, niter, ilsiter = 9000, 64, 8, 2, 3
    X, C, B = make_problem(2300, n, d, m)
    R = None
    if use_R:
        R = np.linalg.qr(np.random.default_rng(1).standard_normal((d, d)))[0].astype(np.float32)
    C1, B1, cbn, Bn, obj = gpu.train_lsq(X, m, 256, R, B, C, niter, ilsiter, 4, True, 4, seed=21)
    Bc = B.copy()
    if use_R:
        Cc = gpu.update_codebooks((X @ R).astype(np.float32), Bc, 256)
    else:
        Cc = gpu.update_codebooks(X, Bc, 256)
    count = 0
    objs = []
    if not use_R:   # with a rotation the rotated-back codebooks differ from a host matmul in the last bit
        for it in range(niter + 1):
            if it > 0:
                objs.append(gpu.qerror(X, Bc, Cc))
                Cc = gpu.update_codebooks(X, Bc, 256)
            for i in range(ilsiter):
                Bc = gpu.encoding_icm(X, Bc, Cc, 4, True, 4, seed=21, ils_iter=count)
                count += 1
        assert np.array_equal(B1, Bc)
        assert np.array_equal(C1, Cc)
        assert np.array_equal(obj, np.asarray(objs, np.float32))
    assert all(b <= a * (1 + 1e-6) for a, b in zip(obj, obj[1:]))
    assert gpu.qerror(X, B1, C1) <= obj[-1] * (1 + 1e-6)
    # norm codebook: ascending centres; B_norms is exactly quantize_norms(B, C, cbnorms) (utils.jl:6-31)
    assert np.all(np.diff(cbn) >= 0)
    assert np.array_equal(Bn, gpu.quantize_norms(B1, C1, cbn))


def test_train_lsq_fast_mode_equals_public_calls(gpu, monkeypatch):
    """LSQ_B200_UNARY=tc (tensor-core unaries) applies to the resident training loop and to the host encode calls
    alike: in fast mode the one-call train_lsq still equals the loop of public calls bit for bit, and its
    objective stays within the north-star's 1e-5 of the exact mode (Gaussian data, so the modes really differ)."""
    n, d, m, niter, ilsiter = 8000, 128, 8, 2, 2
    X, C, B = make_problem(2350, n, d, m, kind="gauss")
    X *= 20.0
    _, _, _, _, obj_exact = gpu.train_lsq(X, m, 256, None, B, None, niter, ilsiter, 4, True, 4, seed=5)
    monkeypatch.setenv("LSQ_B200_UNARY", "tc")
    C1, B1, _, _, obj = gpu.train_lsq(X, m, 256, None, B, None, niter, ilsiter, 4, True, 4, seed=5)
    Bc, Cc, count = B.copy(), gpu.update_codebooks(X, B, 256), 0
    for it in range(niter + 1):
        if it > 0:
            Cc = gpu.update_codebooks(X, Bc, 256)
        for i in range(ilsiter):
            Bc = gpu.encoding_icm(X, Bc, Cc, 4, True, 4, seed=5, ils_iter=count)
            count += 1
    assert np.array_equal(B1, Bc) and np.array_equal(C1, Cc)
    assert np.allclose(obj, obj_exact, rtol=1e-5, atol=0)


def test_train_lsq_vs_oracle_objective(gpu, oracle):
    """Against the restated train_lsq (oracle): the codebook update is parity-unpinned (IterativeSolvers),
    so codes may part ways after the first update; the objective trajectory must agree to 1e-3 relative."""
    n, d, m = 3000, 32, 4
    X, C, B = make_problem(2400, n, d, m)
    R = np.linalg.qr(np.random.default_rng(2).standard_normal((d, d)))[0].astype(np.float32)
    C1, B1, cbn, Bn, obj = gpu.train_lsq(X, m, 256, R, B, C, 3, 2, 4, True, 2, seed=5)
    Co, Bo, cbo, Bno, objo = oracle.train_lsq(X, m, 256, R, (B - 1).astype(np.int16), 3, 2, 4, True, 2, seed=5)
    assert np.allclose(obj, objo, rtol=1e-3)
    assert abs(gpu.qerror(X, B1, C1) - oracle.qerror(X, Bo, Co)) <= 1e-3 * oracle.qerror(X, Bo, Co)


@pytest.mark.parametrize("n,h", [(20000, 256), (5000, 16), (300, 256), (100, 7)])
def test_kmeans1d_matches_oracle(gpu, oracle, n, h):
    """Same deterministic Lloyd iteration on both sides; float64 means are summed in a different (fixed)
    order on the GPU, so centres agree to fp32 rounding, and the iteration counts agree."""
    rng = np.random.default_rng(n + h)
    v = (rng.standard_normal(n) ** 2 * 1000).astype(np.float32)
    cg, itg = gpu.kmeans1d(v, h)
    co, ito = oracle.kmeans1d(v, h)
    assert np.allclose(cg, co, rtol=1e-5, atol=0)
    assert abs(itg - ito) <= 1 or max(itg, ito) == 100
    assert np.all(np.diff(cg) >= 0)


def test_eval_recall_matches_oracle(gpu, oracle):
    rng = np.random.default_rng(3)
    nq, k = 500, 1000
    pred = np.stack([rng.permutation(5000)[:k] for _ in range(nq)]).astype(np.int32) + 1
    gt = np.where(rng.random(nq) < 0.7, pred[np.arange(nq), rng.integers(0, k, nq)], 999999).astype(np.int32)
    pred[7, 3] = pred[7, 9] = gt[7]   # found twice -> counts as a miss (Linscan.jl:94-98)
    for kk in (1, 10, 1000):
        assert np.array_equal(gpu.eval_recall(gt, pred, kk), oracle.eval_recall(gt, pred, kk))
    assert np.array_equal(gpu.eval_recall(gt.astype(np.uint32), pred.astype(np.uint32), 100),
                          oracle.eval_recall(gt, pred, 100))


def test_eval_recall_searches_the_whole_list(gpu, oracle):
    """Linscan.jl:91 runs find() over the WHOLE column: an id inside the top k that occurs again beyond
    position k is a miss, and an id found once beyond k is a miss too."""
    rng = np.random.default_rng(4)
    nq, ld, k = 64, 200, 50
    pred = np.stack([rng.permutation(10000)[:ld] for _ in range(nq)]).astype(np.int32) + 1
    gt = pred[np.arange(nq), rng.integers(0, k, nq)].copy()
    pred[3, 120] = gt[3]            # second occurrence beyond k -> miss in the reference
    gt[5] = pred[5, 150]            # only occurrence beyond k -> miss
    rg, ro = gpu.eval_recall(gt, pred, k), oracle.eval_recall(gt, pred, k)
    assert np.array_equal(rg, ro)
    assert rg[-1] == (nq - 2) / nq


def test_encoding_icm_default_iteration_counter(gpu, oracle):
    """The reference loop `for i = 1:ilsiter; B = encoding_icm(X, B, C, ...)` (LSQ.jl:45-48) ported verbatim:
    with ils_iter left at its default every call must draw a fresh perturbation schedule (iterations
    0, 1, 2, ... of the Philox stream), not replay the first one."""
    X, C, B = make_problem(4100, 3000, 32, 8)
    gpu.reset_ils_counter()
    Bd = B
    for _ in range(3):
        Bd = gpu.encoding_icm(X, Bd, C, 2, True, 4, seed=11)
    Be = B
    for i in range(3):
        Be = gpu.encoding_icm(X, Be, C, 2, True, 4, seed=11, ils_iter=i)
    assert np.array_equal(Bd, Be)
    Bo = (B - 1).astype(np.int16)
    for i in range(3):
        Bo, _ = oracle.encoding_icm(X, Bo, C, 2, True, 4, seed=11, ils_iter=i)
    assert np.array_equal(Bd, Bo + 1)
    Br = B
    for _ in range(3):   # what the old default (always iteration 0) did: a different, weaker result
        Br = gpu.encoding_icm(X, Br, C, 2, True, 4, seed=11, ils_iter=0)
    assert not np.array_equal(Bd, Br)


def test_pageable_and_pinned_host_buffers_agree(gpu):
    """Julia arrays are pageable: the staged host->device copy must deliver the same bytes as a pinned source."""
    import torch
    X, C, B = make_problem(4200, 70000, 128, 8)
    its = [2]
    Bp, op = gpu.encode_icm_cuda(X, B, C, its, 2, 4, True, 1, seed=3)
    Xp, Bpin = torch.from_numpy(X).pin_memory(), torch.from_numpy(B).pin_memory()
    Bq, oq = gpu.encode_icm_cuda(Xp.numpy(), Bpin.numpy(), C, its, 2, 4, True, 1, seed=3)
    assert np.array_equal(Bp[0], Bq[0]) and np.array_equal(op, oq)


# ---------------------------------------------------------------- §8(f4): chain (Viterbi) encoder
@pytest.mark.parametrize("n,d,m,kind", [(2000, 128, 8, "sift"), (777, 64, 7, "sift"), (501, 32, 2, "gauss"),
                                        (300, 16, 16, "sift"), (1, 128, 8, "gauss"), (3, 24, 3, "sift")])
def test_encoding_viterbi_bit_exact(gpu, oracle, n, d, m, kind):
    """encode_chain.jl:1-127 — deterministic, so the codes must equal the oracle's bit for bit (sift-like
    quarter-integer codebooks make exact ties plausible: first minimum wins)."""
    X, C, B = make_problem(3000 + n + m, n, d, m, kind=kind)
    Bg = gpu.encoding_viterbi(X, C)
    Bo = oracle.encoding_viterbi(X, C)
    assert np.array_equal(Bg, Bo + 1)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "viterbi_*.npz"))))
def test_encoding_viterbi_golden(gpu, path):
    g = np.load(path)
    X, C, _ = make_problem(int(g["seed"]), int(g["n"]), int(g["d"]), int(g["m"]), kind=str(g["kind"]))
    assert np.array_equal(gpu.encoding_viterbi(X, C), g["codes"])


@pytest.mark.parametrize("kernel", ["simple", "tma"])
@pytest.mark.parametrize("n,d,m", [(777, 32, 8), (64, 16, 2), (1, 16, 5), (4099, 24, 16)])
def test_encoding_viterbi_both_kernels(gpu, oracle, monkeypatch, kernel, n, d, m):
    """LSQ_B200_VITERBI forces the warp-per-4-vectors kernel or the TMA-pipelined persistent kernel (ragged
    tails, fewer vectors than one CTA pass): both must give the oracle's codes."""
    monkeypatch.setenv("LSQ_B200_VITERBI", kernel)
    X, C, B = make_problem(3200 + n + m, n, d, m)
    assert np.array_equal(gpu.encoding_viterbi(X, C), oracle.encoding_viterbi(X, C) + 1)


def test_encoding_viterbi_large_n_default_dispatch(gpu, oracle):
    """n >= 148 x 64 takes the TMA-pipelined kernel by default."""
    X, C, B = make_problem(3300, 12345, 32, 8)
    assert np.array_equal(gpu.encoding_viterbi(X, C), oracle.encoding_viterbi(X, C) + 1)


def test_encoding_viterbi_properties(gpu, oracle):
    """Exact chain MAP: no single-node change (what ICM tries) can lower the CHAIN energy, the result does
    not depend on how the input is batched, and it beats random codes on the true objective."""
    n, d, m = 4000, 64, 6
    X, C, B = make_problem(3100, n, d, m)
    Bv = gpu.encoding_viterbi(X, C)
    assert np.array_equal(gpu.encoding_viterbi(X[1000:1700], C), Bv[1000:1700])
    assert gpu.qerror(X, Bv, C) < gpu.qerror(X, B, C)
    U = oracle.get_unaries(X[:50], C)
    G, cbi = oracle.get_binaries(C)
    pair = {(int(a), int(b)): G[i] for i, (a, b) in enumerate(cbi)}   # G[idx][b][a], (i<j) 0-based
    def energy(codes):
        e = sum(float(U[i, v, codes[i]]) for i in range(m))
        return e + sum(float(pair[(i, i + 1)][codes[i + 1], codes[i]]) for i in range(m - 1))
    rng = np.random.default_rng(0)
    for v in range(50):
        base = (Bv[v] - 1).astype(int)
        e0 = energy(base)
        for _ in range(20):
            alt = base.copy()
            alt[rng.integers(m)] = rng.integers(256)
            assert energy(alt) >= e0 - 1e-3 * abs(e0)


def test_encoding_viterbi_errors(gpu):
    X, C, B = make_problem(1, 10, 16, 2)
    with pytest.raises(gpu.LsqError):
        gpu.encoding_viterbi(X, C[:1])          # a chain needs two nodes
    assert gpu.encoding_viterbi(X[:0], C).shape == (0, 2)


def test_device_api_viterbi_and_recall(gpu, oracle):
    """lsq_dev_viterbi / lsq_dev_eval_recall on torch-owned device buffers == the host-pointer calls."""
    import torch
    X, C, B = make_problem(3400, 5000, 32, 6)
    codes = gpu.device.viterbi(torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda())
    assert np.array_equal(codes.cpu().numpy().astype(np.int16) + 1, gpu.encoding_viterbi(X, C))
    rng = np.random.default_rng(0)
    pred = np.stack([rng.permutation(2000)[:100] for _ in range(64)]).astype(np.int32)
    gt = pred[np.arange(64), rng.integers(0, 100, 64)].astype(np.int32)
    r = gpu.device.eval_recall(torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda(), 100)
    assert np.array_equal(r.cpu().numpy(), oracle.eval_recall(gt, pred, 100))


def test_memory_bounded_chunking_does_not_change_results(gpu, monkeypatch):
    """train_lsq and encoding_viterbi walk the data in chunks when the unaries would not fit; the chunk size
    is forced small here (LSQ_B200_CHUNK_VECTORS) and nothing may change: vectors are independent and the
    perturbation stream is keyed by the global vector index."""
    X, C, B = make_problem(3500, 7001, 32, 5)
    ref_train = gpu.train_lsq(X, 5, 256, None, B, None, 2, 2, 3, True, 2, seed=8)
    ref_vit = gpu.encoding_viterbi(X, C)
    monkeypatch.setenv("LSQ_B200_CHUNK_VECTORS", "1234")
    got = gpu.train_lsq(X, 5, 256, None, B, None, 2, 2, 3, True, 2, seed=8)
    for a, b in zip(ref_train, got):
        assert np.array_equal(a, b)
    assert np.array_equal(gpu.encoding_viterbi(X, C), ref_vit)


def test_c_level_dropin_against_reference_so(gpu, oracle, tmp_path):
    """examples/dropin_linscan.c: a plain C program dlopen()s liblsq_b200.so and the reference's own
    linscan .so, calls the SAME symbol with the SAME arguments in both, and compares the outputs bit for bit
    — the drop-in boundary exercised without Python or Julia in between."""
    import shutil
    import subprocess
    if not oracle.ref_available() or shutil.which("gcc") is None:
        pytest.skip("needs oracle/_ref (built from the reference) and gcc")
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "dropin_linscan")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(root, "examples", "dropin_linscan.c"), "-ldl"], check=True)
    r = subprocess.run([exe, gpu.lib_path(), os.path.join(root, "oracle", "_ref", "linscan_aqd_pairwise_byte.so")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ids identical, distances identical" in r.stdout


def test_edge_cases_empty_and_degenerate(gpu, oracle):
    """Empty / degenerate inputs across the boundary: nothing to scan, nothing to return, zero outer
    iterations, snapshot slots that no iteration fills (they must come back zeroed even if the caller's
    buffer held garbage, like the reference's preallocated Bs)."""
    import ctypes as ct
    codes, queries, codebooks, norms = make_scan_problem(9, 5000, 4, 16, 4)
    R = np.eye(16, dtype=np.float32)
    d0, i0 = gpu.linscan_lsq(codes, queries[:0], codebooks.reshape(4, 256, 16), norms, R, 10)
    assert d0.shape == (0, 10) and i0.shape == (0, 10)
    d0, i0 = gpu.linscan_lsq(codes, queries, codebooks.reshape(4, 256, 16), norms, R, 0)
    assert d0.shape == (4, 0)
    with pytest.raises(gpu.LsqError, match="nn"):
        gpu.linscan_lsq(codes[:5], queries, codebooks.reshape(4, 256, 16), norms[:5], R, 6)   # nn > ncodes
    # train_lsq with zero outer iterations: initial codebook update + ilsiter encodes only
    X, C, B = make_problem(3600, 800, 16, 3)
    Ct, Bt, cbn, Bn, obj = gpu.train_lsq(X, 3, 256, None, B, None, 0, 2, 2, True, 2, seed=3)
    assert obj.shape == (0,) and Bt.min() >= 1 and Bt.max() <= 256
    Cc = gpu.update_codebooks(X, B, 256)
    Bc = B
    for i in range(2):
        Bc = gpu.encoding_icm(X, Bc, Cc, 2, True, 2, seed=3, ils_iter=i)
    assert np.array_equal(Bt, Bc) and np.array_equal(Ct, Cc)
    # snapshots: [2, 2, 0] -> the first slot holds iteration 2, the duplicate and the zero stay all-zero
    n, m = X.shape[0], 3
    its = np.array([2, 2, 0], np.int64)
    Bs = np.full((3, n, m), 12345, np.int16)          # garbage the library must not leave behind
    objs = np.full(3, -1.0, np.float32)
    P = lambda a: a.ctypes.data_as(ct.c_void_p)
    rc = gpu.lib().lsq_encode_icm_cuda(P(X), 16, ct.c_int64(n), P(B), P(np.ascontiguousarray(C)), m, 256, P(its), 3, 2, 2, 1, 1,
                                       ct.c_uint64(3), ct.c_uint64(0), P(Bs), P(objs), 0)
    assert rc == 0
    ref, _ = gpu.encode_icm_cuda(X, B, C, [2], 2, 2, True, 1, seed=3)
    assert np.array_equal(Bs[0], ref[0]) and not Bs[1].any() and not Bs[2].any()
    assert objs[0] > 0 and objs[1] == 0 and objs[2] == 0
    # eval_recall at k = 1
    assert np.array_equal(gpu.eval_recall(np.array([5, 6]), np.array([[5, 1], [2, 6]]), 1), oracle.eval_recall([5, 6], np.array([[5, 1], [2, 6]]), 1))
