"""CPU suite (-m "not gpu"): the oracle against the golden vectors, the twin and the reference's own
linscan; host logic of the library; the C ABI's symbol table.  No GPU compute anywhere here."""
import ctypes as ct
import glob
import os
import re
import subprocess

import numpy as np
import pytest

from util import make_problem, make_scan_problem

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")


def test_philox_known_answers(oracle):
    # Random123 kat_vectors: philox4x32 10 rounds
    assert [hex(x) for x in oracle.philox([0, 0, 0, 0], [0, 0])] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    assert [hex(x) for x in oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(x) for x in oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "icm_*.npz"))))
def test_oracle_matches_golden_icm(oracle, path):
    g = np.load(path)
    X, C, B1 = make_problem(int(g["seed"]), int(g["n"]), int(g["d"]), int(g["m"]))
    B = (B1 - 1).astype(np.int16)
    for it in range(int(g["iters"])):
        B, cost = oracle.encoding_icm(X, B, C, int(g["niter"]), bool(g["randord"]), int(g["npert"]),
                                      seed=int(g["seed"]), ils_iter=it)
        assert np.array_equal(B + 1, g["codes"][it])
        assert np.array_equal(cost, g["cost"][it])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "linscan_*.npz"))))
def test_oracle_matches_golden_linscan(oracle, path):
    """Golden vectors were produced by the reference's own C++ (oracle/_ref)."""
    g = np.load(path)
    n, nq, d, m, nn = (int(g[k]) for k in ("n", "nq", "d", "m", "nn"))
    codes, queries, codebooks, norms = make_scan_problem(int(g["seed"]), n, nq, d, m)
    if "pq" in os.path.basename(path):
        centers = codebooks[:, : d // m].reshape(m, 256, d // m).copy()
        dists, ids = oracle.linscan_pq(codes, queries, centers, nn)
    else:
        dists, ids = oracle.linscan_lsq(codes, queries, codebooks, norms, nn)
    assert np.array_equal(ids.astype(np.int64), g["ids"])
    assert np.array_equal(dists, g["dists"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "viterbi_*.npz"))))
def test_oracle_matches_golden_viterbi(oracle, path):
    g = np.load(path)
    X, C, _ = make_problem(int(g["seed"]), int(g["n"]), int(g["d"]), int(g["m"]), kind=str(g["kind"]))
    assert np.array_equal(oracle.encoding_viterbi(X, C) + 1, g["codes"])


def test_oracle_vs_twin_viterbi(oracle):
    """encode_chain.jl:1-127: the C restatement and the NumPy twin agree, and the result is the exact chain
    optimum (brute force over all 256^2 / 4^3 configurations on tiny problems)."""
    from oracle import np_twin
    X, C, _ = make_problem(5, 30, 12, 4)
    assert np.array_equal(oracle.encoding_viterbi(X, C), np_twin.encoding_viterbi(X, C))
    X, C, _ = make_problem(6, 6, 8, 2, kind="gauss")
    U = oracle.get_unaries(X, C)
    G, _ = oracle.get_binaries(C)          # G[0][b][a] = 2<C_0[:,a], C_1[:,b]>
    Bv = oracle.encoding_viterbi(X, C)
    for v in range(6):
        E = U[0, v][:, None] + U[1, v][None, :] + G[0].T   # E[a][b]
        a, b = np.unravel_index(np.argmin(E), E.shape)
        assert abs(E[Bv[v, 0], Bv[v, 1]] - E[a, b]) <= 1e-4 * abs(E[a, b])


def test_oracle_vs_twin_icm(oracle):
    from oracle import np_twin
    X, C, B1 = make_problem(3, 20, 24, 5)
    B = (B1 - 1).astype(np.int16)
    a, ca = oracle.encoding_icm(X, B, C, 2, True, 3, seed=77, ils_iter=4, g0=1000)
    b, cb = np_twin.encoding_icm(X, B, C, 2, True, 3, seed=77, ils_iter=4, g0=1000)
    assert np.array_equal(a, b) and np.array_equal(ca, cb)
    assert np.array_equal(oracle.get_unaries(X, C), np_twin.get_unaries(X, C))


def test_oracle_vs_reference_linscan_live(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    from oracle import np_twin
    codes, queries, codebooks, norms = make_scan_problem(5, 4000, 7, 48, 8)
    r = oracle.ref_linscan_lsq(codes, queries, codebooks, norms, 64)
    o = oracle.linscan_lsq(codes, queries, codebooks, norms, 64)
    t = np_twin.linscan_lsq(codes, queries, codebooks, norms, 64)
    assert np.array_equal(r[0], o[0]) and np.array_equal(r[1], o[1])
    assert np.array_equal(r[0], t[0]) and np.array_equal(r[1], t[1])
    centers = codebooks[:, :6].reshape(8, 256, 6).copy()
    r = oracle.ref_linscan_pq(codes, queries, centers, 33)
    o = oracle.linscan_pq(codes, queries, centers, 33)
    assert np.array_equal(r[0], o[0]) and np.array_equal(r[1], o[1])


def test_oracle_linscan_ties_lowest_id(oracle):
    # duplicated code rows -> equal distances -> the smaller id must come first (partial_sort on pairs)
    codes, queries, codebooks, norms = make_scan_problem(6, 600, 3, 16, 4)
    codes[300:] = codes[:300]
    norms[300:] = norms[:300]
    d_, ids = oracle.linscan_lsq(codes, queries, codebooks, norms, 50)
    for q in range(3):
        for j in range(49):
            assert (d_[q, j], ids[q, j]) < (d_[q, j + 1], ids[q, j + 1])


def test_oracle_properties(oracle):
    X, C, B1 = make_problem(8, 300, 32, 8)
    B = (B1 - 1).astype(np.int16)
    prev = oracle.veccost(X, B, C)
    for it in range(3):
        B2, cost = oracle.encoding_icm(X, B, C, 4, True, 4, seed=1, ils_iter=it)
        assert np.all(cost <= prev)                       # monotone per vector (encode_icm.jl:183-186)
        assert np.array_equal(cost, oracle.veccost(X, B2, C))
        assert B2.min() >= 0 and B2.max() < 256
        changed = np.any(B2 != B, axis=1)
        assert np.all(cost[changed] < prev[changed])      # a changed code is strictly better
        B, prev = B2, cost
    # sharding invariance: vectors keyed by global index
    whole, _ = oracle.encoding_icm(X, B, C, 2, True, 4, seed=9, ils_iter=0)
    lo, hi = oracle.splitarray(300, 3)[1]
    part, _ = oracle.encoding_icm(X[lo:hi], B[lo:hi], C, 2, True, 4, seed=9, ils_iter=0, g0=lo)
    assert np.array_equal(whole[lo:hi], part)
    # worker fan-out does not change results
    w4, _ = oracle.encoding_icm(X, B, C, 2, True, 4, seed=9, ils_iter=0, nworkers=4)
    assert np.array_equal(whole, w4)


def test_oracle_ils_snapshots(oracle):
    X, C, B1 = make_problem(9, 64, 16, 4)
    B = (B1 - 1).astype(np.int16)
    Bs, objs = oracle.encode_icm_ils(X, B, C, [1, 3], 2, 2, True, seed=3)
    cur = B
    for i in range(3):
        cur, _ = oracle.encoding_icm(X, cur, C, 2, True, 2, seed=3, ils_iter=i)
        if i == 0:
            assert np.array_equal(Bs[0], cur)
    assert np.array_equal(Bs[1], cur)
    assert abs(objs[1] - oracle.qerror(X, cur, C)) <= 1e-6 * objs[1]
    assert objs[1] <= objs[0]


def test_codebook_update_oracles_agree():
    from oracle import codebook_update as cu
    import oracle as orc
    X, C, B1 = make_problem(10, 3000, 16, 3)
    B = (B1 - 1).astype(np.int16)
    Ce = cu.update_codebooks_exact(X, B, 256)
    Cl = cu.update_codebooks_lsqr(X, B, 256)
    qe, ql = orc.qerror(X, B, Ce), orc.qerror(X, B, Cl)
    assert qe <= ql * (1 + 1e-6)          # the exact LS optimum is never worse than LSQR's iterate
    assert ql <= qe * (1 + 1e-3)          # and LSQR at sqrt(eps) tolerance is close to it
    # unused codes get all-zero codewords (min-norm solution)
    used = np.zeros((3, 256), bool)
    for i in range(3):
        used[i, np.unique(B[:, i])] = True
    assert np.all(Ce[~used] == 0)


def test_splitarray_rule(oracle, lsq):
    for n, p in [(10, 3), (7, 7), (5, 8), (0, 2), (1000003, 8)]:
        parts = oracle.splitarray(n, p)
        assert parts == lsq.splitarray(n, p)
        assert parts[0][0] == 0 and parts[-1][1] == n
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_host_schedule_matches_oracle(oracle, lsq):
    for seed, it, m, npert in [(0, 0, 8, 4), (123456789012345, 7, 16, 7), (5, 2, 7, 7), (1, 1, 3, 0), (2, 9, 1, 1)]:
        assert np.array_equal(oracle.make_to_look(seed, it, m, True), lsq.make_to_look(seed, it, m, True))
        assert np.array_equal(lsq.make_to_look(seed, it, m, False), np.arange(m))
        s1, v1 = oracle.make_perturb(seed, it, 2**33 + 5, 200, m, 256, npert)
        s2, v2 = lsq.make_perturb(seed, it, 2**33 + 5, 200, m, 256, npert)
        assert np.array_equal(s1, s2) and np.array_equal(v1, v2)
        if npert:
            assert np.all(np.diff(s1.astype(int), axis=1) > 0)  # distinct + ascending
            assert s1.max() < m and v1.min() >= 0 and v1.max() < 256


def test_abi_exports_every_declared_symbol(lsq):
    """include/lsq_b200.h <-> liblsq_b200.so <-> EXPORTED_SYMBOLS must agree; loading needs no GPU."""
    header = open(os.path.join(ROOT, "include", "lsq_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b((?:lsq|linscan)_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lsq.EXPORTED_SYMBOLS)
    L = lsq.lib()
    for s in declared:
        assert hasattr(L, s), f"{s} not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", lsq.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T ((?:lsq|linscan)_[a-z0-9_]+)\b", out))
    assert declared <= exported


def test_argument_errors_without_gpu(lsq):
    """Argument validation happens before any device work and mirrors the reference's errors."""
    X = np.zeros((4, 8), np.float32)
    B = np.ones((4, 2), np.int16)
    C = np.zeros((2, 256, 8), np.float32)
    with pytest.raises(lsq.LsqError, match="Codebook update method unknown"):   # codebook_update.jl:59
        lsq.update_codebooks(X, B, 256, False, "cholesky")
    with pytest.raises(lsq.LsqError, match="h must be 256"):
        lsq.encoding_icm(X, B, np.zeros((2, 128, 8), np.float32), 1, True, 1)
    with pytest.raises(lsq.LsqError, match="npert"):
        lsq.encoding_icm(X, B, C, 1, True, 3)
    with pytest.raises(lsq.LsqError, match="m must be in 1..16"):
        lsq.encoding_icm(X, np.ones((4, 17), np.int16), np.zeros((17, 256, 8), np.float32), 1, True, 1)
    with pytest.raises(TypeError):
        lsq.encoding_icm(X, B.astype(np.int32), C, 1, True, 1)
    # the entry points added for SURVEY §8(f): train_lsq, chain encoder, norm codebook, eval_recall
    with pytest.raises(lsq.LsqError, match="npert"):
        lsq.train_lsq(X, 2, 256, None, B, None, 1, 1, 1, True, 3)
    with pytest.raises(lsq.LsqError, match="h must be 256"):
        lsq.train_lsq(X, 2, 128, None, B, None, 1, 1, 1, True, 1)
    with pytest.raises(lsq.LsqError, match="chain needs two nodes"):
        lsq.encoding_viterbi(X, C[:1])
    with pytest.raises(lsq.LsqError, match="kmeans1d"):
        lsq.kmeans1d(np.zeros(0, np.float32), 4)
    with pytest.raises(lsq.LsqError, match="eval_recall"):
        lsq.eval_recall(np.arange(3), np.zeros((3, 5), np.int32), 9)      # k > row length


def test_no_cpu_fallback(lsq):
    """Without a GPU the product must fail loudly, never compute on the host."""
    if lsq.device_count() > 0:
        pytest.skip("GPU present")
    X, C, B = make_problem(1, 8, 8, 2)
    with pytest.raises(lsq.LsqError, match="no CUDA device"):
        lsq.encoding_icm(X, B, C, 1, True, 1)
    with pytest.raises(lsq.LsqError, match="no CUDA device"):
        lsq.update_codebooks(X, B, 256)
    with pytest.raises(lsq.LsqError, match="no CUDA device"):
        lsq.train_lsq(X, 2, 256, None, B, None, 1, 1, 1, True, 1)
    with pytest.raises(lsq.LsqError, match="no CUDA device"):
        lsq.encoding_viterbi(X, C)
    with pytest.raises(lsq.LsqError, match="no CUDA device"):
        lsq.eval_recall(np.arange(3), np.zeros((3, 5), np.int32), 5)


def test_product_never_imports_oracle():
    for path in glob.glob(os.path.join(ROOT, "local-search-quantization_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
            src = open(path, errors="ignore").read()
            assert "import oracle" not in src and "from oracle" not in src and "liblsq_oracle" not in src, path


def test_eval_recall(oracle):
    """Linscan.jl:76-117 restated: rank of the ground-truth id, duplicates count as misses (:94-98)."""
    gt = np.array([3, 9, 5])
    pred = np.array([[3, 1, 2], [1, 9, 2], [7, 8, 6]])
    assert np.allclose(oracle.eval_recall(gt, pred, 3), [1 / 3, 2 / 3, 2 / 3])
    pred[1] = [9, 9, 2]   # found twice -> `length(nn_pos) == 1` fails -> miss
    assert np.allclose(oracle.eval_recall(gt, pred, 3), [1 / 3, 1 / 3, 1 / 3])


def test_oracle_kmeans1d_and_train_lsq(oracle):
    """The norm-codebook stand-in is a Lloyd fixed point; the restated train_lsq alternation never
    increases the objective (LSQ.jl:52-66) and its B_norms follow the quantize_norms rule."""
    rng = np.random.default_rng(5)
    v = (rng.standard_normal(5000) ** 2 * 100).astype(np.float32)
    cent, it = oracle.kmeans1d(v, 16, maxiter=5000)
    assert it < 5000 and np.all(np.diff(cent) > 0)
    assert oracle.kmeans1d(v, 16, maxiter=7)[1] == 7   # the iteration cap (Clustering.jl default: 100)
    lab = np.argmin((v[:, None] - cent[None, :]) ** 2, axis=1)
    for j in range(16):
        assert abs(v[lab == j].astype(np.float64).mean() - cent[j]) <= 1e-3 * max(1.0, abs(cent[j]))
    X, C, B = make_problem(77, 600, 16, 3)
    R = np.linalg.qr(rng.standard_normal((16, 16)))[0].astype(np.float32)
    C1, B1, cbn, Bn, obj = oracle.train_lsq(X, 3, 256, R, (B - 1).astype(np.int16), 3, 2, 2, True, 2, seed=4, nworkers=2)
    assert all(b <= a * (1 + 1e-6) for a, b in zip(obj, obj[1:]))
    assert oracle.qerror(X, B1, C1) <= obj[-1] * (1 + 1e-6)
    assert np.array_equal(Bn, oracle.quantize_norms(B1, C1, cbn))


def test_vecs_wire_formats(lsq, tmp_path):
    """.fvecs/.ivecs/.bvecs round trips and ranges (src/read/*.jl: int32 dim header per vector)."""
    rng = np.random.default_rng(0)
    for name, dt, w, r in [("f", np.float32, lsq.io.fvecs_write, lsq.io.fvecs_read),
                           ("i", np.int32, lsq.io.ivecs_write, lsq.io.ivecs_read),
                           ("b", np.uint8, lsq.io.bvecs_write, lsq.io.bvecs_read)]:
        a = (rng.random((37, 12)) * 200).astype(dt)
        path = str(tmp_path / f"x.{name}vecs")
        w(path, a)
        assert os.path.getsize(path) == 37 * (4 + 12 * np.dtype(dt).itemsize)
        assert np.array_equal(r(path), a)
        assert np.array_equal(r(path, 5), a[:5])
        assert np.array_equal(r(path, (3, 9)), a[2:9])


def test_linscan_path_rule(lsq, monkeypatch):
    """Which main pass linscan_lsq takes is a pure shape rule (no GPU needed to ask)."""
    monkeypatch.delenv("LSQ_B200_ADC", raising=False)
    assert lsq.linscan_path(1_000_000, 10_000, 8, 128) == 1      # BASELINE configs[4]
    assert lsq.linscan_path(1_000_000, 10_000, 16, 64) == 1
    assert lsq.linscan_path(1_000_000, 100, 8, 128) == 0         # too few queries to pay for decoding the base set
    assert lsq.linscan_path(50_000, 10_000, 8, 128) == 0         # small base set
    assert lsq.linscan_path(1_000_000, 10_000, 8, 100) == 0      # d not a multiple of 16
    monkeypatch.setenv("LSQ_B200_ADC", "scan")
    assert lsq.linscan_path(1_000_000, 10_000, 8, 128) == 0
    monkeypatch.setenv("LSQ_B200_ADC", "tc")
    assert lsq.linscan_path(1000, 1, 8, 128) == 1
