"""GPU parity at BASELINE.json's FULL sizes (-m gpu): the oracle cannot encode a million vectors in
seconds, so these tests use size-independent properties of the path plus bit-exact oracle checks on
contiguous blocks cut out of the full problem (possible because vectors never interact and the
perturbation stream is keyed by the GLOBAL vector index).  Nothing here reads /root/reference."""
import numpy as np
import pytest

from util import make_problem, make_scan_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(lsq):
    assert lsq.device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    lsq.init(0)
    return lsq


def _icm_fullsize(gpu, oracle, m, n, block, blocks_at):
    d, ils = 128, 16
    X, C, B = make_problem(4000 + m, n, d, m)
    prev = gpu.veccost(X, B, C)
    Bs, objs = gpu.encode_icm_cuda(X, B, C, [1, ils], 4, 4, True, 2, seed=77)
    B1, B16 = Bs
    assert B16.shape == (n, m) and B16.min() >= 1 and B16.max() <= 256
    # accept rule (encode_icm.jl:183-186): per-vector cost never increases, strictly drops where codes changed
    c1, c16 = gpu.veccost(X, B1, C), gpu.veccost(X, B16, C)
    assert np.all(c1 <= prev) and np.all(c16 <= c1)
    changed = np.any(B16 != B, axis=1)
    assert np.all(c16[changed] < prev[changed])
    # the objective snapshots are the means of those costs (encode_icm_cuda.jl:291-293)
    assert abs(objs[0] - c1.astype(np.float64).mean()) <= 1e-6 * objs[0]
    assert abs(objs[1] - c16.astype(np.float64).mean()) <= 1e-6 * objs[1]
    assert objs[1] < objs[0] < prev.astype(np.float64).mean()
    # bit-exact oracle checks on contiguous blocks of the full problem, and sharding invariance of the
    # library itself (a shard encoded alone, with its global offset, equals that slice of the whole)
    for lo in blocks_at:
        hi = lo + block
        Bo, _ = oracle.encode_icm_ils(X[lo:hi], (B[lo:hi] - 1).astype(np.int16), C, [ils], 4, 4, True, seed=77,
                                      g0=lo, nworkers=oracle.num_threads())
        assert np.array_equal(B16[lo:hi], Bo[0] + 1), f"block at {lo} differs from the oracle"
    lo = n // 3
    sh, _ = gpu.encode_icm_cuda(X[lo:lo + 50000], B[lo:lo + 50000], C, [ils], 4, 4, True, 1, seed=77, g0=lo)
    assert np.array_equal(sh[0], B16[lo:lo + 50000])


def test_icm_config1_full_size_m8(gpu, oracle):
    """BASELINE configs[1]: 1 M base vectors, m = 8, 16 ILS iterations."""
    _icm_fullsize(gpu, oracle, 8, 1_000_000, 1500, [0, 499_123, 998_500])


def test_icm_config2_full_size_m16(gpu, oracle):
    """BASELINE configs[2]: 1 M base vectors, m = 16 (128-bit codes)."""
    _icm_fullsize(gpu, oracle, 16, 1_000_000, 400, [0, 731_001])


@pytest.mark.parametrize("m", [8, 16])
def test_linscan_config4_full_size(gpu, oracle, m):
    """BASELINE configs[4]: 1 M codes x 10 K queries, top-1000.  Properties of every row + the reference's
    own C++ (or the restatement) on a sample of the queries, bit for bit."""
    n, nq, d, nn = 1_000_000, 10_000, 128, 1000
    codes, queries, codebooks, norms = make_scan_problem(4100 + m, n, nq, d, m)
    dists, ids = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), norms, np.eye(d, dtype=np.float32), nn)
    assert dists.shape == (nq, nn) and ids.shape == (nq, nn)
    assert ids.min() >= 1 and ids.max() <= n                       # 1-based (linscan_aqd_pairwise_byte.cpp:75)
    assert np.all(np.diff(dists, axis=1) >= 0)                     # ascending distances
    tie = np.diff(dists, axis=1) == 0
    assert np.all(np.diff(ids, axis=1)[tie] > 0)                   # ties: lower id first
    srt = np.sort(ids, axis=1)
    assert np.all(np.diff(srt, axis=1) > 0)                        # no id twice in a row
    sample = np.r_[0:6, nq // 2:nq // 2 + 5, nq - 5:nq]            # first / middle / last tiles
    fn = oracle.ref_linscan_lsq if oracle.ref_available() else oracle.linscan_lsq
    dr, ir = fn(codes, queries[sample], codebooks, norms, nn)
    assert np.array_equal(ids[sample], ir) and np.array_equal(dists[sample], dr)
