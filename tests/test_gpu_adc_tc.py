"""Tensor-core ADC prefilter (csrc/adc_tc.cu): the filter values against float64, and the whole linscan through
the filter against the reference's own .so and against the lookup scan — ids and distances bit for bit."""
import os

import numpy as np
import pytest

from util import make_scan_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(lsq):
    assert lsq.device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    lsq.init(0)
    return lsq


def gauss_scan_problem(seed, n, nq, d, m, h=256):
    """Signed, non-representable data: every operand has a non-zero bf16 lo part; norms = ||xhat||^2."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, h, size=(n, m)).astype(np.uint8)
    queries = (rng.standard_normal((nq, d)) * 3.0).astype(np.float32)
    codebooks = rng.standard_normal((m * h, d)).astype(np.float32)
    xhat = np.zeros((n, d), np.float32)
    for k in range(m):
        xhat += codebooks[k * h + codes[:, k].astype(np.int64)]
    norms = (xhat.astype(np.float64) ** 2).sum(1).astype(np.float32)
    return codes, queries, codebooks, norms


def _ref(oracle, *a):
    return oracle.ref_linscan_lsq(*a) if oracle.ref_available() else oracle.linscan_lsq(*a)


@pytest.mark.parametrize("n,nq,d,m,kind", [
    (1000, 5, 128, 8, "gauss"),      # one A tile, tail tile (1000 = 7*128 + 104)
    (3000, 300, 128, 16, "gauss"),   # two query groups, second one with a single A tile
    (640, 256, 64, 3, "gauss"),      # both A tiles full, d = 64
    (257, 130, 16, 1, "gauss"),      # d = 16: a single MMA K step
    (5000, 200, 96, 12, "sift"),     # d = 96: K not a multiple of 64
])
@pytest.mark.parametrize("passes", [2, 1])
def test_filter_values_match_float64(gpu, n, nq, d, m, kind, passes, monkeypatch):
    import torch
    monkeypatch.setenv("LSQ_B200_ADC_PASSES", str(passes))
    from lsq_b200 import device as dev
    mk = gauss_scan_problem if kind == "gauss" else make_scan_problem
    codes, queries, codebooks, norms = mk(7000 + n, n, nq, d, m)
    out = dev.adc_filter_values(torch.from_numpy(codes).cuda(), torch.from_numpy(queries).cuda(),
                                torch.from_numpy(codebooks).cuda(), torch.from_numpy(norms).cuda()).cpu().numpy()
    def bf16(x):   # round to nearest even, like __float2bfloat16_rn
        b = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
        return (((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)).view(np.float32)

    xhat32 = np.zeros((n, d), np.float32)   # the decode kernel sums the codewords in fp32, k ascending
    for k in range(m):
        xhat32 += codebooks[k * 256 + codes[:, k].astype(np.int64)]
    xhat = xhat32.astype(np.float64)
    # the filter multiplies hi(q) = bf16(q) with hi(x) + lo(x) (two products) or hi(x) alone (one product); what it
    # drops is covered by the 2 |lo(q)| max|xhat| and 2 |q| max|lo(x)| terms of the margin
    q_hi = bf16(queries)
    x_used = xhat if passes == 2 else bf16(xhat32).astype(np.float64)
    D = norms.astype(np.float64)[None, :] - 2.0 * q_hi.astype(np.float64) @ x_used.T
    got = out[:, :n].astype(np.float64)
    assert np.all(np.isfinite(got)), "filter values missing (an epilogue warp skipped a tile?)"
    assert np.all(out[:, n:] > 1e37), "padding columns must carry a huge distance (they never pass the filter)"
    scale = 2.0 * np.linalg.norm(queries.astype(np.float64), axis=1)[:, None] * np.linalg.norm(xhat, axis=1).max()
    rel = np.abs(got - D) / scale
    print(f"filter: max |d_tc - D| / (2 |q| max|xhat|) = {rel.max():.3e}")
    # the margin in adc_tc.cu allows 2^-13 for the hi/lo split of xhat + fp32 accumulation: stay far below it
    assert rel.max() < 2.0 ** -16, rel.max()


CASES = [
    (200000, 300, 128, 8, 1000, "sift"),
    (150000, 257, 128, 16, 500, "gauss"),
    (70000, 40, 64, 12, 200, "gauss"),
    (66000, 129, 32, 15, 100, "sift"),
    (20000, 30, 16, 4, 50, "gauss"),     # forced below the size gate
    (100000, 64, 128, 8, 3000, "gauss"),
    (70000, 300, 96, 12, 100, "gauss"),  # d = 96: the K tail of the query rows is stored in 8-column pieces
    (300000, 40, 64, 8, 5000, "gauss"),  # large nn: ~7000 survivors per query, the 128 KB sorter takes every query
]


@pytest.mark.parametrize("thresholds", ["lists", "sample_buffer"])
@pytest.mark.parametrize("n,nq,d,m,nn,kind", CASES)
def test_linscan_through_the_filter_is_exact(gpu, oracle, n, nq, d, m, nn, kind, thresholds, monkeypatch):
    if thresholds == "sample_buffer":   # the first design: all sample values written out, r-th smallest selected
        if (n, nq) not in ((200000, 300), (66000, 129)):
            pytest.skip("two shapes are enough for the alternative threshold path")
        monkeypatch.setenv("LSQ_B200_ADC_SBUF", "1")
    elif (n, nq) in ((150000, 257), (70000, 40)):
        # by default the number of products is chosen on the device from the sample; force each here
        for products in ("1", "2"):
            monkeypatch.setenv("LSQ_B200_ADC", "tc")
            monkeypatch.setenv("LSQ_B200_ADC_PASSES", products)
            mk0 = gauss_scan_problem if kind == "gauss" else make_scan_problem
            c0, q0, cb0, n0 = mk0(7100 + m + nn, n, nq, d, m)
            df, idf = gpu.linscan_lsq(c0, q0, cb0.reshape(m, 256, d), n0, np.eye(d, dtype=np.float32), nn)
            dr0, ir0 = _ref(oracle, c0, q0[:8], cb0, n0, nn)
            assert np.array_equal(idf[:8], ir0) and np.array_equal(df[:8], dr0), products
        monkeypatch.delenv("LSQ_B200_ADC_PASSES")
    mk = gauss_scan_problem if kind == "gauss" else make_scan_problem
    codes, queries, codebooks, norms = mk(7100 + m + nn, n, nq, d, m)
    R = np.eye(d, dtype=np.float32)
    monkeypatch.setenv("LSQ_B200_ADC", "tc")
    l0 = gpu.launch_count()
    dt, it = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), norms, R, nn)
    launches_tc = gpu.launch_count() - l0
    monkeypatch.setenv("LSQ_B200_ADC", "scan")
    l0 = gpu.launch_count()
    ds, is_ = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), norms, R, nn)
    launches_scan = gpu.launch_count() - l0
    assert launches_tc > launches_scan, "the tensor-core path did not run"
    assert np.array_equal(it, is_) and np.array_equal(dt, ds)
    k = min(nq, 24)
    dr, ir = _ref(oracle, codes, queries[:k], codebooks, norms, nn)
    assert np.array_equal(it[:k], ir) and np.array_equal(dt[:k], dr)


def test_filter_ties_and_clustered_neighbours(gpu, oracle, monkeypatch):
    """Duplicates (ties -> lower id first) and a base set whose near neighbours all sit at the end: the sampled
    threshold misses, the filter lists overflow or come up short, and the exhaustive path must take over."""
    n, nq, d, m, nn = 80000, 20, 32, 8, 500
    codes, queries, codebooks, norms = make_scan_problem(7300, n, nq, d, m)
    codes[n // 2:] = codes[: n // 2]
    norms[n // 2:] = norms[: n // 2]
    R = np.eye(d, dtype=np.float32)
    monkeypatch.setenv("LSQ_B200_ADC", "tc")
    for nrm in (norms, np.concatenate([norms[:-600], norms[-600:] - 1e6]).astype(np.float32)):
        dr, ir = _ref(oracle, codes, queries, codebooks, nrm, nn)
        dg, ig = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), nrm, R, nn)
        assert np.array_equal(ig, ir) and np.array_equal(dg, dr)


def test_filter_with_several_query_batches(gpu, oracle, monkeypatch):
    """More queries than one batch of the candidate buffers holds (21845 at the 16 K-key capacity): the base image
    is decoded once, every batch runs sample -> thresholds -> LUT rows -> filter -> rescoring -> top-k."""
    n, nq, d, m, nn = 66000, 23000, 16, 4, 10
    codes, queries, codebooks, norms = gauss_scan_problem(7400, n, nq, d, m)
    R = np.eye(d, dtype=np.float32)
    monkeypatch.setenv("LSQ_B200_ADC", "tc")
    dt, it = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), norms, R, nn)
    monkeypatch.setenv("LSQ_B200_ADC", "scan")
    ds, is_ = gpu.linscan_lsq(codes, queries, codebooks.reshape(m, 256, d), norms, R, nn)
    assert np.array_equal(it, is_) and np.array_equal(dt, ds)
    sample = np.r_[0:8, 21840:21856, nq - 8:nq]   # first batch, the batch boundary, the last queries
    dr, ir = _ref(oracle, codes, queries[sample], codebooks, norms, nn)
    assert np.array_equal(it[sample], ir) and np.array_equal(dt[sample], dr)


def _ref_pq(oracle, *a):
    return oracle.ref_linscan_pq(*a) if oracle.ref_available() else oracle.linscan_pq(*a)


@pytest.mark.parametrize("n,nq,d,m,nn,kind", [
    (150000, 50, 128, 8, 1000, "sift"),       # sub-codewords of 16 elements
    (70000, 300, 64, 16, 200, "gauss"),       # 4 elements: shorter than an 8-element K chunk
    (66000, 129, 32, 16, 100, "gauss"),       # 2 elements
    (80000, 40, 64, 4, 50, "gauss"),          # 16 elements, 4 codebooks
])
def test_linscan_pq_through_the_filter_is_exact(gpu, oracle, n, nq, d, m, nn, kind, monkeypatch):
    """PQ / OPQ tables (squared differences per sub-space, linscan_aqd.cpp:37-102) through the tensor-core filter:
    dist = ||q||^2 - 2<q, xhat> + ||xhat||^2 with xhat the concatenated sub-codewords."""
    mk = gauss_scan_problem if kind == "gauss" else make_scan_problem
    subdim = d // m   # Linscan.jl:23
    codes, queries, codebooks, _ = mk(7500 + m + nn, n, nq, d, m)
    centers = np.ascontiguousarray(codebooks[:, :subdim].reshape(m, 256, subdim))
    monkeypatch.setenv("LSQ_B200_ADC", "tc")
    l0 = gpu.launch_count()
    dt, it = gpu.linscan_pq(codes, queries, centers, 8 * m, nn)
    launches_tc = gpu.launch_count() - l0
    monkeypatch.setenv("LSQ_B200_ADC", "scan")
    l0 = gpu.launch_count()
    ds, is_ = gpu.linscan_pq(codes, queries, centers, 8 * m, nn)
    launches_scan = gpu.launch_count() - l0
    assert launches_tc > launches_scan, "the tensor-core path did not run"
    assert np.array_equal(it, is_) and np.array_equal(dt, ds)
    k = min(nq, 24)
    dr, ir = _ref_pq(oracle, codes, queries[:k], centers, nn)
    assert np.array_equal(it[:k].astype(np.int64) - 1, ir.astype(np.int64)) and np.array_equal(dt[:k], dr)
