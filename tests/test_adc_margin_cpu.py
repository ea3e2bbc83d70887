"""The exactness of the tensor-core ADC path (csrc/adc_tc.cu) rests on one inequality: for every (query, base vector)
pair, |filter value - reference value| <= margin_q.  This CPU test restates the filter's operands in NumPy — bf16
round-to-nearest-even hi / lo parts, the fp32 decode sum, the products taken exactly in float64 — and the reference's
fp32 chains (linscan_aqd_pairwise_byte.cpp:42-48, 69-73 and linscan_aqd.cpp:66-74, 84-86), and checks the inequality
on data of very different scales, for one and two products, LSQ and PQ tables.  (What it cannot restate is the rounding
of the tensor core's own fp32 accumulation; the GPU tests measure that at 1e-6 .. 3e-6 of the scale the margin grants
2^-13 of.)"""
import numpy as np
import pytest

U = 2.0 ** -24
H = 256


def bf16(x):
    b = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return (((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)).view(np.float32)


def make(kind, rng, n, nq, d, m):
    if kind == "sift":
        g = lambda *s: np.clip(np.floor(np.abs(rng.standard_normal(s)) * 40.0), 0, 218).astype(np.float32)
        return g(nq, d), (g(m * H, d) / np.float32(m)).astype(np.float32)
    scale, offset = {"gauss": (1.0, 0.0), "tiny": (1e-3, 0.0), "huge": (1e4, 0.0), "offset": (1.0, 50.0)}[kind]
    q = (rng.standard_normal((nq, d)) * scale + offset).astype(np.float32)
    c = (rng.standard_normal((m * H, d)) * scale + offset / m).astype(np.float32)
    return q, c


@pytest.mark.parametrize("kind", ["sift", "gauss", "tiny", "huge", "offset"])
@pytest.mark.parametrize("products", [1, 2])
def test_lsq_margin_bounds_filter_minus_reference(kind, products):
    rng = np.random.default_rng(11)
    n, nq, d, m = 1500, 24, 64, 8
    q, cb = make(kind, rng, n, nq, d, m)
    codes = rng.integers(0, H, size=(n, m))
    # reference: t = t - (2 q_k) c_k in fp32 per table entry, then ((0 + L0) + L1) + ... + dbnorm in fp32
    two_q = (np.float32(2.0) * q).astype(np.float32)
    lut = np.zeros((nq, m * H), np.float32)
    for k in range(d):
        lut = (lut - (two_q[:, k:k + 1] * cb[None, :, k]).astype(np.float32)).astype(np.float32)
    xhat32 = np.zeros((n, d), np.float32)
    for k in range(m):
        xhat32 = (xhat32 + cb[k * H + codes[:, k]]).astype(np.float32)
    norms = (xhat32.astype(np.float64) ** 2).sum(1).astype(np.float32)
    dref = np.zeros((nq, n), np.float32)
    for k in range(m):
        dref = (dref + lut[:, k * H + codes[:, k]]).astype(np.float32)
    dref = (dref + norms[None, :]).astype(np.float32)
    # filter operands
    q_hi = bf16(q)
    x_hi = bf16(xhat32)
    x_lo = bf16(xhat32 - x_hi)
    x_used = x_hi.astype(np.float64) + (x_lo.astype(np.float64) if products == 2 else 0.0)
    dfil = norms.astype(np.float64)[None, :] - 2.0 * q_hi.astype(np.float64) @ x_used.T
    # the margin exactly as adc_filter_kernel computes it (tau = the largest reference value: the loosest case)
    qn = np.sqrt((q.astype(np.float64) ** 2).sum(1)) * 1.00001
    ql = np.sqrt(((q - q_hi).astype(np.float64) ** 2).sum(1)) * 1.00001
    xmax = np.sqrt((xhat32.astype(np.float64) ** 2).sum(1).max()) * 1.00001
    xlo = np.sqrt(((xhat32 - x_hi).astype(np.float64) ** 2).sum(1).max()) * 1.00001 if products == 1 else 0.0
    cmax = np.sqrt((cb.astype(np.float64) ** 2).sum(1).max()) * 1.00001
    nmax = np.abs(norms).max()
    eps_f, eps_r = 1.0 / 8192.0, 2.0 * (d + m + 2) * U
    margin = 2 * ql * xmax + 2 * qn * xlo + eps_f * 2 * qn * xmax + eps_r * (2 * qn * m * cmax + nmax)
    err = np.abs(dfil - dref.astype(np.float64)).max(1)
    assert np.all(err <= margin), (kind, products, (err / margin).max())
    # and the representational part alone leaves room for the accumulation rounding of the tensor core
    assert np.all(err <= margin - 0.5 * eps_f * 2 * qn * xmax), (kind, products)


@pytest.mark.parametrize("kind", ["sift", "gauss", "huge", "offset"])
def test_pq_margin_bounds_filter_minus_reference(kind):
    rng = np.random.default_rng(12)
    n, nq, d, m = 1500, 24, 64, 8
    sub = d // m
    q, cb = make(kind, rng, n, nq, d, m)
    centers = np.ascontiguousarray(cb[:, :sub])          # [m*256][sub]
    codes = rng.integers(0, H, size=(n, m))
    # reference: LUT[k][r] = sum_s sqr(c[s] - q[k*sub + s]) in fp32, dist = ((0 + L0) + L1) + ...
    lut = np.zeros((nq, m * H), np.float32)
    for k in range(m):
        for s in range(sub):
            df = (centers[None, k * H:(k + 1) * H, s] - q[:, k * sub + s][:, None]).astype(np.float32)
            lut[:, k * H:(k + 1) * H] = (lut[:, k * H:(k + 1) * H] + (df * df).astype(np.float32)).astype(np.float32)
    dref = np.zeros((nq, n), np.float32)
    for k in range(m):
        dref = (dref + lut[:, k * H + codes[:, k]]).astype(np.float32)
    xhat = np.concatenate([centers[k * H + codes[:, k]] for k in range(m)], axis=1)   # [n][d], exact
    q_hi, x_hi = bf16(q), bf16(xhat)
    x_lo = bf16(xhat - x_hi)
    n2 = (xhat.astype(np.float64) ** 2).sum(1)
    qn2 = (q.astype(np.float64) ** 2).sum(1)
    dfil = qn2[:, None] + n2[None, :] - 2.0 * q_hi.astype(np.float64) @ (x_hi.astype(np.float64) + x_lo.astype(np.float64)).T
    qn = np.sqrt(qn2) * 1.00001
    ql = np.sqrt(((q - q_hi).astype(np.float64) ** 2).sum(1)) * 1.00001
    xmax = np.sqrt(n2.max()) * 1.00001
    eps_f, eps_r = 1.0 / 8192.0, 2.0 * (sub + m + 4) * U   # the kernel uses 2 (d + m + 2) u >= this
    margin = 2 * ql * xmax + eps_f * 2 * qn * xmax + eps_r * (qn + xmax) ** 2 + eps_f * (qn ** 2 + xmax ** 2)
    err = np.abs(dfil - dref.astype(np.float64)).max(1)
    assert np.all(err <= margin), (kind, (err / margin).max())
