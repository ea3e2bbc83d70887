"""Independent NumPy restatement of the same reference lines as lsq_oracle.c.  TEST INFRASTRUCTURE ONLY.

Written separately from the C oracle (vector-major loops instead of node-major passes, long-double
emulation of the FMA chain instead of fmaf) so that agreement between the two is evidence that both
follow the reference, since Julia itself cannot be run here.  Small sizes only.
"""
import numpy as np

M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c = [int(x) for x in ctr]
    k = [int(x) for x in key]
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k[1]) & M32, p0 & M32]
        k = [(k[0] + 0x9E3779B9) & M32, (k[1] + 0xBB67AE85) & M32]
    return c


def _word(seed, it, g, stream, i):
    ctr = [g & M32, (g >> 32) & M32, it & M32, ((stream << 24) | (i >> 2)) & M32]
    return philox4x32_10(ctr, [seed & M32, (seed >> 32) & M32])[i & 3]


def make_to_look(seed, it, m, randord):
    a = list(range(m))
    if randord:
        w = 0
        for i in range(m - 1, 0, -1):
            j = _word(seed, it, 0, 1, w) % (i + 1)
            a[i], a[j] = a[j], a[i]
            w += 1
    return np.array(a, np.int32)


def make_perturb(seed, it, g0, n, m, h, npert):
    slots = np.zeros((n, npert), np.uint8)
    vals = np.zeros((n, npert), np.int16)
    for v in range(n):
        g = g0 + v
        a = list(range(m))
        for i in range(npert):
            j = i + _word(seed, it, g, 0, i) % (m - i)
            a[i], a[j] = a[j], a[i]
        slots[v] = sorted(a[:npert])
        vals[v] = [_word(seed, it, g, 0, npert + i) % h for i in range(npert)]
    return slots, vals


def fma_dot(A, B):
    """Row-wise sequential fp32 FMA chain: A (..., d), B (..., d) broadcastable -> (...).
    fl32(a*b + acc) evaluated in x87 long double (64-bit mantissa): the product of two fp32 numbers
    is exact and the single rounding to 64 bits before the rounding to 24 bits is harmless at any
    realistic probability."""
    A = np.asarray(A, np.float32)
    B = np.asarray(B, np.float32)
    d = A.shape[-1]
    acc = np.zeros(np.broadcast_shapes(A.shape[:-1], B.shape[:-1]), np.float32)
    for k in range(d):
        acc = (A[..., k].astype(np.longdouble) * B[..., k].astype(np.longdouble)
               + acc.astype(np.longdouble)).astype(np.float32)
    return acc


def get_unaries(X, C):
    """utils.jl:94-122 -> (m, n, h)"""
    m, h, d = C.shape
    n = X.shape[0]
    U = np.zeros((m, n, h), np.float32)
    for i in range(m):
        nrm = fma_dot(C[i], C[i])
        dots = fma_dot(C[i][None, :, :], X[:, None, :])  # (n, h)
        U[i] = (np.float32(-2.0) * dots).astype(np.float32) + nrm[None, :]
    return U


def pair_table(C, j, k):
    """T[b][a] = 2<C_j[:,a], C_k[:,b]>  (utils.jl:137; transposed copy encode_icm.jl:25-28)"""
    return (np.float32(2.0) * fma_dot(C[j][None, :, :], C[k][:, None, :])).astype(np.float32)


def veccost(X, B0, C):
    n, d = X.shape
    m = B0.shape[1]
    out = np.zeros(n, np.float32)
    for v in range(n):
        r = np.zeros(d, np.float32)
        for k in range(m):
            r = (r + C[k, B0[v, k]]).astype(np.float32)
        df = (r - X[v]).astype(np.float32)
        sq = (df * df).astype(np.float32)
        p = np.zeros(32, np.float32)
        for t in range(d):
            p[t & 31] = np.float32(p[t & 31] + sq[t])
        for off in (16, 8, 4, 2, 1):
            p = (p + p[np.arange(32) ^ off]).astype(np.float32)
        out[v] = p[0]
    return out


def encoding_icm(X, oldB0, C, niter, randord, npert, seed=0, ils_iter=0, g0=0):
    """encode_icm.jl:131-189 + 4-127, vector-major."""
    n, d = X.shape
    m, h, _ = C.shape
    U = get_unaries(X, C)
    T = {(j, k): pair_table(C, j, k) for j in range(m) for k in range(m) if j != k}
    to_look = make_to_look(seed, ils_iter, m, randord)
    slots, vals = make_perturb(seed, ils_iter, g0, n, m, h, npert)
    prev = veccost(X, oldB0, C)
    B = oldB0.copy()
    for v in range(n):
        b = B[v].copy()
        for s, x in zip(slots[v], vals[v]):
            b[s] = x
        for _ in range(niter):
            for j in to_look:
                ub = U[j, v].copy()
                for k in range(m):
                    if k != j:
                        ub = (ub + T[(j, k)][b[k]]).astype(np.float32)
                b[j] = int(np.argmin(ub))  # first minimum
        B[v] = b
    new = veccost(X, B, C)
    keep_old = ~(new < prev)
    B[keep_old] = oldB0[keep_old]
    new[keep_old] = prev[keep_old]
    return B, new


def linscan_lsq(codes, queries, codebooks, dbnorms, nn):
    """linscan_aqd_pairwise_byte.cpp:14-93 -> dists, idx (1-based)"""
    n, m = codes.shape
    nq, d = queries.shape
    h = codebooks.shape[0] // m
    dists = np.zeros((nq, nn), np.float32)
    idx = np.zeros((nq, nn), np.int32)
    for q in range(nq):
        lut = np.zeros(m * h, np.float32)
        for k in range(d):
            lut = (lut - ((np.float32(2) * queries[q, k]) * codebooks[:, k]).astype(np.float32)).astype(np.float32)
        s = np.zeros(n, np.float32)
        for k in range(m):
            s = (s + lut[h * k + codes[:, k].astype(np.int64)]).astype(np.float32)
        s = (s + dbnorms).astype(np.float32)
        order = np.lexsort((np.arange(n), s))[:nn]
        dists[q] = s[order]
        idx[q] = order + 1
    return dists, idx


def encoding_viterbi(X, C):
    """encode_chain.jl:1-127 -> (n, m) int16 0-based.  Vectorised over the data points; the per-element
    arithmetic (one fp32 add per (k, j), first minimum over k) is the reference's."""
    m, h, d = C.shape
    n = X.shape[0]
    U = get_unaries(X, C)                       # (m, n, h)
    minidx = np.zeros((m - 1, n, h), np.int64)
    V = U[0].copy()
    for i in range(m - 1):
        bb = pair_table(C, i + 1, i)            # [k][j] = 2<C_{i+1}[:,j], C_i[:,k]> = bb[k, j]
        cost = (V[:, :, None] + bb[None, :, :]).astype(np.float32)   # (n, k, j)
        minidx[i] = np.argmin(cost, axis=1)     # first minimum over k
        mincost = np.take_along_axis(cost, minidx[i][:, None, :], axis=1)[:, 0, :]
        V = (U[i + 1] + mincost).astype(np.float32)
    B = np.zeros((n, m), np.int16)
    cur = np.argmin(V, axis=1)
    B[:, m - 1] = cur
    for i in range(m - 2, -1, -1):
        cur = minidx[i][np.arange(n), cur]
        B[:, i] = cur
    return B
