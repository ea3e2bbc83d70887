"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (local-search-quantization_b200) never does.

Array conventions are C-order numpy arrays that are byte-identical to the Julia column-major arrays
of the reference: X (n, d) float32; codes (n, m) int16 0-based unless a function says otherwise;
C (m, h, d) float32.
"""
import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liblsq_oracle.so")
_REF_DIR = os.path.join(_HERE, "_ref")


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(_HERE, "lsq_oracle.c")
    stale = (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < os.path.getmtime(src)
    need_ref = os.path.isdir("/root/reference") and not os.path.exists(
        os.path.join(_REF_DIR, "linscan_aqd_pairwise_byte.so"))
    if force or stale or need_ref:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ct.CDLL(_LIB)
        _lib.orc_qerror.restype = ct.c_double
        _lib.orc_num_threads.restype = ct.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i16(a):
    return np.ascontiguousarray(a, dtype=np.int16)


def num_threads():
    return lib().orc_num_threads()


def use_all_cores():
    """Make the oracle use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets
    it to 1 in its workers).  -> the thread count."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(np.asarray(ctr, np.uint32)), _p(np.asarray(key, np.uint32)), _p(out))
    return out


def make_to_look(seed, ils_iter, m, randord):
    out = np.zeros(m, np.int32)
    lib().orc_make_to_look(ct.c_uint64(seed), ct.c_uint32(ils_iter), m, int(randord), _p(out))
    return out


def make_perturb(seed, ils_iter, g0, n, m, h, npert):
    slots = np.zeros((n, npert), np.uint8)
    vals = np.zeros((n, npert), np.int16)
    lib().orc_make_perturb(ct.c_uint64(seed), ct.c_uint32(ils_iter), ct.c_uint64(g0), ct.c_int64(n), m, h,
                           npert, _p(slots), _p(vals))
    return slots, vals


def get_norms(C):
    C = _f32(C)
    m, h, d = C.shape
    out = np.zeros((m, h), np.float32)
    lib().orc_get_norms(_p(C), m, h, d, _p(out))
    return out


def get_unaries(X, C):
    X, C = _f32(X), _f32(C)
    n, d = X.shape
    m, h, _ = C.shape
    U = np.zeros((m, n, h), np.float32)
    lib().orc_get_unaries(_p(X), _p(C), ct.c_int64(n), m, h, d, _p(U))
    return U


def get_unary_gemm(X, C):
    """-2*C_i'*X without the norms (the gemm of utils.jl:108 / encode_icm_cuda.jl:92) -> (m, n, h)."""
    X, C = _f32(X), _f32(C)
    n, d = X.shape
    m, h, _ = C.shape
    G = np.zeros((m, n, h), np.float32)
    lib().orc_get_unary_gemm(_p(X), _p(C), ct.c_int64(n), m, h, d, _p(G))
    return G


def get_binaries(C):
    C = _f32(C)
    m, h, d = C.shape
    ncbi = m * (m - 1) // 2
    G = np.zeros((max(ncbi, 1), h, h), np.float32)
    cbi = np.zeros((max(ncbi, 1), 2), np.int32)
    lib().orc_get_binaries(_p(C), m, h, d, _p(G), _p(cbi))
    return G[:ncbi], cbi[:ncbi]


def veccost(X, B0, C):
    X, C, B0 = _f32(X), _f32(C), _i16(B0)
    n, d = X.shape
    m, h, _ = C.shape
    out = np.zeros(n, np.float32)
    lib().orc_veccost(_p(X), _p(B0), _p(C), ct.c_int64(n), m, h, d, _p(out))
    return out


def qerror(X, B0, C):
    X, C, B0 = _f32(X), _f32(C), _i16(B0)
    n, d = X.shape
    m, h, _ = C.shape
    return float(lib().orc_qerror(_p(X), _p(B0), _p(C), ct.c_int64(n), m, h, d))


def reconstruct(B0, C):
    C, B0 = _f32(C), _i16(B0)
    n, m = B0.shape
    _, h, d = C.shape
    out = np.zeros((n, d), np.float32)
    lib().orc_reconstruct(_p(B0), _p(C), ct.c_int64(n), m, h, d, _p(out))
    return out


def splitarray(n, nparts):
    lo, hi = ct.c_int64(), ct.c_int64()
    out = []
    for p in range(nparts):
        lib().orc_splitarray(ct.c_int64(n), nparts, p, ct.byref(lo), ct.byref(hi))
        out.append((lo.value, hi.value))
    return out


def icm_fully(B0, U, G, n, m, h, niter, to_look, slots, vals):
    """encode_icm_fully! on precomputed tables; B0 (n, m) int16 0-based is modified in place."""
    U, G = _f32(U), _f32(G)
    Gt = np.ascontiguousarray(np.transpose(G, (0, 2, 1)))
    npert = slots.shape[1] if slots.size else 0
    lib().orc_icm_fully(_p(B0), _p(U), _p(G), _p(Gt), ct.c_int64(n), m, h, niter,
                        _p(np.asarray(to_look, np.int32)), npert, _p(np.ascontiguousarray(slots)),
                        _p(np.ascontiguousarray(vals)))
    return B0


def encoding_icm_sched(X, oldB0, C, niter, to_look, slots, vals, nworkers=1):
    X, C, oldB0 = _f32(X), _f32(C), _i16(oldB0)
    n, d = X.shape
    m, h, _ = C.shape
    newB = np.zeros_like(oldB0)
    cost = np.zeros(n, np.float32)
    npert = slots.shape[1] if slots.size else 0
    lib().orc_encoding_icm_sched(_p(X), _p(oldB0), _p(newB), _p(C), ct.c_int64(n), m, h, d, niter,
                                 _p(np.asarray(to_look, np.int32)), npert,
                                 _p(np.ascontiguousarray(slots, dtype=np.uint8)),
                                 _p(np.ascontiguousarray(vals, dtype=np.int16)), nworkers, _p(cost))
    return newB, cost


def encoding_icm(X, oldB0, C, niter, randord, npert, seed=0, ils_iter=0, g0=0, nworkers=1):
    """One ILS iteration (encode_icm.jl:131) with the canonical Philox schedule; 0-based codes."""
    X, C, oldB0 = _f32(X), _f32(C), _i16(oldB0)
    n, d = X.shape
    m, h, _ = C.shape
    newB = np.zeros_like(oldB0)
    cost = np.zeros(n, np.float32)
    lib().orc_encoding_icm(_p(X), _p(oldB0), _p(newB), _p(C), ct.c_int64(n), m, h, d, niter, int(randord),
                           npert, ct.c_uint64(seed), ct.c_uint32(ils_iter), ct.c_uint64(g0), nworkers,
                           _p(cost))
    return newB, cost


def encode_icm_ils(X, B0, C, ilsiters, icmiter, npert, randord, seed=0, g0=0, nworkers=1):
    X, C, B0 = _f32(X), _f32(C), _i16(B0)
    n, d = X.shape
    m, h, _ = C.shape
    its = np.asarray(ilsiters, np.int64)
    nr = len(its)
    Bs = np.zeros((nr, n, m), np.int16)
    objs = np.zeros(nr, np.float32)
    lib().orc_encode_icm_ils(_p(X), _p(B0), _p(C), ct.c_int64(n), m, h, d, _p(its), nr, icmiter, npert,
                             int(randord), ct.c_uint64(seed), ct.c_uint64(g0), nworkers, _p(Bs), _p(objs))
    return Bs, objs


def linscan_lsq(codes, queries, codebooks, dbnorms, nn):
    """codes (n, m) uint8 0-based; queries (nq, d); codebooks (m*h, d); -> dists, idx (1-based)."""
    codes = np.ascontiguousarray(codes, np.uint8)
    queries, codebooks, dbnorms = _f32(queries), _f32(codebooks), _f32(dbnorms)
    n, m = codes.shape
    nq, d = queries.shape
    h = codebooks.shape[0] // m
    dists = np.zeros((nq, nn), np.float32)
    idx = np.zeros((nq, nn), np.int32)
    lib().orc_linscan_lsq(_p(dists), _p(idx), _p(codes), _p(queries), _p(codebooks), _p(dbnorms), nq, n, m,
                          h, d, nn)
    return dists, idx


def linscan_pq(codes, queries, centers, K):
    """codes (n, m) uint8; queries (nq, d); centers (m, h, subdim); -> dists, ids (0-based)."""
    codes = np.ascontiguousarray(codes, np.uint8)
    queries, centers = _f32(queries), _f32(centers)
    n, m = codes.shape
    nq, d = queries.shape
    _, h, subdim = centers.shape
    dists = np.zeros((nq, K), np.float32)
    res = np.zeros((nq, K), np.uint32)
    lib().orc_linscan_pq(_p(dists), _p(res), _p(codes), _p(centers), _p(queries), n, nq, m, h, K, d, subdim)
    return dists, res


def quantize_norms(B0, C, cbnorms):
    C, B0, cbnorms = _f32(C), _i16(B0), _f32(cbnorms)
    n, m = B0.shape
    _, h, d = C.shape
    out = np.zeros(n, np.int16)
    lib().orc_quantize_norms(_p(B0), _p(C), _p(cbnorms), ct.c_int64(n), m, h, d, len(cbnorms), _p(out))
    return out


# ---- the reference's own linear scan, compiled unmodified into oracle/_ref --------------------------
_ref = {}


def ref_available():
    return os.path.exists(os.path.join(_REF_DIR, "linscan_aqd_pairwise_byte.so")) and os.path.exists(
        os.path.join(_REF_DIR, "linscan_aqd.so"))


def _ref_lib(name):
    if name not in _ref:
        build()
        _ref[name] = ct.CDLL(os.path.join(_REF_DIR, name))
    return _ref[name]


def ref_linscan_lsq(codes, queries, codebooks, dbnorms, nn):
    """The real linscan_aqd_query_extra_byte (linscan_aqd_pairwise_byte.cpp:97)."""
    codes = np.ascontiguousarray(codes, np.uint8)
    queries, codebooks, dbnorms = _f32(queries), _f32(codebooks), _f32(dbnorms)
    n, m = codes.shape
    nq, d = queries.shape
    h = codebooks.shape[0] // m
    dists = np.zeros((nq, nn), np.float32)
    idx = np.zeros((nq, nn), np.int32)
    _ref_lib("linscan_aqd_pairwise_byte.so").linscan_aqd_query_extra_byte(
        _p(dists), _p(idx), _p(codes), _p(queries), _p(codebooks), _p(dbnorms), nq, n, m, h, d, nn)
    return dists, idx


def ref_linscan_pq(codes, queries, centers, K):
    """The real linscan_aqd_query (linscan_aqd.cpp:107)."""
    codes = np.ascontiguousarray(codes, np.uint8)
    queries, centers = _f32(queries), _f32(centers)
    n, m = codes.shape
    nq, d = queries.shape
    _, h, subdim = centers.shape
    dists = np.zeros((nq, K), np.float32)
    res = np.zeros((nq, K), np.uint32)
    _ref_lib("linscan_aqd.so").linscan_aqd_query(_p(dists), _p(res), _p(codes), _p(centers), _p(queries), n,
                                                 ct.c_uint32(nq), 8 * m, K, m, d, subdim)
    return dists, res


# ---- §8(f) rows: eval_recall, the norm codebook, the train_lsq alternation -------------------------
def eval_recall(ids_gnd, ids_predicted, k):
    """Linscan.jl:76-117: nn_ranks[i] = position of the ground-truth id in query i's WHOLE list (`find`
    over the full column, :91) if it occurs exactly once (:94-98), else k+1;
    recall_at_i[i] = #{rank <= i and rank <= k} / nquery (:111-113).
    ids_predicted is (nq, >= k): row q = column q of the Julia matrix."""
    ids_gnd = np.asarray(ids_gnd).reshape(-1)
    ids_predicted = np.asarray(ids_predicted)
    nquery = ids_predicted.shape[0]
    assert nquery == len(ids_gnd)
    nn_ranks = np.full(nquery, k + 1, np.int64)
    for i in range(nquery):
        pos = np.nonzero(ids_predicted[i] == ids_gnd[i])[0]
        if len(pos) == 1 and pos[0] < k:
            nn_ranks[i] = pos[0] + 1
    nn_ranks.sort()
    return np.searchsorted(nn_ranks, np.arange(1, k + 1), side="right") / nquery


def kmeans1d(values, h, maxiter=100):
    """The deterministic stand-in for `kmeans(dbnorms, h)` (LSQ.jl:80; Clustering.jl is third-party,
    unvendored and seeds from Julia's global RNG -> parity unpinned): Lloyd on the SORTED scalars,
    centres seeded at the (2j+1)/(2h) quantiles, a value moves to the next cluster when it is > the fp32
    midpoint of the two centres, float64 means, empty clusters keep their centre, stop at a fixed point.
    -> (centers float32 (h,), iterations)."""
    s = np.sort(np.asarray(values, np.float32).reshape(-1))
    n = len(s)
    cent = s[((2 * np.arange(h, dtype=np.int64) + 1) * n) // (2 * h)].astype(np.float32)
    bounds = np.full(h + 1, -1, np.int64)
    it = 0
    while it < maxiter:
        mid = (np.float32(0.5) * (cent[:-1] + cent[1:])).astype(np.float32)
        nb = np.concatenate([[0], np.searchsorted(s, mid, side="right"), [n]]).astype(np.int64)
        if np.array_equal(nb, bounds):
            break
        bounds = nb
        for j in range(h):
            if bounds[j + 1] > bounds[j]:
                cent[j] = np.float32(np.sum(s[bounds[j]:bounds[j + 1]].astype(np.float64)) / (bounds[j + 1] - bounds[j]))
        it += 1
    return cent, it


def decoded_norms(B0, C):
    """LSQ.jl:72-77: dbnorms[i] = sum_j CB[j,i]^2, j ascending, fp32."""
    CB = reconstruct(B0, C)
    out = np.zeros(CB.shape[0], np.float32)
    for j in range(CB.shape[1]):
        out = (out + CB[:, j] * CB[:, j]).astype(np.float32)
    return out


def train_lsq(X, m, h, R, B0, niter, ilsiter, icmiter, randord, npert, seed=0, update=None, nworkers=None):
    """src/lsq/LSQ.jl:10-88 restated over the oracle's pieces; 0-based codes.  `update(X, B0, h)` is the
    codebook-update stand-in (default: the exact min-norm least-squares solution).
    -> (C, B0, cbnorms, B_norms (1-based), obj)."""
    from . import codebook_update as cu
    update = update or cu.update_codebooks_exact
    nworkers = nworkers or num_threads()
    X = _f32(X)
    B0 = _i16(B0).copy()
    if R is not None:
        R = _f32(R)
        C = update(_f32(X @ R), B0, h)                   # RX = R'X  (:31), update on RX (:34)
        C = _f32(np.einsum("ij,mhj->mhi", R, C))          # C[i] = R * C[i]  (:37-39)
    else:
        C = update(X, B0, h)
    count = 0
    for _ in range(ilsiter):                             # :44-48
        B0, _c = encoding_icm(X, B0, C, icmiter, randord, npert, seed=seed, ils_iter=count, nworkers=nworkers)
        count += 1
    obj = np.zeros(niter, np.float32)
    for it in range(niter):                              # :52-66
        obj[it] = qerror(X, B0, C)
        C = update(X, B0, h)
        for _ in range(ilsiter):
            B0, _c = encoding_icm(X, B0, C, icmiter, randord, npert, seed=seed, ils_iter=count, nworkers=nworkers)
            count += 1
    cbnorms, _ = kmeans1d(decoded_norms(B0, C), h)       # :68-81
    B_norms = quantize_norms(B0, C, cbnorms)
    return C, B0, cbnorms, B_norms, obj


def encoding_viterbi(X, C):
    """encode_chain.jl:95-127 (ChainQ's exact chain encoder) -> (n, m) int16 0-based codes."""
    X, C = _f32(X), _f32(C)
    n, d = X.shape
    m, h, _ = C.shape
    assert m >= 2
    B0 = np.zeros((n, m), np.int16)
    lib().orc_encoding_viterbi(_p(X), _p(C), ct.c_int64(n), m, h, d, _p(B0))
    return B0
