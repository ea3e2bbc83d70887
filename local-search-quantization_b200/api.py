"""Host-side mirror of the reference API over the C ABI (include/lsq_b200.h).  See package docstring."""
import ctypes as ct
import os

import numpy as np

__all__ = [
    "LsqError", "lib", "lib_path", "have_library", "init", "init_devices", "num_bound_devices", "finalize", "device_count", "launch_count", "linscan_path", "linscan_last_phases", "version",
    "splitarray", "make_to_look", "make_perturb", "get_unaries", "get_binaries", "veccost", "qerror",
    "reconstruct", "quantize_norms", "encoding_icm", "reset_ils_counter", "encoding_icm_sched", "encode_icm_cuda",
    "update_codebooks", "linscan_lsq", "linscan_pq", "linscan_opq", "eval_recall", "randinit", "train_lsq",
    "kmeans1d", "encoding_viterbi",
    "EXPORTED_SYMBOLS",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblsq_b200.so")
_lib = None

# every symbol include/lsq_b200.h declares (tests check the .so exports exactly these)
EXPORTED_SYMBOLS = [
    "lsq_init", "lsq_init_devices", "lsq_num_bound_devices", "lsq_finalize", "lsq_last_error", "lsq_device_count", "lsq_launch_count", "lsq_last_collective_ms", "lsq_version", "lsq_splitarray",
    "lsq_make_to_look", "lsq_make_perturb", "lsq_get_unaries", "lsq_get_binaries", "lsq_veccost",
    "lsq_qerror", "lsq_reconstruct", "lsq_encoding_icm", "lsq_encoding_icm_sched", "lsq_encode_icm_cuda",
    "lsq_update_codebooks", "linscan_aqd_query_extra_byte", "linscan_aqd_query", "lsq_linscan_lsq",
    "lsq_linscan_pq", "lsq_quantize_norms", "lsq_dev_tables_bytes", "lsq_dev_sliced_tables_bytes",
    "lsq_dev_icm_layout", "lsq_dev_build_tables",
    "lsq_dev_build_unaries", "lsq_dev_build_unaries_tc", "lsq_dev_veccost", "lsq_dev_icm_ils", "lsq_dev_icm_visit_counter",
    "lsq_cb_stats_len", "lsq_cb_scale_exp", "lsq_dev_absmax", "lsq_dev_cb_accumulate", "lsq_dev_cb_finalize",
    "lsq_dev_cb_stats", "lsq_dev_cb_solve",
    "lsq_dev_linscan", "lsq_dev_adc_filter_values", "lsq_linscan_path", "lsq_linscan_last_phases", "lsq_train_lsq", "lsq_kmeans1d", "lsq_eval_recall", "lsq_dev_eval_recall",
    "lsq_encoding_viterbi", "lsq_dev_viterbi",
]


class LsqError(RuntimeError):
    """Raised for every non-zero status of the C ABI (and when the library itself is missing)."""

    def __init__(self, code, msg):
        super().__init__(f"lsq_b200 error {code}: {msg}")
        self.code = code


def lib_path():
    return _LIB_PATH


def have_library():
    return os.path.exists(_LIB_PATH)


def lib():
    """The loaded CDLL.  Fails loudly when the CUDA library has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise LsqError(-1, f"{_LIB_PATH} is missing: build it with "
                               f"`python local-search-quantization_b200/build.py` (there is no CPU fallback)")
        L = ct.CDLL(_LIB_PATH)
        L.lsq_last_error.restype = ct.c_char_p
        L.lsq_version.restype = ct.c_char_p
        L.lsq_dev_tables_bytes.restype = ct.c_int64
        L.lsq_dev_sliced_tables_bytes.restype = ct.c_int64
        L.lsq_cb_stats_len.restype = ct.c_int64
        L.lsq_launch_count.restype = ct.c_uint64
        L.lsq_last_collective_ms.restype = ct.c_float
        L.linscan_aqd_query_extra_byte.restype = None
        L.linscan_aqd_query.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise LsqError(rc, lib().lsq_last_error().decode())


def _p(a):
    return a.ctypes.data_as(ct.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _codebooks(C):
    """Vector{Matrix} (list of m arrays (h, d)) or an (m, h, d) array -> (m, h, d) float32."""
    if isinstance(C, (list, tuple)):
        C = np.stack([np.asarray(c, np.float32) for c in C])
    C = _f32(C)
    if C.ndim != 3:
        raise ValueError("C must be m codebooks of shape (h, d)")
    return C


def _codes16(B, n=None):
    B = np.asarray(B)
    if B.dtype != np.int16:
        raise TypeError("B must be a Matrix{Int16} (numpy int16, 1-based)")
    return np.ascontiguousarray(B)


def init(device=0):
    """CudaUtilsModule.init (cudaUtilsModule.jl:41-43) equivalent."""
    _check(lib().lsq_init(int(device)))


def init_devices(devices=None):
    """Bind several GPUs of the box (None = all visible): encode / update_codebooks / train_lsq / linscan calls
    then shard over them inside the library (the reference hard-codes device 0, encode_icm_cuda.jl:59-64)."""
    if devices is None:
        _check(lib().lsq_init_devices(None, 0))
    else:
        devs = np.ascontiguousarray(list(devices), np.int32)
        _check(lib().lsq_init_devices(_p(devs), len(devs)))
    return int(lib().lsq_num_bound_devices())


def num_bound_devices():
    return int(lib().lsq_num_bound_devices())


def finalize():
    """CudaUtilsModule.finit (cudaUtilsModule.jl:37-39) equivalent."""
    _check(lib().lsq_finalize())


def device_count():
    return int(lib().lsq_device_count())


def linscan_path(n, nq, m, d):
    """1 if linscan_lsq runs the tensor-core filter + exact rescoring for n base vectors and nq queries, 0 for the
    lookup-table scan."""
    return int(lib().lsq_linscan_path(ct.c_int64(n), ct.c_int64(nq), int(m), int(d)))


def linscan_last_phases():
    """{phase name: device ms} of the calling thread's most recent linscan call (needs LSQ_B200_ADC_TIMING set)."""
    L = lib()
    out = {}
    n = L.lsq_linscan_last_phases(-1, None, None)
    for i in range(n):
        ms, name = ct.c_float(0), ct.c_char_p()
        L.lsq_linscan_last_phases(i, ct.byref(ms), ct.byref(name))
        out[name.value.decode()] = ms.value
    out["_filter_products"] = int(L.lsq_linscan_last_phases(-2, None, None))
    return out


def launch_count():
    """CUDA kernels launched by the library so far in this process."""
    return int(lib().lsq_launch_count())


def version():
    return lib().lsq_version().decode()


def splitarray(n, nparts):
    """utils.jl:152-177 over the range 0..n-1 -> list of (lo, hi) half-open parts."""
    lo, hi = ct.c_int64(), ct.c_int64()
    out = []
    for p in range(nparts):
        _check(lib().lsq_splitarray(ct.c_int64(n), int(nparts), p, ct.byref(lo), ct.byref(hi)))
        out.append((lo.value, hi.value))
    return out


def randinit(n, m, h, rng=None):
    """initializations.jl:2-8: uniform 1..h Int16 codes, (n, m)."""
    rng = np.random.default_rng() if rng is None else rng
    return rng.integers(1, h + 1, size=(n, m)).astype(np.int16)


def make_to_look(seed, ils_iter, m, randord):
    out = np.zeros(m, np.int32)
    _check(lib().lsq_make_to_look(ct.c_uint64(seed), ct.c_uint32(ils_iter), int(m), int(bool(randord)), _p(out)))
    return out


def make_perturb(seed, ils_iter, g0, n, m, h, npert):
    slots = np.zeros((n, npert), np.uint8)
    vals = np.zeros((n, npert), np.int16)
    _check(lib().lsq_make_perturb(ct.c_uint64(seed), ct.c_uint32(ils_iter), ct.c_uint64(g0), ct.c_int64(n),
                                  int(m), int(h), int(npert), _p(slots), _p(vals)))
    return slots, vals


def get_unaries(X, C, V=False):
    """utils.jl:94-122 -> (m, n, h) float32."""
    X, C = _f32(X), _codebooks(C)
    n, d = X.shape
    m, h, _ = C.shape
    U = np.empty((m, n, h), np.float32)
    _check(lib().lsq_get_unaries(_p(X), d, ct.c_int64(n), _p(C), m, h, _p(U)))
    return U


def get_binaries(C):
    """utils.jl:125-144 -> (binaries (ncbi, h, h) with [idx][b][a], cbi (ncbi, 2) 1-based)."""
    C = _codebooks(C)
    m, h, d = C.shape
    ncbi = m * (m - 1) // 2
    G = np.empty((max(ncbi, 1), h, h), np.float32)
    cbi = np.zeros((max(ncbi, 1), 2), np.int32)
    _check(lib().lsq_get_binaries(_p(C), d, m, h, _p(G), _p(cbi)))
    return G[:ncbi], cbi[:ncbi]


def veccost(X, B, C):
    """utils.jl:225-254 -> (n,) float32."""
    X, C, B = _f32(X), _codebooks(C), _codes16(B)
    n, d = X.shape
    m, h, _ = C.shape
    out = np.empty(n, np.float32)
    _check(lib().lsq_veccost(_p(X), d, ct.c_int64(n), _p(B), _p(C), m, h, _p(out)))
    return out


def qerror(X, B, C):
    """utils.jl:257-285 -> float."""
    X, C, B = _f32(X), _codebooks(C), _codes16(B)
    n, d = X.shape
    m, h, _ = C.shape
    out = ct.c_float()
    _check(lib().lsq_qerror(_p(X), d, ct.c_int64(n), _p(B), _p(C), m, h, ct.byref(out)))
    return float(out.value)


def reconstruct(B, C):
    """utils.jl:203-223 -> (n, d) float32."""
    C, B = _codebooks(C), _codes16(B)
    n, m = B.shape
    _, h, d = C.shape
    out = np.empty((n, d), np.float32)
    _check(lib().lsq_reconstruct(_p(B), ct.c_int64(n), _p(C), d, m, h, _p(out)))
    return out


def quantize_norms(B, C, cbnorms):
    """utils.jl:6-31 -> (n,) int16, 1-based."""
    C, B, cbnorms = _codebooks(C), _codes16(B), _f32(cbnorms)
    n, m = B.shape
    _, h, d = C.shape
    out = np.empty(n, np.int16)
    _check(lib().lsq_quantize_norms(_p(B), ct.c_int64(n), _p(C), d, m, h, _p(cbnorms), len(cbnorms), _p(out)))
    return out


_ils_counter = 0


def reset_ils_counter(value=0):
    """Restart the implicit ILS-iteration counter of encoding_icm (see there)."""
    global _ils_counter
    _ils_counter = int(value)


def encoding_icm(X, oldB, C, niter, randord, npert, V=False, *, seed=0, ils_iter=None, g0=0):
    """encode_icm.jl:131-189: one ILS iteration; returns the new (n, m) int16 1-based codes.

    The reference draws its schedule from Julia's global RNG, so every call perturbs differently; here the
    schedule is Philox(seed, ils_iter, global vector index).  With ils_iter=None (the default) the
    iteration number comes from a module-level counter that advances on every call — like the Julia
    overlay's LSQ_B200_COUNTER — so the reference loop `for i = 1:ilsiter; B = encoding_icm(...)`
    (LSQ.jl:45-48, demo_lsq.jl:48-51) ported verbatim gets fresh perturbations each iteration.  Pass
    ils_iter explicitly for reproducible / sharded runs.
    """
    global _ils_counter
    if ils_iter is None:
        ils_iter = _ils_counter
        _ils_counter += 1
    X, C, oldB = _f32(X), _codebooks(C), _codes16(oldB)
    n, d = X.shape
    m, h, _ = C.shape
    if oldB.shape != (n, m):
        raise ValueError("oldB must be (n, m)")
    newB = np.empty_like(oldB)
    _check(lib().lsq_encoding_icm(_p(X), d, ct.c_int64(n), _p(oldB), _p(newB), _p(C), m, h, int(niter),
                                  int(bool(randord)), int(npert), ct.c_uint64(seed), ct.c_uint32(ils_iter),
                                  ct.c_uint64(g0), int(bool(V))))
    return newB


def encoding_viterbi(X, C, V=False):
    """encode_chain.jl:95-127: exact chain (ChainQ) encoding -> (n, m) int16 1-based codes."""
    X, C = _f32(X), _codebooks(C)
    n, d = X.shape
    m, h, _ = C.shape
    B = np.zeros((n, m), np.int16)
    _check(lib().lsq_encoding_viterbi(_p(X), d, ct.c_int64(n), _p(C), m, h, _p(B), int(bool(V))))
    return B


def encoding_icm_sched(X, oldB, C, niter, to_look, slots, vals, V=False):
    """encoding_icm with an explicit schedule (0-based to_look / slots / vals) — the parity entry."""
    X, C, oldB = _f32(X), _codebooks(C), _codes16(oldB)
    n, d = X.shape
    m, h, _ = C.shape
    to_look = np.ascontiguousarray(to_look, np.int32)
    slots = np.ascontiguousarray(slots, np.uint8).reshape(n, -1)
    vals = np.ascontiguousarray(vals, np.int16).reshape(n, -1)
    npert = slots.shape[1]
    newB = np.empty_like(oldB)
    _check(lib().lsq_encoding_icm_sched(_p(X), d, ct.c_int64(n), _p(oldB), _p(newB), _p(C), m, h, int(niter),
                                        _p(to_look), npert, _p(slots), _p(vals), int(bool(V))))
    return newB


def encode_icm_cuda(RX, B, C, ilsiters, icmiter, npert, randord, nsplits=2, V=False, *, seed=0, g0=0):
    """encode_icm_cuda.jl:253-296 -> (Bs: list of (n, m) int16, objs: (nr,) float32)."""
    RX, C, B = _f32(RX), _codebooks(C), _codes16(B)
    n, d = RX.shape
    m, h, _ = C.shape
    its = np.ascontiguousarray(ilsiters, np.int64)
    nr = len(its)
    Bs = np.zeros((nr, n, m), np.int16)
    objs = np.zeros(nr, np.float32)
    _check(lib().lsq_encode_icm_cuda(_p(RX), d, ct.c_int64(n), _p(B), _p(C), m, h, _p(its), nr, int(icmiter),
                                     int(npert), int(bool(randord)), int(nsplits), ct.c_uint64(seed),
                                     ct.c_uint64(g0), _p(Bs), _p(objs), int(bool(V))))
    return [Bs[r] for r in range(nr)], objs


def update_codebooks(X, B, h, V=False, codebook_upd_method="lsqr"):
    """codebook_update.jl:52-86 -> (m, h, d) float32."""
    if codebook_upd_method not in ("lsmr", "lsqr"):
        raise LsqError(1, "Codebook update method unknown")  # codebook_update.jl:59
    X, B = _f32(X), _codes16(B)
    n, d = X.shape
    m = B.shape[1]
    C = np.empty((m, h, d), np.float32)
    _check(lib().lsq_update_codebooks(_p(X), d, ct.c_int64(n), _p(B), m, int(h), _p(C),
                                      codebook_upd_method.encode(), int(bool(V))))
    return C


def linscan_lsq(B, X, C, dbnorms, R, k=10000):
    """Linscan.jl:46-73.  B (n, m) uint8 0-based; X (nq, d) queries; R (d, d) rotation (RX = R'X).
    Returns (dists (nq, k) float32, idx (nq, k) int32 1-based)."""
    B = np.ascontiguousarray(B)
    if B.dtype != np.uint8:
        raise TypeError("B must be Matrix{UInt8}")
    C = _codebooks(C)
    RX = _f32(_f32(X) @ _f32(R))
    dbnorms = _f32(dbnorms)
    n, m = B.shape
    nq, d = RX.shape
    h = C.shape[1]
    dists = np.zeros((nq, k), np.float32)
    res = np.zeros((nq, k), np.int32)
    _check(lib().lsq_linscan_lsq(_p(dists), _p(res), _p(B), _p(RX), _p(C), _p(dbnorms), nq, n, m, h, d, int(k)))
    return dists, res


def linscan_pq(B, X, C, b, k=10000):
    """Linscan.jl:5-27.  C: m codebooks (h, d/m).  Returns (dists, ids 1-based like `res .+= 1`)."""
    B = np.ascontiguousarray(B)
    if B.dtype != np.uint8:
        raise TypeError("B must be Matrix{UInt8}")
    C = _codebooks(C)
    X = _f32(X)
    n, m = B.shape
    nq, d = X.shape
    dists = np.zeros((nq, k), np.float32)
    res = np.zeros((nq, k), np.uint32)
    _check(lib().lsq_linscan_pq(_p(dists), _p(res), _p(B), _p(C), _p(X), n, ct.c_uint32(nq), int(b), int(k), m, d,
                                d // m))
    res += 1
    return dists, res


def linscan_opq(B, X, C, b, R, k=10000):
    """Linscan.jl:30-43: rotate the queries, then linscan_pq."""
    return linscan_pq(B, _f32(_f32(X) @ _f32(R)), C, b, k)


def eval_recall(ids_gnd, ids_predicted, k):
    """Linscan.jl:76-117 -> recall@i for i = 1..k (fractions, float64).  ids_predicted is (nq, >= k):
    row q is query q's ranked id list (a column of the Julia matrix)."""
    ids_gnd = np.ascontiguousarray(np.asarray(ids_gnd).reshape(-1).astype(np.int64).astype(np.int32))
    pred = np.asarray(ids_predicted)
    pred = np.ascontiguousarray(pred.astype(np.int64).astype(np.int32) if pred.dtype != np.int32 else pred)
    nq, ld = pred.shape
    assert nq == len(ids_gnd)   # Linscan.jl:85
    out = np.zeros(k, np.float64)
    _check(lib().lsq_eval_recall(_p(ids_gnd), _p(pred), nq, ld, int(k), _p(out)))
    return out


def kmeans1d(values, h, maxiter=100):
    """Norm-codebook stand-in for `kmeans(dbnorms, h)` (LSQ.jl:80) -> (centers (h,) ascending, iterations)."""
    v = _f32(np.asarray(values).reshape(-1))
    out = np.zeros(h, np.float32)
    it = ct.c_int()
    _check(lib().lsq_kmeans1d(_p(v), ct.c_int64(len(v)), int(h), int(maxiter), _p(out), ct.byref(it)))
    return out, it.value


def train_lsq(X, m, h, R, B, C, niter, ilsiter, icmiter, randord, npert, V=False, *, seed=0):
    """src/lsq/LSQ.jl:10-88 -> (C (m, h, d), B (n, m) int16, cbnorms (h,), B_norms (n,) int16, obj (niter,)).
    X (n, d); R (d, d) with R[i, j] = R(i, j) (or None for the identity); B initial codes, 1-based.
    The C argument is accepted for signature parity and ignored, exactly as the reference overwrites it
    (LSQ.jl:34) before its first use."""
    X, B = _f32(X), _codes16(B).copy()
    n, d = X.shape
    if B.shape != (n, m):
        raise ValueError("B must be (n, m)")
    Rm = None if R is None else np.ascontiguousarray(_f32(R).T)   # column-major bytes of the Julia matrix
    Cout = np.zeros((m, h, d), np.float32)
    cbnorms = np.zeros(h, np.float32)
    B_norms = np.zeros(n, np.int16)
    obj = np.zeros(max(niter, 1), np.float32)
    _check(lib().lsq_train_lsq(_p(X), d, ct.c_int64(n), int(m), int(h), _p(Rm), _p(B), _p(Cout), int(niter),
                               int(ilsiter), int(icmiter), int(bool(randord)), int(npert), ct.c_uint64(seed),
                               _p(cbnorms), _p(B_norms), _p(obj), int(bool(V))))
    return Cout, B, cbnorms, B_norms, obj[:niter]
