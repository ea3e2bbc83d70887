"""Device-resident sessions over the lsq_dev_* entry points (torch = device memory + streams only)."""
import ctypes as ct

import numpy as np
import torch

from . import api


def _ptr(t):
    return ct.c_void_p(t.data_ptr()) if t is not None else None


def _stream(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return ct.c_void_p(s.cuda_stream)


class EncodeSession:
    """X, codebooks, unary/pair tables, codes and costs resident in HBM across ILS iterations
    (the reference re-uploads X and codes every call: encode_icm_cuda.jl:79,134,146-153)."""

    def __init__(self, X, C, codes, g0=0, sliced=None, unary="exact"):
        assert X.is_cuda and X.dtype == torch.float32 and X.is_contiguous()
        assert C.is_cuda and C.dtype == torch.float32 and C.is_contiguous() and C.shape[1] == 256
        assert codes.is_cuda and codes.dtype == torch.uint8 and codes.is_contiguous()
        self.X, self.C, self.codes = X, C, codes
        self.n, self.d = X.shape
        self.m = C.shape[0]
        self.g0 = int(g0)
        self.unary = unary  # "exact" (parity) or "tc" (tensor cores, fast mode)
        L = api.lib()
        dev = X.device
        self.sliced = int(L.lsq_dev_icm_layout(self.m, ct.c_int64(self.n))) if sliced is None else int(sliced)
        self.T = torch.empty(L.lsq_dev_tables_bytes(self.m) // 4, dtype=torch.float32, device=dev)
        self.Ts = (torch.empty(L.lsq_dev_sliced_tables_bytes(self.m) // 4, dtype=torch.float32, device=dev)
                   if self.sliced else None)
        self.U = torch.empty((self.m, self.n, 256), dtype=torch.float32, device=dev)  # or [m][8][n][32]
        self.cost = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.set_codebooks(C)

    def set_codebooks(self, C):
        """(Re)build pair tables, unaries and the cost of the current codes for new codebooks."""
        L = api.lib()
        self.C = C
        api._check(L.lsq_dev_build_tables(_ptr(C), self.d, self.m, _ptr(self.T), _ptr(self.Ts), _stream()))
        if self.unary == "tc" and not self.sliced:
            api._check(L.lsq_dev_build_unaries_tc(_ptr(self.X), self.d, ct.c_int64(self.n), _ptr(C), self.m,
                                                  _ptr(self.U), _stream()))
        else:
            api._check(L.lsq_dev_build_unaries(_ptr(self.X), self.d, ct.c_int64(self.n), _ptr(C), self.m,
                                               _ptr(self.U), self.sliced, _stream()))
        self.refresh_cost()

    def refresh_cost(self):
        api._check(api.lib().lsq_dev_veccost(_ptr(self.X), self.d, ct.c_int64(self.n), _ptr(self.codes),
                                             _ptr(self.C), self.m, _ptr(self.cost), _stream()))

    def ils(self, niters, icmiter, npert, randord, seed=0, ils_iter0=0, slots=None, vals=None, orders=None):
        """`niters` ILS iterations in one launch; codes/cost updated in place."""
        if orders is None:
            orders = np.stack([api.make_to_look(seed, ils_iter0 + i, self.m, randord) for i in range(niters)])
        orders = np.ascontiguousarray(orders, np.int8)
        api._check(api.lib().lsq_dev_icm_ils(
            _ptr(self.X), self.d, ct.c_int64(self.n), _ptr(self.C), self.m, _ptr(self.U), _ptr(self.T),
            _ptr(self.Ts), self.sliced, _ptr(self.codes), _ptr(self.cost), int(icmiter), int(npert), orders.ctypes.data_as(ct.c_void_p),
            _ptr(slots), _ptr(vals), ct.c_uint64(seed), ct.c_uint32(ils_iter0), int(niters),
            ct.c_uint64(self.g0), None, None, None, _stream()))

    def qerror(self):
        return float(self.cost.double().mean().item())


def absmax(X):
    """max|x| of a float32 device tensor (one pass; needed once per data set, not per iteration)."""
    out = torch.zeros(1, dtype=torch.float32, device=X.device)
    api._check(api.lib().lsq_dev_absmax(_ptr(X), ct.c_int64(X.numel()), _ptr(out), _stream()))
    return float(out.item())


def cb_scale_exp(absmax_all, n_total):
    """Scale exponent of the fixed-point statistics from the GLOBAL max|x| and vector count."""
    return int(api.lib().lsq_cb_scale_exp(ct.c_float(absmax_all), ct.c_int64(n_total)))


def cb_accumulate(X, codes, m, scale_exp, stats=None):
    """Exact integer codebook-update statistics of one shard: int64 device tensor [mh*mh + mh*d] (counts,
    then fixed-point sums).  Shards add up with ONE all-reduce of this buffer, in any order, bit-exactly."""
    n, d = X.shape
    if stats is None:
        stats = torch.zeros(int(api.lib().lsq_cb_stats_len(m, d)), dtype=torch.int64, device=X.device)
    api._check(api.lib().lsq_dev_cb_accumulate(_ptr(X), d, ct.c_int64(n), _ptr(codes), m, int(scale_exp),
                                               _ptr(stats), _stream()))
    return stats


def cb_finalize(stats, m, d, scale_exp):
    """Summed statistics -> (Gram (mh, mh), Rhs (mh, d)) float64 for cb_solve."""
    mh = m * 256
    gram = torch.empty((mh, mh), dtype=torch.float64, device=stats.device)
    rhs = torch.empty((mh, d), dtype=torch.float64, device=stats.device)
    api._check(api.lib().lsq_dev_cb_finalize(_ptr(stats), m, d, int(scale_exp), _ptr(gram), _ptr(rhs), _stream()))
    return gram, rhs


def cb_stats(X, codes, m):
    """Single-shard convenience: (Gram, Rhs) float64 of this X alone."""
    n, d = X.shape
    e = cb_scale_exp(absmax(X), n)
    return cb_finalize(cb_accumulate(X, codes, m, e), m, d, e)


def cb_solve(gram, rhs, m, max_iter=0, tol=0.0):
    d = rhs.shape[1]
    C = torch.empty((m, 256, d), dtype=torch.float32, device=gram.device)
    iters = ct.c_int(0)
    api._check(api.lib().lsq_dev_cb_solve(_ptr(gram), _ptr(rhs), m, d, _ptr(C), int(max_iter), ct.c_double(tol),
                                          ct.byref(iters), _stream()))
    return C, iters.value


def linscan(codes, queries, codebooks, dbnorms, nn, lut_kind=0, subdim=0):
    """Device ADC scan; lut_kind 0 = LSQ (ids 1-based), 1 = PQ (ids 0-based)."""
    n, m = codes.shape
    nq, d = queries.shape
    dists = torch.empty((nq, nn), dtype=torch.float32, device=codes.device)
    ids = torch.empty((nq, nn), dtype=torch.int32, device=codes.device)
    api._check(api.lib().lsq_dev_linscan(_ptr(codes), ct.c_int64(n), m, _ptr(queries), nq, d, _ptr(codebooks),
                                         _ptr(dbnorms), int(lut_kind), int(subdim), int(nn), _ptr(dists), _ptr(ids),
                                         _stream()))
    return dists, ids


def adc_filter_values(codes, queries, codebooks, dbnorms):
    """Test hook for the tensor-core ADC prefilter (csrc/adc_tc.cu): the values dbnorm[v] - 2<q, xhat_v> as the
    bf16 hi/lo tcgen05 GEMM computes them, (nq, 128*ceil(n/128)) float32 (columns >= n are padding)."""
    n, m = codes.shape
    nq, d = queries.shape
    ld = 128 * ((n + 127) // 128)
    out = torch.full((nq, ld), float("nan"), dtype=torch.float32, device=codes.device)
    api._check(api.lib().lsq_dev_adc_filter_values(_ptr(codes), ct.c_int64(n), m, _ptr(queries), nq, d, _ptr(codebooks),
                                                   _ptr(dbnorms), _ptr(out), ct.c_int64(ld), _stream()))
    return out


def viterbi(X, C):
    """Chain (ChainQ) encoder with device tensors: builds the tables and unaries, runs the DP
    (encode_chain.jl:95-127); returns uint8 0-based codes (n, m).  The unary buffer is scratch."""
    n, d = X.shape
    m = C.shape[0]
    L = api.lib()
    T = torch.empty(L.lsq_dev_tables_bytes(m) // 4, dtype=torch.float32, device=X.device)
    U = torch.empty((m, n, 256), dtype=torch.float32, device=X.device)
    codes = torch.empty((n, m), dtype=torch.uint8, device=X.device)
    api._check(L.lsq_dev_build_tables(_ptr(C), d, m, _ptr(T), None, _stream()))
    api._check(L.lsq_dev_build_unaries(_ptr(X), d, ct.c_int64(n), _ptr(C), m, _ptr(U), 0, _stream()))
    api._check(L.lsq_dev_viterbi(_ptr(U), ct.c_int64(n), m, _ptr(T), _ptr(codes), _stream()))
    return codes


def eval_recall(ids_gnd, ids_predicted, k):
    """recall@1..k (float64 device tensor) from int32 device tensors: ground truth (nq,), ranked ids (nq, >= k)."""
    nq, ld = ids_predicted.shape
    out = torch.empty(k, dtype=torch.float64, device=ids_predicted.device)
    api._check(api.lib().lsq_dev_eval_recall(_ptr(ids_gnd), _ptr(ids_predicted), nq, ld, int(k), _ptr(out), _stream()))
    return out
