"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` for the collective.

The path shards naturally:
  * encoding  — vectors are independent: contiguous `splitarray` shards (utils.jl:152-177), no
    communication; the perturbation RNG is keyed by the GLOBAL vector index (`g0` + local index), so the
    codes are identical for any number of GPUs.
  * codebook update — the single exchange step of the path: every rank accumulates its shard's
    Gram = A'A (counts) and Rhs = A'X' (fixed-point sums) as exact int64, ONE all-reduce (NCCL over
    NVLink/NVSwitch on the GPU box, gloo in the CPU tests) sums them, then every rank runs the same
    deterministic solve and ends up with identical codebooks (no broadcast needed) — the same bits as a
    single-GPU run, because integer sums do not depend on the order.
This module is the one-process-per-GPU (torchrun) plumbing; a single process can instead bind several GPUs
with lsq_b200.init_devices() and the library shards the host-pointer calls internally.
Nothing here touches the oracle; the statistics/solve functions are injected by the caller so that the
host logic can be exercised on CPU tensors with gloo.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, size=None, rank=None):
    """[lo, hi) of this rank's contiguous shard of 0..n-1 (the reference's splitarray rule)."""
    r, s = world()
    rank = r if rank is None else rank
    size = s if size is None else size
    return api.splitarray(n, size)[rank]


def allreduce_stats(stats):
    """The one collective of the path: sum the codebook-update statistics over all ranks, in place.
    `stats` is ONE buffer (device.cb_accumulate: int64 counts + fixed-point sums), reduced by ONE all-reduce.
    Integer sums are exact, so the result — and the codebooks solved from it — do not depend on the number
    of ranks or on the reduction order.  (A (gram, rhs) pair of float tensors, as the CPU tests inject, is
    packed into one buffer first.)"""
    _, size = world()
    if isinstance(stats, (tuple, list)):
        gram, rhs = stats
        if size > 1:
            flat = torch.cat([gram.reshape(-1), rhs.reshape(-1).to(gram.dtype)])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            gram.copy_(flat[: gram.numel()].view_as(gram))
            rhs.copy_(flat[gram.numel():].view_as(rhs).to(rhs.dtype))
        return gram, rhs
    if size > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def global_scale_exp(X, n_total=None):
    """Scale exponent of the fixed-point statistics: needs max|x| and the vector count over ALL ranks (two
    scalars, exchanged once per data set — X does not change between outer iterations)."""
    from . import device as dev
    t = torch.tensor([dev.absmax(X)], dtype=torch.float32, device=X.device)
    c = torch.tensor([X.shape[0]], dtype=torch.int64, device=X.device)
    _, size = world()
    if size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if n_total is None:
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return dev.cb_scale_exp(float(t.item()), int(n_total if n_total is not None else c.item()))


def global_mean(local_sum, local_count, device=None):
    """Mean over all ranks of a per-vector quantity given this rank's sum and count (for qerror)."""
    t = torch.tensor([float(local_sum), float(local_count)], dtype=torch.float64, device=device)
    _, size = world()
    if size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0] / max(float(t[1]), 1.0))


def gather_codes(codes):
    """All shards' codes, concatenated in rank order, on every rank (shards may differ by one row)."""
    if codes.dtype == torch.int16 and world()[1] > 1:  # gloo has no int16 collectives
        return gather_rows(codes.to(torch.int32)).to(torch.int16)
    return gather_rows(codes)


def gather_rows(t):
    """Concatenate per-rank row blocks (shards may differ by one row) in rank order on every rank."""
    _, size = world()
    if size == 1:
        return t
    counts = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(size)]
    dist.all_gather(counts, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    nmax = int(max(int(c) for c in counts))
    pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: int(c)] for p, c in zip(parts, counts)])


def linscan_sharded(queries, scan_fn, gather=True):
    """ADC scan over several GPUs by QUERY partitioning: codes, norms and codebooks are
    replicated (12-20 MB per million vectors), every rank scans its contiguous `splitarray` slice of the
    queries with `scan_fn(queries_slice) -> (dists, ids)`, and the per-query results are disjoint, so no
    merge is needed — only an optional all-gather of the (nq, nn) outputs."""
    lo, hi = shard_bounds(queries.shape[0])
    dists, ids = scan_fn(queries[lo:hi])
    if not gather:
        return dists, ids, (lo, hi)
    return gather_rows(dists), gather_rows(ids), (lo, hi)


def update_codebooks_sharded(X, codes, m, stats_fn=None, solve_fn=None, scale_exp=None):
    """update_codebooks (codebook_update.jl:52-86) over sharded data: local statistics -> ONE all-reduce
    -> replicated solve.  Defaults run the CUDA kernels (exact integer statistics: bit-identical codebooks for
    any number of ranks); stats_fn / solve_fn let the CPU tests inject stand-ins."""
    if stats_fn is not None or solve_fn is not None:
        gram, rhs = allreduce_stats(tuple(stats_fn(X, codes)))
        return solve_fn(gram, rhs)
    from . import device as dev
    if scale_exp is None:
        scale_exp = global_scale_exp(X)
    stats = allreduce_stats(dev.cb_accumulate(X, codes, m, scale_exp))
    gram, rhs = dev.cb_finalize(stats, m, X.shape[1], scale_exp)
    return dev.cb_solve(gram, rhs, m)[0]


def train_lsq_sharded(X, codes, C, niter, ilsiter, icmiter, randord, npert, seed=0, g0=0, verbose=False):
    """The alternation of train_lsq (LSQ.jl:57-66) with X/codes/tables resident on each GPU:
    qerror -> update_codebooks (one all-reduce) -> ilsiter ILS iterations (no communication).
    X (n_local, d) float32 cuda; codes (n_local, m) uint8 cuda, 0-based; C (m, 256, d) float32 cuda.
    Returns (C, codes, objective history)."""
    from . import device as dev
    m = C.shape[0]
    sess = dev.EncodeSession(X, C, codes, g0=g0)
    scale_exp = global_scale_exp(X)   # once: X is fixed for the whole alternation
    obj = []
    it_count = 0
    for it in range(niter):
        obj.append(global_mean(float(sess.cost.double().sum().item()), X.shape[0], device=X.device))
        if verbose and world()[0] == 0:
            print(f"{it + 1:3d} {obj[-1]:e}")
        C = update_codebooks_sharded(X, codes, m, scale_exp=scale_exp)
        sess.set_codebooks(C)
        sess.ils(ilsiter, icmiter, npert, randord, seed=seed, ils_iter0=it_count)
        it_count += ilsiter
    obj.append(global_mean(float(sess.cost.double().sum().item()), X.shape[0], device=X.device))
    return C, codes, np.asarray(obj)
