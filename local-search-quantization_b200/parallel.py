"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` for the collective.

The path shards naturally:
  * encoding  — vectors are independent: contiguous `splitarray` shards (utils.jl:152-177), no
    communication; the perturbation RNG is keyed by the GLOBAL vector index (`g0` + local index), so the
    codes are identical for any number of GPUs.
  * codebook update — the single exchange step of the path: every rank accumulates its shard's
    Gram = A'A and Rhs = A'X' (float64), ONE all-reduce (NCCL over NVLink/NVSwitch on the GPU box, gloo
    in the CPU tests) sums them, then every rank runs the same deterministic solve and ends up with
    identical codebooks (no broadcast needed).
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


def allreduce_stats(gram, rhs):
    """The one collective of the path: sum the codebook-update statistics over all ranks, in place.
    Gram and Rhs travel in ONE all-reduce: `device.cb_stats` allocates them back to back in one buffer,
    which is then reduced as a whole; separately allocated tensors are packed into one buffer first."""
    _, size = world()
    if size > 1:
        adjacent = (gram.is_contiguous() and rhs.is_contiguous() and gram.dtype == rhs.dtype
                    and gram.untyped_storage().data_ptr() == rhs.untyped_storage().data_ptr()
                    and rhs.storage_offset() == gram.storage_offset() + gram.numel())
        if adjacent:
            flat = torch.as_strided(gram, (gram.numel() + rhs.numel(),), (1,), gram.storage_offset())
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        else:
            flat = torch.cat([gram.reshape(-1), rhs.reshape(-1).to(gram.dtype)])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            gram.copy_(flat[: gram.numel()].view_as(gram))
            rhs.copy_(flat[gram.numel():].view_as(rhs).to(rhs.dtype))
    return gram, rhs


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


def update_codebooks_sharded(X, codes, m, stats_fn=None, solve_fn=None):
    """update_codebooks (codebook_update.jl:52-86) over sharded data: local statistics -> all-reduce
    -> replicated solve.  Defaults run the CUDA kernels (lsq_dev_cb_stats / lsq_dev_cb_solve)."""
    if stats_fn is None or solve_fn is None:
        from . import device as dev
        stats_fn = stats_fn or (lambda X_, c_: dev.cb_stats(X_, c_, m))
        solve_fn = solve_fn or (lambda g_, r_: dev.cb_solve(g_, r_, m)[0])
    gram, rhs = stats_fn(X, codes)
    allreduce_stats(gram, rhs)
    return solve_fn(gram, rhs)


def train_lsq_sharded(X, codes, C, niter, ilsiter, icmiter, randord, npert, seed=0, g0=0, verbose=False):
    """The alternation of train_lsq (LSQ.jl:57-66) with X/codes/tables resident on each GPU:
    qerror -> update_codebooks (one all-reduce) -> ilsiter ILS iterations (no communication).
    X (n_local, d) float32 cuda; codes (n_local, m) uint8 cuda, 0-based; C (m, 256, d) float32 cuda.
    Returns (C, codes, objective history)."""
    from . import device as dev
    m = C.shape[0]
    sess = dev.EncodeSession(X, C, codes, g0=g0)
    obj = []
    it_count = 0
    for it in range(niter):
        obj.append(global_mean(float(sess.cost.double().sum().item()), X.shape[0], device=X.device))
        if verbose and world()[0] == 0:
            print(f"{it + 1:3d} {obj[-1]:e}")
        C = update_codebooks_sharded(X, codes, m)
        sess.set_codebooks(C)
        sess.ils(ilsiter, icmiter, npert, randord, seed=seed, ils_iter0=it_count)
        it_count += ilsiter
    obj.append(global_mean(float(sess.cost.double().sum().item()), X.shape[0], device=X.device))
    return C, codes, np.asarray(obj)
