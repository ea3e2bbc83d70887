"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard bounds, the statistics all-reduce, code
gathering and sharding invariance of the schedule.  The per-shard compute here is done by the oracle
(this is a test); on the GPU box the same plumbing carries the CUDA kernels (tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, size, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        import lsq_b200
        import oracle
        from oracle import codebook_update as cu
        from util import make_problem
        par = lsq_b200.parallel
        n, d, m = 1001, 16, 3
        X, C, B = make_problem(5, n, d, m)
        B0 = (B - 1).astype(np.int16)
        lo, hi = par.shard_bounds(n)
        assert (lo, hi) == oracle.splitarray(n, size)[rank]
        # --- encode: shard keyed by global index, no communication ---
        part, _ = oracle.encoding_icm(X[lo:hi], B0[lo:hi], C, 2, True, 2, seed=4, ils_iter=1, g0=lo)
        allc = par.gather_codes(torch.from_numpy(part.astype(np.int16))).numpy()
        whole, _ = oracle.encoding_icm(X, B0, C, 2, True, 2, seed=4, ils_iter=1)
        assert np.array_equal(allc, whole)
        # --- codebook update: local stats -> one all-reduce -> replicated solve ---
        def stats(Xs, cs):
            G, R = cu.gram_stats(Xs, cs, 256)
            return torch.from_numpy(G), torch.from_numpy(R)
        def solve(G, R):
            K = np.linalg.pinv(G.numpy(), rcond=1e-12, hermitian=True) @ R.numpy()
            return K.reshape(m, 256, d).astype(np.float32)
        Cs = par.update_codebooks_sharded(X[lo:hi], whole[lo:hi], m, stats_fn=stats, solve_fn=solve)
        Cw = cu.update_codebooks_exact(X, whole, 256)
        assert np.allclose(Cs, Cw, rtol=0, atol=1e-4 * np.abs(Cw).max())
        # integer statistics (the layout of device.cb_accumulate) travel in ONE all-reduce of ONE int64 buffer,
        # and the result is exactly the statistics of the whole set (integer sums are order-independent)
        def int_stats(Xs, cs, e):
            G, R = cu.gram_stats(Xs, cs, 256)   # R exact here: the test data are small integers
            return torch.from_numpy(np.concatenate([G.reshape(-1), np.rint(R * 2.0 ** e).reshape(-1)]).astype(np.int64))
        calls = []
        orig = dist.all_reduce
        dist.all_reduce = lambda t, *a, **k: (calls.append((t.numel(), t.dtype)), orig(t, *a, **k))[1]
        try:
            S = par.allreduce_stats(int_stats(X[lo:hi], whole[lo:hi], 20))
        finally:
            dist.all_reduce = orig
        assert calls == [((m * 256) ** 2 + m * 256 * d, torch.int64)], calls   # exactly one collective
        assert torch.equal(S, int_stats(X, whole, 20))
        # every rank holds the same codebooks (replicated solve, no broadcast)
        t = torch.from_numpy(Cs.copy())
        ts = [torch.empty_like(t) for _ in range(size)]
        dist.all_gather(ts, t)
        assert all(torch.equal(ts[0], x) for x in ts)
        # --- global objective from per-shard sums ---
        cost = oracle.veccost(X[lo:hi], whole[lo:hi], Cs)
        q = par.global_mean(float(cost.astype(np.float64).sum()), hi - lo)
        assert abs(q - oracle.qerror(X, whole, Cs)) <= 1e-9 * q
        # --- linscan: queries partitioned, codes replicated, no merge ---
        from util import make_scan_problem
        codes, queries, codebooks, norms = make_scan_problem(9, 3000, 11, 16, 4)
        def scan(qs):
            dd, ii = oracle.linscan_lsq(codes, qs, codebooks, norms, 20)
            return torch.from_numpy(dd), torch.from_numpy(ii)
        dd, ii, (qlo, qhi) = par.linscan_sharded(queries, scan)
        dw, iw = oracle.linscan_lsq(codes, queries, codebooks, norms, 20)
        assert (qlo, qhi) == oracle.splitarray(11, size)[rank]
        assert np.array_equal(ii.numpy(), iw) and np.array_equal(dd.numpy(), dw)
        out[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        out[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import oracle
    oracle.build()  # before forking, so the workers do not race on the build
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(out.get(r) == "ok" for r in range(2)), dict(out)
