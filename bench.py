#!/usr/bin/env python
"""bench.py — headline measurement of the LSQ hot path on B200.

Metric (BASELINE.json): ICM encode vectors/sec at d=128, m=8, h=256.  One "step" = one full encode of
the base shard: pair tables + unary tables + initial cost + `--ils` ILS iterations (icmiter=4, npert=4,
random visit order) — BASELINE configs[1] ("1M base d=128, ICM 16 iters, 1xB200").

  value   : whole-job vectors/s with X, codebooks and the initial codes already resident in HBM
  e2e     : the same through the reference-facing C-ABI call lsq_encode_icm_cuda with HOST buffers
            (pinned), H2D of X/codes and D2H of the codes inside the timed region
  roofline: the ICM sweep kernel alone (CUDA events around its launch), algorithmic bytes per
            vector-ILS-iteration = icmiter*m*h*4 + 2m + 4d + 8 against the measured HBM peak
  cpu_baseline: the oracle port (oracle/lsq_oracle.c, reference loop structure) on the host cores, on a
            bounded sample, rank 0 only

`--impl reference` times that CPU port alone (Julia is not installable here).
Multi-GPU (torchrun): every rank encodes its own n-vector shard, no data-path collective: weak scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

D, M, H = 128, 8, 256
ICMITER, NPERT = 4, 4


def algorithmic_bytes_per_vec_iter():
    return ICMITER * M * H * 4 + 2 * M + 4 * D + 8


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel, n, ils):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, if it was taken on this
    launch shape (profiles/r*/traffic.json); else None."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", "traffic.json"))):
        try:
            t = json.load(open(path)).get(kernel)
        except (OSError, ValueError):
            continue
        if t and t.get("launch", "").startswith(f"n={n}, {ils} ILS"):
            best = t["dram_bytes_read"] + t["dram_bytes_write"]
    return best


def recall_probe(L, seed=0, ntrain=30000, nbase=200000, nquery=500, m=7, d=D, knn=10):
    """recall@1 of the full flow (train -> encode base -> quantise norms -> ADC scan) on a small
    synthetic problem, m = 7 codebooks + 1 norm byte like demos/demo_lsq.jl:14.  Not timed."""
    rng = np.random.default_rng(seed)
    W = (rng.standard_normal((24, d)) * 12).astype(np.float32)

    def gen(n):
        x = rng.standard_normal((n, 24)).astype(np.float32) @ W + rng.standard_normal((n, d)).astype(np.float32) * 2.0
        return np.clip(np.floor(np.abs(x)), 0, 255).astype(np.float32)

    xt, xb, xq = gen(ntrain), gen(nbase), gen(nquery)
    B = L.randinit(ntrain, m, H, rng)
    C = L.update_codebooks(xt, B, H)
    it = 0
    for outer in range(4):
        for _ in range(4):
            B = L.encoding_icm(xt, B, C, ICMITER, True, NPERT, seed=7, ils_iter=it)
            it += 1
        C = L.update_codebooks(xt, B, H)
    norms = (L.reconstruct(B, C) ** 2).sum(1)
    cbn = np.quantile(norms, (np.arange(256) + 0.5) / 256).astype(np.float32)
    Bb = L.encode_icm_cuda(xb, L.randinit(nbase, m, H, rng), C, [8], ICMITER, NPERT, True, 1, seed=8)[0][-1]
    dbn = cbn[L.quantize_norms(Bb, C, cbn) - 1]
    _, idx = L.linscan_lsq((Bb - 1).astype(np.uint8), xq, C, dbn, np.eye(d, dtype=np.float32), knn)
    bn = (xb.astype(np.float64) ** 2).sum(1)
    gt = np.argmin(bn[None, :] - 2.0 * xq.astype(np.float64) @ xb.T.astype(np.float64), axis=1) + 1
    rec = L.eval_recall(gt, idx, knn)
    return {"recall@1": float(rec[0]), f"recall@{knn}": float(rec[-1]),
            "config": f"synthetic rank-24 descriptors, {ntrain} train / {nbase} base / {nquery} queries, m={m}+norm byte"}


def cpu_port_rate(n_sample, ils_total, seed=1):
    """Oracle port on all host threads: one ILS iteration over n_sample vectors -> vectors/s for a
    full `ils_total`-iteration encode (per-vector work is independent and identical per iteration)."""
    import oracle
    from util import make_problem
    X, C, B = make_problem(seed, n_sample, D, M)
    B0 = (B - 1).astype(np.int16)
    threads = oracle.use_all_cores()   # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    oracle.encoding_icm(X[:256], B0[:256], C, ICMITER, True, NPERT, seed=seed, ils_iter=0, nworkers=threads)
    t0 = time.perf_counter()
    oracle.encoding_icm(X, B0, C, ICMITER, True, NPERT, seed=seed, ils_iter=0, nworkers=threads)
    dt = time.perf_counter() - t0
    return n_sample / (dt * ils_total), threads, dt


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; Julia absent) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ns = args.cpu_sample
    rates = []
    for _ in range(args.warmup):
        cpu_port_rate(min(ns, 2000), args.ils)
    t_steps = []
    for _ in range(args.steps):
        r, threads, dt = cpu_port_rate(ns, args.ils)
        rates.append(r); t_steps.append(dt)
    value = statistics.median(rates)
    sample = f"{ns} vectors x 1 ILS iteration per step, scaled to {args.ils} iterations (vectors independent)"
    print(json.dumps({
        "impl": "reference", "metric": "icm_encode_vectors_per_sec", "value": value, "unit": "vectors/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.median(t_steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"LSQ ICM encode m={M} h={H} d={D}, {args.ils} ILS iters x icmiter={ICMITER}, npert={NPERT}",
                   "n_per_gpu": args.n, "cpu_sample": ns},
        "cpu_baseline": {"value": value, "unit": "vectors/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="base vectors per GPU (weak) or in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --n vectors per GPU (the contract's default); strong: --n vectors split over the GPUs")
    ap.add_argument("--ils", type=int, default=16, help="ILS iterations per encode (LSQ-16)")
    ap.add_argument("--cpu-sample", type=int, default=400000)
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-recall", action="store_true", help="skip the untimed recall@1 probe of the full flow")
    ap.add_argument("--m", type=int, default=8, help="codebooks (BASELINE configs[2] uses 16)")
    ap.add_argument("--unary", default="exact", choices=["exact", "tc"],
                    help="exact = parity path (sequential fp32 chain); tc = tcgen05 3xTF32 fast mode (tolerance-checked)")
    args = ap.parse_args()
    global M
    M = args.m
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import lsq_b200
    from lsq_b200 import device as lsqdev
    from util import make_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lsq_b200.init(local)
    dev = torch.device("cuda", local)
    n, ils = args.n, args.ils
    if args.scaling == "strong":  # fixed total, contiguous splitarray shards (utils.jl:152-177)
        lo, hi = lsq_b200.splitarray(args.n, world)[rank]
        n, g0 = hi - lo, lo
    else:
        g0 = rank * n  # global index of this shard's first vector
    n_total = args.n if args.scaling == "strong" else world * n

    # synthetic SIFT-shaped shard; codebooks identical on every rank
    _, C_h, _ = make_problem(0, 16, D, M)
    rng = np.random.default_rng(1000 + rank)
    from util import sift_like
    X_h = sift_like(rng, n, D)
    B_h = rng.integers(1, H + 1, size=(n, M)).astype(np.int16)
    X = torch.from_numpy(X_h).to(dev)
    C = torch.from_numpy(C_h).to(dev)
    codes0 = torch.from_numpy((B_h - 1).astype(np.uint8)).to(dev)
    codes = codes0.clone()
    sess = lsqdev.EncodeSession(X, C, codes, g0=g0, unary=args.unary)
    if args.unary == "tc":
        os.environ["LSQ_B200_UNARY"] = "tc"  # the host-API (e2e) leg follows the same mode
    orders = np.stack([lsq_b200.make_to_look(1, i, M, True) for i in range(ils)])

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kern_ms = []

    def step(timed):
        codes.copy_(codes0)
        sess.set_codebooks(C)              # pair tables + norms + unaries + cost of the initial codes
        if timed:
            a, b = ev(), ev()
            a.record()
        sess.ils(ils, ICMITER, NPERT, True, seed=1, ils_iter0=0, orders=orders)
        if timed:
            b.record()
            kern_ms.append((a, b))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        step(True)
    t1.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    k_ms = torch.tensor([statistics.mean(a.elapsed_time(b) for a, b in kern_ms)], device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(k_ms, op=dist.ReduceOp.MAX)
    ms_per_step = total_ms.item() / args.steps
    value = n_total / (ms_per_step * 1e-3)
    qerr = sess.qerror()

    # ---- e2e: the C-ABI host call with pinned host buffers, copies inside the timed region ----
    Xp = torch.from_numpy(X_h).pin_memory()
    Bp = torch.from_numpy(B_h).pin_memory()
    its = np.array([ils], np.int64)
    e2e_t = []
    e2e_out = None
    for i in range(1 + args.e2e_steps if args.e2e_steps > 0 else 0):
        sync()
        w0 = time.perf_counter()
        e2e_out, _ = lsq_b200.encode_icm_cuda(Xp.numpy(), Bp.numpy(), C_h, its, ICMITER, NPERT, True, 1, seed=1, g0=g0)
        if i > 0:
            e2e_t.append(time.perf_counter() - w0)
    e2e_ms = torch.tensor([1e3 * statistics.median(e2e_t) if e2e_t else float("nan")], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    same = bool(np.array_equal(e2e_out[0], codes.cpu().numpy().astype(np.int16) + 1)) if e2e_out else None

    if rank == 0:
        peak, peak_src = measured_peak()
        abytes = algorithmic_bytes_per_vec_iter() * n * ils
        achieved = abytes / (k_ms.item() * 1e-3) / 1e9
        out = {
            "metric": "icm_encode_vectors_per_sec", "value": value, "unit": "vectors/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"LSQ ICM encode m={M} h={H} d={D}, {ils} ILS iters x icmiter={ICMITER}, npert={NPERT} (BASELINE configs[1] base-set encode)",
                       "n_per_gpu": n, "parallelism": f"shard{world}", "l2": f"inputs larger than L2 ({M * n * 1024 / 1e9:.0f} GB unaries per GPU)",
                       "step": "pair tables + unaries + cost + all ILS iterations",
                       "unary_mode": "exact fp32 chain (parity)" if args.unary == "exact" else "tcgen05 3xTF32 (fast mode, not bit-exact)"},
            "vector_ils_iters_per_sec": value * ils,
            "qerror": qerr, "e2e_codes_equal_resident_codes": same,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(f"icm_ils_warp_kernel<{M}>", n, ils), "kernel": f"icm_ils_warp_kernel<{M}>", "kernel_ms": k_ms.item(),
                         "peak_source": peak_src, "bytes_per_vector_iter": algorithmic_bytes_per_vec_iter(),
                         "algorithmic_bytes_per_launch": abytes},
            "e2e": {"value": n_total / (e2e_ms.item() * 1e-3), "unit": "vectors/s",
                    "h2d_bytes_per_step": int(X_h.nbytes + B_h.nbytes + C_h.nbytes), "d2h_bytes_per_step": int(B_h.nbytes),
                    "ms_per_step": e2e_ms.item(), "api": "lsq_encode_icm_cuda (host pointers)"},
            "gpu_launches": 5 * args.steps,
            "clocks": clocks,
        }
        if not args.no_recall:
            out["recall"] = recall_probe(lsq_b200)
        if not args.no_cpu:
            r, threads, dt = cpu_port_rate(args.cpu_sample, ils)
            out["cpu_baseline"] = {"value": r, "unit": "vectors/s", "cores": threads, "kind": "port",
                                   "sample": f"{args.cpu_sample} vectors x 1 ILS iteration ({dt:.1f} s), scaled to {ils} iterations"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
