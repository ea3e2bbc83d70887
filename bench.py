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


def workload_config(args, world):
    """`config` of the JSON line — identical for the b200 arm and the reference arm."""
    n_per_gpu = args.n // world if args.scaling == "strong" else args.n
    return {"workload": f"LSQ ICM encode m={M} h={H} d={D}, {args.ils} ILS iters x icmiter={ICMITER}, npert={NPERT} "
                        f"(BASELINE configs[1] base-set encode)",
            "n_total": args.n if args.scaling == "strong" else args.n * world, "n_per_gpu": n_per_gpu,
            "parallelism": f"shard{world}", "l2": f"inputs larger than L2 ({M * n_per_gpu * 1024 / 1e9:.1f} GB unaries per GPU)",
            "step": "pair tables + unaries + cost + all ILS iterations",
            "unary_mode": "exact fp32 chain (parity)" if args.unary == "exact" else "tcgen05 3xTF32 (fast mode, not bit-exact)"}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; Julia absent) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
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
        "ms_per_step": 1e3 * statistics.median(t_steps), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "vectors/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))



# ----------------------------------------------------------------------------------------------------
# Secondary workloads (python bench.py --workload adc|chain|legacy ...): they live in this file because
# their CPU legs run the oracle / the reference's own binaries, which only bench.py and tests/ may do.

# ----------------------------------------------------------------------------------------------------
# ADC linear scan (BASELINE configs[4]): n base codes x nq queries, top-nn.  Reports queries/s, the effective
# scan bandwidth nq*n*(m+4)/t against the measured HBM peak (an *effective* figure — the
# query-tiled kernel re-serves the code array from L2), the LUT lookup rate against the shared-memory bound
# 148 SMs x 32 banks x f_SM, and the reference's own C++ (oracle/_ref, OpenMP) on a query subsample.
def adc_measure(m, n, nq, nn, d=128, reps=5, cpu_queries=64, check_queries=16, rank=0, world=1, cpu_leg=True):
    """One ADC configuration on the current device.  With world > 1 the QUERIES are partitioned over the ranks
    (codes replicated, no merge: SURVEY.md §8e) and the time is the max over ranks."""
    import torch
    import torch.distributed as dist
    import lsq_b200
    from lsq_b200 import device as dev
    from util import make_scan_problem
    peak, _ = measured_peak()
    codes, queries, codebooks, norms = make_scan_problem(50 + m, n, nq, d, m)
    qlo, qhi = lsq_b200.splitarray(nq, world)[rank]
    dc, dq = torch.from_numpy(codes).cuda(), torch.from_numpy(queries[qlo:qhi]).cuda()
    dcb, dn = torch.from_numpy(codebooks).cuda(), torch.from_numpy(norms).cuda()
    for _ in range(2):
        dd, di = dev.linscan(dc, dq, dcb, dn, nn)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    l0 = lsq_b200.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    each = []
    for _ in range(reps):   # every call timed on its own: the median is immune to a host hiccup between launches
        a.record()
        dd, di = dev.linscan(dc, dq, dcb, dn, nn)
        b.record()
        torch.cuda.synchronize()
        each.append(a.elapsed_time(b))
    launches = (lsq_b200.launch_count() - l0) // reps
    t = torch.tensor([float(np.median(each))], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    out = {"m": m, "n": n, "nq": nq, "nn": nn, "ms": ms, "queries_per_s": nq / (ms * 1e-3), "launches_per_call": launches,
           "ms_each_call": [round(x, 3) for x in each], "timing": "median of the calls, each bracketed by CUDA events"}
    tc = bool(lsq_b200.linscan_path(n, qhi - qlo, m, d))
    out["path"] = ("tcgen05 bf16 filter GEMM (queries resident in TMEM, norm and threshold folded in, sign-bit epilogue) + "
                   "exact rescoring of the survivors (csrc/adc_tc.cu)") if tc else "lookup-table scan (csrc/linscan.cu)"
    eff = nq * n * (m + 4) / (ms * 1e-3) / 1e9
    lookups = nq * n * m / (ms * 1e-3)
    out.update({"pairs_per_s": nq * n / (ms * 1e-3), "effective_scan_GBps": eff, "effective_frac_of_hbm_peak": eff / peak / world})
    if tc:
        # two products of K = d plus one extra K step of 16 per (query, base vector) pair; the whole call is charged
        flops = 2.0 * nq * n * (2 * d + 16)
        bf16_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1387.9)) \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1387.9
        out.update({"filter_tflops_whole_call": flops / (ms * 1e-3) / 1e12,
                    "frac_of_measured_bf16_peak_whole_call": flops / (ms * 1e-3) / 1e12 / bf16_peak / world,
                    "equivalent_lookups_per_s": lookups})
        # device time of every phase of one more call (CUDA events inside the library), and the filter kernel alone
        # against the tensor-core peak: 2 (2 d + 16) flop per (query, base vector) pair
        os.environ["LSQ_B200_ADC_TIMING"] = "1"
        try:
            dev.linscan(dc, dq, dcb, dn, nn)
            torch.cuda.synchronize()
            ph = lsq_b200.linscan_last_phases()
        finally:
            del os.environ["LSQ_B200_ADC_TIMING"]
        products = int(ph.pop("_filter_products", 2)) or 2   # 1: hi(q).hi(x) only, 2: + hi(q).lo(x); chosen on the device per call
        out["phases_ms"] = {k: round(v, 3) for k, v in ph.items()}
        out["filter_products"] = products
        flops = 2.0 * nq * n * (products * d + 16)
        out.update({"filter_tflops_whole_call": flops / (ms * 1e-3) / 1e12,
                    "frac_of_measured_bf16_peak_whole_call": flops / (ms * 1e-3) / 1e12 / bf16_peak / world})
        if ph.get("filter"):
            ftf = 2.0 * (qhi - qlo) * n * (products * d + 16) / (ph["filter"] * 1e-3) / 1e12
            out["filter_kernel"] = {"ms": ph["filter"], "tflops": ftf, "frac_of_measured_bf16_peak": ftf / bf16_peak,
                                    "frac_of_nominal_bf16_peak": ftf / 2250.0, "bound": f"tensor ({products * (d // 16) + 1} tcgen05.mma of 128x128x16 per 128 queries x 128 base vectors)"}
        # the lookup-table scan on the same problem (the path every shape took before this round's tensor-core filter)
        os.environ["LSQ_B200_ADC"] = "scan"
        try:
            dev.linscan(dc, dq, dcb, dn, nn)
            torch.cuda.synchronize()
            a.record()
            ds, is_ = dev.linscan(dc, dq, dcb, dn, nn)
            b.record()
            torch.cuda.synchronize()
        finally:
            del os.environ["LSQ_B200_ADC"]
        out["lookup_scan_ms"] = a.elapsed_time(b)
        out["equal_to_lookup_scan"] = bool(torch.equal(ds, dd) and torch.equal(is_, di))
        # PQ / OPQ tables (linscan_aqd.cpp) on the same shape: the same filter (||q||^2 - 2<q,xhat> + ||xhat||^2) vs the lookup kernel
        if d % m == 0:
            sub = d // m
            dcen = torch.from_numpy(np.ascontiguousarray(codebooks[:, :sub].reshape(m, 256, sub))).cuda()
            pq = {}
            for mode in ("tc", "scan"):
                os.environ["LSQ_B200_ADC"] = mode
                try:
                    dev.linscan(dc, dq, dcen, None, nn, lut_kind=1, subdim=sub)
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(3):
                        a.record()
                        r_pq = dev.linscan(dc, dq, dcen, None, nn, lut_kind=1, subdim=sub)
                        b.record()
                        torch.cuda.synchronize()
                        ts.append(a.elapsed_time(b))
                finally:
                    del os.environ["LSQ_B200_ADC"]
                pq[mode] = (float(np.median(ts)), r_pq)
            out["pq"] = {"subdim": sub, "ms": pq["tc"][0], "lookup_scan_ms": pq["scan"][0],
                         "equal_to_lookup_scan": bool(torch.equal(pq["tc"][1][0], pq["scan"][1][0]) and
                                                      torch.equal(pq["tc"][1][1], pq["scan"][1][1]))}
    else:
        out.update({"lookups_per_s": lookups, "lookup_frac_of_smem_bound": lookups / (world * 148 * 32 * 1.965e9)})
    if cpu_leg and rank == 0:
        import oracle
        # exactness spot check against the reference .so (or the oracle restatement), which is also the CPU baseline
        fn = oracle.ref_linscan_lsq if oracle.ref_available() else oracle.linscan_lsq
        t0 = time.perf_counter()
        dr, ir = fn(codes, queries[:cpu_queries], codebooks, norms, nn)
        cpu_s = time.perf_counter() - t0
        k = min(check_queries, qhi - qlo)
        out["exact_vs_reference"] = bool(np.array_equal(ir[:k], di[:k].cpu().numpy()) and np.array_equal(dr[:k], dd[:k].cpu().numpy()))
        out["cpu_baseline"] = {"kind": "reference" if oracle.ref_available() else "port", "cores": oracle.num_threads(),
                               "queries_per_s": cpu_queries / cpu_s, "sample": f"{cpu_queries} queries x {n} codes"}
    return out


def run_adc(argv):
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--nn", type=int, default=1000)
    ap.add_argument("--m", type=int, nargs="+", default=[8, 16])
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-queries", type=int, default=64)
    ap.add_argument("--check-queries", type=int, default=16)
    args = ap.parse_args(argv)
    import lsq_b200
    lsq_b200.init(0)
    for m in args.m:
        r = adc_measure(m, args.n, args.nq, args.nn, args.d, args.reps, args.cpu_queries, args.check_queries)
        r.update({"metric": "adc_scan_queries_per_sec", "value": r["queries_per_s"], "unit": "queries/s"})
        print(json.dumps(r))


# Chain (Viterbi / ChainQ) encoder: resident-data timing of the kernel, the whole host call, and the CPU
# oracle on a subsample.  (m-1)*65536 candidate transitions per vector, one FADD + one FMNMX each: bound by
# the ALU pipe (one FMNMX per 2 cycles per SM sub-partition).
def run_chain(argv):
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, nargs="+", default=[8])
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-n", type=int, default=4000)
    args = ap.parse_args(argv)
    import ctypes as ct
    import torch
    import lsq_b200
    import oracle
    from util import make_problem
    lsq_b200.init(0)
    L = lsq_b200.lib()
    for m in args.m:
        X, C, _ = make_problem(60 + m, args.n, args.d, m)
        dX, dC = torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda()
        dU = torch.empty((m, args.n, 256), dtype=torch.float32, device="cuda")
        dT = torch.empty((m, m, 256, 256), dtype=torch.float32, device="cuda")
        codes = torch.empty((args.n, m), dtype=torch.uint8, device="cuda")
        st = ct.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: ct.c_void_p(t.data_ptr())
        assert L.lsq_dev_build_tables(P(dC), args.d, m, P(dT), None, st) == 0
        # lsq_dev_viterbi consumes dU (forward messages overwrite it in place): rebuild it before every run
        # and time only the chain kernel
        ms = 0.0
        for rep in range(args.reps + 1):
            assert L.lsq_dev_build_unaries(P(dX), args.d, ct.c_int64(args.n), P(dC), m, P(dU), 0, st) == 0
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            assert L.lsq_dev_viterbi(P(dU), ct.c_int64(args.n), m, P(dT), P(codes), st) == 0
            b.record()
            torch.cuda.synchronize()
            if rep > 0:
                ms += a.elapsed_time(b) / args.reps
        t0 = time.perf_counter()
        Bh = lsq_b200.encoding_viterbi(X, C)
        host_s = time.perf_counter() - t0
        same = bool(np.array_equal(Bh, codes.cpu().numpy().astype(np.int16) + 1))
        t0 = time.perf_counter()
        Bo = oracle.encoding_viterbi(X[: args.cpu_n], C)
        cpu_s = time.perf_counter() - t0
        exact = bool(np.array_equal(Bo + 1, Bh[: args.cpu_n]))
        pairs = args.n * (m - 1) * 65536
        print(json.dumps({
            "metric": "viterbi_encode_vectors_per_sec", "value": args.n / (ms * 1e-3), "unit": "vectors/s", "m": m,
            "n": args.n, "kernel_ms": ms, "host_call_s": host_s, "host_equals_device": same, "exact_vs_oracle": exact,
            "pairs_per_s": pairs / (ms * 1e-3), "alu_pipe_frac_at_2_cycles_per_pair": pairs / 32 * 2 / (ms * 1e-3) / (148 * 4 * 1.965e9),
            "cpu_baseline": {"kind": "port", "cores": oracle.num_threads(), "vectors_per_s": args.cpu_n / cpu_s,
                             "sample": f"{args.cpu_n} vectors"},
        }))


# The reference's own GPU kernel on the same B200: src/encodings/cuda/cudautils.cu compiled UNMODIFIED for
# sm_100a (oracle/Makefile -> oracle/_ref/cudautils_sm100a.cubin) and launched as encode_icm_cuda.jl:158-186
# launches it, with the pair tables already resident and without its per-visit H2D upload / host sync /
# perturb / veccost: a generous lower bound on its time.  A timing baseline only, not an oracle.
def run_legacy(argv, n=1_000_000, m=8, d=128, ils=16, icmiter=4):
    import ctypes as ct
    import torch
    from cuda.bindings import driver as cu
    import lsq_b200 as L
    from lsq_b200 import device as dev
    from util import make_problem, sift_like
    cubin = os.path.join(ROOT, "oracle", "_ref", "cudautils_sm100a.cubin")
    if not os.path.exists(cubin):
        print(json.dumps({"legacy_kernel": "unavailable", "why": "oracle/_ref/cudautils_sm100a.cubin not built"}))
        return
    L.init(0)
    torch.cuda.init()
    _, C_h, _ = make_problem(0, 16, d, m)
    rng = np.random.default_rng(1000)
    X = torch.from_numpy(sift_like(rng, n, d)).cuda()
    C = torch.from_numpy(C_h).cuda()
    codes = torch.from_numpy(rng.integers(0, 256, size=(n, m)).astype(np.uint8)).cuda()
    sess = dev.EncodeSession(X, C, codes.clone(), sliced=0)
    torch.cuda.synchronize()
    # --- ours: the ILS kernel alone ---
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sess.ils(ils, icmiter, 4, True, seed=1)
    sess.codes.copy_(codes); sess.refresh_cost()
    a.record(); sess.ils(ils, icmiter, 4, True, seed=1); b.record(); torch.cuda.synchronize()
    ours_ms = a.elapsed_time(b)
    # --- legacy: condition_icm3 per node visit ---
    err, mod = cu.cuModuleLoad(cubin.encode())
    assert err == cu.CUresult.CUDA_SUCCESS, err
    err, fn = cu.cuModuleGetFunction(mod, b"condition_icm3")
    assert err == cu.CUresult.CUDA_SUCCESS, err
    T = sess.T.view(m, m, 256, 256)
    bbs = [torch.cat([T[k, l] for l in range(m) if l != k]).contiguous() for k in range(m)]  # cat(2, bbs...)
    codes_soa = codes.t().contiguous()  # d_codek[i_idx + n*i]
    U = sess.U  # [m][n][256]
    stream = torch.cuda.current_stream().cuda_stream

    def visit(k):
        args = [ct.c_void_p(U[k].data_ptr()), ct.c_void_p(bbs[k].data_ptr()), ct.c_void_p(codes_soa.data_ptr()),
                ct.c_int(k), ct.c_int(m), ct.c_int(n)]
        arr = (ct.c_void_p * len(args))(*[ct.cast(ct.pointer(x), ct.c_void_p) for x in args])
        (e,) = cu.cuLaunchKernel(fn, n, 1, 1, 1, 256, 1, 0, stream, ct.addressof(arr), 0)
        assert e == cu.CUresult.CUDA_SUCCESS, e

    orders = [L.make_to_look(1, i, m, True) for i in range(ils)]
    for k in range(m):
        visit(k)
    torch.cuda.synchronize()
    a.record()
    for i in range(ils):
        for _ in range(icmiter):
            for k in orders[i]:
                visit(int(k))
    b.record(); torch.cuda.synchronize()
    legacy_ms = a.elapsed_time(b)
    print(json.dumps({
        "workload": f"n={n} m={m} d={d}, {ils} ILS iterations x icmiter={icmiter}",
        "lsq_b200_icm_kernel_ms": ours_ms,
        "legacy_condition_icm3_ms": legacy_ms, "legacy_launches": ils * icmiter * m,
        "legacy_note": "tables pre-resident, no perturb/veccost/H2D/sync (generous to the legacy path)",
        "speedup_kernel_only": legacy_ms / ours_ms,
    }))


def _ev():
    import torch
    return torch.cuda.Event(enable_timing=True)


def _max_over_ranks(x, world, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def train_step_block(rank, world, dev, n_per_gpu, ilsiter=8, steps=2):
    """One outer iteration of train_lsq on BASELINE configs[3]'s shape (10 M vectors over 8 GPUs = 1.25 M per GPU,
    sharded): local statistics -> ONE all-reduce of the integer statistics (NCCL) -> replicated solve -> tables +
    unaries -> `ilsiter` ILS iterations.  Each phase is timed with CUDA events on the launching stream; the
    figures are the max over ranks of the per-step mean."""
    import torch
    import torch.distributed as dist
    import lsq_b200
    from lsq_b200 import device as lsqdev, parallel as par
    from util import make_problem, sift_like
    _, C_h, _ = make_problem(0, 16, D, M)
    rng = np.random.default_rng(2000 + rank)
    X = torch.from_numpy(sift_like(rng, n_per_gpu, D)).to(dev)
    codes = torch.from_numpy(rng.integers(0, H, size=(n_per_gpu, M)).astype(np.uint8)).to(dev)
    sess = lsqdev.EncodeSession(X, torch.from_numpy(C_h).to(dev), codes, g0=rank * n_per_gpu)
    scale_exp = par.global_scale_exp(X)
    stats = torch.zeros(int(lsq_b200.lib().lsq_cb_stats_len(M, D)), dtype=torch.int64, device=dev)
    # outputs of finalize / solve are allocated once: no allocator call inside the timed phases
    import ctypes as ct
    L = lsq_b200.lib()
    gram = torch.empty((M * H, M * H), dtype=torch.float64, device=dev)
    rhs = torch.empty((M * H, D), dtype=torch.float64, device=dev)
    C = torch.empty((M, H, D), dtype=torch.float32, device=dev)
    cg = ct.c_int(0)
    stream = lambda: ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    phases = {k: [] for k in ("cb_accumulate_ms", "allreduce_ms", "finalize_solve_ms", "tables_unaries_ms", "ils_ms", "total_ms")}
    it = 0
    warm = 2
    for s in range(steps + warm):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e = [_ev() for _ in range(6)]
        e[0].record()
        stats.zero_()
        lsqdev.cb_accumulate(X, codes, M, scale_exp, stats=stats)
        e[1].record()
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        e[2].record()
        lsq_b200.api._check(L.lsq_dev_cb_finalize(ct_ptr(stats), M, D, scale_exp, ct_ptr(gram), ct_ptr(rhs), stream()))
        lsq_b200.api._check(L.lsq_dev_cb_solve(ct_ptr(gram), ct_ptr(rhs), M, D, ct_ptr(C), 0, ct.c_double(0.0), ct.byref(cg), stream()))
        cg_iters = cg.value
        e[3].record()
        sess.set_codebooks(C)
        e[4].record()
        sess.ils(ilsiter, ICMITER, NPERT, True, seed=3, ils_iter0=it)
        e[5].record()
        it += ilsiter
        torch.cuda.synchronize()
        if s >= warm:
            for k, i in zip(list(phases)[:5], range(5)):
                phases[k].append(e[i].elapsed_time(e[i + 1]))
            phases["total_ms"].append(e[0].elapsed_time(e[5]))
    out = {k: _max_over_ranks(statistics.mean(v), world, dev) for k, v in phases.items()}
    out.update({"workload": f"train_lsq outer iteration, m={M}, {n_per_gpu} vectors per GPU x {world} GPUs, ilsiter={ilsiter} "
                            f"(BASELINE configs[3] shape: 10 M over 8 GPUs)",
                "allreduce": (f"torch.distributed NCCL all_reduce(SUM) of {stats.numel() * 8 / 1e6:.1f} MB int64" if world > 1 else None),
                "cg_iterations": cg_iters, "qerror_after": sess.qerror(),
                "vectors_per_s": world * n_per_gpu / (out["total_ms"] * 1e-3)})
    return out


def trained_codebooks_block(dev, X, codes0, ils, steps=3, ntrain=100_000):
    """BASELINE configs[1] literally: codebooks from a short train_lsq on 100 K training vectors (same
    distribution), then the resident-data encode of the base shard with THOSE codebooks.  The skip rate of the
    ICM kernel is a property of the data, so this is reported beside the random-codebook headline."""
    import torch
    import lsq_b200
    from lsq_b200 import device as lsqdev
    from util import sift_like
    rng = np.random.default_rng(4242)
    Xt = sift_like(rng, ntrain, D)
    Bt = lsq_b200.randinit(ntrain, M, H, rng)
    t0 = time.perf_counter()
    Ct, _, _, _, obj = lsq_b200.train_lsq(Xt, M, H, None, Bt, None, 4, 4, ICMITER, True, NPERT, seed=5)
    train_s = time.perf_counter() - t0
    C = torch.from_numpy(Ct).to(dev)
    codes = codes0.clone()
    sess = lsqdev.EncodeSession(X, C, codes)
    n = X.shape[0]
    orders = np.stack([lsq_b200.make_to_look(1, i, M, True) for i in range(ils)])
    visits = torch.zeros(1, dtype=torch.int64, device=dev)
    ms, kms = [], []
    for s in range(steps + 1):
        codes.copy_(codes0)
        torch.cuda.synchronize()
        a, b, c = _ev(), _ev(), _ev()
        a.record()
        sess.set_codebooks(C)
        b.record()
        if s == 0:
            lsq_b200.lib().lsq_dev_icm_visit_counter(ct_ptr(visits))
        sess.ils(ils, ICMITER, NPERT, True, seed=1, ils_iter0=0, orders=orders)
        if s == 0:
            lsq_b200.lib().lsq_dev_icm_visit_counter(None)
        c.record()
        torch.cuda.synchronize()
        if s > 0:
            ms.append(a.elapsed_time(c)); kms.append(b.elapsed_time(c))
    peak, _ = measured_peak()
    k = statistics.mean(kms)
    return {"value_per_gpu": n / (statistics.mean(ms) * 1e-3), "unit": "vectors/s", "n_per_gpu": n, "ms_per_step": statistics.mean(ms),
            "kernel_ms": k, "roofline_frac": algorithmic_bytes_per_vec_iter() * n * ils / (k * 1e-3) / 1e9 / peak,
            "visits_per_vector_iter": visits.item() / (n * ils), "qerror": sess.qerror(),
            "codebooks": f"lsq_train_lsq on {ntrain} vectors, 4 outer x 4 ILS iterations ({train_s:.2f} s, objective {float(obj[0]):.1f} -> {float(obj[-1]):.1f})"}


def fast_mode_block(dev, X, codes0, C, ils, orders, qerr_exact, codes_exact, steps=3):
    """The same step with the unary tables built on the tensor cores (tcgen05 kind::tf32, 3xTF32 split, fp32
    accumulation in TMEM): not bit-exact with the sequential fp32 chain of the parity path, so it is opt-in
    (LSQ_B200_UNARY=tc) and held to the north-star's tolerance (quantisation error within 1e-5 relative)."""
    import torch
    from lsq_b200 import device as lsqdev
    codes = codes0.clone()
    sess = lsqdev.EncodeSession(X, C, codes, unary="tc")
    ms, ums = [], []
    for s in range(steps + 1):
        codes.copy_(codes0)
        torch.cuda.synchronize()
        a, b, c = _ev(), _ev(), _ev()
        a.record()
        sess.set_codebooks(C)
        b.record()
        sess.ils(ils, ICMITER, NPERT, True, seed=1, ils_iter0=0, orders=orders)
        c.record()
        torch.cuda.synchronize()
        if s > 0:
            ms.append(a.elapsed_time(c)); ums.append(a.elapsed_time(b))
    n = X.shape[0]
    q = sess.qerror()
    return {"value_per_gpu": n / (statistics.mean(ms) * 1e-3), "unit": "vectors/s", "ms_per_step": statistics.mean(ms),
            "tables_unaries_cost_ms": statistics.mean(ums), "qerror": q, "qerror_rel_diff_vs_exact": abs(q - qerr_exact) / qerr_exact,
            "codes_equal_to_exact_frac": float((codes == codes_exact).all(dim=1).float().mean().item()),
            "unary": "tcgen05 kind::tf32 3xTF32, codebook operand in TMEM, TMA raw-tile ring (csrc/unary_tc.cu)"}


def ct_ptr(t):
    import ctypes as ct
    return ct.c_void_p(t.data_ptr())


def main():
    pre = argparse.ArgumentParser(add_help=False)
    pre.add_argument("--workload", default="icm", choices=["icm", "adc", "chain", "legacy"])
    ns, rest = pre.parse_known_args()
    if ns.workload != "icm":
        return {"adc": run_adc, "chain": run_chain, "legacy": run_legacy}[ns.workload](rest)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="icm", help="icm (headline, default) | adc | chain | legacy")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="base vectors in total (strong) or per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): --n vectors split over the GPUs, the north-star's 1 M-vector encode; "
                         "weak: --n vectors per GPU.  Identical at 1 GPU.")
    ap.add_argument("--ils", type=int, default=16, help="ILS iterations per encode (LSQ-16)")
    ap.add_argument("--cpu-sample", type=int, default=400000)
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-recall", action="store_true", help="skip the untimed recall@1 probe of the full flow")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary blocks (adc, train_step, trained_codebooks, weak, in-library multi-GPU)")
    ap.add_argument("--train-n", type=int, default=1_250_000, help="vectors per GPU of the train_step block (configs[3])")
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
    from util import make_problem, sift_like

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")   # host-side barrier for the leg where rank 0 drives every GPU
    lsq_b200.init(local)
    dev = torch.device("cuda", local)
    ils = args.ils

    def shard_of(n_arg, scaling):
        if scaling == "strong":  # fixed total, contiguous splitarray shards (utils.jl:152-177)
            lo, hi = lsq_b200.splitarray(n_arg, world)[rank]
            return hi - lo, lo, n_arg
        return n_arg, rank * n_arg, world * n_arg

    _, C_h, _ = make_problem(0, 16, D, M)   # codebooks identical on every rank
    C = torch.from_numpy(C_h).to(dev)
    orders = np.stack([lsq_b200.make_to_look(1, i, M, True) for i in range(ils)])
    if args.unary == "tc":
        os.environ["LSQ_B200_UNARY"] = "tc"  # the host-API (e2e) leg follows the same mode

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def measure(n, g0, warmup, steps, sampler=None):
        """Resident-data steps on this rank's shard: returns (ms_per_step, kernel_ms, state)."""
        rng = np.random.default_rng(1000 + rank)
        X_h = sift_like(rng, n, D)
        B_h = rng.integers(1, H + 1, size=(n, M)).astype(np.int16)
        X = torch.from_numpy(X_h).to(dev)
        codes0 = torch.from_numpy((B_h - 1).astype(np.uint8)).to(dev)
        codes = codes0.clone()
        sess = lsqdev.EncodeSession(X, C, codes, g0=g0, unary=args.unary)
        kern = []

        def step(timed):
            codes.copy_(codes0)
            sess.set_codebooks(C)              # pair tables + norms + unaries + cost of the initial codes
            if timed:
                a, b = _ev(), _ev()
                a.record()
            sess.ils(ils, ICMITER, NPERT, True, seed=1, ils_iter0=0, orders=orders)
            if timed:
                b.record()
                kern.append((a, b))

        for _ in range(warmup):
            step(False)
        sync()
        if sampler is not None:
            sampler.start()
        l0 = lsq_b200.launch_count()
        t0, t1 = _ev(), _ev()
        t0.record()
        for _ in range(steps):
            step(True)
        t1.record()
        sync()
        launches = lsq_b200.launch_count() - l0
        clocks = sampler.stop() if sampler is not None else None
        total = _max_over_ranks(t0.elapsed_time(t1), world, dev)
        k_ms = _max_over_ranks(statistics.mean(a.elapsed_time(b) for a, b in kern), world, dev)
        return total / steps, k_ms, dict(X_h=X_h, B_h=B_h, X=X, codes0=codes0, codes=codes, sess=sess, launches=launches,
                                         clocks=clocks, step=step)

    n, g0, n_total = shard_of(args.n, args.scaling)
    ms_per_step, k_ms, S = measure(n, g0, args.warmup, args.steps, ClockSampler(local) if rank == 0 else None)
    value = n_total / (ms_per_step * 1e-3)
    sess, X_h, B_h, codes = S["sess"], S["X_h"], S["B_h"], S["codes"]
    qerr = sess.qerror()

    # executed node visits per vector and ILS iteration (device counter; one extra untimed step)
    visits = torch.zeros(1, dtype=torch.int64, device=dev)
    lsq_b200.lib().lsq_dev_icm_visit_counter(ct_ptr(visits))
    S["step"](False)
    torch.cuda.synchronize()
    lsq_b200.lib().lsq_dev_icm_visit_counter(None)
    visits_pvi = visits.item() / (n * ils)

    # ---- e2e: the C-ABI host call, host<->device copies inside the timed region; pinned and pageable sources ----
    its = np.array([ils], np.int64)

    def e2e_leg(Xsrc, Bsrc):
        ts, out = [], None
        for i in range(1 + args.e2e_steps if args.e2e_steps > 0 else 0):
            sync()
            w0 = time.perf_counter()
            out, _ = lsq_b200.encode_icm_cuda(Xsrc, Bsrc, C_h, its, ICMITER, NPERT, True, 1, seed=1, g0=g0)
            if i > 0:
                ts.append(time.perf_counter() - w0)
        ms = _max_over_ranks(1e3 * statistics.median(ts) if ts else float("nan"), world, dev)
        return ms, out, [round(1e3 * t, 2) for t in ts]

    Xp, Bp = torch.from_numpy(X_h).pin_memory(), torch.from_numpy(B_h).pin_memory()
    e2e_ms, e2e_out, e2e_all = e2e_leg(Xp.numpy(), Bp.numpy())
    same = bool(np.array_equal(e2e_out[0], codes.cpu().numpy().astype(np.int16) + 1)) if e2e_out else None
    e2e_pg_ms, pg_out, e2e_pg_all = e2e_leg(X_h, B_h)      # plain numpy arrays: pageable, like Julia's
    same_pg = bool(np.array_equal(pg_out[0], e2e_out[0])) if pg_out else None
    del Xp, Bp

    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        abytes = algorithmic_bytes_per_vec_iter() * n * ils
        achieved = abytes / (k_ms * 1e-3) / 1e9
        out = {
            "metric": "icm_encode_vectors_per_sec", "value": value, "unit": "vectors/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "vector_ils_iters_per_sec": value * ils,
            "qerror": qerr, "e2e_codes_equal_resident_codes": same, "pageable_codes_equal_pinned_codes": same_pg,
            "visits_per_vector_iter": visits_pvi, "nominal_visits_per_vector_iter": ICMITER * M,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(f"icm_ils_warp_kernel<{M}>", n, ils), "kernel": f"icm_ils_warp_kernel<{M}>", "kernel_ms": k_ms,
                         "peak_source": peak_src, "bytes_per_vector_iter": algorithmic_bytes_per_vec_iter(),
                         "algorithmic_bytes_per_launch": abytes,
                         "note": "effective figure (SURVEY.md 8d streamed-unary model); the physical limiter is the SM<->L2 gather path, see profiles/"},
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "vectors/s",
                    "h2d_bytes_per_step": int(X_h.nbytes + B_h.nbytes + C_h.nbytes), "d2h_bytes_per_step": int(B_h.nbytes),
                    "ms_per_step": e2e_ms, "ms_each_call_rank0": e2e_all,
                    "api": "lsq_encode_icm_cuda (host pointers, pinned source), median of the timed calls, max over ranks"},
            "e2e_pageable": {"value": n_total / (e2e_pg_ms * 1e-3), "unit": "vectors/s", "ms_per_step": e2e_pg_ms, "ms_each_call_rank0": e2e_pg_all,
                             "api": "lsq_encode_icm_cuda (plain pageable numpy arrays, as a Julia caller passes them)",
                             "h2d": os.environ.get("LSQ_B200_H2D", "staged")},
            "gpu_launches": S["launches"],
            "clocks": S["clocks"],
        }

    # free the headline shard before the secondary blocks
    codes0_keep, X_keep = S["codes0"], S["X"]
    del sess, S
    torch.cuda.empty_cache()

    if not args.no_extras:
        # the same base shard with codebooks trained on 100 K vectors (configs[1] as written); rank 0's figure
        tr = trained_codebooks_block(dev, X_keep, codes0_keep, ils)
        if rank == 0:
            out["trained_codebooks"] = tr
        if args.unary == "exact" and D % 8 == 0:
            fm = fast_mode_block(dev, X_keep, codes0_keep, C, ils, orders, qerr, codes)
            if rank == 0:
                out["fast_mode"] = fm
    del codes0_keep, X_keep
    torch.cuda.empty_cache()

    if not args.no_extras:
        if world > 1 and args.scaling == "strong":
            # the weak-scaling figure (fixed work per GPU), kept as an extra key
            wn, wg0, wtot = shard_of(args.n, "weak")
            wms, wk, WS = measure(wn, wg0, 1, 2)
            del WS
            torch.cuda.empty_cache()
            if rank == 0:
                out["weak"] = {"value": wtot / (wms * 1e-3), "unit": "vectors/s", "n_per_gpu": wn, "ms_per_step": wms, "kernel_ms": wk}
        # train_lsq outer iteration at the configs[3] shape, with the all-reduce timed
        ts = train_step_block(rank, world, dev, args.train_n)
        torch.cuda.empty_cache()
        if rank == 0:
            out["train_step"] = ts
        # ADC scan, configs[4]; with several ranks the queries are partitioned
        adc = [adc_measure(m_, 1_000_000, 10_000, 1000, rank=rank, world=world, cpu_leg=not args.no_cpu) for m_ in (8, 16)]
        torch.cuda.empty_cache()
        if rank == 0:
            out["adc"] = adc
        # one process driving ALL GPUs through the C ABI (lsq_init_devices): what a Julia caller gets.  Rank 0 runs
        # it while the other ranks wait at a HOST barrier (no kernel of theirs is resident).
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)
            if rank == 0:
                out["inlib_multi_gpu"] = inlib_block(world, args.n, ils, C_h, args.e2e_steps)
            dist.barrier(group=cpu_group)

    if rank == 0:
        if not args.no_recall:
            out["recall"] = recall_probe(lsq_b200)
        if not args.no_cpu:
            r, threads, dt = cpu_port_rate(args.cpu_sample, ils)
            out["cpu_baseline"] = {"value": r, "unit": "vectors/s", "cores": threads, "kind": "port",
                                   "sample": f"{args.cpu_sample} vectors x 1 ILS iteration ({dt:.1f} s), scaled to {ils} iterations"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


def inlib_block(ndev, n, ils, C_h, e2e_steps):
    """lsq_init_devices(0..ndev-1) in THIS process, then the same host-pointer encode of all n vectors: the library
    shards it over the devices from the one calling thread.  Wall-clock, copies included."""
    import lsq_b200
    from util import sift_like
    rng = np.random.default_rng(999)
    X = sift_like(rng, n, D)
    B = rng.integers(1, H + 1, size=(n, M)).astype(np.int16)
    its = np.array([ils], np.int64)
    res = {"devices": ndev, "n": n}
    for label, devs in (("one_device", [0]), ("all_devices", list(range(ndev)))):
        lsq_b200.finalize()
        lsq_b200.init_devices(devs)
        ts, outc = [], None
        for i in range(1 + max(1, e2e_steps)):
            w0 = time.perf_counter()
            outc, _ = lsq_b200.encode_icm_cuda(X, B, C_h, its, ICMITER, NPERT, True, 1, seed=1)
            if i > 0:
                ts.append(time.perf_counter() - w0)
        res[label] = {"ms": 1e3 * statistics.median(ts), "vectors_per_s": n / statistics.median(ts)}
        res.setdefault("codes", outc[0])
        res["codes_equal"] = bool(np.array_equal(res["codes"], outc[0]))
    del res["codes"]
    res["speedup"] = res["one_device"]["ms"] / res["all_devices"]["ms"]
    res["source"] = "pageable numpy arrays"
    # the same call from pinned host memory (what the pageable path is up against)
    import torch
    Xp, Bp = torch.from_numpy(X).pin_memory(), torch.from_numpy(B).pin_memory()
    ts = []
    for i in range(1 + max(1, e2e_steps)):
        w0 = time.perf_counter()
        lsq_b200.encode_icm_cuda(Xp.numpy(), Bp.numpy(), C_h, its, ICMITER, NPERT, True, 1, seed=1)
        if i > 0:
            ts.append(time.perf_counter() - w0)
    res["all_devices_pinned_source"] = {"ms": 1e3 * statistics.median(ts), "vectors_per_s": n / statistics.median(ts),
                                        "speedup_vs_one_device_pageable": res["one_device"]["ms"] / (1e3 * statistics.median(ts))}
    del Xp, Bp
    # lsq_train_lsq sharded inside the library: the statistics exchange by NCCL (ncclAllReduce on the clique built with
    # ncclCommInitAll) and by the library's own fused peer-memory kernel (the finalize kernel reads every device's
    # statistics over NVLink and sums them itself); identical results, device-timed on the primary GPU
    tr = {}
    ref_out = None
    for backend in ("nccl", "p2p"):
        os.environ["LSQ_B200_ALLREDUCE"] = backend
        lsq_b200.finalize()
        lsq_b200.init_devices(list(range(ndev)))
        ts = []
        for i in range(2):
            w0 = time.perf_counter()
            Ct, Bt, _, _, obj = lsq_b200.train_lsq(X, M, H, None, B, None, 2, 2, ICMITER, True, NPERT, seed=3)
            ts.append(time.perf_counter() - w0)
        tr[backend] = {"train_lsq_ms": 1e3 * min(ts), "exchange_plus_finalize_ms": float(lsq_b200.lib().lsq_last_collective_ms())}
        if ref_out is None:
            ref_out = (Ct, Bt)
        else:
            tr["backends_bit_identical"] = bool(np.array_equal(ref_out[0], Ct) and np.array_equal(ref_out[1], Bt))
    os.environ.pop("LSQ_B200_ALLREDUCE", None)
    tr["workload"] = f"lsq_train_lsq on {n} vectors, 2 outer iterations x 2 ILS iterations, {ndev} devices, host call incl. copies"
    res["train_lsq"] = tr
    lsq_b200.finalize()
    lsq_b200.init(0)
    return res


if __name__ == "__main__":
    main()
