#!/usr/bin/env python
"""ADC linear-scan benchmark (BASELINE configs[4]): n base codes x nq queries, top-nn, m in {8, 16}.

Reports, per m: queries/s, effective scan bandwidth nq*n*(m+4)/t against the measured HBM peak
(an *effective* figure — the query-tiled kernel re-serves the code array from L2), the
LUT lookup rate nq*n*m/t against the shared-memory bound 148 SMs x 32 banks x f_SM, and the
reference's own C++ (oracle/_ref, OpenMP, all host cores) on a query subsample as the CPU baseline.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--nn", type=int, default=1000)
    ap.add_argument("--m", type=int, nargs="+", default=[8, 16])
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-queries", type=int, default=64)
    ap.add_argument("--check-queries", type=int, default=16)
    args = ap.parse_args()
    import torch
    import lsq_b200
    from lsq_b200 import device as dev
    import oracle
    from util import make_scan_problem
    lsq_b200.init(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for m in args.m:
        codes, queries, codebooks, norms = make_scan_problem(50 + m, args.n, args.nq, args.d, m)
        dc, dq = torch.from_numpy(codes).cuda(), torch.from_numpy(queries).cuda()
        dcb, dn = torch.from_numpy(codebooks).cuda(), torch.from_numpy(norms).cuda()
        for _ in range(2):
            dd, di = dev.linscan(dc, dq, dcb, dn, args.nn)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            dd, di = dev.linscan(dc, dq, dcb, dn, args.nn)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.reps
        # exactness spot check against the reference .so (or the oracle restatement)
        k = args.check_queries
        fn = oracle.ref_linscan_lsq if oracle.ref_available() else oracle.linscan_lsq
        t0 = time.perf_counter()
        dr, ir = fn(codes, queries[: args.cpu_queries], codebooks, norms, args.nn)
        cpu_s = time.perf_counter() - t0
        exact = bool(np.array_equal(ir[:k], di[:k].cpu().numpy()) and np.array_equal(dr[:k], dd[:k].cpu().numpy()))
        eff = args.nq * args.n * (m + 4) / (ms * 1e-3) / 1e9
        lookups = args.nq * args.n * m / (ms * 1e-3)
        print(json.dumps({
            "metric": "adc_scan_queries_per_sec", "value": args.nq / (ms * 1e-3), "unit": "queries/s", "m": m,
            "n": args.n, "nq": args.nq, "nn": args.nn, "ms": ms, "exact_vs_reference": exact,
            "effective_scan_GBps": eff, "effective_frac_of_hbm_peak": eff / peak,
            "lookups_per_s": lookups, "lookup_frac_of_smem_bound": lookups / (148 * 32 * 1.965e9),
            "cpu_baseline": {"kind": "reference" if oracle.ref_available() else "port", "cores": oracle.num_threads(),
                             "queries_per_s": args.cpu_queries / cpu_s,
                             "sample": f"{args.cpu_queries} queries x {args.n} codes"},
        }))


if __name__ == "__main__":
    main()
