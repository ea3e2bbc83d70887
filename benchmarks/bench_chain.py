#!/usr/bin/env python
"""Chain (Viterbi / ChainQ) encoder benchmark: n vectors, m codebooks; resident-data timing of the
kernel, the whole host call, and the CPU oracle on a subsample.  (m-1)*65536 candidate transitions per
vector, one FADD + one FMNMX each: bound by the ALU pipe (one FMNMX per 2 cycles per SM sub-partition),
reported as a fraction of 148 SMs x 4 sub-partitions x f_SM / 2."""
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
    ap.add_argument("--m", type=int, nargs="+", default=[8])
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-n", type=int, default=4000)
    args = ap.parse_args()
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


if __name__ == "__main__":
    main()
