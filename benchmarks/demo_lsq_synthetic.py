#!/usr/bin/env python
"""The flow of demos/demo_lsq_gpu.jl on synthetic data, end to end through the reference-named API:
train (update_codebooks <-> encoding_icm, LSQ.jl:57-66), encode the base set with ILS
(encode_icm_cuda), quantise norms (quantize_norms), search (linscan_lsq), recall@N (eval_recall).
m = 7 codebooks + 1 norm byte = 64-bit codes, as in the demo (demo_lsq.jl:14)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def clustered(rng, n, d, W):
    """SIFT-like non-negative data of low intrinsic dimension (rank-24 latent + small noise), so that a
    64-bit code has something to learn and recall@1 lands in the range real descriptors give."""
    z = rng.standard_normal((n, W.shape[0])).astype(np.float32)
    x = z @ W + rng.standard_normal((n, d)).astype(np.float32) * 2.0
    return np.clip(np.floor(np.abs(x)), 0, 255).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ntrain", type=int, default=100_000)
    ap.add_argument("--nbase", type=int, default=1_000_000)
    ap.add_argument("--nquery", type=int, default=1000)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--m", type=int, default=7)
    ap.add_argument("--niter", type=int, default=10)
    ap.add_argument("--ilsiter", type=int, default=8)
    ap.add_argument("--ils-base", type=int, default=16)
    ap.add_argument("--knn", type=int, default=100)
    args = ap.parse_args()
    import lsq_b200 as L
    L.init(0)
    rng = np.random.default_rng(0)
    d, m, h = args.d, args.m, 256
    centers = (rng.standard_normal((24, d)) * 12).astype(np.float32)  # latent -> descriptor map
    x_train = clustered(rng, args.ntrain, d, centers)
    x_base = clustered(rng, args.nbase, d, centers)
    x_query = clustered(rng, args.nquery, d, centers)
    t = {}

    # --- train_lsq (LSQ.jl:10-88) with random initial codes: (a) the caller-side loop over the public
    #     calls, which re-sends X on every call like the reference does; (b) ONE lsq_train_lsq call with
    #     everything resident on the GPU.  Same schedule, same seed: identical codes and codebooks. ---
    B0 = L.randinit(args.ntrain, m, h, rng)
    t0 = time.perf_counter()
    B = B0
    C = L.update_codebooks(x_train, B, h)
    obj = []
    it = 0
    for i in range(args.ilsiter):
        B = L.encoding_icm(x_train, B, C, 4, True, 4, seed=1, ils_iter=it); it += 1
    for outer in range(args.niter):
        obj.append(L.qerror(x_train, B, C))
        C = L.update_codebooks(x_train, B, h)
        for i in range(args.ilsiter):
            B = L.encoding_icm(x_train, B, C, 4, True, 4, seed=1, ils_iter=it); it += 1
    obj.append(L.qerror(x_train, B, C))
    t["train_loop_of_public_calls_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    Cf, Bf, cbnorms, B_norms, objf = L.train_lsq(x_train, m, h, None, B0, None, args.niter, args.ilsiter, 4, True, 4, seed=1)
    t["train_s"] = time.perf_counter() - t0
    fused_equal = bool(np.array_equal(Bf, B) and np.array_equal(Cf, C))
    # the norm codebook comes from train_lsq (1-D k-means, LSQ.jl:79-84)
    # --- encode the base set (demo_lsq_gpu.jl:43-51) ---
    t0 = time.perf_counter()
    B_base = L.randinit(args.nbase, m, h, rng)
    Bs, objs = L.encode_icm_cuda(x_base, B_base, C, [args.ils_base], 4, 4, True, 1, seed=2)
    B_base = Bs[-1]
    t["encode_base_s"] = time.perf_counter() - t0
    # --- norms (demo_lsq.jl:55-57) ---
    nb = L.quantize_norms(B_base, C, cbnorms)
    db_norms = cbnorms[nb - 1]
    # --- search + recall (demo_lsq.jl:59-77) ---
    t0 = time.perf_counter()
    dists, idx = L.linscan_lsq((B_base - 1).astype(np.uint8), x_query, C, db_norms, np.eye(d, dtype=np.float32), args.knn)
    t["search_s"] = time.perf_counter() - t0
    # brute-force ground truth (1-based ids like the demo's gt + 1)
    gt = np.empty(args.nquery, np.int64)
    bn = (x_base.astype(np.float64) ** 2).sum(1)
    for q0 in range(0, args.nquery, 100):
        q = x_query[q0:q0 + 100].astype(np.float64)
        gt[q0:q0 + 100] = np.argmin(bn[None, :] - 2.0 * q @ x_base.T.astype(np.float64), axis=1) + 1
    rec = L.eval_recall(gt, idx, args.knn)
    out = {"m": m, "ntrain": args.ntrain, "nbase": args.nbase, "nquery": args.nquery,
           "train_qerror": [float(o) for o in obj], "fused_train_equals_loop": fused_equal, "base_qerror": float(objs[0]),
           "recall@1": float(rec[0]), "recall@10": float(rec[9]), f"recall@{args.knn}": float(rec[-1]), "seconds": t}
    print(json.dumps(out))
    assert all(b <= a * (1 + 1e-6) for a, b in zip(obj, obj[1:])), "training objective must not increase"


if __name__ == "__main__":
    main()
