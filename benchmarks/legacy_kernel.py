#!/usr/bin/env python
"""The reference's own GPU kernel on the same B200: src/encodings/cuda/cudautils.cu compiled UNMODIFIED
for sm_100a (oracle/Makefile -> oracle/_ref/cudautils_sm100a.cubin; it no longer builds for the
compute_35 of compile.sh:3) and launched exactly as encode_icm_cuda.jl:158-186 launches it
(condition_icm3, grid n, block (1,256), one launch per node visit, icmiter*m launches per ILS
iteration) — but with the pair tables already resident, i.e. WITHOUT the reference's per-visit H2D
upload and host synchronisation, and without perturb / veccost: a generous lower bound on its time.
Reported next to lsq-b200's kernel on the same inputs.  A timing baseline only (non-lowest-index tie
breaking, clock() seeded RNG: not an oracle)."""
import ctypes as ct
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(n=1_000_000, m=8, d=128, ils=16, icmiter=4):
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


if __name__ == "__main__":
    main()
