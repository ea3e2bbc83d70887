"""PQ / OPQ scan (linscan_aqd.cpp), 1 M codes x 10 K queries, top-1000: tensor-core filter vs the lookup kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import lsq_b200
from lsq_b200 import device as dev
from util import make_scan_problem

lsq_b200.init(0)
for m in (8, 16):
    n, nq, d, nn = 1_000_000, 10_000, 128, 1000
    codes, queries, codebooks, _ = make_scan_problem(60 + m, n, nq, d, m)
    sub = d // m
    centers = np.ascontiguousarray(codebooks[:, :sub].reshape(m, 256, sub))
    dc, dq, dcb = torch.from_numpy(codes).cuda(), torch.from_numpy(queries).cuda(), torch.from_numpy(centers).cuda()
    out = {}
    for mode in ("tc", "scan"):
        os.environ["LSQ_B200_ADC"] = mode
        for _ in range(2):
            r = dev.linscan(dc, dq, dcb, None, nn, lut_kind=1, subdim=sub)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = dev.linscan(dc, dq, dcb, None, nn, lut_kind=1, subdim=sub); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out[mode] = (float(np.median(ts)), r)
    del os.environ["LSQ_B200_ADC"]
    same = torch.equal(out["tc"][1][0], out["scan"][1][0]) and torch.equal(out["tc"][1][1], out["scan"][1][1])
    print(f"PQ m={m} subdim={sub}: filter {out['tc'][0]:.2f} ms, lookup scan {out['scan'][0]:.2f} ms, identical results: {same}")
