import os, sys, ctypes as ct
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repository root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lsq_b200
from util import make_problem, sift_like
lsq_b200.init(0)
L = lsq_b200.lib()
n, D = int(os.environ.get("N", 1_000_000)), 128
P = lambda t: ct.c_void_p(t.data_ptr())
st = lambda: ct.c_void_p(torch.cuda.current_stream().cuda_stream)
for M in (8, 16):
    _, C_h, _ = make_problem(0, 16, D, M)
    X = torch.from_numpy(sift_like(np.random.default_rng(1), n, D)).cuda()
    C = torch.from_numpy(C_h).cuda()
    U = torch.empty((M, n, 256), dtype=torch.float32, device="cuda")
    def run(kind):
        if kind == "exact":
            return L.lsq_dev_build_unaries(P(X), D, ct.c_int64(n), P(C), M, P(U), 0, st())
        os.environ["LSQ_B200_UNARY_TC"] = kind
        return L.lsq_dev_build_unaries_tc(P(X), D, ct.c_int64(n), P(C), M, P(U), st())
    for kind in ("exact", "pipelined", "serial"):
        for _ in range(2):
            assert run(kind) == 0
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            run(kind)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        gb = (M * n * 1024 + n * D * 4) / 1e9
        print(f"m={M} n={n} {kind}: {ms:.3f} ms  {gb / ms:.0f} GB/s effective ({gb / ms / 6550.7 * 1000 * 100 / 1000:.1f}% of HBM peak)  {3 if kind != 'exact' else 1}x{2 * M * 256 * D * n / ms / 1e9:.0f} TFLOP/s", flush=True)
    del U
