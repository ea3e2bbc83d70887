import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lsq_b200
from lsq_b200 import device as dev, parallel as par
from util import make_problem, sift_like
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lsq_b200.init(local)
M, D, n = 8, 128, 300000
rng = np.random.default_rng(rank)
X = torch.from_numpy(sift_like(rng, n, D)).cuda()
codes = torch.from_numpy(rng.integers(0, 256, size=(n, M)).astype(np.uint8)).cuda()
e = par.global_scale_exp(X)
stats = dev.cb_accumulate(X, codes, M, e)
def ev(): return torch.cuda.Event(enable_timing=True)
for trial in range(6):
    do_ar = world > 1 and trial >= 3
    torch.cuda.synchronize()
    a, b, c = ev(), ev(), ev()
    w0 = time.perf_counter()
    if do_ar:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    a.record()
    g, r = dev.cb_finalize(stats, M, D, e)
    b.record()
    w1 = time.perf_counter()
    C, it = dev.cb_solve(g, r, M)
    c.record()
    w2 = time.perf_counter()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"trial {trial} allreduce={do_ar} finalize {a.elapsed_time(b):.2f} ms solve {b.elapsed_time(c):.2f} ms (host {1e3*(w2-w1):.2f} ms) iters {it}", flush=True)
if world > 1:
    dist.destroy_process_group()
