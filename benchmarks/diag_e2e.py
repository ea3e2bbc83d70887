import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import lsq_b200
lsq_b200.init(lr)
from util import make_problem
n = 1_000_000
X, C, B = make_problem(1, n, 128, 8)
Xp = torch.from_numpy(X).pin_memory(); Bp = torch.from_numpy(B).pin_memory()
dX = torch.empty_like(Xp, device="cuda")
def bw():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): dX.copy_(Xp, non_blocking=True)
    torch.cuda.synchronize(); return 4 * Xp.numel() * 4 / (time.perf_counter() - t0) / 1e9
if world > 1: dist.barrier()
print(f"rank {rank} H2D pinned GB/s (all ranks at once): {bw():.1f}", flush=True)
if world > 1: dist.barrier()
for parts in (8, 1):
    os.environ["LSQ_B200_PIPELINE_PARTS"] = str(parts)
    ts = []
    for i in range(4):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out, _ = lsq_b200.encode_icm_cuda(Xp.numpy(), Bp.numpy(), C, [16], 4, 4, True, 1, seed=1, g0=rank * n)
        ts.append(time.perf_counter() - t0)
    print(f"rank {rank} parts={parts} e2e ms: " + " ".join(f"{1e3*t:.1f}" for t in ts), flush=True)
# solo: only rank 0 runs while the others wait
if world > 1: dist.barrier()
if rank == 0:
    os.environ["LSQ_B200_PIPELINE_PARTS"] = "8"
    ts = []
    for i in range(3):
        t0 = time.perf_counter()
        out, _ = lsq_b200.encode_icm_cuda(Xp.numpy(), Bp.numpy(), C, [16], 4, 4, True, 1, seed=1)
        ts.append(time.perf_counter() - t0)
    print("rank 0 SOLO e2e ms: " + " ".join(f"{1e3*t:.1f}" for t in ts), " H2D solo GB/s %.1f" % bw(), flush=True)
if world > 1: dist.barrier()
