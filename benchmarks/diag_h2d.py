"""Host-side copy diagnostics on the multi-GPU box (scratch, not product)."""
import os, sys, time, threading, ctypes as ct
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repository root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import lsq_b200
from util import make_problem, sift_like
ng = torch.cuda.device_count()
print("gpus", ng, "cores", os.cpu_count(), flush=True)
# (1) aggregate host memcpy bandwidth into pinned memory
src = np.random.default_rng(0).random(64 * 1024 * 1024 // 8 * 8)  # 512 MB
for nthreads in (1, 4, 8, 16, 32):
    per = src.size // nthreads
    dsts = [torch.empty(per, dtype=torch.float64).pin_memory().numpy() for _ in range(nthreads)]
    def work(i):
        np.copyto(dsts[i], src[i * per:(i + 1) * per])
    for rep in range(2):
        th = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
        t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt = time.perf_counter() - t0
    print(f"memcpy pageable->pinned {nthreads} threads: {src.nbytes / dt / 1e9:.1f} GB/s", flush=True)
    del dsts
# (2) cudaHostRegister cost
cudart = ct.CDLL("libcudart.so.12")
buf = np.random.default_rng(1).random(64 * 1024 * 1024 // 8)  # 64 MB
for rep in range(3):
    t0 = time.perf_counter(); rc = cudart.cudaHostRegister(ct.c_void_p(buf.ctypes.data), ct.c_size_t(buf.nbytes), 0); t1 = time.perf_counter()
    cudart.cudaHostUnregister(ct.c_void_p(buf.ctypes.data)); t2 = time.perf_counter()
    print(f"cudaHostRegister 64 MB rc={rc}: {1e3*(t1-t0):.2f} ms, unregister {1e3*(t2-t1):.2f} ms", flush=True)
big = src
t0 = time.perf_counter(); rc = cudart.cudaHostRegister(ct.c_void_p(big.ctypes.data), ct.c_size_t(big.nbytes), 0); t1 = time.perf_counter()
cudart.cudaHostUnregister(ct.c_void_p(big.ctypes.data)); t2 = time.perf_counter()
print(f"cudaHostRegister 512 MB rc={rc}: {1e3*(t1-t0):.2f} ms, unregister {1e3*(t2-t1):.2f} ms", flush=True)
# (3) in-library encode on all devices: pinned vs pageable vs direct
n, M, D = 1_000_000, 8, 128
_, C, _ = make_problem(0, 16, D, M)
rng = np.random.default_rng(999)
X = sift_like(rng, n, D); B = rng.integers(1, 257, size=(n, M)).astype(np.int16)
Xp, Bp = torch.from_numpy(X).pin_memory(), torch.from_numpy(B).pin_memory()
its = np.array([16], np.int64)
for devs in ([0], list(range(ng))):
    lsq_b200.finalize(); lsq_b200.init_devices(devs)
    for label, xs, bs in (("pinned", Xp.numpy(), Bp.numpy()), ("pageable", X, B)):
        ts = []
        for i in range(6):
            t0 = time.perf_counter(); lsq_b200.encode_icm_cuda(xs, bs, C, its, 4, 4, True, 1, seed=1); ts.append(1e3 * (time.perf_counter() - t0))
        print(f"in-lib {len(devs)} devices {label}: {[round(t, 1) for t in ts]}", flush=True)
