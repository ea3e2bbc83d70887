import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, lsq_b200 as L
from lsq_b200 import device as dev
from util import make_problem
L.init(0)
for n,m in [(100000,8),(1000000,8),(100000,16)]:
    X,C,B=make_problem(3,n,128,m)
    Xd=torch.from_numpy(X).cuda(); cd=torch.from_numpy((B-1).astype(np.uint8)).cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.perf_counter()
        g,r=dev.cb_stats(Xd,cd,m); torch.cuda.synchronize(); t1=time.perf_counter()
        Cn,iters=dev.cb_solve(g,r,m); torch.cuda.synchronize(); t2=time.perf_counter()
    print(f"n={n} m={m}: stats {1e3*(t1-t0):.2f} ms, solve {1e3*(t2-t1):.2f} ms, CG iters {iters}")
