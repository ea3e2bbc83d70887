import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, lsq_b200 as L
from util import make_problem, make_scan_problem
L.init(0)
X,C,B=make_problem(1,600,128,8)
for k in ("warp","slice"):
    os.environ["LSQ_B200_ICM_KERNEL"]=k
    Bs,o=L.encode_icm_cuda(X,B,C,[2],2,4,True,1,seed=1)
os.environ["LSQ_B200_UNARY"]="tc"; os.environ["LSQ_B200_ICM_KERNEL"]="warp"
Bs,o=L.encode_icm_cuda(X,B,C,[1],2,4,True,1,seed=1)
C2=L.update_codebooks(X,Bs[0],256)
codes,q,cb,nr=make_scan_problem(2,20000,8,64,8)
d,i=L.linscan_lsq(codes,q,cb.reshape(8,256,64),nr,np.eye(64,dtype=np.float32),50)
codes,q,cb,nr=make_scan_problem(2,3000,5,64,16)
d,i=L.linscan_lsq(codes,q,cb.reshape(16,256,64),nr,np.eye(64,dtype=np.float32),20)
print("done")
