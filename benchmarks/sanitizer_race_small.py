import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, lsq_b200 as L
from util import make_problem, make_scan_problem
L.init(0)
Xs, Cs, Bsm = make_problem(5, 150, 16, 3)
os.environ["LSQ_B200_VITERBI"] = "tma"
Bv = L.encoding_viterbi(Xs, Cs)
codes, q, cb, nr = make_scan_problem(2, 20000, 30, 32, 8)
d, i = L.linscan_lsq(codes, q, cb.reshape(8, 256, 32), nr, np.eye(32, dtype=np.float32), 50)
print("done")
