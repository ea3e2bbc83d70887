"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lsq_b200 as L
from util import make_problem, make_scan_problem

L.init(0)
X, C, B = make_problem(1, 600, 128, 8)
for k in ("warp", "slice"):
    os.environ["LSQ_B200_ICM_KERNEL"] = k
    Bs, o = L.encode_icm_cuda(X, B, C, [2], 2, 4, True, 1, seed=1)
os.environ["LSQ_B200_UNARY"] = "tc"
os.environ["LSQ_B200_ICM_KERNEL"] = "warp"
Bs, o = L.encode_icm_cuda(X, B, C, [1], 2, 4, True, 1, seed=1)
del os.environ["LSQ_B200_UNARY"]
C2 = L.update_codebooks(X, Bs[0], 256)
# linscan: sampled path (m = 8: 28-query tiles, select top-k), 16-query tiles (m = 12), 2 queries/lane (m = 16,
# exhaustive path), PQ LUT
codes, q, cb, nr = make_scan_problem(2, 20000, 30, 64, 8)
d, i = L.linscan_lsq(codes, q, cb.reshape(8, 256, 64), nr, np.eye(64, dtype=np.float32), 50)
codes, q, cb, nr = make_scan_problem(3, 20000, 17, 64, 12)
d, i = L.linscan_lsq(codes, q, cb.reshape(12, 256, 64), nr, np.eye(64, dtype=np.float32), 50)
codes, q, cb, nr = make_scan_problem(2, 3000, 5, 64, 16)
d, i = L.linscan_lsq(codes, q, cb.reshape(16, 256, 64), nr, np.eye(64, dtype=np.float32), 20)
codes, q, cb, nr = make_scan_problem(4, 20000, 9, 64, 8)
d, i = L.linscan_pq(codes, q, cb[:, :8].reshape(8, 256, 8).copy(), 64, 30)
# chain encoder, both kernels (ragged tails), train_lsq with a rotation, norm codebook, eval_recall
Xs, Cs, Bsm = make_problem(5, 333, 32, 5)
for k in ("simple", "tma"):
    os.environ["LSQ_B200_VITERBI"] = k
    Bv = L.encoding_viterbi(Xs, Cs)
R = np.linalg.qr(np.random.default_rng(0).standard_normal((32, 32)))[0].astype(np.float32)
out = L.train_lsq(Xs, 5, 256, R, Bsm, None, 2, 2, 2, True, 2, seed=3)
rec = L.eval_recall(np.arange(1, 6), np.tile(np.arange(1, 41), (5, 1)), 40)
print("done")
