"""Small run of the tensor-core ADC path (csrc/adc_tc.cu) for compute-sanitizer (memcheck / synccheck / racecheck):
filter values in both product modes, and whole scans through the filter (one and two A tiles, tail tiles, d = 32 / 96 /
128), each compared with the lookup scan."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import lsq_b200 as L
from lsq_b200 import device as dev
from util import make_scan_problem

L.init(0)
for passes in ("2", "1"):
    os.environ["LSQ_B200_ADC_PASSES"] = passes
    codes, q, cb, nr = make_scan_problem(2, 1000, 140, 96, 12)
    v = dev.adc_filter_values(torch.from_numpy(codes).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(cb).cuda(),
                              torch.from_numpy(nr).cuda())
    torch.cuda.synchronize()
    for (n, nq, d, m, nn) in [(20000, 30, 128, 8, 50), (18000, 200, 32, 16, 20), (17000, 5, 96, 3, 10)]:
        codes, q, cb, nr = make_scan_problem(3 + m, n, nq, d, m)
        os.environ["LSQ_B200_ADC"] = "tc"
        d1, i1 = L.linscan_lsq(codes, q, cb.reshape(m, 256, d), nr, np.eye(d, dtype=np.float32), nn)
        os.environ["LSQ_B200_ADC"] = "scan"
        d2, i2 = L.linscan_lsq(codes, q, cb.reshape(m, 256, d), nr, np.eye(d, dtype=np.float32), nn)
        del os.environ["LSQ_B200_ADC"]
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
# PQ / OPQ tables through the filter, and the sample-buffer thresholds (the alternative to the list-based ones)
codes, q, cb, nr = make_scan_problem(9, 20000, 40, 64, 16)
centers = np.ascontiguousarray(cb[:, :4].reshape(16, 256, 4))
os.environ["LSQ_B200_ADC"] = "tc"
d1, i1 = L.linscan_pq(codes, q, centers, 128, 30)
os.environ["LSQ_B200_ADC_SBUF"] = "1"
d3, i3 = L.linscan_pq(codes, q, centers, 128, 30)
del os.environ["LSQ_B200_ADC_SBUF"]
os.environ["LSQ_B200_ADC"] = "scan"
d2, i2 = L.linscan_pq(codes, q, centers, 128, 30)
del os.environ["LSQ_B200_ADC"]
assert np.array_equal(i1, i2) and np.array_equal(d1, d2) and np.array_equal(i3, i2) and np.array_equal(d3, d2)
print("done")
