"""Mints the golden vectors under tests/golden/.  Run in the BUILD container (needs /root/reference for
the linscan vectors, which come from the reference's own C++ compiled unmodified into oracle/_ref).

The reference ships no tests or fixtures, so:
  * icm_*.npz     — outputs of oracle/lsq_oracle.c, accepted only if oracle/np_twin.py (an independent
                    restatement of the same reference lines) agrees bit for bit;
  * viterbi_*.npz — chain-encoder codes (encode_chain.jl), same rule: oracle and twin must agree;
  * linscan_*.npz — outputs of the REAL reference linscan_aqd_query[_extra_byte].
Inputs are regenerated from the stored seeds by tests/util.py, so the files stay tiny.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from oracle import np_twin  # noqa: E402
from util import make_problem, make_scan_problem  # noqa: E402

ICM_CASES = [  # name, seed, n, d, m, niter, npert, randord, ils_iters
    ("icm_m4_d16", 11, 48, 16, 4, 3, 2, True, 2),
    ("icm_m8_d128", 12, 24, 128, 8, 4, 4, True, 2),
    ("icm_m7_d32_noshuffle", 13, 32, 32, 7, 2, 3, False, 1),
    ("icm_m16_d32", 14, 12, 32, 16, 2, 4, True, 1),
]
VITERBI_CASES = [  # name, seed, n, d, m, kind
    ("viterbi_m8_d128", 31, 40, 128, 8, "sift"),
    ("viterbi_m16_d32", 32, 30, 32, 16, "sift"),
    ("viterbi_m3_d16_gauss", 33, 64, 16, 3, "gauss"),
]
SCAN_CASES = [  # name, seed, n, nq, d, m, nn
    ("linscan_lsq_m8", 21, 3000, 6, 32, 8, 25),
    ("linscan_lsq_m7", 22, 2000, 5, 16, 7, 10),
    ("linscan_lsq_m16", 23, 1500, 4, 32, 16, 40),
    ("linscan_pq_m8", 24, 2500, 6, 32, 8, 30),
]


def main():
    oracle.build()
    for name, seed, n, d, m, niter, npert, randord, iters in ICM_CASES:
        X, C, B1 = make_problem(seed, n, d, m)
        B = (B1 - 1).astype(np.int16)
        Bt = B.copy()
        outs, costs = [], []
        for it in range(iters):
            B, cost = oracle.encoding_icm(X, B, C, niter, randord, npert, seed=seed, ils_iter=it)
            Bt, cost_t = np_twin.encoding_icm(X, Bt, C, niter, randord, npert, seed=seed, ils_iter=it)
            assert np.array_equal(B, Bt) and np.array_equal(cost, cost_t), f"{name}: oracle != twin"
            outs.append(B + 1)
            costs.append(cost)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, n=n, d=d, m=m, niter=niter, npert=npert,
                            randord=randord, iters=iters, codes=np.stack(outs).astype(np.int16),
                            cost=np.stack(costs))
        print("wrote", name)
    for name, seed, n, d, m, kind in VITERBI_CASES:
        X, C, _ = make_problem(seed, n, d, m, kind=kind)
        a, b = oracle.encoding_viterbi(X, C), np_twin.encoding_viterbi(X, C)
        assert np.array_equal(a, b), f"{name}: oracle != twin"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, n=n, d=d, m=m, kind=kind,
                            codes=(a + 1).astype(np.int16))
        print("wrote", name)
    assert oracle.ref_available(), "oracle/_ref missing: run `make -C oracle` with /root/reference present"
    for name, seed, n, nq, d, m, nn in SCAN_CASES:
        codes, queries, codebooks, norms = make_scan_problem(seed, n, nq, d, m)
        if "pq" in name:
            centers = codebooks[:, : d // m].reshape(m, 256, d // m).copy()
            dists, ids = oracle.ref_linscan_pq(codes, queries, centers, nn)
        else:
            dists, ids = oracle.ref_linscan_lsq(codes, queries, codebooks, norms, nn)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, n=n, nq=nq, d=d, m=m, nn=nn, dists=dists,
                            ids=ids.astype(np.int64))
        print("wrote", name)


if __name__ == "__main__":
    main()
