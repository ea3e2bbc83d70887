"""Multi-GPU checks (-m gpu).

1. IN-LIBRARY sharding (lsq_init_devices): one process, one calling thread, the host-pointer calls split their
   input over the bound devices; every result must equal the single-device result BIT FOR BIT (codes, snapshots,
   objectives, codebooks, norm codebook).  With >= 2 GPUs it runs on devices [0, 1] with both all-reduce backends
   (NCCL and the peer-memory kernel); on a single-GPU box it still runs, on two "virtual" devices [0, 0] (testing
   hook LSQ_B200_ALLOW_DUPLICATE_DEVICES; NCCL refuses such a clique, so that leg uses the peer-memory kernel).
2. One process per GPU (torchrun + NCCL, parallel.py): needs >= 2 GPUs, skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import make_problem, make_scan_problem

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _calls(L, seed=31):
    """A bundle of host-pointer calls whose results must not depend on the bound device set."""
    out = {}
    n, d, m = 30011, 64, 8
    X, C, B = make_problem(seed, n, d, m, kind="gauss")
    Bs, objs = L.encode_icm_cuda(X, B, C, [1, 3], 4, 4, True, 2, seed=5)
    out["enc_B1"], out["enc_B3"], out["enc_obj"] = Bs[0], Bs[1], objs
    out["icm"] = L.encoding_icm(X, B, C, 3, True, 3, seed=6, ils_iter=2)
    slots, vals = L.make_perturb(7, 0, 0, n, m, 256, 2)
    out["sched"] = L.encoding_icm_sched(X, B, C, 2, np.arange(m)[::-1].copy(), slots, vals)
    out["upd"] = L.update_codebooks(X, B, 256)
    rng = np.random.default_rng(seed)
    R = np.linalg.qr(rng.standard_normal((d, d)))[0].astype(np.float32)
    Ct, Bt, cbn, Bn, obj = L.train_lsq(X, m, 256, R, B, None, 2, 2, 3, True, 3, seed=8)
    out["tr_C"], out["tr_B"], out["tr_cbn"], out["tr_Bn"], out["tr_obj"] = Ct, Bt, cbn, Bn, obj
    # m = 16 (two code words per vector) and an n that does not divide evenly
    X2, C2, B2 = make_problem(seed + 1, 9001, 32, 16, kind="sift")
    out["icm16"] = L.encoding_icm(X2, B2, C2, 2, True, 4, seed=9, ils_iter=0)
    sc, sq, scb, sn = make_scan_problem(seed + 2, 60000, 77, 64, 8)
    dd, ii = L.linscan_lsq(sc, sq, scb.reshape(8, 256, 64), sn, np.eye(64, dtype=np.float32), 150)
    out["scan_d"], out["scan_i"] = dd, ii
    return out


def _same(a, b):
    for k in a:
        assert np.array_equal(a[k], b[k]), f"{k} differs between the single-device and the multi-device run"


def test_inlibrary_sharding_matches_single_device(lsq, monkeypatch):
    assert lsq.device_count() > 0
    ngpu = lsq.device_count()
    lsq.finalize()
    lsq.init(0)
    assert lsq.num_bound_devices() == 1
    single = _calls(lsq)
    legs = [([0, 1], "nccl"), ([0, 1], "p2p")] if ngpu >= 2 else [([0, 0], "p2p")]
    if ngpu >= 4:
        legs.append(([0, 1, 2, 3], "nccl"))
    try:
        for devs, backend in legs:
            lsq.finalize()
            monkeypatch.setenv("LSQ_B200_ALLREDUCE", backend)
            if len(set(devs)) < len(devs):
                monkeypatch.setenv("LSQ_B200_ALLOW_DUPLICATE_DEVICES", "1")
            assert lsq.init_devices(devs) == len(devs)
            _same(single, _calls(lsq))
        # three shards of unequal size
        if ngpu == 1:
            lsq.finalize()
            assert lsq.init_devices([0, 0, 0]) == 3
            _same(single, _calls(lsq))
    finally:
        lsq.finalize()
        lsq.init(0)


def test_error_on_one_shard_fails_the_call_without_deadlock(lsq, monkeypatch):
    """An invalid code in the LAST shard: every sharded call must return the error (no worker may be left at a
    barrier or inside the collective), and the library must stay usable afterwards."""
    ngpu = lsq.device_count()
    devs = [0, 1] if ngpu >= 2 else [0, 0]
    lsq.finalize()
    if ngpu < 2:
        monkeypatch.setenv("LSQ_B200_ALLOW_DUPLICATE_DEVICES", "1")
    try:
        assert lsq.init_devices(devs) == 2
        n, d, m = 20000, 32, 4
        X, C, B = make_problem(91, n, d, m)
        bad = B.copy()
        bad[n - 7, 2] = 300
        for call in (lambda: lsq.encoding_icm(X, bad, C, 2, True, 2, seed=1, ils_iter=0),
                     lambda: lsq.encode_icm_cuda(X, bad, C, [2], 2, 2, True, 1, seed=1),
                     lambda: lsq.update_codebooks(X, bad, 256),
                     lambda: lsq.train_lsq(X, m, 256, None, bad, None, 1, 1, 2, True, 2, seed=1)):
            with pytest.raises(lsq.LsqError, match="1-based"):
                call()
        # still healthy: the same calls with valid codes
        good = lsq.encoding_icm(X, B, C, 2, True, 2, seed=1, ils_iter=0)
        Cn = lsq.update_codebooks(X, good, 256)
        assert lsq.qerror(X, good, Cn) <= lsq.qerror(X, good, C) * (1 + 1e-6)
    finally:
        lsq.finalize()
        lsq.init(0)


def test_device_set_from_environment(lsq, monkeypatch):
    """LSQ_B200_DEVICES selects the device set of a process that never calls lsq_init (an unchanged Julia program)."""
    lsq.finalize()
    monkeypatch.setenv("LSQ_B200_ALLOW_DUPLICATE_DEVICES", "1")
    monkeypatch.setenv("LSQ_B200_DEVICES", "0,0" if lsq.device_count() < 2 else "all")
    try:
        X, C, B = make_problem(92, 9000, 32, 4)
        out = lsq.encoding_icm(X, B, C, 2, True, 2, seed=1, ils_iter=0)      # binds lazily
        assert lsq.num_bound_devices() == max(2, lsq.device_count())
        monkeypatch.delenv("LSQ_B200_DEVICES")
        lsq.finalize()
        lsq.init(0)
        assert np.array_equal(out, lsq.encoding_icm(X, B, C, 2, True, 2, seed=1, ils_iter=0))
    finally:
        lsq.finalize()
        lsq.init(0)


def test_update_codebooks_is_deterministic_and_chunk_invariant(lsq):
    """Exact integer statistics: run-to-run identical bits, and accumulating a shard in pieces (what sharding over
    GPUs does) gives the same integers as one pass — on non-integer data, where float64 atomics would not."""
    import torch
    from lsq_b200 import device as dev
    lsq.init(0)
    n, d, m = 50000, 128, 8
    X, C, B = make_problem(55, n, d, m, kind="gauss")
    Xd = torch.from_numpy(X).cuda()
    cd = torch.from_numpy((B - 1).astype(np.uint8)).cuda()
    e = dev.cb_scale_exp(dev.absmax(Xd), n)
    assert dev.absmax(Xd) == float(np.abs(X).max())
    whole = dev.cb_accumulate(Xd, cd, m, e)
    again = dev.cb_accumulate(Xd, cd, m, e)
    assert torch.equal(whole, again)
    parts = None
    for lo, hi in lsq.splitarray(n, 3):
        parts = dev.cb_accumulate(Xd[lo:hi], cd[lo:hi], m, e, stats=parts)
    assert torch.equal(whole, parts)
    # the integers are what they claim to be: counts, and sums within one rounding per element of float64
    mh = m * 256
    G = whole[: mh * mh].view(mh, mh).cpu().numpy()
    assert G.trace() == n * m and np.array_equal(G, G.T)
    rhs = whole[mh * mh:].view(mh, d).cpu().numpy().astype(np.float64) * 2.0 ** (-e)
    truth = np.zeros((mh, d))
    np.add.at(truth, (B[:, 0] - 1).astype(np.int64), X.astype(np.float64))
    assert np.allclose(rhs[:256], truth[:256], rtol=0, atol=n * 2.0 ** (-e))
    C1 = lsq.update_codebooks(X, B, 256)
    C2 = lsq.update_codebooks(X, B, 256)
    assert np.array_equal(C1, C2)


SCRIPT = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, %(here)r)
import lsq_b200
from lsq_b200 import device as dev, parallel as par
from util import make_problem
rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lsq_b200.init(local)
n, d, m = 40000, 64, 8
X, C, B = make_problem(77, n, d, m, kind="gauss")
lo, hi = par.shard_bounds(n)
Xs = torch.from_numpy(X[lo:hi]).cuda(); cs = torch.from_numpy((B[lo:hi] - 1).astype(np.uint8)).cuda()
C1, codes, obj = par.train_lsq_sharded(Xs, cs, torch.from_numpy(C).cuda(), 2, 3, 4, True, 4, seed=9, g0=lo)
allc = par.gather_codes(codes).cpu().numpy()
# linscan by query partitioning: replicated codes, disjoint outputs, all-gather over NCCL
from util import make_scan_problem
sc, sq, scb, sn = make_scan_problem(78, 50000, 37, 64, 8)
dc, dcb, dn = torch.from_numpy(sc).cuda(), torch.from_numpy(scb).cuda(), torch.from_numpy(sn).cuda()
sd, si, _ = par.linscan_sharded(torch.from_numpy(sq).cuda(), lambda qs: dev.linscan(dc, qs.contiguous(), dcb, dn, 100))
if rank == 0:
    # single-process reference run of the same loop through the host API
    Bc, Cc = B.copy(), C
    for it in range(2):
        Cc = lsq_b200.update_codebooks(X, Bc, 256)
        for i in range(3):
            Bc = lsq_b200.encoding_icm(X, Bc, Cc, 4, True, 4, seed=9, ils_iter=3 * it + i)
    same = np.array_equal(allc.astype(np.int16) + 1, Bc)
    same_C = np.array_equal(C1.cpu().numpy(), Cc)
    q = lsq_b200.qerror(X, Bc, Cc)
    wd, wi = dev.linscan(dc, torch.from_numpy(sq).cuda(), dcb, dn, 100)
    scan_same = bool(torch.equal(wd, sd) and torch.equal(wi, si))
    print("RESULT", same, abs(obj[-1] - q) / q, scan_same, same_C)
dist.destroy_process_group()
'''


def test_two_process_nccl_training_matches_single(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    path = tmp_path / "multi_rank_script.py"
    path.write_text(SCRIPT % {"root": ROOT, "here": HERE})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(path)],
                       capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    assert line, r.stdout + r.stderr
    parts = line[0].split()
    # exact integer statistics summed by one NCCL all-reduce: codebooks AND codes identical to one GPU, on
    # Gaussian (non-integer) data
    assert parts[1] == "True"
    assert float(parts[2]) < 1e-6
    assert parts[3] == "True"   # sharded scan == single-GPU scan, bit for bit
    assert parts[4] == "True"   # codebooks bit-identical for 1 and 2 ranks
