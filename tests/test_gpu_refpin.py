"""Pins the ICM node visit, the unary '+ norms' step and veccost to the REFERENCE'S OWN CUDA kernels (-m gpu).

oracle/_ref/cudautils_sm100a.cubin is src/encodings/cuda/cudautils.cu of the reference compiled UNMODIFIED for
sm_100a (oracle/Makefile; built in the build container, it travels to the GPU box — nothing here reads
/root/reference).  Its kernels are launched exactly as encode_icm_cuda.jl launches them:
  vec_add        grid n, block (1, h)            encode_icm_cuda.jl:96
  condition_icm3 grid n, block (1, h)            encode_icm_cuda.jl:180-182  (cudautils.cu:236-339)
  veccost2       grid n, block (1, d), smem 4d   encode_icm_cuda.jl:134,199  (cudautils.cu:145-183)
On tie-free (Gaussian) inputs one condition_icm3 launch must return exactly the argmin the oracle and our
kernel compute for that node visit: the same ascending-k separate fp32 adds (cudautils.cu:259-268) on the same
table bits.  (The reference kernel's tie-breaking differs — a pairwise tree, not lowest index — which is why
only tie-free data pins it.)"""
import ctypes as ct
import os

import numpy as np
import pytest

from util import make_problem

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CUBIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "cudautils_sm100a.cubin")


class RefKernels:
    def __init__(self):
        import torch
        from cuda.bindings import driver as cu
        torch.cuda.init()
        torch.zeros(1, device="cuda")
        self.cu, self.torch = cu, torch
        err, self.mod = cu.cuModuleLoad(CUBIN.encode())
        assert err == cu.CUresult.CUDA_SUCCESS, err
        self.fn = {}
        for name in (b"vec_add", b"condition_icm3", b"veccost2"):
            err, f = cu.cuModuleGetFunction(self.mod, name)
            assert err == cu.CUresult.CUDA_SUCCESS, (name, err)
            self.fn[name.decode()] = f

    def launch(self, name, grid, block, smem, *args):
        cu = self.cu
        arr = (ct.c_void_p * len(args))(*[ct.cast(ct.pointer(a), ct.c_void_p) for a in args])
        stream = self.torch.cuda.current_stream().cuda_stream
        (e,) = cu.cuLaunchKernel(self.fn[name], grid, 1, 1, block[0], block[1], 1, smem, stream, ct.addressof(arr), 0)
        assert e == cu.CUresult.CUDA_SUCCESS, (name, e)


@pytest.fixture(scope="module")
def ref(lsq):
    if not os.path.exists(CUBIN):
        pytest.skip("oracle/_ref/cudautils_sm100a.cubin not built (needs /root/reference + nvcc at build time)")
    assert lsq.device_count() > 0
    lsq.init(0)
    return RefKernels()


P = lambda t: ct.c_void_p(t.data_ptr())


@pytest.mark.parametrize("n,d,m,sweeps", [(3000, 128, 8, 2), (1500, 64, 7, 1), (700, 32, 16, 2), (257, 96, 3, 3)])
def test_node_visit_pinned_to_reference_kernels(lsq, oracle, ref, n, d, m, sweeps):
    import torch
    from lsq_b200 import device as dev
    h = 256
    X, C, B = make_problem(4000 + n, n, d, m, kind="gauss")
    B0 = (B - 1).astype(np.int16)

    # ---- unaries: the reference adds the norms with its own vec_add kernel to the -2*C'X gemm ----
    gemm = torch.from_numpy(oracle.get_unary_gemm(X, C)).cuda()            # [m][n][h]
    norms = torch.from_numpy(oracle.get_norms(C)).cuda()                    # [m][h]
    for j in range(m):
        ref.launch("vec_add", n, (1, h), 0, P(gemm[j]), P(norms[j]), ct.c_int(n), ct.c_int(h))
    torch.cuda.synchronize()
    U_ref = gemm.cpu().numpy()
    U_ours = lsq.get_unaries(X, C)
    assert np.array_equal(U_ref, U_ours)                                     # ours == reference vec_add output
    assert np.array_equal(U_ref, oracle.get_unaries(X, C))

    # ---- pair tables: ours, bit-identical to the oracle's, concatenated per node like cat(2, bbs...) ----
    Xd, Cd = torch.from_numpy(X).cuda(), torch.from_numpy(C).cuda()
    codes0 = torch.from_numpy(B0.astype(np.uint8)).cuda()
    sess = dev.EncodeSession(Xd, Cd, codes0.clone(), sliced=0)
    T = sess.T.view(m, m, h, h)
    G, cbi = oracle.get_binaries(C)
    for idx, (i, j) in enumerate(cbi):
        assert np.array_equal(T[i, j].cpu().numpy(), G[idx])                 # binaries[(i,j)]
        assert np.array_equal(T[j, i].cpu().numpy(), G[idx].T)               # binaries_t[(i,j)]
    bbs = [torch.cat([T[k, l] for l in range(m) if l != k]).contiguous() for k in range(m)]

    # ---- `sweeps` block-ICM sweeps in natural order with the reference's condition_icm3 ----
    codes_soa = codes0.t().contiguous()                                      # d_codek[i_idx + n*i]
    for _ in range(sweeps):
        for k in range(m):
            ref.launch("condition_icm3", n, (1, h), 0, P(gemm[k]), P(bbs[k]), P(codes_soa), ct.c_int(k), ct.c_int(m),
                       ct.c_int(n))
    torch.cuda.synchronize()
    B_ref = codes_soa.t().contiguous().cpu().numpy().astype(np.int16)

    # oracle: encode_icm_fully! on the same tables, no perturbation, same visit order
    none = np.zeros((n, 0), np.uint8)
    B_orc = oracle.icm_fully(B0.copy(), U_ours, G, n, m, h, sweeps, np.arange(m, dtype=np.int32), none,
                             np.zeros((n, 0), np.int16))
    assert np.array_equal(B_ref, B_orc), f"{np.mean(np.any(B_ref != B_orc, axis=1)):.4f} of the vectors differ"

    # ours: one ILS iteration without perturbation = the same sweeps, followed by the accept rule
    # (encode_icm.jl:178-186), so a vector either carries the reference kernel's codes or kept its old ones
    sess.ils(1, sweeps, 0, False, orders=np.arange(m, dtype=np.int8)[None, :])
    B_ours = sess.codes.cpu().numpy().astype(np.int16)
    moved = np.any(B_ours != B0, axis=1)
    assert np.array_equal(B_ours[moved], B_ref[moved])
    kept = ~moved & np.any(B_ref != B0, axis=1)                              # ICM moved them, the accept rule did not
    assert kept.mean() < 1e-3

    # ---- veccost2 of the reference on its own result vs ours (serial sum vs lane partials: ~1 ulp apart) ----
    cost_ref = torch.empty(n, dtype=torch.float32, device="cuda")
    ref.launch("veccost2", n, (1, d), 4 * d, P(Xd), P(Cd), P(codes_soa), P(cost_ref), ct.c_int(d), ct.c_int(m), ct.c_int(n))
    torch.cuda.synchronize()
    cost_ours = lsq.veccost(X, (B_ref + 1).astype(np.int16), C)
    assert np.allclose(cost_ref.cpu().numpy(), cost_ours, rtol=2e-6, atol=0)
    rec = C[np.arange(m)[None, :], B_ref.astype(np.int64)].astype(np.float64).sum(1)   # float64 truth
    truth = ((rec - X.astype(np.float64)) ** 2).sum(1)
    assert np.allclose(cost_ours, truth, rtol=2e-6, atol=0) and np.allclose(cost_ref.cpu().numpy(), truth, rtol=2e-6, atol=0)
