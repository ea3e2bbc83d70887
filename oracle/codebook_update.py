"""CPU oracle for update_codebooks (codebook_update.jl:52-86).  TEST INFRASTRUCTURE ONLY.

The reference solves  min_K || X - K * onehot(B)' ||_F  as d independent sparse least-squares problems
with IterativeSolvers.lsqr (codebook_update.jl:15-18).  IterativeSolvers.jl is a third-party package
that is NOT vendored under /root/reference and is listed without a version (README.md:27), so the
arithmetic at this boundary is "parity unpinned".  Its published algorithm is Paige & Saunders' LSQR
started from x0 = 0, which converges to the MINIMUM-NORM least-squares solution; for Float32 inputs
its default tolerances are atol = btol = sqrt(eps(Float32)).  Two stand-ins:

  * update_codebooks_lsqr : scipy.sparse.linalg.lsqr with those tolerances (same published algorithm);
  * update_codebooks_exact: the exact minimum-norm solution in float64 (pseudo-inverse), the truth the
    CUDA path is compared with.  Acceptance: qerror(X, B, C_cuda) <= qerror(X, B, C_oracle)*(1+1e-5).
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import lsqr


def sparsify_codes(B0, h):
    """utils.jl:50-69: n-by-(m*h) one-hot matrix; column of code (i, b) is i*h + b (0-based)."""
    n, m = B0.shape
    rows = np.tile(np.arange(n), m)
    cols = (B0.astype(np.int64) + np.arange(m)[None, :] * h).T.reshape(-1)
    return sp.csr_matrix((np.ones(n * m, np.float32), (rows, cols)), shape=(n, m * h))


def K2vec(K, m, h):
    """utils.jl:72-87: d-by-(m*h) -> (m, h, d) (C[i] = K[:, i*h:(i+1)*h])"""
    d = K.shape[0]
    return np.ascontiguousarray(K.T.reshape(m, h, d)).astype(np.float32)


def update_codebooks_lsqr(X, B0, h):
    n, d = X.shape
    m = B0.shape[1]
    A = sparsify_codes(B0, h)
    tol = float(np.sqrt(np.finfo(np.float32).eps))
    K = np.zeros((d, m * h), np.float32)
    for i in range(d):
        K[i] = lsqr(A, X[:, i].astype(np.float64), atol=tol, btol=tol, conlim=1.0 / tol,
                    iter_lim=max(A.shape))[0]
    return K2vec(K, m, h)


def gram_stats(X, B0, h):
    """Normal-equation statistics: G = A'A (integer co-occurrence counts), R = A'X, both float64."""
    A = sparsify_codes(B0, h).astype(np.float64)
    G = (A.T @ A).toarray()
    R = np.asarray(A.T @ X.astype(np.float64))
    return G, R


def update_codebooks_exact(X, B0, h):
    n, d = X.shape
    m = B0.shape[1]
    G, R = gram_stats(X, B0, h)
    K = np.linalg.pinv(G, rcond=1e-12, hermitian=True) @ R  # (m*h, d), min-norm
    return np.ascontiguousarray(K.reshape(m, h, d)).astype(np.float32)
