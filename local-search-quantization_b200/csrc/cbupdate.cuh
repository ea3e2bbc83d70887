// cbupdate.cuh — least-squares codebook update (update_codebooks, codebook_update.jl:52-86).
#pragma once
#include "common.cuh"

namespace lsq {

// Gram[mh][mh] += co-occurrence counts, Rhs[mh][d] += per-code sums of X (both float64, device).
int cb_stats(const float* dX, int d, int64_t n, const uint8_t* dcodes, int m, double* dGram, double* dRhs,
             cudaStream_t st);
// conjugate gradients on Gram*K = Rhs from K0 = 0 (-> minimum-norm solution); Cout float [m][256][d].
int cb_solve(const double* dGram, const double* dRhs, int m, int d, float* dCout, int max_iter, double tol,
             int* iters_out, cudaStream_t st);

}  // namespace lsq
