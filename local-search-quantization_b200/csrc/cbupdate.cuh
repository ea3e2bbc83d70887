// cbupdate.cuh — least-squares codebook update (update_codebooks, codebook_update.jl:52-86).
#pragma once
#include "common.cuh"

namespace lsq {

// statistics buffer S: int64 [mh*mh + mh*d] (counts, then fixed-point sums scaled by 2^scale_exp); see cbupdate.cu
int cb_absmax(const float* dX, int64_t count, float* dmax, cudaStream_t st);   // *dmax = max(*dmax, max|x|)
int cb_scale_exp(float absmax, int64_t n_total);                               // host: e with n*max|x|*2^e < 2^62
int cb_accumulate(const float* dX, int d, int64_t n, const uint8_t* dcodes, int m, int scale_exp, int64_t* dS,
                  cudaStream_t st);                                            // S += shard
int cb_finalize(const int64_t* dS, int m, int d, int scale_exp, double* dGram, double* dRhs, cudaStream_t st);
// fused peer-memory form of "all-reduce + finalize": Gram / Rhs from the statistics buffers of k devices, read in place
// through NVLink peer memory and summed in rank order (exact integers: the order is irrelevant for the value)
struct PeerPtrs;
int cb_finalize_peers(const PeerPtrs& peers, int k, int m, int d, int scale_exp, double* dGram, double* dRhs, cudaStream_t st);
// single shard: absmax + accumulate + finalize; OVERWRITES Gram[mh][mh] and Rhs[mh][d] (float64, device)
int cb_stats(const float* dX, int d, int64_t n, const uint8_t* dcodes, int m, double* dGram, double* dRhs,
             cudaStream_t st);
// conjugate gradients on Gram*K = Rhs from K0 = 0 (-> minimum-norm solution); Cout float [m][256][d].
int cb_solve(const double* dGram, const double* dRhs, int m, int d, float* dCout, int max_iter, double tol,
             int* iters_out, cudaStream_t st);

}  // namespace lsq
