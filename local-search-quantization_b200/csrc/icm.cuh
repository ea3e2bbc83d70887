// icm.cuh — parameter block shared by the ICM / ILS kernels and their host launchers.
#pragma once
#include "common.cuh"

namespace lsq {

constexpr int ICM_MAX_ITERS_PER_LAUNCH = 64;

// Everything one launch of `niters` ILS iterations needs.  The visit orders and snapshot map travel
// in the parameter block (constant bank), so a launch needs no schedule upload.
struct IcmParams {
  const float* X;   // [n][d]
  const float* C;   // [m][256][d]
  const float* U;   // [m][n][256] unaries (utils.jl:94-122)
  const float* T;   // [m][m][256][256] pair tables, T[j][k][b][a] = 2<C_j[:,a], C_k[:,b]>
  uint8_t* codes;   // [n][m] accepted codes, in/out
  float* cost;      // [n] cost of `codes`, in/out
  const uint8_t* slots;  // explicit perturbation slots [niters][n][npert] or nullptr (-> Philox)
  const uint8_t* vals;   // explicit perturbation values, same shape
  uint8_t* snap;    // [nsnap][n][m] or nullptr
  float* snapcost;  // [nsnap][n] cost of each snapshot, or nullptr
  int64_t n;
  uint64_t seed;
  uint64_t g0;      // global index of vector 0 (sharding invariance)
  uint32_t ils_iter0;
  int d, m, icmiter, npert, niters;
  int16_t snap_of_iter[ICM_MAX_ITERS_PER_LAUNCH];
  int8_t orders[ICM_MAX_ITERS_PER_LAUNCH][LSQ_MAXM];
};

int launch_icm_warp(const IcmParams& p, cudaStream_t st);
int launch_veccost(const float* dX, int d, int64_t n, const uint8_t* dcodes, const float* dC, int m,
                   float* dcost, cudaStream_t st);
int launch_reconstruct(const uint8_t* dcodes, int64_t n, const float* dC, int d, int m, float* dCB,
                       cudaStream_t st);
int launch_sum_f32_to_f64(const float* dv, int64_t n, double* dout, cudaStream_t st);
int launch_quantize_norms(const uint8_t* dcodes, int64_t n, const float* dC, int d, int m,
                          const float* dcbnorms, int hn, int16_t* dout1, cudaStream_t st);
int launch_codes_i16_to_u8(const int16_t* d16, uint8_t* d8, int64_t count, int* derr, cudaStream_t st);
int launch_codes_u8_to_i16(const uint8_t* d8, int16_t* d16, int64_t count, cudaStream_t st);

int build_norms(const float* dC, int d, int m, float* dnorms, cudaStream_t st);
int build_unaries(const float* dX, int d, int64_t n, const float* dC, int m, const float* dnorms, float* dU,
                  cudaStream_t st);
int build_tables(const float* dC, int d, int m, float* dT, cudaStream_t st);

}  // namespace lsq
