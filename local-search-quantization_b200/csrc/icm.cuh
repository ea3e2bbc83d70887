// icm.cuh — parameter block shared by the ICM / ILS kernels and their host launchers.
#pragma once
#include "common.cuh"

namespace lsq {

constexpr int ICM_MAX_ITERS_PER_LAUNCH = 64;
constexpr int ICM_SLICES = 8;         // 256 candidates = 8 slices of 32 (one 128-byte row each)
constexpr int ICM_SLICE_W = 32;
constexpr int ICM_SLICE_MAX_M = 8;    // (m-1) * 256 * 32 * 4 B <= 224 KB of shared memory

// Everything one launch of `niters` ILS iterations needs.  The visit orders and snapshot map travel
// in the parameter block (constant bank), so a launch needs no schedule upload.
struct IcmParams {
  const float* X;   // [n][d]
  const float* C;   // [m][256][d]
  const float* U;   // unaries (utils.jl:94-122): [m][n][256], or sliced [m][8][n][32] for the slice kernel
  const float* T;   // [m][m][256][256] pair tables, T[j][k][b][a] = 2<C_j[:,a], C_k[:,b]>
  const float* Ts;  // sliced pair tables [m][8][m-1][256][32] (slice kernel only)
  uint8_t* codes;   // [n][m] accepted codes, in/out
  float* cost;      // [n] cost of `codes`, in/out
  const uint8_t* slots;  // explicit perturbation slots [niters][n][npert] or nullptr (-> Philox)
  const uint8_t* vals;   // explicit perturbation values, same shape
  uint8_t* snap;    // [nsnap][n][m] or nullptr
  float* snapcost;  // [nsnap][n] cost of each snapshot, or nullptr
  // slice-kernel scratch, all [n]-sized (allocated by the launcher)
  unsigned long long* wcodes;   // working codes, packed one 64-bit word per vector
  unsigned long long* rbest;    // running (value, index) minimum across candidate slices
  uint16_t* clean;              // accepted-state clean mask
  uint16_t* wclean;             // working clean mask
  int* act;                     // per-CTA active lists (stored in the CTA's own vector range)
  // warp kernel: dynamic work distribution — a zero-initialised counter from which every warp draws its
  // next vector (vectors are independent, so the processing order cannot change any result); nullptr =
  // static grid-stride assignment
  unsigned long long* next_vector;
  // warp kernel: optional device counter that receives the number of node visits actually executed (the
  // clean-node skip makes it data dependent); nullptr = not counted
  unsigned long long* visits;
  int64_t n;
  uint64_t seed;
  uint64_t g0;      // global index of vector 0 (sharding invariance)
  uint32_t ils_iter0;
  int d, m, icmiter, npert, niters;
  int16_t snap_of_iter[ICM_MAX_ITERS_PER_LAUNCH];
  int8_t orders[ICM_MAX_ITERS_PER_LAUNCH][LSQ_MAXM];
};

// which kernel serves (m, n): 1 = slice kernel (shared-memory table slices, sliced unary layout),
// 0 = warp-per-vector kernel (tables gathered from L2, plain unary layout)
int icm_use_slices(int m, int64_t n);

int launch_icm_warp(const IcmParams& p, cudaStream_t st);
// executed-visit counter picked up by every later launch_icm_warp of the calling thread (nullptr = off)
void set_icm_visit_counter(unsigned long long* dcounter);
int launch_icm_slice(IcmParams p, cudaStream_t st);  // allocates its scratch from the pool
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
// sliced = 0: U[m][n][256]; sliced = 1: U[m][8][n][32]
int build_unaries(const float* dX, int d, int64_t n, const float* dC, int m, const float* dnorms, float* dU,
                  int sliced, cudaStream_t st);
// tensor-core (tcgen05, 3xTF32) build of U[m][n][256]: fast mode, tolerance-checked, d % 8 == 0, d <= 128
int build_unaries_tc(const float* dX, int d, int64_t n, const float* dC, int m, const float* dnorms, float* dU,
                     cudaStream_t st);
int build_tables(const float* dC, int d, int m, float* dT, cudaStream_t st);
// Ts[j][s][kk][b][32] from T, kk enumerating k != j in ascending order
int build_sliced_tables(const float* dT, int m, float* dTs, cudaStream_t st);

}  // namespace lsq
