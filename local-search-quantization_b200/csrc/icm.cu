// icm.cu — ICM / ILS encoding kernels (encode_icm_fully! encode_icm.jl:4-127, encoding_icm :131-189,
// the ILS loop of encode_icm_cuda_single encode_icm_cuda.jl:124-222, replacing perturb / veccost2 /
// condition_icm3 of cuda/cudautils.cu).
//
// Kernel `icm_ils_warp_kernel` — one warp per database vector, ALL requested ILS iterations in one
// launch (vectors never interact, Appendix A.1 of SURVEY.md), codes and costs live in registers:
//   perturb (Philox or explicit)  ->  icmiter sweeps over the m nodes in the iteration's visit order
//   ->  veccost of the new codes  ->  keep iff strictly better (encode_icm.jl:183-186).
// A node visit streams the vector's 1 KB unary row from HBM (2 x LDG.128 per lane, fully coalesced),
// adds the m-1 conditioning columns T[j][k][code_k][:] (1 KB each, L2-resident tables) in ascending k
// with separate fp32 adds (the reference's order, encode_icm.jl:84-101), and takes the first strict
// minimum of the 256 candidates: 8 candidates per lane in ascending order, then a shuffle-xor
// lexicographic (value, index) reduction.  Lane l owns candidates 4l..4l+3 and 128+4l..128+4l+3.
#include "icm.cuh"

#include <stdlib.h>

#include <algorithm>

// Occupancy of the warp kernel (measured on B200, kernel ms for 1 M vectors x 16 ILS iterations, m = 8 / m = 16):
//   32 warps per SM, 64 registers (round 1): 103.4 / 465      36 warps, 56 registers: 97.9 / 460
//   40 warps, 48 registers: 99.0 / 430                         48 warps, 40 registers: 104.4 / 448     64 warps, 32: 134 / 530
// The dependent chain of a visit is hidden by resident warps, not by loads in flight per warp, until the register
// budget starts to spill inside the visit loop.  Blocks of 4 warps; 9 per SM for m <= 8, 10 beyond.
#ifndef LSQ_ICM_LB_THREADS
#define LSQ_ICM_LB_THREADS 128
#endif
#ifndef LSQ_ICM_LB_BLOCKS
#define LSQ_ICM_LB_BLOCKS(M) ((M) <= 8 ? 9 : 10)
#endif
#ifndef LSQ_ICM_SHORTCUT
#define LSQ_ICM_SHORTCUT(M) ((M) <= 8)
#endif

namespace lsq {

template <int M>
__device__ __forceinline__ uint32_t get_code(uint64_t lo, uint64_t hi, int k) {
  return (M <= 8 || k < 8) ? (uint32_t)(lo >> (8 * (k & 7))) & 0xFFu : (uint32_t)(hi >> (8 * (k & 7))) & 0xFFu;
}

template <int M>
__device__ __forceinline__ uint32_t get_code_dyn(uint64_t lo, uint64_t hi, int j) {
  const int sh = 8 * (j & 7);
  return (uint32_t)(((M <= 8 || j < 8) ? lo : hi) >> sh) & 0xFFu;
}

template <int M>
__device__ __forceinline__ void set_code(uint64_t& lo, uint64_t& hi, int j, uint32_t val) {
  const int sh = 8 * (j & 7);
  if (M <= 8 || j < 8) lo = (lo & ~(0xFFull << sh)) | ((uint64_t)val << sh);
  else hi = (hi & ~(0xFFull << sh)) | ((uint64_t)val << sh);
}

// veccost of one vector by one warp (utils.jl:225-254, canonical reduction order (2) of the oracle)
template <int M>
__device__ __forceinline__ float warp_veccost(const float* __restrict__ x, const float* __restrict__ C, int d,
                                              uint64_t lo, uint64_t hi, int lane) {
  float p = 0.0f;
  for (int t = lane; t < d; t += 32) {
    float r = 0.0f;
#pragma unroll
    for (int k = 0; k < M; k++)
      r = __fadd_rn(r, __ldg(C + ((size_t)k * LSQ_H + get_code<M>(lo, hi, k)) * d + t));
    const float df = __fsub_rn(r, __ldg(x + t));
    p = __fadd_rn(p, __fmul_rn(df, df));
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(0xFFFFFFFFu, p, off));
  return p;
}

// Cache policy note (measured, 1 M x 16 iterations): default caching for both the unary rows and the
// pair-table columns = 136 ms; table columns with L1::no_allocate = 159 ms; unary rows with
// L1::no_allocate = 147 ms.  Later (104.8 ms baseline): unary rows with L1::evict_last = no change, table
// columns additionally L1::evict_first = +23 %; a RUN-TIME switch between such load flavours costs 20 ms
// by itself (the volatile asm loads stop ptxas from batching the 16 loads of a visit).  Plain __ldg everywhere.
__device__ __forceinline__ void add4(float4& a, const float4 g) {
  a.x = __fadd_rn(a.x, g.x); a.y = __fadd_rn(a.y, g.y); a.z = __fadd_rn(a.z, g.z); a.w = __fadd_rn(a.w, g.w);
}

__device__ __forceinline__ float4 lds128_u(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// USMEM: the warp's m unary rows (m KB) are staged ONCE per vector into shared memory by TMA bulk copies
// (cp.async.bulk + one mbarrier per warp) and re-read from there on every node visit of every ILS
// iteration, which takes the unary share (1/m-th... 1 KB of every 8 KB visit at m = 8) off the saturated
// SM<->L2 path.  Used for m <= 8 (m KB per warp must leave room for >= 24 warps per SM).
// rows of a vector's unary table kept in shared memory by the staged variant: all of them up to m = 6, else 6 —
// 8 warps x 6 KB per block still lets 4 blocks (32 warps) share an SM, which full staging (m KB per warp) did not
template <int M>
__host__ __device__ constexpr int icm_staged_rows() { return M < 6 ? M : 6; }

// COUNT: the measurement variant that adds the executed node visits to *p.visits.  It is a separate
// instantiation on purpose: the kernel sits exactly at its 64-register budget (4 blocks per SM), and one more
// live counter in the visit loop cost 2 % at m = 8 and 10 % at m = 16 (measured against the same build without it).
template <int M, bool USMEM, bool COUNT = false>
__global__ void __launch_bounds__(LSQ_ICM_LB_THREADS, LSQ_ICM_LB_BLOCKS(M)) icm_ils_warp_kernel(const __grid_constant__ IcmParams p) {
  constexpr int UR = icm_staged_rows<M>();
  extern __shared__ __align__(128) unsigned char icm_smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;                                  // warp in block
  uint64_t* ubar = reinterpret_cast<uint64_t*>(icm_smem) + wib;      // 8 barriers, then 8 x M KB of rows
  const uint32_t urow = smem_u32(icm_smem + 128 + (size_t)wib * UR * 1024) + (uint32_t)lane * 16u;
  uint32_t uphase = 0;
  if (USMEM) {
    if (lane == 0) { mbar_init(ubar, 1); fence_mbar_init(); }
    __syncwarp();
  }
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  // The time a vector takes depends on how many node visits it needs; with a static grid-stride split the
  // early finishers idle (ncu: 6 % of the resident warp slots empty on average).  Each warp therefore draws
  // its next vector from a global counter: the first `nwarps` vectors are assigned statically, the rest
  // on demand.
  int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (; v < p.n;) {
    uint64_t lo = 0, hi = 0;
    {
      const uint8_t* cp = p.codes + v * M;
#pragma unroll
      for (int k = 0; k < M; k++) set_code<M>(lo, hi, k, cp[k]);
    }
    float prev = p.cost[v];
    const float* xv = p.X + (size_t)v * p.d;
    if (USMEM) {
      __syncwarp();  // every lane is done with the previous vector's rows
      if (lane == 0) {
        mbar_expect_tx(ubar, (uint32_t)UR * 1024u);
#pragma unroll
        for (int j = 0; j < UR; j++)
          bulk_g2s(icm_smem + 128 + ((size_t)wib * UR + j) * 1024, p.U + ((size_t)j * p.n + v) * LSQ_H, 1024u, ubar);
      }
      mbar_wait(ubar, uphase);
      uphase ^= 1u;
    }

    // `clean` bit j: code_j is already the argmin of node j given the other current codes, so a visit
    // would recompute exactly the same code (the node update is a deterministic function of the other
    // codes).  Such visits are skipped — results are bit-identical, the work drops ~2x because ICM
    // reaches a fixed point after about two sweeps.  Any code change clears every other node's bit.
    constexpr uint32_t ALL_CLEAN = (M == 32) ? 0xFFFFFFFFu : ((1u << M) - 1u);
    uint32_t clean = 0;
    uint32_t nvis = 0;  // node visits executed for this vector (COUNT variant only)
    for (int it = 0; it < p.niters; it++) {
      uint64_t wlo = lo, whi = hi;
      uint32_t wclean = clean;
      // ---- perturbation (encode_icm.jl:56-70) ----
      if (p.slots != nullptr) {
        const size_t base = ((size_t)it * p.n + v) * p.npert;
        for (int i = 0; i < p.npert; i++) {
          const int s = p.slots[base + i];
          const uint32_t x = p.vals[base + i];
          if (get_code_dyn<M>(wlo, whi, s) != x) { set_code<M>(wlo, whi, s, x); wclean = 0; }
        }
      } else if (p.npert > 0) {
        uint8_t s[LSQ_MAXM], x[LSQ_MAXM];
        make_perturb_one(p.seed, p.ils_iter0 + it, p.g0 + (uint64_t)v, M, LSQ_H, p.npert, s, x);
        for (int i = 0; i < p.npert; i++)
          if (get_code_dyn<M>(wlo, whi, s[i]) != x[i]) { set_code<M>(wlo, whi, s[i], x[i]); wclean = 0; }
      }
      // ---- block-ICM sweeps (encode_icm.jl:72-125) ----
      for (int sweep = 0; sweep < p.icmiter && wclean != ALL_CLEAN; sweep++) {
        for (int jj = 0; jj < M; jj++) {
          const int j = p.orders[it][jj];
          if ((wclean >> j) & 1u) continue;
          // Every OTHER node already carries its accepted code and the accepted state is known to be minimal at
          // node j: the visit would compute exactly what it computed then — the accepted code of j (the update is
          // a deterministic function of the other codes).  Typical case: the last perturbed node of a relaxing
          // perturbation.  No memory traffic; afterwards the working codes equal the accepted ones.
          if (LSQ_ICM_SHORTCUT(M) && ((clean >> j) & 1u)) {
            const int sh = 8 * (j & 7);
            const uint64_t mlo = (M <= 8 || j < 8) ? ~(0xFFull << sh) : ~0ull;
            const uint64_t mhi = (M <= 8 || j < 8) ? ~0ull : ~(0xFFull << sh);
            if (((wlo ^ lo) & mlo) == 0 && ((whi ^ hi) & mhi) == 0) {
              wlo = lo; whi = hi;
              wclean = clean;
              continue;
            }
          }
          if (COUNT) nvis++;
          float4 a0, a1;
          if (USMEM && j < UR) {
            a0 = lds128_u(urow + (uint32_t)j * 1024u);
            a1 = lds128_u(urow + (uint32_t)j * 1024u + 512u);
          } else {
            const float4* up = reinterpret_cast<const float4*>(p.U + ((size_t)j * p.n + v) * LSQ_H);
            a0 = __ldg(up + lane);
            a1 = __ldg(up + 32 + lane);
          }
          const float* tj = p.T + (size_t)j * M * LSQ_H * LSQ_H;
#pragma unroll
          for (int k = 0; k < M; k++) {
            if (k == j) continue;
            const float4* tp =
                reinterpret_cast<const float4*>(tj + ((size_t)k * LSQ_H + get_code<M>(wlo, whi, k)) * LSQ_H);
            const float4 g0 = __ldg(tp + lane);
            const float4 g1 = __ldg(tp + 32 + lane);
            add4(a0, g0);
            add4(a1, g1);
          }
          // first strict minimum (encode_icm.jl:105-119)
          float best = a0.x;
          int bi = 4 * lane;
          if (a0.y < best) { best = a0.y; bi = 4 * lane + 1; }
          if (a0.z < best) { best = a0.z; bi = 4 * lane + 2; }
          if (a0.w < best) { best = a0.w; bi = 4 * lane + 3; }
          if (a1.x < best) { best = a1.x; bi = 128 + 4 * lane; }
          if (a1.y < best) { best = a1.y; bi = 128 + 4 * lane + 1; }
          if (a1.z < best) { best = a1.z; bi = 128 + 4 * lane + 2; }
          if (a1.w < best) { best = a1.w; bi = 128 + 4 * lane + 3; }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const float ov = __shfl_xor_sync(0xFFFFFFFFu, best, off);
            const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, off);
            if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
          }
          if ((uint32_t)bi != get_code_dyn<M>(wlo, whi, j)) {
            set_code<M>(wlo, whi, j, (uint32_t)bi);
            // back at the accepted codes (the usual end of a rejected perturbation): everything already
            // known about that state applies again — typically "every node clean", which ends the sweeps
            wclean = (wlo == lo && whi == hi) ? clean : 0u;
          }
          wclean |= 1u << j;
        }
      }
      // ---- accept iff strictly better (encode_icm.jl:178-186) ----
      if (wlo == lo && whi == hi) {
        clean |= wclean;  // same codes: cannot be strictly better; keep what the sweeps learned
      } else {
        const float newc = warp_veccost<M>(xv, p.C, p.d, wlo, whi, lane);
        if (newc < prev) { prev = newc; lo = wlo; hi = whi; clean = wclean; }
      }
      const int sn = p.snap_of_iter[it];
      if (sn >= 0) {
        if (lane < M) p.snap[((size_t)sn * p.n + v) * M + lane] = (uint8_t)get_code<M>(lo, hi, lane);
        if (lane == 0 && p.snapcost != nullptr) p.snapcost[(size_t)sn * p.n + v] = prev;
      }
    }
    if (lane < M) p.codes[v * M + lane] = (uint8_t)get_code<M>(lo, hi, lane);
    if (lane == 0) p.cost[v] = prev;
    if (COUNT && lane == 0) atomicAdd(p.visits, (unsigned long long)nvis);
    if (p.next_vector != nullptr) {
      unsigned long long t = 0;
      if (lane == 0) t = atomicAdd(p.next_vector, 1ull);
      v = nwarps + (int64_t)__shfl_sync(0xFFFFFFFFu, t, 0);
    } else {
      v += nwarps;
    }
  }
}

static thread_local unsigned long long* g_visit_counter = nullptr;
void set_icm_visit_counter(unsigned long long* dcounter) { g_visit_counter = dcounter; }

template <int M>
static int launch_icm_warp_m(const IcmParams& p, cudaStream_t st) {
  int dev = 0, sms = LSQ_NUM_SMS_HINT;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  constexpr bool kCanUsmem = (M <= 8);
  // Measured (1 M x 16 iterations, m = 8): staging the unary rows in shared memory removes 1/8 of the L2
  // traffic but caps residency at 24 warps/SM: 140.3 ms vs 136.6 ms with 32 warps and everything from
  // L2.  Occupancy wins, so the staged variant is opt-in (LSQ_B200_ICM_USMEM=1).
  const char* ue = getenv("LSQ_B200_ICM_USMEM");
  const bool usmem = kCanUsmem && ue != nullptr && atoi(ue) != 0;
  // warps per block: 8, or 4 when the staged rows of 8 warps would leave room for a single block per SM (m > 8)
  int wpb = LSQ_ICM_LB_THREADS / 32;
  if (const char* we = getenv("LSQ_B200_ICM_WARPS_PER_BLOCK")) wpb = std::max(1, std::min(8, atoi(we)));
  const int64_t blocks_needed = ceil_div(p.n, wpb);
  const size_t smem = usmem ? 128 + (size_t)wpb * icm_staged_rows<M>() * 1024 : 0;
  int per_sm = usmem ? (int)std::min<size_t>(64 / wpb, (size_t)(227 * 1024) / (smem + 1024)) : 2 * LSQ_ICM_LB_BLOCKS(M);
  if (const char* e = getenv("LSQ_B200_ICM_BLOCKS_PER_SM")) per_sm = std::max(1, atoi(e));  // tuning override
  const int64_t cap = (int64_t)sms * per_sm;
  const unsigned grid = (unsigned)(blocks_needed < cap ? blocks_needed : cap);
  // m > 8: the pair tables (m*m*256 KB = 67 MB at m = 16) compete with the streaming unary rows for L2
  // (hit rate 77 % measured); pin them with a persisting access-policy window for this launch.
  const size_t tbytes = (size_t)M * M * LSQ_H * LSQ_H * sizeof(float);
  bool window = false;
  const char* we2 = getenv("LSQ_B200_ICM_L2WINDOW");   // A/B switch (default on for m > 8)
  if (M > 8 && !(we2 != nullptr && atoi(we2) == 0)) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_persist > 0 && max_window > 0) {
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
      cudaStreamAttrValue attr;
      memset(&attr, 0, sizeof(attr));
      attr.accessPolicyWindow.base_ptr = const_cast<float*>(p.T);
      attr.accessPolicyWindow.num_bytes = std::min(tbytes, (size_t)max_window);
      attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)max_persist / (double)tbytes);
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      window = (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess);
      cudaGetLastError();
    }
  }
  // dynamic work distribution (LSQ_B200_ICM_STATIC=1 restores the static split for A/B runs)
  IcmParams q = p;
  q.visits = g_visit_counter;
  DevBuf<unsigned long long> counter;
  const char* se = getenv("LSQ_B200_ICM_STATIC");
  if (!(se != nullptr && atoi(se) != 0) && blocks_needed > cap) {
    counter.st = st;
    LSQ_CUDA(counter.alloc(1));
    LSQ_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), st));
    q.next_vector = counter.p;
  } else {
    q.next_vector = nullptr;
  }
  if (usmem) {
    LSQ_CUDA(cudaFuncSetAttribute(icm_ils_warp_kernel<M, kCanUsmem>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    note_launch();
    icm_ils_warp_kernel<M, kCanUsmem><<<grid, 32 * wpb, smem, st>>>(q);
  } else if (q.visits != nullptr) {
    note_launch();
    icm_ils_warp_kernel<M, false, true><<<grid, 32 * wpb, 0, st>>>(q);
  } else {
    note_launch();
    icm_ils_warp_kernel<M, false><<<grid, 32 * wpb, 0, st>>>(q);
  }
  if (window) {
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.num_bytes = 0;  // disable for whatever runs next on this stream
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
  }
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

int launch_icm_warp(const IcmParams& p, cudaStream_t st) {
  if (p.n == 0 || p.niters == 0) return LSQ_OK;
  switch (p.m) {
#define LSQ_CASE(MM) case MM: return launch_icm_warp_m<MM>(p, st);
    LSQ_CASE(1) LSQ_CASE(2) LSQ_CASE(3) LSQ_CASE(4) LSQ_CASE(5) LSQ_CASE(6) LSQ_CASE(7) LSQ_CASE(8)
    LSQ_CASE(9) LSQ_CASE(10) LSQ_CASE(11) LSQ_CASE(12) LSQ_CASE(13) LSQ_CASE(14) LSQ_CASE(15) LSQ_CASE(16)
#undef LSQ_CASE
  }
  set_error("m must be in 1..16");
  return LSQ_ERR_ARG;
}

// ------------------------------------------------------------------------------------------------
// stand-alone cost / decode kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) veccost_kernel(const float* __restrict__ X, int d, int64_t n,
                                                      const uint8_t* __restrict__ codes,
                                                      const float* __restrict__ C, int m,
                                                      float* __restrict__ cost) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t v = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); v < n; v += nwarps) {
    const uint8_t* cp = codes + v * m;
    float p = 0.0f;
    for (int t = lane; t < d; t += 32) {
      float r = 0.0f;
      for (int k = 0; k < m; k++) r = __fadd_rn(r, __ldg(C + ((size_t)k * LSQ_H + cp[k]) * d + t));
      const float df = __fsub_rn(r, __ldg(X + (size_t)v * d + t));
      p = __fadd_rn(p, __fmul_rn(df, df));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(0xFFFFFFFFu, p, off));
    if (lane == 0) cost[v] = p;
  }
}

static unsigned warp_grid(int64_t n) {
  const int64_t b = ceil_div(n, 8);
  const int64_t cap = (int64_t)LSQ_NUM_SMS_HINT * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

int launch_veccost(const float* dX, int d, int64_t n, const uint8_t* dcodes, const float* dC, int m,
                   float* dcost, cudaStream_t st) {
  if (n == 0) return LSQ_OK;
  note_launch();
  veccost_kernel<<<warp_grid(n), 256, 0, st>>>(dX, d, n, dcodes, dC, m, dcost);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

// reconstruct (utils.jl:203-223): CB[v][t] = ((0 + C_0[b_0][t]) + C_1[b_1][t]) + ...
__global__ void __launch_bounds__(256) reconstruct_kernel(const uint8_t* __restrict__ codes, int64_t n,
                                                          const float* __restrict__ C, int d, int m,
                                                          float* __restrict__ CB) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t v = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); v < n; v += nwarps) {
    const uint8_t* cp = codes + v * m;
    for (int t = lane; t < d; t += 32) {
      float r = 0.0f;
      for (int k = 0; k < m; k++) r = __fadd_rn(r, __ldg(C + ((size_t)k * LSQ_H + cp[k]) * d + t));
      CB[(size_t)v * d + t] = r;
    }
  }
}

int launch_reconstruct(const uint8_t* dcodes, int64_t n, const float* dC, int d, int m, float* dCB,
                       cudaStream_t st) {
  if (n == 0) return LSQ_OK;
  note_launch();
  reconstruct_kernel<<<warp_grid(n), 256, 0, st>>>(dcodes, n, dC, d, m, dCB);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

// deterministic float64 sum of a float32 vector: fixed 1024-block partials, then one block.
__global__ void __launch_bounds__(256) sum_partial_kernel(const float* __restrict__ v, int64_t n,
                                                          double* __restrict__ partial) {
  __shared__ double sm[256];
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per;
  const int64_t hi = (lo + per < n) ? lo + per : n;
  double acc = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc += (double)v[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s >= 1; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void sum_final_kernel(const double* __restrict__ partial, int np, double* __restrict__ out) {
  __shared__ double sm[1024];
  sm[threadIdx.x] = ((int)threadIdx.x < np) ? partial[threadIdx.x] : 0.0;
  __syncthreads();
  for (int s = 512; s >= 1; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sm[0];
}

// dout: [1 + 1024] doubles; dout[0] receives the sum, the rest is scratch
int launch_sum_f32_to_f64(const float* dv, int64_t n, double* dout, cudaStream_t st) {
  note_launch();
  sum_partial_kernel<<<1024, 256, 0, st>>>(dv, n, dout + 1);
  note_launch();
  sum_final_kernel<<<1, 1024, 0, st>>>(dout + 1, 1024, dout);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

// quantize_norms (utils.jl:6-31): one thread per vector, sequential arithmetic as in the oracle.
__global__ void __launch_bounds__(128) quantize_norms_kernel(const uint8_t* __restrict__ codes, int64_t n,
                                                             const float* __restrict__ C, int d, int m,
                                                             const float* __restrict__ cbnorms, int hn,
                                                             int16_t* __restrict__ out1) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const uint8_t* cp = codes + v * m;
  float nrm = 0.0f;
  for (int t = 0; t < d; t++) {
    float r = 0.0f;
    for (int k = 0; k < m; k++) r = __fadd_rn(r, __ldg(C + ((size_t)k * LSQ_H + cp[k]) * d + t));
    nrm = __fadd_rn(nrm, __fmul_rn(r, r));
  }
  float e0 = __fsub_rn(nrm, cbnorms[0]);
  float best = __fmul_rn(e0, e0);
  int bi = 0;
  for (int j = 1; j < hn; j++) {
    const float e = __fsub_rn(nrm, cbnorms[j]);
    const float dd = __fmul_rn(e, e);
    if (dd < best) { best = dd; bi = j; }
  }
  out1[v] = (int16_t)(bi + 1);
}

int launch_quantize_norms(const uint8_t* dcodes, int64_t n, const float* dC, int d, int m,
                          const float* dcbnorms, int hn, int16_t* dout1, cudaStream_t st) {
  if (n == 0) return LSQ_OK;
  note_launch();
  quantize_norms_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, st>>>(dcodes, n, dC, d, m, dcbnorms, hn, dout1);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

// Int16 1-based (Julia) <-> uint8 0-based (device); out-of-range codes raise *derr.
__global__ void i16_to_u8_kernel(const int16_t* __restrict__ s, uint8_t* __restrict__ o, int64_t count,
                                 int* __restrict__ derr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int c = (int)s[i] - 1;
  if (c < 0 || c >= LSQ_H) { *derr = 1; o[i] = 0; }
  else o[i] = (uint8_t)c;
}
__global__ void u8_to_i16_kernel(const uint8_t* __restrict__ s, int16_t* __restrict__ o, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) o[i] = (int16_t)((int)s[i] + 1);
}

int launch_codes_i16_to_u8(const int16_t* d16, uint8_t* d8, int64_t count, int* derr, cudaStream_t st) {
  if (count == 0) return LSQ_OK;
  note_launch();
  i16_to_u8_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(d16, d8, count, derr);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}
int launch_codes_u8_to_i16(const uint8_t* d8, int16_t* d16, int64_t count, cudaStream_t st) {
  if (count == 0) return LSQ_OK;
  note_launch();
  u8_to_i16_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(d8, d16, count);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

}  // namespace lsq
