// linscan.cu — ADC lookup-table linear scan with exact top-k (replaces src/linscan/cpp/
// linscan_aqd_pairwise_byte.cpp:14-93 and linscan_aqd.cpp:37-102).
//
// Bit-exactness contract (checked against the reference's own .so in tests/):
//   LUT   LSQ: t -= (2*q[k])*c[k], k ascending, separate mul/sub   (linscan_aqd_pairwise_byte.cpp:42-48)
//         PQ : t += sqr(c[s]-q[s]),  s ascending                   (linscan_aqd.cpp:66-74)
//   dist  ((0 + LUT_0[c0]) + LUT_1[c1]) + ... (+ dbnorm), fp32 adds in that order       (:69-73 / :84-86)
//   top-k ascending by (distance, id), the order std::partial_sort gives pair<float,int> (:81 / :91)
//
// Data flow on the GPU:
//   1. lut_kernel      LUT tiles, query-fastest: lut[tile][k*256+c][QT]  (QT queries per tile)
//   2. scan_kernel     one CTA = one query tile x one slice of the base set.  The tile's whole LUT
//                      (m*256*QT floats, up to 224 KB) is staged in shared memory by TMA bulk copies.
//                      A group of 8 lanes scores one base vector for the whole tile: each lane serves 4
//                      consecutive queries with one LDS.128 of the (k, code) row, so a quarter-warp
//                      phase reads one contiguous row — conflict-free by construction — and a warp
//                      instruction scores 4 base vectors.  Code rows are loaded once per 32 base
//                      vectors (one row per lane) and broadcast by warp shuffles.
//   3. exact top-k     a strided sample of the base set gives each query a threshold tau that bounds
//                      its nn-th distance from above (checked, never assumed: a query whose candidate
//                      count ends up < nn or > capacity is re-run on the exhaustive path); the main
//                      pass appends only candidates with dist <= tau; a per-query CTA then sorts the
//                      64-bit keys (ordered(dist) << 32 | id) in shared memory (bitonic), or, when the
//                      candidate list is longer than the sorter, radix-selects the nn-th key first.
#include "adc_tc.cuh"
#include "linscan.cuh"
#include "runtime.cuh"

#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

namespace lsq {

#ifndef LSQ_SCAN_THREADS
#define LSQ_SCAN_THREADS 768  // 24 warps per SM (77 registers): m = 8 14.6 -> 13.9 ms; 1024 threads: same; 384: 15.9 ms
#endif
constexpr int SCAN_THREADS = LSQ_SCAN_THREADS;
constexpr int LUT_SMEM_BUDGET = 229376;  // 224 KB of the 227 KB per-CTA limit
constexpr int SAMPLE_MAX = 16384;
constexpr int SORT_CAP = LINSCAN_MAX_NN;  // keys the shared-memory bitonic sorter holds (128 KB)

enum { MODE_SAMPLE = 0, MODE_MAIN = 1, MODE_ALL = 2 };
enum { ST_OK = 0, ST_REDO = 1, ST_BIG = 2 };
constexpr int SEL_CAP = 5120;    // candidate keys the select kernel holds (40 KB of shared memory): 5 sigma above the ~3000 expected at nn = 1000, so the 128 KB sorter stays idle
constexpr int SEL_SORT = 1024;   // survivors it sorts (8 KB)

// Tile geometry.  A lane serves QPL consecutive queries of the tile with ONE vector shared-memory load
// per codebook (LDS.128 for QPL = 4, LDS.64 for QPL = 2); GW lanes form a group that scores one base
// vector, so a warp instruction serves 32/GW base vectors.  The tile holds as many LUTs as fit 224 KB
// of shared memory:  m <= 7: 32 queries,  m = 8: 28 (7 of 8 lanes of a group active),  m = 9: 24,
// m = 10..14: 16 (groups of 4 lanes),  m = 15, 16: 14 (two queries per lane).
struct TileCfg { int qt, qpl, gw; };
__host__ __device__ constexpr TileCfg tile_cfg(int m) {
  return (LUT_SMEM_BUDGET / (m * LSQ_H * 4)) >= 32 ? TileCfg{32, 4, 8}
       : (LUT_SMEM_BUDGET / (m * LSQ_H * 4)) >= 24 ? TileCfg{(LUT_SMEM_BUDGET / (m * LSQ_H * 4)) & ~3, 4, 8}
       : (LUT_SMEM_BUDGET / (m * LSQ_H * 4)) >= 16 ? TileCfg{16, 4, 4}
       : TileCfg{(LUT_SMEM_BUDGET / (m * LSQ_H * 4)) & ~1, 2, 8};
}
__host__ __device__ constexpr int tile_queries(int m) { return tile_cfg(m).qt; }

// ------------------------------------------------------------------------------------------------
// 1. LUT construction.  block (32, 8): x = query lane within the tile, y*8.. = 64 table rows.
// ------------------------------------------------------------------------------------------------
constexpr int LUT_KC = 64;
constexpr int LUT_JPT = 8;

template <int KIND>
__global__ void __launch_bounds__(256) lut_kernel(const float* __restrict__ queries, int nq, int qstride,
                                                  const float* __restrict__ cb, int m, int kd, int QT,
                                                  float* __restrict__ lut) {
  __shared__ float qs[LUT_KC][33];
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int tile = blockIdx.x;
  const int jblock = blockIdx.y * (8 * LUT_JPT);
  const int q0 = tile * QT;
  const int qoff = (KIND == LUT_PQ) ? (jblock / LSQ_H) * kd : 0;
  float acc[LUT_JPT];
#pragma unroll
  for (int i = 0; i < LUT_JPT; i++) acc[i] = 0.0f;

  for (int k0 = 0; k0 < kd; k0 += LUT_KC) {
    const int kc = (kd - k0 < LUT_KC) ? (kd - k0) : LUT_KC;
    __syncthreads();
    for (int e = ty * 32 + lane; e < 32 * LUT_KC; e += 256) {
      const int qq = e / LUT_KC, kk = e % LUT_KC;
      float v = 0.0f;
      if (qq < QT && q0 + qq < nq && kk < kc) v = queries[(size_t)(q0 + qq) * qstride + qoff + k0 + kk];
      qs[kk][qq] = (KIND == LUT_LSQ) ? __fmul_rn(2.0f, v) : v;  // LSQ: the factor (2*q[k]) of :45-47, exact
    }
    __syncthreads();
    const float* cbase = cb + (size_t)(jblock + ty * LUT_JPT) * kd + k0;
    if ((kd & 3) == 0 && (reinterpret_cast<uintptr_t>(cb) & 15) == 0) {
      // k outer (4 at a time), the thread's 8 rows inner: one LDS per k and one LDG.128 per row and 4 k;
      // every (query, row) accumulator still sees k in ascending order with separate mul / sub
      for (int kk = 0; kk < kc; kk += 4) {
        const float q0v = qs[kk][lane], q1v = qs[kk + 1][lane], q2v = qs[kk + 2][lane], q3v = qs[kk + 3][lane];
#pragma unroll
        for (int i = 0; i < LUT_JPT; i++) {
          const float4 cv = __ldg(reinterpret_cast<const float4*>(cbase + (size_t)i * kd + kk));
          float t = acc[i];
          if (KIND == LUT_LSQ) {
            t = __fsub_rn(t, __fmul_rn(q0v, cv.x));
            t = __fsub_rn(t, __fmul_rn(q1v, cv.y));
            t = __fsub_rn(t, __fmul_rn(q2v, cv.z));
            t = __fsub_rn(t, __fmul_rn(q3v, cv.w));
          } else {
            float df = __fsub_rn(cv.x, q0v); t = __fadd_rn(t, __fmul_rn(df, df));
            df = __fsub_rn(cv.y, q1v); t = __fadd_rn(t, __fmul_rn(df, df));
            df = __fsub_rn(cv.z, q2v); t = __fadd_rn(t, __fmul_rn(df, df));
            df = __fsub_rn(cv.w, q3v); t = __fadd_rn(t, __fmul_rn(df, df));
          }
          acc[i] = t;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < LUT_JPT; i++) {
        const float* c = cbase + (size_t)i * kd;
        float t = acc[i];
        for (int kk = 0; kk < kc; kk++) {
          const float cv = __ldg(c + kk);
          const float qv = qs[kk][lane];
          if (KIND == LUT_LSQ) {
            t = __fsub_rn(t, __fmul_rn(qv, cv));
          } else {
            const float df = __fsub_rn(cv, qv);
            t = __fadd_rn(t, __fmul_rn(df, df));
          }
        }
        acc[i] = t;
      }
    }
  }
  if (lane < QT) {
#pragma unroll
    for (int i = 0; i < LUT_JPT; i++) {
      const int j = jblock + ty * LUT_JPT + i;
      lut[((size_t)tile * m * LSQ_H + j) * QT + lane] = acc[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 2. scan
// ------------------------------------------------------------------------------------------------
struct ScanParams {
  const uint8_t* codes;      // [n][m]
  const float* norms;        // [n] or nullptr
  const float* lut;          // tiles
  const float* tau;          // [ntiles*32] thresholds (MODE_MAIN)
  unsigned long long* cand;  // [nq][cap] keys
  int* cnt;                  // [nq]
  uint32_t* sbuf;            // MODE_SAMPLE: [(tile*count + t)*32 + lane] ordered distances
  int64_t stride, count;     // base index of step t = t*stride, t < count
  int64_t cap;
  int nq, mode, id_base;
};

// shared-memory vector loads from a 32-bit shared address (ptxas folds constant offsets into the LDS)
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_v2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

template <int M, bool NORM>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(const __grid_constant__ ScanParams p) {
  constexpr TileCfg CFG = tile_cfg(M);
  constexpr int QT = CFG.qt, QPL = CFG.qpl, GW = CFG.gw;
  constexpr int VPW = 32 / GW;  // base vectors scored per warp instruction
  constexpr int U = 2;          // steps per vote group
  static_assert(GW % U == 0 && QT % QPL == 0 && QT <= GW * QPL, "tile geometry");
  constexpr uint32_t LUT_BYTES = (uint32_t)M * LSQ_H * QT * 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* lut = reinterpret_cast<float*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + LUT_BYTES);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int W = SCAN_THREADS / 32;
  const int tile = blockIdx.x;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) bulk_load_issue(lut, p.lut + (size_t)tile * (LUT_BYTES / 4), LUT_BYTES, bar);

  // Lane geometry: group `grp` scores base vector (step*VPW + grp) of the chunk for the query slots
  // qs .. qs+QPL-1.  A group's lanes read one contiguous LUT row (QT floats): with GW = 8 and LDS.128 a
  // quarter-warp phase is exactly one group, so every shared-memory wavefront is conflict-free by
  // construction; lanes past the tile (qs >= QT) re-read the row start (a broadcast) and never store.
  const int grp = lane / GW;
  const int qs = (lane % GW) * QPL;
  const bool lane_on = qs < QT;
  const int q0 = tile * QT + qs;
  float tau[QPL];
#pragma unroll
  for (int i = 0; i < QPL; i++) {
    tau[i] = -INFINITY;  // slots without a query never pass `dist <= tau`
    if (p.mode == MODE_MAIN && lane_on && q0 + i < p.nq) tau[i] = p.tau[tile * 32 + qs + i];
  }
  const uint32_t lane_base = smem_u32(lut) + (lane_on ? qs * 4 : 0);  // 32-bit shared address: one IMAD per row

  // slice of the step range handled by this CTA (multiple of 32 steps)
  int64_t per = (p.count + gridDim.y - 1) / gridDim.y;
  per = (per + 31) & ~(int64_t)31;
  const int64_t t_begin = (int64_t)blockIdx.y * per;
  const int64_t t_end = (t_begin + per < p.count) ? (t_begin + per) : p.count;

  mbar_wait(bar, 0);

  // lane l fetches the code row (and norm) of step c0 + l
  auto load_rows = [&](int64_t c0, uint64_t& lo, uint64_t& hi, float& nrm) {
    lo = 0; hi = 0; nrm = 0.0f;
    const int64_t t = c0 + lane;
    if (t < t_end) {
      const int64_t i = t * p.stride;
      const uint8_t* cp = p.codes + i * M;
      if (M == 8) {
        lo = *reinterpret_cast<const uint64_t*>(cp);
      } else if (M == 16) {
        const uint2 a = *reinterpret_cast<const uint2*>(cp);
        const uint2 b = *reinterpret_cast<const uint2*>(cp + 8);
        lo = ((uint64_t)a.y << 32) | a.x;
        hi = ((uint64_t)b.y << 32) | b.x;
      } else {
#pragma unroll
        for (int k = 0; k < M; k++) {
          const uint64_t c = cp[k];
          if (k < 8) lo |= c << (8 * k);
          else hi |= c << (8 * (k - 8));
        }
      }
      if (NORM) nrm = p.norms[i];
    }
  };

  uint64_t lo, hi, nlo, nhi;
  float nrm, nnrm;
  load_rows(t_begin + (int64_t)warp * 32, nlo, nhi, nnrm);
  for (int64_t c0 = t_begin + (int64_t)warp * 32; c0 < t_end; c0 += (int64_t)W * 32) {
    // the rows of the NEXT chunk are requested before this chunk is scored (hides the L2 latency)
    lo = nlo; hi = nhi; nrm = nnrm;
    load_rows(c0 + (int64_t)W * 32, nlo, nhi, nnrm);
    // Inner loop: per codebook one PRMT (byte extract), one IMAD (row offset + lane base), ONE vector LDS
    // with the codebook offset folded into the immediate, and QPL FADDs — (3 + QPL) / QPL issue slots
    // per lookup instead of 4, which moves the kernel from the issue limit to the shared-memory
    // wavefront limit (one wavefront per base vector and codebook).  U steps per vote group: all
    // shuffles first, then one warp-uniform vote decides whether anything has to be stored.
    const bool full = (c0 + 32 <= t_end);
    const int nvalid = full ? 32 : (int)(t_end - c0);
#pragma unroll 2
    for (int s0 = 0; s0 < GW; s0 += U) {
      float dist[U][QPL];
      int bsrc[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int b = (s0 + u) * VPW + grp;  // chunk-relative index of the base vector this lane scores
        bsrc[u] = b;
        const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, (uint32_t)lo, b);
        const uint32_t w1 = (M > 4) ? __shfl_sync(0xFFFFFFFFu, (uint32_t)(lo >> 32), b) : 0u;
        const uint32_t w2 = (M > 8) ? __shfl_sync(0xFFFFFFFFu, (uint32_t)hi, b) : 0u;
        const uint32_t w3 = (M > 12) ? __shfl_sync(0xFFFFFFFFu, (uint32_t)(hi >> 32), b) : 0u;
        float acc[QPL];
#pragma unroll
        for (int i = 0; i < QPL; i++) acc[i] = 0.0f;
#pragma unroll
        for (int k = 0; k < M; k++) {
          const uint32_t w = (k < 4) ? w0 : (k < 8) ? w1 : (k < 12) ? w2 : w3;
          const uint32_t c = __byte_perm(w, 0u, 0x4440u | (uint32_t)(k & 3));
          const uint32_t row = lane_base + c * (uint32_t)(QT * 4);
          if (QPL == 4) {
            const float4 v = lds_v4(row + (uint32_t)(k * LSQ_H * QT * 4));
            acc[0] = __fadd_rn(acc[0], v.x);
            acc[1] = __fadd_rn(acc[1], v.y);
            acc[QPL - 2] = __fadd_rn(acc[QPL - 2], v.z);
            acc[QPL - 1] = __fadd_rn(acc[QPL - 1], v.w);
          } else {
            const float2 v = lds_v2(row + (uint32_t)(k * LSQ_H * QT * 4));
            acc[0] = __fadd_rn(acc[0], v.x);
            acc[1] = __fadd_rn(acc[1], v.y);
          }
        }
        if (NORM) {
          const float nb = __shfl_sync(0xFFFFFFFFu, nrm, b);
#pragma unroll
          for (int i = 0; i < QPL; i++) acc[i] = __fadd_rn(acc[i], nb);
        }
#pragma unroll
        for (int i = 0; i < QPL; i++) dist[u][i] = acc[i];
      }
      if (p.mode == MODE_MAIN) {
        bool hit = false;
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
          for (int i = 0; i < QPL; i++) hit |= (dist[u][i] <= tau[i]);
        if (__any_sync(0xFFFFFFFFu, hit)) {
#pragma unroll
          for (int u = 0; u < U; u++) {
            if (full || bsrc[u] < nvalid) {
              const uint32_t id = (uint32_t)(c0 + bsrc[u] + p.id_base);
#pragma unroll
              for (int i = 0; i < QPL; i++) {
                if (dist[u][i] <= tau[i]) {
                  const unsigned long long key = ((unsigned long long)float_to_ordered(dist[u][i]) << 32) | id;
                  const int pos = atomicAdd(&p.cnt[q0 + i], 1);
                  if (pos < p.cap) p.cand[(size_t)(q0 + i) * p.cap + pos] = key;
                }
              }
            }
          }
        }
      } else if (lane_on) {
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int64_t t = c0 + bsrc[u];
          if (full || bsrc[u] < nvalid) {
#pragma unroll
            for (int i = 0; i < QPL; i++) {
              if (q0 + i < p.nq) {
                if (p.mode == MODE_SAMPLE) {
                  p.sbuf[((size_t)tile * p.count + t) * 32 + qs + i] = float_to_ordered(dist[u][i]);
                } else {
                  const uint32_t id = (uint32_t)(t * p.stride + p.id_base);
                  p.cand[(size_t)(q0 + i) * p.cap + t] =
                      ((unsigned long long)float_to_ordered(dist[u][i]) << 32) | id;
                }
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 3a. per-query threshold: the r-th smallest (1-based) of the tile's sample, lane = query.
//     MSB-first radix select, 4 passes of 8 bits over the ordered-uint distances.
// ------------------------------------------------------------------------------------------------
constexpr int THR_SMALL_R = 56;   // fast path of threshold_kernel: r <= 56 (of the 64 kept minima)
constexpr int THR_LCAP = 191;     // per-query list it may collect (about 70 at r = 49) before falling back to the radix select

__global__ void __launch_bounds__(1024) threshold_kernel(const uint32_t* __restrict__ sbuf, int64_t s, int r,
                                                         float* __restrict__ tau, int qt) {
  __shared__ int hist[32][257];       // radix path; the fast path reuses it as mins[64][32] + list[32][256]
  __shared__ uint32_t prefix[32];
  __shared__ int rank[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const uint32_t* src = sbuf + (size_t)tile * s * 32 + lane;

  if (r <= THR_SMALL_R) {
    static_assert(8 * 257 >= 64 * 32 && 8 * 257 + 32 * (THR_LCAP + 1) <= 32 * 257, "fast-path buffers must fit the histogram");
    // r is tiny against s (49 of 16384 for top-1000 of 1 M): two passes instead of four, no histogram.
    // Pass 1: every (warp, query) keeps the 2 smallest of its 1/32 of the sample; the r-th smallest of those 64
    // values bounds the r-th smallest of the whole sample from above (they are 64 distinct sample elements).
    // Pass 2: the few values <= that bound are collected per query and ranked exactly.
    uint32_t (*mins)[32] = reinterpret_cast<uint32_t (*)[32]>(&hist[0][0]);                   // [64][32]
    uint32_t (*list)[THR_LCAP + 1] = reinterpret_cast<uint32_t (*)[THR_LCAP + 1]>(&hist[8][0]);   // [32][192], behind mins
    uint32_t a = 0xFFFFFFFFu, b = 0xFFFFFFFFu;   // a <= b: the two smallest so far
    const int64_t steps = (lane < qt) ? (s - warp + 31) / 32 : 0;   // lanes past the tile's queries hold no data
    auto keep2 = [&](uint32_t v) {
      if (v < b) {
        if (v < a) { b = a; a = v; } else { b = v; }
      }
    };
    int64_t i = 0;
    for (; i + 8 <= steps; i += 8) {   // eight independent loads in flight per thread
      uint32_t v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = __ldg(src + (warp + 32 * (i + u)) * 32);
#pragma unroll
      for (int u = 0; u < 8; u++) keep2(v[u]);
    }
    for (; i < steps; i++) keep2(__ldg(src + (warp + 32 * i) * 32));
    mins[2 * warp][lane] = a;
    mins[2 * warp + 1][lane] = b;
    if (tid < 32) rank[tid] = 0;   // list fill counters
    __syncthreads();
    {
      int ra = 0, rb = 0;   // ranks of a and b among the 64 values of this query (ties by position)
      for (int j = 0; j < 64; j++) {
        const uint32_t v = mins[j][lane];
        ra += (v < a) || (v == a && j < 2 * warp);
        rb += (v < b) || (v == b && j < 2 * warp + 1);
      }
      if (ra == r - 1) prefix[lane] = a;
      if (rb == r - 1) prefix[lane] = b;
    }
    __syncthreads();
    const uint32_t bound = prefix[lane];
    __syncthreads();   // mins is dead from here on: list overlaps nothing of it, but keep the phases apart
    auto collect = [&](uint32_t v) {
      if (v <= bound) {
        const int pos = atomicAdd(&rank[lane], 1);
        if (pos <= THR_LCAP) list[lane][pos] = v;
      }
    };
    for (i = 0; i + 8 <= steps; i += 8) {
      uint32_t v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = __ldg(src + (warp + 32 * (i + u)) * 32);
#pragma unroll
      for (int u = 0; u < 8; u++) collect(v[u]);
    }
    for (; i < steps; i++) collect(__ldg(src + (warp + 32 * i) * 32));
    __syncthreads();
    const int c = rank[warp];   // warp w ranks the list of query w
    const bool overflow = (warp < qt) && (c > THR_LCAP + 1);
    if (!overflow) {
      for (int e = lane; e < c; e += 32) {
        const uint32_t v = list[warp][e];
        int rk = 0;
        for (int j = 0; j < c; j++) {
          const uint32_t u = list[warp][j];
          rk += (u < v) || (u == v && j < e);
        }
        if (rk == r - 1) tau[tile * 32 + warp] = ordered_to_float(v);
      }
    }
    if (!__syncthreads_or(overflow)) return;
    // a query collected more than the list holds (heavy ties, or the sample order defeated pass 1): radix select
  }

  if (tid < 32) { prefix[tid] = 0; rank[tid] = r; }
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    const uint32_t pmask = (pass == 0) ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int e = tid; e < 32 * 257; e += 1024) (&hist[0][0])[e] = 0;
    __syncthreads();
    const uint32_t pre = prefix[lane];
    for (int64_t t = warp; t < s; t += 32) {
      const uint32_t v = sbuf[((size_t)tile * s + t) * 32 + lane];
      if ((v & pmask) == pre) atomicAdd(&hist[lane][(v >> shift) & 0xFFu], 1);
    }
    __syncthreads();
    if (tid < 32) {
      int rr = rank[tid], dsel = 255;
      for (int b = 0; b < 256; b++) {
        const int c = hist[tid][b];
        if (rr <= c) { dsel = b; break; }
        rr -= c;
      }
      rank[tid] = rr;
      prefix[tid] |= (uint32_t)dsel << shift;
    }
    __syncthreads();
  }
  if (tid < 32) tau[tile * 32 + tid] = ordered_to_float(prefix[tid]);
}

// ------------------------------------------------------------------------------------------------
// 3b. per-query exact top-nn over its candidate keys
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort_smem(unsigned long long* keys, int N, int tid, int nthreads) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < N; i += nthreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// CAPK = keys the CTA's shared-memory sorter holds, NT = threads.  phase 0 (exhaustive path): handle every
// query; phase 2: only the queries topk_select_kernel flagged ST_BIG (more candidates or a larger nn than it
// holds), taken from its worklist.
template <int CAPK, int NT>
__global__ void __launch_bounds__(NT) topk_kernel(const unsigned long long* __restrict__ cand, int64_t cap,
                                                  const int* __restrict__ cnt, int64_t fixed_count, int nn,
                                                  const int* __restrict__ scatter, float* __restrict__ dists,
                                                  int32_t* __restrict__ ids, int* __restrict__ status, int phase,
                                                  const int* __restrict__ worklist, const int* __restrict__ nwork) {
  extern __shared__ __align__(16) unsigned long long keys[];  // CAPK keys
  __shared__ int hist[256];
  __shared__ unsigned long long sh_prefix;
  __shared__ int sh_rank, sh_fill;
  const int tid = threadIdx.x;
  // without a worklist: one query per CTA (q = blockIdx.x); with one (phase 2): the flagged queries only,
  // a few persistent CTAs instead of a launch of one mostly idle 128 KB CTA per query
  const int nitems = worklist ? *nwork : (int)gridDim.x;
  for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
  const int q = worklist ? worklist[w] : w;
  const int64_t c = (fixed_count >= 0) ? fixed_count : (int64_t)cnt[q];
  const int qo = scatter ? scatter[q] : q;
  const int prior = (phase == 2) ? status[q] : ST_BIG;
  __syncthreads();  // flag read by every thread before thread 0 may overwrite it; shared memory of the previous item is free
  if (prior != ST_BIG) continue;
  if (c < nn || c > cap) {
    if (tid == 0) status[q] = ST_REDO;
    continue;
  }
  if (tid == 0) status[q] = ST_OK;
  const unsigned long long* src = cand + (size_t)q * cap;
  int N;
  if (c <= CAPK) {
    N = 2;
    while (N < c) N <<= 1;
    for (int i = tid; i < N; i += NT) keys[i] = (i < c) ? src[i] : ~0ull;
    __syncthreads();
  } else {
    // radix-select the nn-th smallest key (keys are unique: the id is part of the key)
    if (tid == 0) { sh_prefix = 0ull; sh_rank = nn; }
    for (int pass = 0; pass < 8; pass++) {
      const int shift = 56 - 8 * pass;
      const unsigned long long pmask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const unsigned long long pre = sh_prefix;
      for (int64_t i = tid; i < c; i += NT) {
        const unsigned long long v = src[i];
        if ((v & pmask) == pre) atomicAdd(&hist[(int)((v >> shift) & 0xFFull)], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int rr = sh_rank, dsel = 255;
        for (int b = 0; b < 256; b++) {
          const int h = hist[b];
          if (rr <= h) { dsel = b; break; }
          rr -= h;
        }
        sh_rank = rr;
        sh_prefix = pre | ((unsigned long long)dsel << shift);
      }
      __syncthreads();
    }
    const unsigned long long kth = sh_prefix;
    if (tid == 0) sh_fill = 0;
    N = 2;
    while (N < nn) N <<= 1;
    for (int i = tid; i < N; i += NT) keys[i] = ~0ull;
    __syncthreads();
    for (int64_t i = tid; i < c; i += NT) {
      const unsigned long long v = src[i];
      if (v <= kth) keys[atomicAdd(&sh_fill, 1)] = v;
    }
    __syncthreads();
  }
  bitonic_sort_smem(keys, N, tid, NT);
  for (int j = tid; j < nn; j += NT) {
    const unsigned long long k = keys[j];
    dists[(size_t)qo * nn + j] = ordered_to_float((uint32_t)(k >> 32));
    ids[(size_t)qo * nn + j] = (int32_t)(uint32_t)(k & 0xFFFFFFFFull);
  }
  }
}

// Phase-1 top-k for the common case (a few thousand candidates, nn <= SORTN): the bitonic sort of every
// candidate was issue-bound (4096 keys x 78 stages per query); here the nn-th smallest key is found by
// an MSB-first radix select over the keys held in shared memory (early exit as soon as the selected
// bucket is taken whole), the nn survivors are compacted and only they are sorted.  Keys are unique
// (the id is part of the key), so exactly nn keys are <= the nn-th.
template <int CAPK, int SORTN, int NT>
__global__ void __launch_bounds__(NT) topk_select_kernel(const unsigned long long* __restrict__ cand, int64_t cap,
                                                         const int* __restrict__ cnt, int nn,
                                                         float* __restrict__ dists, int32_t* __restrict__ ids,
                                                         int* __restrict__ status, int* __restrict__ biglist,
                                                         int* __restrict__ nbig) {
  extern __shared__ __align__(16) unsigned long long sel_smem[];  // CAPK candidate keys, then SORTN survivors
  unsigned long long* keys = sel_smem;
  unsigned long long* out = sel_smem + CAPK;
  __shared__ int hist[256];
  __shared__ unsigned long long sh_prefix;
  __shared__ int sh_rank, sh_fill, sh_done;
  const int tid = threadIdx.x;
  const int q = blockIdx.x;
  const int64_t c64 = (int64_t)cnt[q];
  if (c64 < nn || c64 > cap) {
    if (tid == 0) status[q] = ST_REDO;
    return;
  }
  if (c64 > CAPK || nn > SORTN) {
    if (tid == 0) { status[q] = ST_BIG; biglist[atomicAdd(nbig, 1)] = q; }
    return;
  }
  if (tid == 0) status[q] = ST_OK;
  const int c = (int)c64;
  const unsigned long long* src = cand + (size_t)q * cap;
  for (int i = tid; i < c; i += NT) keys[i] = src[i];
  if (tid == 0) { sh_prefix = 0ull; sh_rank = nn; sh_done = 0; sh_fill = 0; }
  int N = 2;
  while (N < nn) N <<= 1;
  for (int i = tid; i < N; i += NT) out[i] = ~0ull;
  unsigned long long below = ~0ull;  // mask of the bits not yet decided
  for (int pass = 0; pass < 8; pass++) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    if (sh_done) break;
    const unsigned long long pre = sh_prefix;
    const unsigned long long pmask = ~below;
    for (int i = tid; i < c; i += NT) {
      const unsigned long long v = keys[i];
      if ((v & pmask) == pre) atomicAdd(&hist[(int)((v >> shift) & 0xFFull)], 1);
    }
    __syncthreads();
    below = (shift == 0) ? 0ull : (~0ull >> (64 - shift));
    if (tid < 32) {
      // warp scan over 32 groups of 8 buckets, then the owning lane walks its 8 buckets
      int part = 0;
#pragma unroll
      for (int b = 0; b < 8; b++) part += hist[tid * 8 + b];
      const int rr = sh_rank;
      __syncwarp();  // every lane has read the rank before the owning lane overwrites it below
      int incl = part;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (tid >= off) incl += t;
      }
      const int excl = incl - part;
      if (rr > excl && rr <= incl) {
        int r2 = rr - excl, dsel = tid * 8 + 7;
#pragma unroll
        for (int b = 0; b < 8; b++) {
          const int hcount = hist[tid * 8 + b];
          if (r2 <= hcount) { dsel = tid * 8 + b; break; }
          r2 -= hcount;
        }
        unsigned long long npre = pre | ((unsigned long long)dsel << shift);
        // the whole bucket is taken: every key <= (prefix | remaining ones) is a survivor
        if (r2 == hist[dsel]) { npre |= below; sh_done = 1; }
        sh_rank = r2;
        sh_prefix = npre;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  const unsigned long long kth = sh_prefix;
  for (int i = tid; i < c; i += NT) {
    const unsigned long long v = keys[i];
    if (v <= kth) out[atomicAdd(&sh_fill, 1)] = v;
  }
  __syncthreads();
  // bitonic sort of the N survivors, one compare-exchange per thread and step
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (N >> 1); t += NT) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
        const int ixj = i | j;
        const unsigned long long a = out[i], b = out[ixj];
        const bool up = ((i & k) == 0);
        if ((a > b) == up) { out[i] = b; out[ixj] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < nn; j += NT) {
    const unsigned long long k = out[j];
    dists[(size_t)q * nn + j] = ordered_to_float((uint32_t)(k >> 32));
    ids[(size_t)q * nn + j] = (int32_t)(uint32_t)(k & 0xFFFFFFFFull);
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ rows, int nrows, int d,
                                   float* __restrict__ dst) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)nrows * d) return;
  const int r = (int)(e / d), k = (int)(e % d);
  dst[e] = src[(size_t)rows[r] * d + k];
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
template <int M>
static int launch_scan_m(const ScanParams& p, int ntiles, int nsplit, cudaStream_t st) {
  constexpr int QT = tile_queries(M);
  constexpr size_t smem = (size_t)M * LSQ_H * QT * 4 + 16;
  dim3 grid(ntiles, nsplit, 1);
  if (p.norms != nullptr) {
    LSQ_CUDA(cudaFuncSetAttribute(scan_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    note_launch();
    scan_kernel<M, true><<<grid, SCAN_THREADS, smem, st>>>(p);
  } else {
    LSQ_CUDA(cudaFuncSetAttribute(scan_kernel<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    note_launch();
    scan_kernel<M, false><<<grid, SCAN_THREADS, smem, st>>>(p);
  }
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

static int launch_scan(int m, const ScanParams& p, int ntiles, cudaStream_t st) {
  // Every CTA does the same amount of work, so the pass takes ceil(CTAs / resident slots) waves of
  // (1 / nsplit) each: pick the split of the base set that wastes the least of the last wave (358 tiles:
  // 2 slices = 4.84 waves -> 5, 3.2 % idle; 7 slices = 16.93 -> 17, 0.4 %), slices of at least 4096 steps.
  const size_t smem = (size_t)m * LSQ_H * tile_queries(m) * 4 + 16;
  const int64_t slots = (int64_t)LSQ_NUM_SMS_HINT * std::max<int64_t>(1, (int64_t)(227 * 1024) / (int64_t)(smem + 1024));
  const int64_t max_split = std::max<int64_t>(1, std::min<int64_t>(1024, p.count / 4096));  // few tiles (small query batches): up to one slice per SM slot
  int nsplit = 1;
  double best = 1e30;
  const double staging = 6000.0 / ((double)std::max<int64_t>(p.count, 1) * m);  // LUT load vs one CTA scanning everything
  for (int64_t sp = 1; sp <= max_split; sp++) {
    const double cost = (double)ceil_div((int64_t)ntiles * sp, slots) / (double)sp + staging * (double)sp;
    if (cost < best - 1e-12) { best = cost; nsplit = (int)sp; }
  }
  switch (m) {
#define LSQ_CASE(MM) case MM: return launch_scan_m<MM>(p, ntiles, nsplit, st);
    LSQ_CASE(1) LSQ_CASE(2) LSQ_CASE(3) LSQ_CASE(4) LSQ_CASE(5) LSQ_CASE(6) LSQ_CASE(7) LSQ_CASE(8)
    LSQ_CASE(9) LSQ_CASE(10) LSQ_CASE(11) LSQ_CASE(12) LSQ_CASE(13) LSQ_CASE(14) LSQ_CASE(15) LSQ_CASE(16)
#undef LSQ_CASE
  }
  set_error("linscan: m must be in 1..16");
  return LSQ_ERR_ARG;
}

static int launch_lut(int lut_kind, const float* dq, int nq, int qstride, const float* dcb, int m, int kd, int QT,
                      float* dlut, cudaStream_t st) {
  const int ntiles = (int)ceil_div(nq, QT);
  dim3 grid(ntiles, m * LSQ_H / (8 * LUT_JPT), 1), block(32, 8, 1);
  note_launch();
  if (lut_kind == LUT_LSQ) lut_kernel<LUT_LSQ><<<grid, block, 0, st>>>(dq, nq, qstride, dcb, m, kd, QT, dlut);
  else lut_kernel<LUT_PQ><<<grid, block, 0, st>>>(dq, nq, qstride, dcb, m, kd, QT, dlut);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

static int configure_topk() {
  LSQ_CUDA(cudaFuncSetAttribute(topk_kernel<SORT_CAP, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_CAP * 8));
  LSQ_CUDA(cudaFuncSetAttribute(topk_select_kernel<SEL_CAP, SEL_SORT, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (SEL_CAP + SEL_SORT) * 8));
  return LSQ_OK;
}

// LSQ_B200_ADC_TIMING=1: device time of every phase of a linscan call (CUDA events on the call's stream), printed
// to stderr at the end of the call.  Measurement aid; off by default (no events, no extra synchronisation).
static thread_local std::vector<std::pair<std::string, float>> g_last_phases;
static thread_local int g_last_products = 0;   // products of the tensor-core filter in the last timed call (0: lookup scan)   // of the calling thread's last timed call

struct PhaseTimer {
  bool on;
  cudaStream_t st;
  std::vector<std::pair<const char*, cudaEvent_t>> ev;
  PhaseTimer(cudaStream_t s) : on(getenv("LSQ_B200_ADC_TIMING") != nullptr), st(s) { mark("start"); }
  void mark(const char* name) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.emplace_back(name, e);
  }
  ~PhaseTimer() {
    if (!on) return;
    cudaStreamSynchronize(st);
    std::string line = "linscan phases (ms):";
    for (size_t i = 1; i < ev.size(); i++) {
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
      char buf[96];
      snprintf(buf, sizeof(buf), " %s %.3f", ev[i].first, ms);
      line += buf;
    }
    float tot = 0.0f;
    if (ev.size() > 1) cudaEventElapsedTime(&tot, ev.front().second, ev.back().second);
    fprintf(stderr, "%s | total %.3f\n", line.c_str(), tot);
    g_last_phases.clear();
    for (size_t i = 1; i < ev.size(); i++) {
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
      g_last_phases.emplace_back(ev[i].first, ms);
    }
    for (auto& e : ev) cudaEventDestroy(e.second);
  }
};

struct ScanCtx {
  const uint8_t* dcodes; int64_t n; int m;
  const float* dcb; const float* dnorms;
  int lut_kind, kd, qstride, nn, id_base;
  float* ddists; int32_t* dids;
  cudaStream_t st;
};

// exhaustive path: every base vector becomes a candidate.  `dq` holds nqc query rows (stride qstride);
// results go to output row scatter[i] (device int array) or i.
static int scan_exhaustive(const ScanCtx& S, const float* dq, int nqc, const int* dscatter) {
  const int QT = tile_queries(S.m);
  // batch so that the candidate buffer stays <= 1 GiB
  int64_t batch = ((int64_t)1 << 27) / std::max<int64_t>(S.n, 1);
  batch = std::max<int64_t>(1, std::min<int64_t>(batch, nqc));
  // MODE_ALL writes candidate rows only for q < nq, so only the LUT needs whole query tiles: rounding the
  // key buffer up to a tile (up to 32 rows of n keys) would exceed the 1 GiB budget 32-fold for large n
  DevBuf<unsigned long long> dcand;
  DevBuf<float> dlut;
  DevBuf<int> dstatus;
  LSQ_CUDA(dcand.alloc((size_t)batch * S.n));
  LSQ_CUDA(dlut.alloc((size_t)ceil_div(batch, QT) * S.m * LSQ_H * QT));
  LSQ_CUDA(dstatus.alloc(batch));
  LSQ_TRY(configure_topk());
  for (int64_t q0 = 0; q0 < nqc; q0 += batch) {
    const int nb = (int)std::min<int64_t>(batch, nqc - q0);
    const int ntiles = (int)ceil_div(nb, QT);
    LSQ_TRY(launch_lut(S.lut_kind, dq + (size_t)q0 * S.qstride, nb, S.qstride, S.dcb, S.m, S.kd, QT, dlut.p, S.st));
    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.codes = S.dcodes; p.norms = S.dnorms; p.lut = dlut.p; p.cand = dcand.p;
    p.stride = 1; p.count = S.n; p.cap = S.n; p.nq = nb; p.mode = MODE_ALL; p.id_base = S.id_base;
    LSQ_TRY(launch_scan(S.m, p, ntiles, S.st));
    // outputs of this batch: rows q0.. (or scattered)
    if (dscatter) {
      note_launch();
      topk_kernel<SORT_CAP, 1024><<<nb, 1024, SORT_CAP * 8, S.st>>>(dcand.p, S.n, nullptr, S.n, S.nn, dscatter + q0,
                                                                     S.ddists, S.dids, dstatus.p, 0, nullptr, nullptr);
    } else {
      note_launch();
      topk_kernel<SORT_CAP, 1024><<<nb, 1024, SORT_CAP * 8, S.st>>>(dcand.p, S.n, nullptr, S.n, S.nn, nullptr,
                                                                     S.ddists + (size_t)q0 * S.nn,
                                                                     S.dids + (size_t)q0 * S.nn, dstatus.p, 0, nullptr, nullptr);
    }
    LSQ_CUDA(cudaGetLastError());
  }
  LSQ_CUDA(cudaStreamSynchronize(S.st));
  return LSQ_OK;
}

int linscan_device(const uint8_t* dcodes, int64_t n, int m, const float* dqueries, int nq, int d,
                   const float* dcodebooks, const float* dbnorms, int lut_kind, int subdim, int nn,
                   float* ddists, int32_t* dids, cudaStream_t st) {
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM, "linscan: m must be in 1..16");
  LSQ_CHECK_ARG(nq >= 0 && n >= 0 && d >= 1, "linscan: bad sizes");
  LSQ_CHECK_ARG(n < ((int64_t)1 << 31) - 1, "linscan: n must fit an int32 id");
  LSQ_CHECK_ARG(nn >= 0 && nn <= n, "linscan: need 0 <= nn <= ncodes");
  if (nn > LINSCAN_MAX_NN) {
    set_error("linscan: nn > 16384 exceeds the top-k sorter capacity");
    return LSQ_ERR_LIMIT;
  }
  if (lut_kind == LUT_PQ) LSQ_CHECK_ARG(subdim >= 1 && (int64_t)subdim * m <= d, "linscan_pq: subdim*m must be <= d");
  if (nq == 0 || nn == 0) return LSQ_OK;

  ScanCtx S;
  S.dcodes = dcodes; S.n = n; S.m = m; S.dcb = dcodebooks;
  S.dnorms = (lut_kind == LUT_LSQ) ? dbnorms : nullptr;
  S.lut_kind = lut_kind; S.kd = (lut_kind == LUT_LSQ) ? d : subdim; S.qstride = d; S.nn = nn;
  S.id_base = (lut_kind == LUT_LSQ) ? 1 : 0;
  S.ddists = ddists; S.dids = dids; S.st = st;

  // sampling plan
  const int64_t s = std::min<int64_t>(SAMPLE_MAX, n);
  const int64_t stride = n / s;
  const double mu = (double)nn * (double)s / (double)n;
  const int64_t r = (int64_t)ceil(mu + 6.0 * sqrt(mu) + 8.0);
  const double expected = (double)r * (double)n / (double)s;
  int64_t cap = SORT_CAP;
  while ((double)cap < 2.0 * expected) cap <<= 1;
  if (n <= SORT_CAP || r >= s || cap >= n) return scan_exhaustive(S, dqueries, nq, nullptr);

  const int QT = tile_queries(m);
  LSQ_TRY(configure_topk());
  // query batches so that candidates + sample stay within ~4 GiB
  int64_t qbatch = ((int64_t)1 << 32) / (cap * 8 + s * 4);
  qbatch = std::max<int64_t>(QT, std::min<int64_t>(qbatch, nq));
  qbatch = ceil_div(qbatch, QT) * QT;
  const int max_tiles = (int)ceil_div(qbatch, QT);

  PhaseTimer timer(st);
  DevBuf<float> dlut, dtau;
  DevBuf<uint32_t> dsbuf;
  DevBuf<unsigned long long> dcand;
  DevBuf<int> dcnt, dstatus, dbig;  // dbig: worklist of the queries the select kernel flags + its length
  // LSQ tables of an inner product: the main pass runs as a tensor-core filter + exact rescoring of the
  // survivors (adc_tc.cu); everything around it (LUT, sample pass, thresholds, top-k, re-runs) is unchanged
  // tensor-core dimension: d for the LSQ tables, m * subdim for PQ / OPQ (queries keep their row stride d)
  const int td = (lut_kind == LUT_LSQ) ? d : m * subdim;
  bool use_tc = ((lut_kind == LUT_LSQ) ? (dbnorms != nullptr) : (d % 4 == 0)) &&
                adc_tc_applicable(dcodes, n, nq, m, td, dqueries, dcodebooks);
  AdcTcBase tcbase;
  DevBuf<uint32_t> dcandidx;   // filter survivors of the main pass; before that, the sample lists of the thresholds
  DevBuf<int> dccnt;
  DevBuf<float> dbound, dinfl;
  DevBuf<int> dnpass;   // products of the main filter, chosen on the device from the sample
  bool two_stage = false;
  int64_t r0 = 0;
  int lcap = 0;
  if (use_tc && (adc_tc_prepare(dcodes, n, m, dcodebooks, td, S.dnorms, s, stride, (lut_kind == LUT_LSQ) ? 0 : subdim, d, st,
                                tcbase) != LSQ_OK ||
                 dcandidx.alloc((size_t)qbatch * cap) != cudaSuccess)) {
    cudaGetLastError();   // no room for the operand images / survivor lists: the lookup scan needs neither
    use_tc = false;
  }
  if (use_tc) {
    LSQ_CUDA(dccnt.alloc(qbatch));
    timer.mark("decode");
    // list-based thresholds: of the r smallest sample values about r/8 fall into the 1/8 sub-sample; its r0-th
    // smallest (7 sigma + 6 above that) bounds the r-th smallest of the sample except with probability ~1e-10, and
    // leaves lists of about 8 r0 sample positions per query
    if (tcbase.s1count > 0 && getenv("LSQ_B200_ADC_SBUF") == nullptr) {
      const double mu1 = (double)r / 8.0;
      r0 = (int64_t)ceil(mu1 + 7.0 * sqrt(mu1) + 6.0);
      lcap = 1024;
      while (lcap < 3 * 8 * r0) lcap <<= 1;
      two_stage = (r0 < tcbase.s1count) && (lcap <= 8192) && (lcap <= cap);
      if (two_stage) {
        LSQ_CUDA(dbound.alloc((size_t)ceil_div(qbatch, 32) * 32));   // threshold_kernel writes whole 32-query tiles
        LSQ_CUDA(dinfl.alloc(qbatch));
        LSQ_CUDA(dnpass.alloc(1));
      }
    }
  }
  LSQ_CUDA(dbig.alloc(qbatch + 1));
  LSQ_CUDA(dlut.alloc((size_t)max_tiles * m * LSQ_H * QT));
  LSQ_CUDA(dtau.alloc((size_t)max_tiles * 32));
  LSQ_CUDA(dsbuf.alloc((size_t)max_tiles * s * 32));
  LSQ_CUDA(dcand.alloc((size_t)qbatch * cap));
  LSQ_CUDA(dcnt.alloc(qbatch));
  LSQ_CUDA(dstatus.alloc(qbatch));
  std::vector<int> hstatus(qbatch), redo;

  for (int64_t q0 = 0; q0 < nq; q0 += qbatch) {
    const int nb = (int)std::min<int64_t>(qbatch, nq - q0);
    const int ntiles = (int)ceil_div(nb, QT);
    const float* dq = dqueries + (size_t)q0 * d;
    if (use_tc) {
      // thresholds from the tensor-core values of the sample (any tau is valid: a query whose candidate count
      // ends up < nn or > capacity is re-run), exact LUT rows, filter + exact rescoring of the survivors
      const int tiles32 = (int)ceil_div(nb, 32);
      LSQ_TRY(adc_tc_lut_rows(tcbase, dq, nb, td, dcodebooks, m, dlut.p, st));
      timer.mark("lut");
      if (two_stage) {
        // coarse bound from the 1/8 sub-sample (r0-th smallest filter value), then the sample positions below it,
        // scored exactly: tau = the exact r-th smallest distance of the sample, no 655 MB sample buffer
        LSQ_TRY(adc_tc_sample(tcbase, true, dq, nb, td, m, dsbuf.p, st));
        note_launch();
        threshold_kernel<<<tiles32, 1024, 0, st>>>(dsbuf.p, tcbase.s1count, (int)r0, dbound.p, 32);
        LSQ_CUDA(cudaGetLastError());
        timer.mark("bound");
        LSQ_TRY(adc_tc_sample_tau(tcbase, dcodes, m, dq, nb, td, S.dnorms, dlut.p, dbound.p, dcandidx.p, dccnt.p, lcap, (int)r,
                                  dtau.p, dinfl.p, dnpass.p, st));
        timer.mark("threshold");
      } else {
        LSQ_TRY(adc_tc_sample(tcbase, false, dq, nb, td, m, dsbuf.p, st));
        timer.mark("sample");
        note_launch();
        threshold_kernel<<<tiles32, 1024, 0, st>>>(dsbuf.p, s, (int)r, dtau.p, 32);
        LSQ_CUDA(cudaGetLastError());
        timer.mark("threshold");
      }
      LSQ_CUDA(cudaMemsetAsync(dbig.p + qbatch, 0, sizeof(int), st));
      LSQ_TRY(adc_tc_main_pass(tcbase, dcodes, n, m, dq, nb, td, S.dnorms, dlut.p, dtau.p, dcandidx.p, dccnt.p, cap,
                               nullptr, nullptr, cap, S.id_base, nullptr, 0, two_stage ? dnpass.p : nullptr, st));
      timer.mark("filter");
      LSQ_TRY(adc_tc_rescore(dcodes, n, m, nb, S.dnorms, dlut.p, dtau.p, dcandidx.p, dccnt.p, cap, dcand.p, dcnt.p, cap,
                             S.id_base, st));
      timer.mark("rescore");
    } else {
      LSQ_TRY(launch_lut(lut_kind, dq, nb, d, dcodebooks, m, S.kd, QT, dlut.p, st));
      ScanParams p;
      memset(&p, 0, sizeof(p));
      p.codes = dcodes; p.norms = S.dnorms; p.lut = dlut.p; p.tau = dtau.p; p.cand = dcand.p; p.cnt = dcnt.p;
      p.sbuf = dsbuf.p; p.cap = cap; p.nq = nb; p.id_base = S.id_base;
      // sample pass -> thresholds
      p.mode = MODE_SAMPLE; p.stride = stride; p.count = s;
      LSQ_TRY(launch_scan(m, p, ntiles, st));
      note_launch();
      threshold_kernel<<<ntiles, 1024, 0, st>>>(dsbuf.p, s, (int)r, dtau.p, QT);
      LSQ_CUDA(cudaGetLastError());
      // main pass
      LSQ_CUDA(cudaMemsetAsync(dcnt.p, 0, (size_t)nb * sizeof(int), st));
      LSQ_CUDA(cudaMemsetAsync(dbig.p + qbatch, 0, sizeof(int), st));
      timer.mark("lut+sample+threshold");
      p.mode = MODE_MAIN; p.stride = 1; p.count = n;
      LSQ_TRY(launch_scan(m, p, ntiles, st));
      timer.mark("scan");
    }
    // most queries end up with a few thousand candidates and nn <= 1024: select + sort of the survivors
    // in 40 KB of shared memory (5 CTAs per SM); the 128 KB sorter only runs for the queries it flags
    note_launch();
    topk_select_kernel<SEL_CAP, SEL_SORT, 256><<<nb, 256, (SEL_CAP + SEL_SORT) * 8, st>>>(
        dcand.p, cap, dcnt.p, nn, ddists + (size_t)q0 * nn, dids + (size_t)q0 * nn, dstatus.p, dbig.p, dbig.p + qbatch);
    note_launch();
    topk_kernel<SORT_CAP, 1024><<<std::min(nb, LSQ_NUM_SMS_HINT), 1024, SORT_CAP * 8, st>>>(dcand.p, cap, dcnt.p, -1, nn, nullptr,
                                                                 ddists + (size_t)q0 * nn, dids + (size_t)q0 * nn,
                                                                 dstatus.p, 2, dbig.p, dbig.p + qbatch);
    LSQ_CUDA(cudaGetLastError());
    timer.mark("topk");
    if (timer.on) {   // measurement aid: how many products did the filter run with?
      g_last_products = 0;
      if (use_tc) {
        const char* e = getenv("LSQ_B200_ADC_PASSES");
        g_last_products = (e != nullptr) ? ((atoi(e) == 1) ? 1 : 2) : 2;
        if (e == nullptr && two_stage) LSQ_CUDA(cudaMemcpyAsync(&g_last_products, dnpass.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      }
    }
    LSQ_CUDA(cudaMemcpyAsync(hstatus.data(), dstatus.p, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    LSQ_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < nb; i++)
      if (hstatus[i] != ST_OK) redo.push_back((int)(q0 + i));
  }

  if (!redo.empty()) {
    // thresholds that missed (or overflowed): exhaustive re-run of just those queries
    const int nr = (int)redo.size();
    DevBuf<int> drows;
    DevBuf<float> dqsel;
    LSQ_CUDA(drows.alloc(nr));
    LSQ_CUDA(dqsel.alloc((size_t)nr * d));
    LSQ_CUDA(cudaMemcpyAsync(drows.p, redo.data(), (size_t)nr * sizeof(int), cudaMemcpyHostToDevice, st));
    note_launch();
    gather_rows_kernel<<<(unsigned)ceil_div((int64_t)nr * d, 256), 256, 0, st>>>(dqueries, drows.p, nr, d, dqsel.p);
    LSQ_CUDA(cudaGetLastError());
    LSQ_TRY(scan_exhaustive(S, dqsel.p, nr, drows.p));
  }
  return LSQ_OK;
}

// host-pointer front end shared by the four exported symbols.  With several bound devices the QUERIES are
// partitioned (splitarray rule), codes / norms / codebooks are replicated (12-20 MB per million vectors): the
// per-query results are disjoint, so there is no merge and no communication.
static int linscan_host(float* dists, int32_t* ids, const unsigned char* codes, const float* queries,
                        const float* codebooks, size_t cb_floats, const float* dbnorms, int nq, int64_t n, int m,
                        int d, int lut_kind, int subdim, int nn) {
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM, "linscan: m must be in 1..16");
  LSQ_CHECK_ARG(nq >= 0 && n >= 0 && d >= 1 && nn >= 0, "linscan: bad sizes");
  LSQ_TRY(rt_ensure_init());
  const int k = rt_devices_for(nq, 32);
  return rt_parallel(k, [&](int r) -> int {
    LSQ_TRY(rt_bind(r));
    const cudaStream_t st = rt_ctx(r).st;
    int64_t qlo = 0, qhi = nq;
    lsq_splitarray(nq, k, r, &qlo, &qhi);
    const int nql = (int)(qhi - qlo);
    DevBuf<uint8_t> dcodes;
    DevBuf<float> dq, dcb, dnorm, dd;
    DevBuf<int32_t> di;
    LSQ_CUDA(dcodes.alloc((size_t)n * m));
    LSQ_CUDA(dq.alloc((size_t)nql * d));
    LSQ_CUDA(dcb.alloc(cb_floats));
    LSQ_CUDA(dd.alloc((size_t)nql * nn));
    LSQ_CUDA(di.alloc((size_t)nql * nn));
    LSQ_TRY(rt_h2d(dcodes.p, codes, (size_t)n * m, st));
    LSQ_TRY(rt_h2d(dq.p, queries + (size_t)qlo * d, (size_t)nql * d * 4, st));
    LSQ_CUDA(cudaMemcpyAsync(dcb.p, codebooks, cb_floats * 4, cudaMemcpyHostToDevice, st));
    if (lut_kind == LUT_LSQ) {
      LSQ_CUDA(dnorm.alloc(n));
      LSQ_TRY(rt_h2d(dnorm.p, dbnorms, (size_t)n * 4, st));
    }
    int rc = linscan_device(dcodes.p, n, m, dq.p, nql, d, dcb.p, dnorm.p, lut_kind, subdim, nn, dd.p, di.p, st);
    if (rc == LSQ_OK && nql > 0 && nn > 0) {
      cudaError_t e = cudaMemcpyAsync(dists + (size_t)qlo * nn, dd.p, (size_t)nql * nn * 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(ids + (size_t)qlo * nn, di.p, (size_t)nql * nn * 4, cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) rc = cuda_fail(e, "linscan result copy", __FILE__, __LINE__);
    }
    cudaStreamSynchronize(st);  // also on the error path: the buffers above are about to be freed
    return rc;
  });
}

}  // namespace lsq

using namespace lsq;

extern "C" {

int lsq_linscan_lsq(float* dists, int* idx, const unsigned char* codes, const float* queries,
                    const float* codebooks, const float* dbnorms, int nqueries, int ncodes, int m, int h, int d,
                    int nn) {
  LSQ_CHECK_ARG(h == LSQ_H, "linscan: h must be 256");
  return linscan_host(dists, idx, codes, queries, codebooks, (size_t)m * h * d, dbnorms, nqueries, ncodes, m, d,
                      LUT_LSQ, 0, nn);
}

int lsq_linscan_pq(float* dists, unsigned int* res, const unsigned char* codes, const float* centers,
                   const float* queries, int N, unsigned int NQ, int B, int K, int dim1codes, int dim1queries,
                   int subdim) {
  const int m = B / 8;
  LSQ_CHECK_ARG(B % 8 == 0 && m == dim1codes, "linscan_pq: expected B = 8*dim1codes (one byte per codebook)");
  return linscan_host(dists, reinterpret_cast<int32_t*>(res), codes, queries, centers, (size_t)m * LSQ_H * subdim,
                      nullptr, (int)NQ, N, m, dim1queries, LUT_PQ, subdim, K);
}

static void die(const char* fn) {
  fprintf(stderr, "liblsq_b200: %s failed: %s\n", fn, lsq_last_error());
  abort();
}

// The reference symbols return void and have no error channel (Linscan.jl:19-23, 63-69): fail loudly.
void linscan_aqd_query_extra_byte(float* dists, int* idx, unsigned char* codes, float* queries, float* codebooks,
                                  float* dbnorms, int nqueries, int ncodes, int m, int h, int d, int nn) {
  if (lsq_linscan_lsq(dists, idx, codes, queries, codebooks, dbnorms, nqueries, ncodes, m, h, d, nn) != LSQ_OK)
    die("linscan_aqd_query_extra_byte");
}

void linscan_aqd_query(float* dists, unsigned int* res, unsigned char* codes, float* centers, float* queries, int N,
                       unsigned int NQ, int B, int K, int dim1codes, int dim1queries, int subdim) {
  if (lsq_linscan_pq(dists, res, codes, centers, queries, N, NQ, B, K, dim1codes, dim1queries, subdim) != LSQ_OK)
    die("linscan_aqd_query");
}

// Measurement aid (with LSQ_B200_ADC_TIMING set): device time of phase i of the calling thread's most recent
// linscan call; name (may be NULL) receives a pointer valid until the next call.  Returns the number of phases.
int lsq_linscan_last_phases(int i, float* ms, const char** name) {
  if (i == -2) return g_last_products;   // products of the tensor-core filter in that call (0: lookup scan)
  if (i >= 0 && i < (int)g_last_phases.size()) {
    if (ms) *ms = g_last_phases[i].second;
    if (name) *name = g_last_phases[i].first.c_str();
  }
  return (int)g_last_phases.size();
}

int lsq_dev_linscan(const uint8_t* dcodes, int64_t n, int m, const float* dqueries, int nq, int d,
                    const float* dcodebooks, const float* dbnorms, int lut_kind, int subdim, int nn, float* ddists,
                    int32_t* dids, void* stream) {
  set_alloc_stream((cudaStream_t)stream);
  return linscan_device(dcodes, n, m, dqueries, nq, d, dcodebooks, dbnorms, lut_kind, subdim, nn, ddists, dids,
                        (cudaStream_t)stream);
}

}  // extern "C"
