// adc_tc.cu — tensor-core PREFILTER for the exact ADC scan (linscan_lsq, linscan_aqd_pairwise_byte.cpp:14-93).
//
// The reference scores every (query, base vector) pair with m table lookups.  On B200 that scan is bound by
// shared-memory wavefronts and, for m >= 9, by shared-memory CAPACITY (only 14 query LUTs of 16 KB fit one SM at
// m = 16).  But the quantity the lookups add up is an inner product:
//     dist(q, v) = dbnorm[v] - 2 <q, xhat_v>,   xhat_v = sum_k C_k[code_vk]
// so the pass that decides WHICH pairs can be among a query's nn nearest is GEMM-shaped, and the result only has
// to be bit-exact for the pairs that survive.  This file therefore splits the main pass of linscan.cu in two:
//
//   1. adc_decode_kernel   xhat_v in fp32, split into bf16 hi + lo, written as ready-made UMMA operand images
//                          (K-major, no swizzle; one block per 128 base vectors), plus max ||xhat||.  The hi image
//                          carries 16 extra K elements per vector: the 3-term bf16 split of -dbnorm/2 and three 1s.
//   2. adc_filter_kernel   tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM): hi(q).hi(x), preceded by hi(q).lo(x)
//                          when the sample says the wider one-product margin would cost too many survivors.  A CTA
//                          keeps 256 queries (two 128-row A tiles) resident in TENSOR MEMORY for its whole life —
//                          each row extended by three 1s and the 3-term split of (tau_q + margin_q)/2 — and streams
//                          base tiles through a 3-stage TMA ring.  With the extra K elements the accumulator IS
//                              <hi(q), x> - dbnorm/2 + (tau_q + margin_q)/2  =  ((tau_q + margin_q) - dist) / 2,
//                          so "this pair may be among the nn nearest" is its SIGN BIT: the epilogue warps gather
//                          sign bits straight out of TMEM (one funnel shift per pair, no load, no compare) and only
//                          the ids of the pairs that pass are written.
//   3. adc_rescore_kernel  the survivors (a few thousand per query) are scored EXACTLY like the reference —
//                          ((0 + LUT_0[c0]) + LUT_1[c1]) + ... + dbnorm, fp32 adds in that order — and appended as
//                          keys iff dist <= tau_q.
//
// The keys that reach the top-k kernels are therefore the same SET the thresholded scan produces, provided
// margin_q bounds |filter value - reference value|.  Bound used (u = 2^-24; all norms Euclidean):
//     reference:  |d_ref - D| <= (d+m+2) u (2 ||q|| m cmax + max|dbnorm|)        (fp32 chains, Cauchy-Schwarz)
//     dropped lo(q) = q - bf16(q):      2 ||lo(q)|| xmax          (||lo(q)|| computed exactly per query; 0 for
//                                                                  8-bit data such as SIFT descriptors)
//     one-product mode drops hi(q).lo(x) as well:  2 ||q|| max_v ||lo(xhat_v)||   (computed exactly at decode)
//     bf16 hi+lo split of xhat (2^-16), 3-term splits of the scalars (2^-24), fp32 accumulation in the tensor
//     core:       2^-13 (2 ||q|| xmax + max|dbnorm| + |tau|)    — measured 1e-6 .. 3e-6 relative
//                 (tests/test_gpu_adc_tc.py), i.e. > 40x headroom
// A wider margin only lets a few more pairs through to the exact rescoring; it never changes the result.
// Anything unusual (NaN, candidate overflow, fewer than nn survivors) ends on linscan.cu's exhaustive path,
// exactly as before.  The thresholds tau_q themselves come from the same kernel run on a strided sample of the
// base set: its values on a 1/8 sub-sample give a coarse bound, the sample positions below that bound are listed
// (filter mode) and scored exactly, and tau_q is the exact r-th smallest sample distance (adc_sample_tau_kernel; any
// tau is valid, see linscan.cu).  tests/: bit-identical ids and distances against the reference's
// own .so and against the lookup scan (LSQ_B200_ADC=scan).
//
// PQ / OPQ tables (linscan_aqd.cpp:37-102: dist = sum_k sum_s sqr(c_k[s] - q[k*subdim + s])) take the same path:
// dist = ||q||^2 - 2 <q, xhat> + ||xhat||^2 with xhat the CONCATENATION of the sub-codewords, so ||xhat||^2 stands
// where dbnorm stood and ||q||^2 goes into the threshold term; the reference's sums of squares are good to
// (subdim + m + 4) u (||q|| + max||xhat||)^2, which replaces the first line of the bound above.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "adc_tc.cuh"

namespace lsq {

constexpr int AT_M = 128;         // queries per A tile (UMMA M)
constexpr int AT_NA = 2;          // A tiles resident in tensor memory per CTA
constexpr int AT_N = 128;         // base vectors per tile (UMMA N)
constexpr int AT_STAGES = 6;      // max depth of the shared-memory ring of base tiles (as many as fit ~200 KB: 5 x 36 KB in one-product mode)
// epilogue warps: SPLIT warp sets share the 128 columns of a product (4 warps per set = the 4 lane quadrants)
__host__ __device__ constexpr int at_epi_warps(int split) { return 4 * AT_NA * split; }
__host__ __device__ constexpr int at_threads(int split) { return 32 * (at_epi_warps(split) + 2); }
// K-major, no swizzle, 2-byte elements: core matrix = 8 rows x 8 elements (16 B per row, 128 B, contiguous);
// 8-row groups are SBO apart, K-adjacent core matrices LBO apart
constexpr uint32_t AT_SBO = 128u;
constexpr uint32_t AT_LBO = (AT_N / 8) * 128u;
// tensor-memory columns: A tile a at 72 a (d <= 128 bf16 of hi(q) = 64 columns, then 8 columns = 16 extra K
// elements); two accumulators of 128 columns at 144 + 128 b, used alternately by the (base tile, A tile) products
constexpr int AT_NACC = 2;
constexpr uint32_t AT_COL_A = 0u, AT_A_STRIDE = 72u, AT_COL_ACC = 144u;
constexpr float AT_BIG = 1.5e38f;   // "never passes": -AT_BIG as the folded norm of padding vectors / threshold of padding rows

// operand images of one base tile: hi part = d/8 + 2 K chunks (the last two carry the 16 extra K elements), lo part = d/8
__host__ __device__ inline uint32_t at_hi_bytes(int d) { return (uint32_t)(d / 8 + 2) * AT_LBO; }
__host__ __device__ inline uint32_t at_lo_bytes(int d) { return (uint32_t)(d / 8) * AT_LBO; }
__host__ __device__ inline uint32_t at_tile_bytes(int d) { return at_hi_bytes(d) + at_lo_bytes(d); }

// x = a + b + c with a, b, c bf16 (24 significant bits: exact for normal fp32 values)
__device__ __forceinline__ void bf16_split3(float x, uint32_t& a, uint32_t& b, uint32_t& c) {
  const __nv_bfloat16 ha = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(ha);
  const __nv_bfloat16 hb = __float2bfloat16_rn(r1);
  const __nv_bfloat16 hc = __float2bfloat16_rn(r1 - __bfloat162float(hb));
  a = (uint32_t)__bfloat16_as_ushort(ha);
  b = (uint32_t)__bfloat16_as_ushort(hb);
  c = (uint32_t)__bfloat16_as_ushort(hc);
}

__device__ __forceinline__ void bf16_split(float x, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = (uint32_t)__bfloat16_as_ushort(h);
  lo = (uint32_t)__bfloat16_as_ushort(l);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// 0. largest squared codeword norm (for the reference-rounding part of the margin)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adc_cbnorm_kernel(const float* __restrict__ C, int rows, int d, AdcStats* stats) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  float s = 0.0f;
  if (r < rows)
    for (int k = 0; k < d; k++) { const float c = C[(size_t)r * d + k]; s = fmaf(c, c, s); }
  s = warp_max(s);
  if ((threadIdx.x & 31) == 0) atomicMax(&stats->cmax2_bits, __float_as_uint(s));
}

// ------------------------------------------------------------------------------------------------
// 1. decode + split.  CTA = one tile of 128 base vectors, 8 warps x 16 vectors.  A warp sums TWO vectors at a
//    time (rows 2i and 2i+1 of the tile, one per half-warp): lane = 32-byte piece of the d-float row, so every
//    codeword row is one coalesced 512-byte load per half-warp, the per-vector address arithmetic is shared by two
//    vectors, and the two half-warps write ADJACENT 16-byte rows of the same core matrices (whole 32-byte sectors;
//    the first version gathered 32-byte pieces per lane and was bound by the LSU: 1.8 ms at m = 16, the second
//    wrote half sectors: 0.74 ms).  Vector `sidx` of the image is base vector sidx * stride (stride > 1: the
//    strided sample for the thresholds).  The lane with K chunk 0 adds the 16 extra K elements of its row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adc_decode_kernel(const uint8_t* __restrict__ codes, int m,
                                                         const float* __restrict__ C, int d,
                                                         const float* __restrict__ norms, unsigned char* __restrict__ img,
                                                         AdcStats* stats, int64_t count, int64_t stride, int subdim) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, kc = lane & 15;   // which vector of the pair, which K chunk of 8 elements
  const int64_t tile = blockIdx.x;
  const bool lane_on = kc < d / 8;
  const uint32_t hi_bytes = at_hi_bytes(d);
  unsigned char* base = img + (size_t)tile * at_tile_bytes(d);
  const bool vec16 = (m == 16) && ((reinterpret_cast<uintptr_t>(codes) & 15) == 0);
  const bool vec8 = (m == 8) && ((reinterpret_cast<uintptr_t>(codes) & 7) == 0);
  uint32_t xb = 0u, nb = 0u, lb = 0u;   // running maxima (bit patterns of non-negative floats)
#pragma unroll 2
  for (int i = 0; i < 8; i++) {
    const int r = warp * 16 + 2 * i + half;
    const int64_t sidx = tile * AT_N + r;
    const bool valid = sidx < count;
    const int64_t v = sidx * stride;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; e++) acc[e] = 0.0f;
    if (valid) {
      const uint8_t* cp = codes + (size_t)v * m;   // the same bytes in the 16 lanes of a half-warp: broadcast loads
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (vec16) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(cp));
        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
      } else if (vec8) {
        const uint2 x = __ldg(reinterpret_cast<const uint2*>(cp));
        w[0] = x.x; w[1] = x.y;
      } else {
#pragma unroll
        for (int k = 0; k < LSQ_MAXM; k++)
          if (k < m) w[k >> 2] |= (uint32_t)cp[k] << (8 * (k & 3));
      }
      if (lane_on && subdim == 0) {   // LSQ: xhat = sum of m full-dimensional codewords
#pragma unroll
        for (int k = 0; k < LSQ_MAXM; k++) {
          if (k < m) {
            const uint32_t c = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            const float4* row = reinterpret_cast<const float4*>(C + ((size_t)k * LSQ_H + c) * d) + 2 * kc;
            const float4 a = __ldg(row), b = __ldg(row + 1);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
          }
        }
      } else if (lane_on) {           // PQ: xhat = concatenation of m sub-codewords of `subdim` elements
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const int i = 8 * kc + e, k = i / subdim, t = i - k * subdim;
          acc[e] = __ldg(C + ((size_t)k * LSQ_H + cp[k]) * subdim + t);
        }
      }
    }
    float n2 = 0.0f, lo2 = 0.0f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      uint32_t h0, l0, h1, l1;
      bf16_split(acc[2 * e], h0, l0);
      bf16_split(acc[2 * e + 1], h1, l1);
      hi[e] = h0 | (h1 << 16);   // element k in the low half, k + 1 in the high half (little endian)
      lo[e] = l0 | (l1 << 16);
      // ||x - hi(x)||^2 bounds the product a one-pass filter drops; x - hi(x) is exact in fp32
      const float e0 = acc[2 * e] - __uint_as_float(h0 << 16), e1 = acc[2 * e + 1] - __uint_as_float(h1 << 16);
      n2 = fmaf(acc[2 * e], acc[2 * e], fmaf(acc[2 * e + 1], acc[2 * e + 1], n2));
      lo2 = fmaf(e0, e0, fmaf(e1, e1, lo2));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {   // over the 16 lanes of the half-warp
      n2 += __shfl_xor_sync(0xFFFFFFFFu, n2, o);
      lo2 += __shfl_xor_sync(0xFFFFFFFFu, lo2, o);
    }
    const uint32_t row_off = (uint32_t)(r >> 3) * AT_SBO + (uint32_t)(r & 7) * 16u;
    if (lane_on) {
      const uint32_t off = (uint32_t)kc * AT_LBO + row_off;   // elements 8 kc .. 8 kc + 7 of row r: one 16-byte core-matrix row
      *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(base + hi_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (kc == 0) {
      float half_neg = -AT_BIG;   // padding columns of the last tile never pass the filter
      if (valid) {
        const float nv = (norms != nullptr) ? norms[v] : n2;   // PQ: the "norm" of the expansion is ||xhat||^2 itself
        half_neg = -0.5f * nv;
        nb = max(nb, __float_as_uint(fabsf(nv)));   // a NaN is the largest pattern: it survives and poisons the margin
        xb = max(xb, __float_as_uint(n2));
        lb = max(lb, __float_as_uint(lo2));
      }
      // extra K elements d .. d+15 of the hi image: -dbnorm/2 as three bf16 terms, three 1s (they meet the
      // threshold terms of the query rows), zeros
      uint32_t ca, cb, cc;
      bf16_split3(half_neg, ca, cb, cc);
      const uint32_t one = 0x3F80u;
      *reinterpret_cast<uint4*>(base + (uint32_t)(d / 8) * AT_LBO + row_off) = make_uint4(ca | (cb << 16), cc | (one << 16), one | (one << 16), 0u);
      *reinterpret_cast<uint4*>(base + (uint32_t)(d / 8 + 1) * AT_LBO + row_off) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (stats != nullptr && kc == 0) {
    atomicMax(&stats->xmax2_bits, xb);
    atomicMax(&stats->nmax_bits, nb);
    atomicMax(&stats->xlo2_bits, lb);
  }
}

// ------------------------------------------------------------------------------------------------
// 1b. exact LUT rows for the rescoring: lutq[q][j] = -(2 q).c_j accumulated like the reference
//     (t -= (2*q[k])*c[k], k ascending, separate multiply and subtract, linscan_aqd_pairwise_byte.cpp:42-48).
//     Register tile 4 queries x 8 rows per thread, both operands of the 64 x 128 block tile in shared memory,
//     k-major, so every k step is 3 LDS.128 for 64 multiply-subtract pairs.
// ------------------------------------------------------------------------------------------------
constexpr int LR_Q = 64, LR_J = 128;

__global__ void __launch_bounds__(256) adc_lut_rows_kernel(const float* __restrict__ queries, int nq, int d,
                                                           const float* __restrict__ cb, int rows,
                                                           float* __restrict__ lutq) {
  extern __shared__ __align__(16) float lr_smem[];
  float* qs = lr_smem;                 // [d][LR_Q]   (2 * q)
  float* cs = lr_smem + d * LR_Q;      // [d][LR_J]
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * LR_Q, j0 = blockIdx.y * LR_J;
  const int d4 = d / 4;
  for (int e = tid; e < LR_Q * d4; e += 256) {
    const int qq = e % LR_Q, k4 = e / LR_Q;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + qq < nq) x = __ldg(reinterpret_cast<const float4*>(queries + (size_t)(q0 + qq) * d) + k4);
    qs[(4 * k4 + 0) * LR_Q + qq] = __fmul_rn(2.0f, x.x);   // the factor (2*q[k]) of :45-47, exact
    qs[(4 * k4 + 1) * LR_Q + qq] = __fmul_rn(2.0f, x.y);
    qs[(4 * k4 + 2) * LR_Q + qq] = __fmul_rn(2.0f, x.z);
    qs[(4 * k4 + 3) * LR_Q + qq] = __fmul_rn(2.0f, x.w);
  }
  for (int e = tid; e < LR_J * d4; e += 256) {
    const int jj = e % LR_J, k4 = e / LR_J;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j0 + jj < rows) x = __ldg(reinterpret_cast<const float4*>(cb + (size_t)(j0 + jj) * d) + k4);
    cs[(4 * k4 + 0) * LR_J + jj] = x.x;
    cs[(4 * k4 + 1) * LR_J + jj] = x.y;
    cs[(4 * k4 + 2) * LR_J + jj] = x.z;
    cs[(4 * k4 + 3) * LR_J + jj] = x.w;
  }
  __syncthreads();
  const int tq = tid & 15, tj = tid >> 4;
  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 8; b++) acc[a][b] = 0.0f;
#pragma unroll 4
  for (int k = 0; k < d; k++) {
    const float4 qv = *reinterpret_cast<const float4*>(qs + k * LR_Q + 4 * tq);
    const float4 c0 = *reinterpret_cast<const float4*>(cs + k * LR_J + 8 * tj);
    const float4 c1 = *reinterpret_cast<const float4*>(cs + k * LR_J + 8 * tj + 4);
    const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
    const float ca[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 8; b++) acc[a][b] = __fsub_rn(acc[a][b], __fmul_rn(qa[a], ca[b]));
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int q = q0 + 4 * tq + a;
    if (q < nq && j0 + 8 * tj < rows) {
      float4* dst = reinterpret_cast<float4*>(lutq + (size_t)q * rows + j0 + 8 * tj);
      dst[0] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      dst[1] = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
    }
  }
}

// PQ / OPQ tables: lutq[q][k*256 + r] = sum_s sqr(c[s] - q[k*subdim + s]), s ascending, separate subtract, multiply
// and add (linscan_aqd.cpp:66-74).  grid (queries, codebooks), thread = table row.
__global__ void __launch_bounds__(LSQ_H) adc_lut_rows_pq_kernel(const float* __restrict__ queries, int qstride, int subdim,
                                                                const float* __restrict__ centers, int m,
                                                                float* __restrict__ lutq) {
  const int q = blockIdx.x, k = blockIdx.y, r = threadIdx.x;
  const float* c = centers + ((size_t)k * LSQ_H + r) * subdim;
  const float* qv = queries + (size_t)q * qstride + k * subdim;
  float t = 0.0f;
  for (int s = 0; s < subdim; s++) {
    const float df = __fsub_rn(__ldg(c + s), __ldg(qv + s));
    t = __fadd_rn(t, __fmul_rn(df, df));
  }
  lutq[((size_t)q * m + k) * LSQ_H + r] = t;
}

// ------------------------------------------------------------------------------------------------
// 2. the filter GEMM
// ------------------------------------------------------------------------------------------------
struct AdcFilterParams {
  const float* queries;      // [nq][d]
  const unsigned char* img;  // [ntiles][at_tile_bytes(d)]
  const float* tau;          // [nq] (threshold_kernel's layout with 32-query tiles)
  const AdcStats* stats;
  uint32_t* candidx;         // [nq][ccap] 0-based base indices that passed
  int* ccnt;                 // [nq]
  float* dbg;                // values mode: [nq][dbg_ld] filter values (tests)
  uint32_t* sbuf;            // sample mode: ordered filter values, threshold_kernel's layout [(q/32 * scount + t) * 32 + q%32]
  int64_t n, ntiles, ccap, dbg_ld, scount;
  const int* npass_dev;      // optional: number of products chosen on the device (adc_choose_passes_kernel); overrides npass
  int nq, d, m, npass, nstages, qstride, pq;
  uint32_t stage_stride;     // bytes between ring slots (sized for the larger of the two modes)
};

__device__ __forceinline__ void at_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void at_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void at_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem], bf16 operands.  Called by every lane of the issuing warp in uniform control
// flow; one elected lane issues (see unary_tc.cu: keeping the election inside the asm keeps the descriptors in
// uniform registers).
__device__ __forceinline__ void at_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void at_commit_elected(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

// bit i of the result = SIGN bit of accumulator column i (set = the pair does NOT pass).  One funnel shift per
// value, four independent chains of eight.
__device__ __forceinline__ uint32_t at_signs32(const uint32_t (&r)[32]) {
  uint32_t s[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    uint32_t x = 0u;
#pragma unroll
    for (int i = 7; i >= 0; i--) x = __funnelshift_l(r[8 * c + i], x, 1);   // x = (x << 1) | sign(r)
    s[c] = x;
  }
  return (s[0] | (s[1] << 8)) | ((s[2] << 16) | (s[3] << 24));
}

enum { AT_FILTER = 0, AT_SAMPLE = 1, AT_VALUES = 2 };   // what the epilogue does with the products

template <int SPLIT, int MODE>
__global__ void __launch_bounds__(at_threads(SPLIT), 1) adc_filter_kernel(const __grid_constant__ AdcFilterParams p) {
  constexpr int EPI_WARPS = at_epi_warps(SPLIT), W_TMA = EPI_WARPS, W_MMA = EPI_WARPS + 1;
  constexpr int HPW = 2 / SPLIT;   // 64-column halves of a product drained by one warp
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar_full[AT_STAGES], bar_empty[AT_STAGES], bar_acc_full[AT_NACC], bar_acc_empty[AT_NACC];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = p.d, ksteps = d / 16;
  const int npass = (p.npass_dev != nullptr) ? *p.npass_dev : p.npass;
  const uint32_t hi_bytes = at_hi_bytes(d), tile_bytes = at_tile_bytes(d);
  const uint32_t load_bytes = (npass == 2) ? tile_bytes : hi_bytes;   // one product: the lo image stays in HBM
  const int nstages = p.nstages;
  const int qbase = blockIdx.x * (AT_NA * AT_M);
  const int na = (p.nq - qbase > AT_M) ? 2 : 1;   // A tiles in use
  // this CTA's slice of the base tiles
  const int64_t per = (p.ntiles + gridDim.y - 1) / gridDim.y;
  const int64_t t_lo = (int64_t)blockIdx.y * per;
  const int64_t t_hi = (t_lo + per < p.ntiles) ? (t_lo + per) : p.ntiles;
  const int64_t my_tiles = (t_hi > t_lo) ? (t_hi - t_lo) : 0;

  if (tid == 0) {
    for (int s = 0; s < AT_STAGES; s++) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    for (int b = 0; b < AT_NACC; b++) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_empty[b], 4 * SPLIT); }
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_slot;

  // ---- stationary operand -> tensor memory: thread = query row (lane of its warp's quadrant); hi(q) in bf16,
  //      two K elements per 32-bit column, then the 16 extra K elements: 1, 1, 1, the three bf16 terms of
  //      (tau + margin) / 2, zeros.  Done once per CTA by the first warp set. ----
  const int a_mine = (warp >> 2) & 1, quad = warp & 3, half0 = (warp >> 3) * HPW;
  const int row = a_mine * AT_M + quad * 32 + lane;
  const int q = qbase + row;
  const bool is_epi = warp < EPI_WARPS && a_mine < na;
  const bool q_valid = is_epi && (q < p.nq);
  if (warp < 8 && a_mine < na) {
    const float* qrow = p.queries + (size_t)(q_valid ? q : 0) * p.qstride;
    const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16) + AT_COL_A + (uint32_t)a_mine * AT_A_STRIDE;
    float qn2 = 0.0f, ql2 = 0.0f;   // ||q||^2 and ||q - hi(q)||^2
    for (int k0 = 0; k0 < d; k0 += 64) {
      uint32_t hi[32];
#pragma unroll
      for (int i4 = 0; i4 < 16; i4++) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q_valid && k0 + 4 * i4 < d) x = __ldg(reinterpret_cast<const float4*>(qrow + k0) + i4);
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y);
        const __nv_bfloat16 h2 = __float2bfloat16_rn(x.z), h3 = __float2bfloat16_rn(x.w);
        hi[2 * i4] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        hi[2 * i4 + 1] = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
        const float r0 = x.x - __bfloat162float(h0), r1 = x.y - __bfloat162float(h1);   // exact in fp32
        const float r2 = x.z - __bfloat162float(h2), r3 = x.w - __bfloat162float(h3);
        ql2 = fmaf(r0, r0, ql2); ql2 = fmaf(r1, r1, ql2); ql2 = fmaf(r2, r2, ql2); ql2 = fmaf(r3, r3, ql2);
        qn2 = fmaf(x.x, x.x, qn2); qn2 = fmaf(x.y, x.y, qn2); qn2 = fmaf(x.z, x.z, qn2); qn2 = fmaf(x.w, x.w, qn2);
      }
      if (d - k0 >= 64) {
        at_st32(lane_base + (uint32_t)(k0 >> 1), hi);
      } else {   // d % 64 != 0: the tail in 8-column pieces, so that the extra columns right behind it stay free
        for (int c0 = 0; c0 < (d - k0) / 2; c0 += 8) {
          uint32_t part[8];
#pragma unroll
          for (int i = 0; i < 8; i++) part[i] = hi[c0 + i];
          at_st8(lane_base + (uint32_t)(k0 >> 1) + (uint32_t)c0, part);
        }
      }
    }
    // threshold: the filter accumulates <hi(q), x> - dbnorm/2 + half_thr, which is >= 0 iff dist <= tau + margin
    float half_thr = (MODE == AT_FILTER) ? -AT_BIG : 0.0f;   // rows without a query never pass; sample / values: plain values
    if (MODE != AT_FILTER && p.pq && q_valid) half_thr = -0.5f * qn2;   // PQ: the ||q||^2 term of the expansion
    if (q_valid && MODE == AT_FILTER) {
      const float tau = p.tau[q];
      const float qn = sqrtf(qn2) * 1.00001f, ql = sqrtf(ql2) * 1.00001f;
      const float xmax = sqrtf(__uint_as_float(p.stats->xmax2_bits)) * 1.00001f;
      const float cmax = sqrtf(__uint_as_float(p.stats->cmax2_bits)) * 1.00001f;
      const float nmax = __uint_as_float(p.stats->nmax_bits);
      // one product only (hi(q).hi(x)): the dropped hi(q).lo(x) is bounded by 2 ||q|| max||lo(x)||, exactly like lo(q)
      const float xlo = (npass == 1) ? sqrtf(__uint_as_float(p.stats->xlo2_bits)) * 1.00001f : 0.0f;
      const float eps_f = 1.0f / 8192.0f;                                   // 2^-13, see the header
      const float eps_r = 2.0f * (float)(d + p.m + 2) * 5.9604645e-8f;      // 2 (d+m+2) u
      float margin = 2.0f * ql * xmax + 2.0f * qn * xlo + eps_f * 2.0f * qn * xmax;
      if (p.pq) {
        // PQ tables: dist = sum_i (xhat_i - q_i)^2 = ||q||^2 - 2 <q, xhat> + ||xhat||^2; the reference's fp32 sums of
        // squares are good to (subdim + m + 4) u relative to the distance, which is at most (||q|| + max||xhat||)^2;
        // ||q||^2 and ||xhat||^2 are fp32 sums here (d u relative each, far inside eps_f)
        const float span = (qn + xmax) * (qn + xmax);
        margin += eps_r * span + eps_f * (qn * qn + xmax * xmax);
        margin += 1.01f * eps_f * (fabsf(tau) + margin);
        half_thr = 0.5f * (tau + margin - qn2);
      } else {
        margin += eps_r * (2.0f * qn * (float)p.m * cmax + nmax);
        margin += 1.01f * eps_f * (nmax + fabsf(tau) + margin);
        half_thr = 0.5f * (tau + margin);
      }
      if (!(fabsf(half_thr) < AT_BIG)) half_thr = __uint_as_float(0x7FC00000u);   // NaN / overflow: poison the row -> exhaustive re-run
    }
    {
      uint32_t ta, tb, tc, ext[8];
      bf16_split3(half_thr, ta, tb, tc);
      const uint32_t one = 0x3F80u;
      ext[0] = one | (one << 16); ext[1] = one | (ta << 16); ext[2] = tb | (tc << 16);
#pragma unroll
      for (int i = 3; i < 8; i++) ext[i] = 0u;
      at_st8(lane_base + (uint32_t)(d >> 1), ext);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == W_TMA) {
    // ===================== TMA: base tiles (operand images), up to AT_STAGES ahead =====================
    if (lane == 0) {
      int s = 0;            // ring slot and the parity of its current use, kept incrementally (no division per tile)
      uint32_t par = 0u;
      for (int64_t t = 0; t < my_tiles; t++) {
        mbar_wait(&bar_empty[s], par ^ 1u);
        bulk_load_issue(smem_raw + (size_t)s * p.stage_stride, p.img + (size_t)(t_lo + t) * tile_bytes, load_bytes, &bar_full[s]);
        if (++s == nstages) { s = 0; par ^= 1u; }
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (whole warp, uniform; one elected lane issues) =====================
    // instruction descriptor: D = f32 (1 << 4), A = B = bf16 (1 << 7, 1 << 10), K-major both, N >> 3 at bit 17,
    // M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_N >> 3) << 17) | ((uint32_t)(AT_M >> 4) << 24);
    const uint32_t sB_u = smem_u32(smem_raw);
    const uint64_t desc_hi = (uint64_t)((AT_SBO >> 4) | (1u << 14)) << 32;   // SBO, descriptor version 1
    const uint32_t desc_lbo = (AT_LBO >> 4) << 16;
    const uint32_t kstep_enc = (2u * AT_LBO) >> 4;                           // one MMA consumes K = 16 = 2 core matrices
    int64_t j = 0;        // product index: (base tile t, A tile a) -> accumulator j % 2, its (j / 2)-th use
    int s = 0;            // ring slot and the parity of its current use
    uint32_t par = 0u;
    for (int64_t t = 0; t < my_tiles; t++) {
      mbar_wait(&bar_full[s], par);
      const uint32_t stage_u = sB_u + (uint32_t)s * p.stage_stride;
      const uint32_t lo_hiX = desc_lbo | (stage_u >> 4), lo_loX = desc_lbo | ((stage_u + hi_bytes) >> 4);
      for (int a = 0; a < na; a++) {
        const int b = (int)(j % AT_NACC);
        const uint32_t bu = (uint32_t)((j / AT_NACC) & 1);
        mbar_wait(&bar_acc_empty[b], bu ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dcol = tmem + AT_COL_ACC + (uint32_t)b * AT_N;
        const uint32_t a_hi = tmem + AT_COL_A + (uint32_t)a * AT_A_STRIDE;
        uint32_t accum = 0u;
        if (npass == 2) {
#pragma unroll 4
          for (int k = 0; k < ksteps; k++) {   // hi(q).lo(x)  (small terms first)
            at_mma_bf16_ts(dcol, a_hi + (uint32_t)(8 * k), desc_hi | (lo_loX + (uint32_t)k * kstep_enc), idesc, accum);
            accum = 1u;
          }
        }
#pragma unroll 3
        for (int k = 0; k <= ksteps; k++) {    // hi(q).hi(x), and the extra K step: - dbnorm/2 + (tau + margin)/2
          at_mma_bf16_ts(dcol, a_hi + (uint32_t)(8 * k), desc_hi | (lo_hiX + (uint32_t)k * kstep_enc), idesc, accum);
          accum = 1u;
        }
        if (a == na - 1) at_commit_elected(&bar_empty[s]);   // the stage may be refilled once these MMAs have read it
        at_commit_elected(&bar_acc_full[b]);
        j++;
      }
      if (++s == nstages) { s = 0; par ^= 1u; }
    }
  } else if (is_epi) {
    // ===================== epilogue: 4 * SPLIT warps drain each product of their A tile =====================
    int* my_cnt = p.ccnt + (q_valid ? q : 0);
    uint32_t* my_list = p.candidx + (size_t)(q_valid ? q : 0) * p.ccap;
    uint32_t* my_sbuf = (MODE == AT_SAMPLE && q_valid) ? (p.sbuf + ((size_t)(q >> 5) * p.scount) * 32 + (q & 31)) : nullptr;
    float* my_dbg = (MODE == AT_VALUES && q_valid) ? (p.dbg + (size_t)q * p.dbg_ld) : nullptr;
    // The list position of a tile's hits comes from an atomicAdd whose round trip (~1 us) must not sit in the
    // per-tile dependency chain: the add is issued at the end of tile t and its result is consumed, together
    // with the saved hit masks, after tile t + 1 has been drained.
    uint32_t pmask[2 * HPW];
#pragma unroll
    for (int c = 0; c < 2 * HPW; c++) pmask[c] = 0u;
    int ppos = 0;
    int64_t pv0 = 0;
    const int ccap32 = (int)((p.ccap < 0x7FFFFFFF) ? p.ccap : 0x7FFFFFFF);
    auto flush = [&]() {
      asm volatile("" : "+r"(ppos));   // the first use of the atomic's result stays HERE (not right behind the add)
#pragma unroll
      for (int c = 0; c < 2 * HPW; c++) {
        uint32_t mk = pmask[c];
        while (mk) {
          const int bit = __ffs(mk) - 1;
          mk &= mk - 1u;
          if (ppos < ccap32) my_list[ppos] = (uint32_t)(pv0 + c * 32 + bit);
          ppos++;
        }
      }
    };
    for (int64_t t = 0; t < my_tiles; t++) {
      const int64_t j = t * na + a_mine;   // product index -> accumulator j % 2, its (j / 2)-th use
      const int b = (int)(j % AT_NACC);
      const uint32_t bu = (uint32_t)((j / AT_NACC) & 1);
      const int64_t v0 = (t_lo + t) * AT_N + half0 * 64;   // first base vector of this warp's columns
      mbar_wait(&bar_acc_full[b], bu);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + AT_COL_ACC + (uint32_t)b * AT_N + (uint32_t)(half0 * 64);
      uint32_t mask[2 * HPW];
#pragma unroll
      for (int c = 0; c < 2 * HPW; c++) mask[c] = 0u;
#pragma unroll
      for (int h = 0; h < HPW; h++) {
        uint32_t r0[32], r1[32];
        at_ld32(taddr + (uint32_t)(h * 64), r0);
        at_ld32(taddr + (uint32_t)(h * 64 + 32), r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (h == HPW - 1) {   // every column of this warp is in registers: its share of the accumulator is free
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
        }
        if (MODE == AT_FILTER) {
          mask[2 * h] = ~at_signs32(r0);
          mask[2 * h + 1] = ~at_signs32(r1);
        } else {
          // the accumulator is <hi(q), x> - dbnorm/2: the filter's estimate of the distance is -2 * acc
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const int64_t c0 = v0 + h * 64 + i, c1 = c0 + 32;
            const float d0 = -2.0f * __uint_as_float(r0[i]), d1 = -2.0f * __uint_as_float(r1[i]);
            if (MODE == AT_SAMPLE) {   // ordered for the radix select of threshold_kernel
              if (my_sbuf != nullptr && c0 < p.scount) my_sbuf[(size_t)c0 * 32] = float_to_ordered(d0);
              if (my_sbuf != nullptr && c1 < p.scount) my_sbuf[(size_t)c1 * 32] = float_to_ordered(d1);
            } else if (my_dbg != nullptr) {
              my_dbg[c0] = d0;
              my_dbg[c1] = d1;
            }
          }
        }
      }
      if (MODE != AT_FILTER) continue;
      flush();   // hits of the previous tile: their atomicAdd has had a whole tile to return
      int hits = 0;
#pragma unroll
      for (int c = 0; c < 2 * HPW; c++) { hits += __popc(mask[c]); pmask[c] = mask[c]; }
      pv0 = v0;
      if (hits > 0) ppos = atomicAdd(my_cnt, hits);   // rows without a query carry -AT_BIG: never here
    }
    if (MODE == AT_FILTER) flush();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------
// 3. exact rescoring of the survivors: one CTA per query, its LUT column in shared memory.
//    dist = ((0 + LUT_0[c0]) + LUT_1[c1]) + ... + dbnorm  — the reference's order (:69-73) — appended iff <= tau,
//    i.e. exactly what scan_kernel's MODE_MAIN appends.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adc_rescore_kernel(const uint8_t* __restrict__ codes, int64_t n, int m,
                                                          const float* __restrict__ norms, const float* __restrict__ lutq,
                                                          const float* __restrict__ tau,
                                                          const uint32_t* __restrict__ candidx, const int* __restrict__ ccnt,
                                                          int64_t ccap, unsigned long long* __restrict__ cand,
                                                          int* __restrict__ cnt, int64_t cap, int id_base) {
  __shared__ __align__(16) float lut[LSQ_MAXM * LSQ_H];
  __shared__ int sh_n;
  const int q = blockIdx.x, tid = threadIdx.x;
  const float4* lsrc = reinterpret_cast<const float4*>(lutq + (size_t)q * m * LSQ_H);
  for (int j = tid; j < m * LSQ_H / 4; j += 256) reinterpret_cast<float4*>(lut)[j] = __ldg(lsrc + j);
  if (tid == 0) sh_n = 0;
  __syncthreads();
  const float tq = tau[q];
  const bool vec16 = (reinterpret_cast<uintptr_t>(codes) & 15) == 0, vec8 = (reinterpret_cast<uintptr_t>(codes) & 7) == 0;
  const int64_t c_all = ccnt[q];
  const int64_t c = (c_all < ccap) ? c_all : ccap;
  const uint32_t* list = candidx + (size_t)q * ccap;
  unsigned long long* out = cand + (size_t)q * cap;
  for (int64_t i = tid; i < c; i += 256) {
    const uint32_t v = list[i];
    if ((int64_t)v >= n) continue;
    const uint8_t* cp = codes + (size_t)v * m;
    float acc = 0.0f;
    // the code row in ONE gather (16 / 8 / 4 bytes): the kernel is bound by L1 gather wavefronts, not by arithmetic
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (m == 16 && vec16) {
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(cp));
      w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
    } else if (m == 8 && vec8) {
      const uint2 x = __ldg(reinterpret_cast<const uint2*>(cp));
      w[0] = x.x; w[1] = x.y;
    } else if ((m & 3) == 0) {
#pragma unroll
      for (int k4 = 0; k4 < LSQ_MAXM; k4 += 4)
        if (k4 < m) w[k4 >> 2] = *reinterpret_cast<const uint32_t*>(cp + k4);
    } else {
#pragma unroll
      for (int k = 0; k < LSQ_MAXM; k++)
        if (k < m) w[k >> 2] |= (uint32_t)cp[k] << (8 * (k & 3));
    }
#pragma unroll
    for (int k = 0; k < LSQ_MAXM; k++)
      if (k < m) acc = __fadd_rn(acc, lut[k * LSQ_H + ((w[k >> 2] >> (8 * (k & 3))) & 0xFFu)]);
    if (norms != nullptr) acc = __fadd_rn(acc, norms[v]);   // PQ tables carry no norm term
    if (acc <= tq) {
      const int pos = atomicAdd(&sh_n, 1);
      if (pos < cap) out[pos] = ((unsigned long long)float_to_ordered(acc) << 32) | (uint32_t)(v + (uint32_t)id_base);
    }
  }
  __syncthreads();
  // a filter list that overflowed may have lost true neighbours: report "too many" so that the query is re-run
  if (tid == 0) cnt[q] = (c_all > ccap) ? (int)((cap + 1 < 0x7FFFFFFF) ? cap + 1 : 0x7FFFFFFF) : sh_n;
}

// ------------------------------------------------------------------------------------------------
// 3b. thresholds without a sample buffer.  The filter kernel, run on the strided sample with a coarse bound from a
//     1/8 sub-sample, leaves each query a short list of sample positions (a few hundred of 16 K); this kernel scores
//     them exactly (same arithmetic as the rescoring) and returns the r-th smallest: the very tau the lookup-scan
//     path derives from its exact sample pass.  A list that overflowed or holds fewer than r entries (the coarse
//     bound was too tight: probability ~1e-10 per query) gives tau = +inf, which poisons the query's filter row
//     and sends the query to the exhaustive path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) adc_sample_tau_kernel(const uint8_t* __restrict__ codes, int m,
                                                             const float* __restrict__ norms, const float* __restrict__ lutq,
                                                             const uint32_t* __restrict__ list, const int* __restrict__ lcnt,
                                                             int lcap, int64_t stride, int64_t scount, int r,
                                                             float* __restrict__ tau, const float* __restrict__ queries,
                                                             int qstride, int d, const AdcStats* __restrict__ stats,
                                                             float* __restrict__ infl) {
  __shared__ float sh_q2[4];
  __shared__ float sh_tau;
  __shared__ int sh_c1;
  extern __shared__ __align__(16) unsigned char st_smem[];
  float* lut = reinterpret_cast<float*>(st_smem);                                   // [m * 256]
  uint32_t* dist = reinterpret_cast<uint32_t*>(st_smem + (size_t)m * LSQ_H * 4);    // [lcap] ordered distances
  const int q = blockIdx.x, tid = threadIdx.x;
  const float4* lsrc = reinterpret_cast<const float4*>(lutq + (size_t)q * m * LSQ_H);
  for (int j = tid; j < m * LSQ_H / 4; j += 128) reinterpret_cast<float4*>(lut)[j] = __ldg(lsrc + j);
  const int c_all = lcnt[q];
  const int c = (c_all < lcap) ? c_all : lcap;
  __syncthreads();
  const uint32_t* mine = list + (size_t)q * lcap;
  for (int i = tid; i < c; i += 128) {
    const uint32_t pos = mine[i];
    uint32_t key = 0xFFFFFFFFu;   // padding positions of the last sample tile never pass; be safe anyway
    if ((int64_t)pos < scount) {
      const int64_t v = (int64_t)pos * stride;
      const uint8_t* cp = codes + (size_t)v * m;
      float acc = 0.0f;
      for (int k = 0; k < m; k++) acc = __fadd_rn(acc, lut[k * LSQ_H + cp[k]]);
      if (norms != nullptr) acc = __fadd_rn(acc, norms[v]);
      key = float_to_ordered(acc);
    }
    dist[i] = key;
  }
  __syncthreads();
  if (c_all > lcap || c < r) {
    if (tid == 0) { tau[q] = INFINITY; if (infl != nullptr) infl[q] = 1.0f; }
    return;
  }
  for (int e = tid; e < c; e += 128) {   // rank by counting (ties by position): exactly one element has rank r - 1
    const uint32_t v = dist[e];
    int rk = 0;
    for (int j = 0; j < c; j++) {
      const uint32_t u = dist[j];
      rk += (u < v) || (u == v && j < e);
    }
    if (rk == r - 1) { tau[q] = ordered_to_float(v); sh_tau = ordered_to_float(v); }
  }
  if (infl == nullptr) return;
  // How many more pairs would a ONE-product filter let through?  Its margin is wider by 2 ||q|| max||lo(xhat)||; the
  // sample says how many distances lie within that of tau (the list reaches a good deal beyond tau: the coarse bound
  // sits near the 1 % quantile, tau near 0.3 %; where it does not, the count saturates and the estimate errs high).
  float q2 = 0.0f;
  for (int k = tid; k < d; k += 128) { const float x = queries[(size_t)q * qstride + k]; q2 = fmaf(x, x, q2); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q2 += __shfl_xor_sync(0xFFFFFFFFu, q2, o);
  if ((tid & 31) == 0) sh_q2[tid >> 5] = q2;
  if (tid == 0) sh_c1 = 0;
  __syncthreads();
  const float qn = sqrtf(sh_q2[0] + sh_q2[1] + sh_q2[2] + sh_q2[3]);
  const float m1 = 2.0f * qn * sqrtf(__uint_as_float(stats->xlo2_bits));
  const uint32_t lim = float_to_ordered(sh_tau + m1);
  int mine_c = 0;
  for (int e = tid; e < c; e += 128) mine_c += (dist[e] <= lim);
  atomicAdd(&sh_c1, mine_c);
  __syncthreads();
  if (tid == 0) infl[q] = (float)sh_c1 / (float)r;
}

// one product or two?  mean over the queries of the estimated growth of the survivor lists; one product (9 instead of
// 17 MMAs per tile pair, half the TMA bytes) pays as long as the extra survivors cost less in the rescoring and top-k
// than the filter saves: measured break-even is beyond 2x, the switch sits at 1.6x
__global__ void __launch_bounds__(256) adc_choose_passes_kernel(const float* __restrict__ infl, int nq, int* __restrict__ npass) {
  __shared__ float part[8];
  float s = 0.0f;
  for (int i = threadIdx.x; i < nq; i += 256) s += infl[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f;
    for (int w = 0; w < 8; w++) tot += part[w];
    const float mean = tot / (float)(nq > 0 ? nq : 1);
    *npass = (mean <= 1.6f) ? 1 : 2;   // a NaN (poisoned statistics) compares false: two products
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// shape rule only (what lsq_linscan_path reports): LSQ tables, d a multiple of 16 up to 128, enough base vectors —
// below ~64 K the lookup scan is launch-bound anyway — and enough queries to pay for decoding the base set
// (0.6-0.9 ms per million vectors, once per call): measured break-even against the lookup kernel at 1 M base vectors is
// ~940 queries at m = 8 and ~250 at m = 16, i.e. nq * m of 4-8 K.  LSQ_B200_ADC=scan|tc overrides the size rules.
bool adc_tc_shape_ok(int64_t n, int64_t nq, int m, int d) {
  const char* mode = getenv("LSQ_B200_ADC");
  if (mode != nullptr && strcmp(mode, "scan") == 0) return false;
  if (d % 16 != 0 || d < 16 || d > 128 || m < 1 || m > LSQ_MAXM) return false;
  const bool forced = (mode != nullptr && strcmp(mode, "tc") == 0);
  return forced || (n >= 65536 && nq * m >= 8000);
}

bool adc_tc_applicable(const uint8_t* dcodes, int64_t n, int64_t nq, int m, int d, const float* dqueries,
                       const float* dcodebooks) {
  if (!adc_tc_shape_ok(n, nq, m, d)) return false;
  if ((reinterpret_cast<uintptr_t>(dqueries) & 15) || (reinterpret_cast<uintptr_t>(dcodebooks) & 15) ||
      (reinterpret_cast<uintptr_t>(dcodes) & 3))
    return false;
  // above the memory gate the images would crowd out the caller (about 4 d bytes per base vector)
  const size_t need = (size_t)ceil_div(n, AT_N) * at_tile_bytes(d);
  if (need <= ((size_t)4 << 30)) return true;   // up to ~8 M vectors: no need to ask the driver (cudaMemGetInfo costs ~0.1 ms)
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
  return need < free_b / 4;
}

int adc_tc_prepare(const uint8_t* dcodes, int64_t n, int m, const float* dcodebooks, int d, const float* dbnorms,
                   int64_t scount, int64_t sstride, int subdim, int qstride, cudaStream_t st, AdcTcBase& B) {
  B.subdim = subdim;      // 0: LSQ (codewords of d elements, summed); > 0: PQ (sub-codewords of subdim elements, d = m * subdim)
  B.qstride = qstride;
  B.ntiles = ceil_div(n, AT_N);
  B.scount = scount;
  B.stiles = ceil_div(scount, AT_N);
  B.sstride = sstride;
  // 1/8 sub-sample of the sample (coarse bounds for the list-based thresholds); off for small samples
  B.s1count = (scount >= 8 * 1024) ? scount / 8 : 0;
  B.s1tiles = ceil_div(B.s1count, AT_N);
  LSQ_CUDA(B.img.alloc((size_t)B.ntiles * at_tile_bytes(d)));
  LSQ_CUDA(B.simg.alloc((size_t)B.stiles * at_tile_bytes(d)));
  LSQ_CUDA(B.s1img.alloc((size_t)std::max<int64_t>(B.s1tiles, 1) * at_tile_bytes(d)));
  LSQ_CUDA(B.stats.alloc(1));
  LSQ_CUDA(cudaMemsetAsync(B.stats.p, 0, sizeof(AdcStats), st));
  note_launch();
  adc_cbnorm_kernel<<<(unsigned)ceil_div((int64_t)m * LSQ_H, 256), 256, 0, st>>>(dcodebooks, m * LSQ_H, subdim > 0 ? subdim : d, B.stats.p);
  note_launch();
  adc_decode_kernel<<<(unsigned)B.ntiles, 256, 0, st>>>(dcodes, m, dcodebooks, d, dbnorms, B.img.p, B.stats.p, n, 1, subdim);
  if (B.stiles > 0) {
    note_launch();
    adc_decode_kernel<<<(unsigned)B.stiles, 256, 0, st>>>(dcodes, m, dcodebooks, d, dbnorms, B.simg.p, nullptr,
                                                         scount, sstride, subdim);
  }
  if (B.s1tiles > 0) {
    note_launch();
    adc_decode_kernel<<<(unsigned)B.s1tiles, 256, 0, st>>>(dcodes, m, dcodebooks, d, dbnorms, B.s1img.p, nullptr,
                                                          B.s1count, sstride * 8, subdim);
  }
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

static int filter_slices(int groups, int64_t ntiles) {
  // every CTA does the same work: the pass takes ceil(CTAs / SMs) waves of 1/slices each; slices of >= 16 tiles
  int dev = 0, sms = LSQ_NUM_SMS_HINT;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t max_s = std::max<int64_t>(1, std::min<int64_t>(65535, ntiles / 16));
  int best_s = 1;
  double best = 1e30;
  for (int64_t s = 1; s <= max_s; s++) {
    const double cost = (double)ceil_div((int64_t)groups * s, sms) / (double)s + 2e-4 * (double)s;  // + per-CTA set-up
    if (cost < best - 1e-12) { best = cost; best_s = (int)s; }
  }
  return best_s;
}

static int launch_filter(AdcFilterParams& p, cudaStream_t st) {
  const int groups = (int)ceil_div(p.nq, AT_NA * AT_M);
  const int slices = filter_slices(groups, p.ntiles);
  const size_t stage = (p.npass == 2 || p.npass_dev != nullptr) ? at_tile_bytes(p.d) : at_hi_bytes(p.d);
  p.stage_stride = (uint32_t)stage;
  p.nstages = (int)std::min<size_t>(AT_STAGES, (size_t)(210 * 1024) / stage);
  const size_t smem = (size_t)p.nstages * stage;
  int split = 2;
  if (const char* e = getenv("LSQ_B200_ADC_SPLIT")) split = (atoi(e) == 1) ? 1 : 2;   // A/B switch of the epilogue width
  const int mode = (p.sbuf != nullptr) ? AT_SAMPLE : (p.dbg != nullptr) ? AT_VALUES : AT_FILTER;
  note_launch();
#define LSQ_AT_LAUNCH(SP, MD)                                                                                      \
  do {                                                                                                             \
    LSQ_CUDA(cudaFuncSetAttribute(adc_filter_kernel<SP, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    adc_filter_kernel<SP, MD><<<dim3(groups, slices, 1), at_threads(SP), smem, st>>>(p);                           \
  } while (0)
  if (mode == AT_SAMPLE) LSQ_AT_LAUNCH(1, AT_SAMPLE);
  else if (mode == AT_VALUES) LSQ_AT_LAUNCH(1, AT_VALUES);
  else if (split == 2) LSQ_AT_LAUNCH(2, AT_FILTER);
  else LSQ_AT_LAUNCH(1, AT_FILTER);
#undef LSQ_AT_LAUNCH
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

// filter values of the strided sample -> dsbuf in threshold_kernel's layout (32-query tiles, `scount` steps)
int adc_tc_sample(const AdcTcBase& B, bool subsample, const float* dq, int nb, int d, int m, uint32_t* dsbuf,
                  cudaStream_t st) {
  AdcFilterParams p;
  memset(&p, 0, sizeof(p));
  p.queries = dq; p.sbuf = dsbuf; p.nq = nb; p.d = d; p.m = m; p.npass = 2; p.qstride = B.qstride; p.pq = B.subdim > 0;
  if (subsample) { p.img = B.s1img.p; p.scount = B.s1count; p.ntiles = B.s1tiles; }
  else { p.img = B.simg.p; p.scount = B.scount; p.ntiles = B.stiles; }
  p.n = p.scount;
  return launch_filter(p, st);
}

// sample positions whose filter value is <= dbound (+ margin) -> dlist / dlcnt, then the exact r-th smallest -> dtau
int adc_tc_sample_tau(const AdcTcBase& B, const uint8_t* dcodes, int m, const float* dq, int nb, int d,
                      const float* dbnorms, const float* dlutq, const float* dbound, uint32_t* dlist, int* dlcnt,
                      int lcap, int r, float* dtau, float* dinfl, int* dnpass, cudaStream_t st) {
  AdcFilterParams p;
  memset(&p, 0, sizeof(p));
  p.queries = dq; p.img = B.simg.p; p.tau = dbound; p.stats = B.stats.p; p.candidx = dlist; p.ccnt = dlcnt;
  p.n = B.scount; p.ntiles = B.stiles; p.ccap = lcap; p.nq = nb; p.d = d; p.m = m; p.npass = 2; p.qstride = B.qstride;
  p.pq = B.subdim > 0;
  LSQ_CUDA(cudaMemsetAsync(dlcnt, 0, (size_t)nb * sizeof(int), st));
  LSQ_TRY(launch_filter(p, st));
  const size_t smem = (size_t)m * LSQ_H * 4 + (size_t)lcap * 4;
  LSQ_CUDA(cudaFuncSetAttribute(adc_sample_tau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  note_launch();
  adc_sample_tau_kernel<<<nb, 128, smem, st>>>(dcodes, m, dbnorms, dlutq, dlist, dlcnt, lcap, B.sstride, B.scount, r, dtau, dq,
                                               B.qstride, d, B.stats.p, dinfl);
  if (dinfl != nullptr && dnpass != nullptr) {
    note_launch();
    adc_choose_passes_kernel<<<1, 256, 0, st>>>(dinfl, nb, dnpass);
  }
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

int adc_tc_lut_rows(const AdcTcBase& B, const float* dq, int nb, int d, const float* dcodebooks, int m, float* dlutq,
                    cudaStream_t st) {
  if (B.subdim > 0) {
    note_launch();
    adc_lut_rows_pq_kernel<<<dim3((unsigned)nb, (unsigned)m, 1), LSQ_H, 0, st>>>(dq, B.qstride, B.subdim, dcodebooks, m, dlutq);
    LSQ_CUDA(cudaGetLastError());
    return LSQ_OK;
  }
  const size_t smem = (size_t)d * (LR_Q + LR_J) * sizeof(float);
  LSQ_CUDA(cudaFuncSetAttribute(adc_lut_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  note_launch();
  adc_lut_rows_kernel<<<dim3((unsigned)ceil_div(nb, LR_Q), (unsigned)ceil_div((int64_t)m * LSQ_H, LR_J), 1), 256, smem, st>>>(
      dq, nb, d, dcodebooks, m * LSQ_H, dlutq);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

int adc_tc_main_pass(const AdcTcBase& B, const uint8_t* dcodes, int64_t n, int m, const float* dq, int nb, int d,
                     const float* dbnorms, const float* dlutq, const float* dtau, uint32_t* dcandidx, int* dccnt,
                     int64_t ccap, unsigned long long* dcand, int* dcnt, int64_t cap, int id_base, float* ddbg,
                     int64_t dbg_ld, const int* dnpass, cudaStream_t st) {
  AdcFilterParams p;
  memset(&p, 0, sizeof(p));
  p.queries = dq; p.img = B.img.p; p.tau = dtau; p.stats = B.stats.p;
  p.candidx = dcandidx; p.ccnt = dccnt; p.dbg = ddbg; p.n = n; p.ntiles = B.ntiles; p.ccap = ccap; p.dbg_ld = dbg_ld;
  p.nq = nb; p.d = d; p.m = m; p.qstride = B.qstride; p.pq = B.subdim > 0;
  // One product hi(q).hi(x) (9 MMAs per tile pair, margin wider by 2 ||q|| max||lo(x)||) or two, hi(q).lo(x) + hi(q).hi(x)
  // (17 MMAs)?  On the 1 M x 10 K benchmark: 2.34 + 0.43 ms (filter + rescoring of 1.3x the survivors) against 3.06 +
  // 0.33 ms.  The choice is made on the device from the sample (adc_choose_passes_kernel) when the caller passes
  // dnpass; LSQ_B200_ADC_PASSES=1|2 fixes it.  Without either: two products (the tighter filter).
  p.npass = 2;
  p.npass_dev = dnpass;
  if (const char* e = getenv("LSQ_B200_ADC_PASSES")) { p.npass = (atoi(e) == 1) ? 1 : 2; p.npass_dev = nullptr; }
  LSQ_CUDA(cudaMemsetAsync(dccnt, 0, (size_t)nb * sizeof(int), st));
  LSQ_TRY(launch_filter(p, st));
  if (dcand != nullptr) LSQ_TRY(adc_tc_rescore(dcodes, n, m, nb, dbnorms, dlutq, dtau, dcandidx, dccnt, ccap, dcand, dcnt, cap, id_base, st));
  return LSQ_OK;
}

int adc_tc_rescore(const uint8_t* dcodes, int64_t n, int m, int nb, const float* dbnorms, const float* dlutq,
                   const float* dtau, const uint32_t* dcandidx, const int* dccnt, int64_t ccap, unsigned long long* dcand,
                   int* dcnt, int64_t cap, int id_base, cudaStream_t st) {
  note_launch();
  adc_rescore_kernel<<<nb, 256, 0, st>>>(dcodes, n, m, dbnorms, dlutq, dtau, dcandidx, dccnt, ccap, dcand, dcnt, cap, id_base);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

}  // namespace lsq

using namespace lsq;

extern "C" int lsq_linscan_path(int64_t n, int64_t nq, int m, int d) { return adc_tc_shape_ok(n, nq, m, d) ? 1 : 0; }

// Test hook: the filter values  dbnorm[v] - 2 <q, xhat_v>  as the tensor cores compute them, for every pair.
// dout: device float [nq][ld], ld >= 128 * ceil(n / 128).  Nothing passes the filter (NaN thresholds).
extern "C" int lsq_dev_adc_filter_values(const uint8_t* dcodes, int64_t n, int m, const float* dqueries, int nq, int d,
                                         const float* dcodebooks, const float* dbnorms, float* dout, int64_t ld,
                                         void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  set_alloc_stream(st);
  LSQ_CHECK_ARG(n >= 1 && nq >= 1 && m >= 1 && m <= LSQ_MAXM && d % 16 == 0 && d >= 16 && d <= 128, "adc filter: bad sizes");
  LSQ_CHECK_ARG(ld >= 128 * ceil_div(n, 128), "adc filter: ld too small");
  AdcTcBase B;
  LSQ_TRY(adc_tc_prepare(dcodes, n, m, dcodebooks, d, dbnorms, 0, 1, 0, d, st, B));
  DevBuf<float> dtau;
  DevBuf<int> dccnt;
  LSQ_CUDA(dtau.alloc(nq));
  LSQ_CUDA(dccnt.alloc(nq));
  LSQ_CUDA(cudaMemsetAsync(dtau.p, 0xFF, (size_t)nq * sizeof(float), st));  // NaN thresholds: nothing passes
  return adc_tc_main_pass(B, dcodes, n, m, dqueries, nq, d, dbnorms, nullptr, dtau.p, nullptr, dccnt.p, 0, nullptr,
                          nullptr, 0, 0, dout, ld, nullptr, st);
}
