// chain.cu — exact chain encoder (ChainQ): encoding_viterbi / encode_viterbi! (src/encodings/
// encode_chain.jl:1-127), SURVEY.md §8 row f4.  Same unary and pair tables as the ICM path, a chain
// 1-2-...-m instead of the full graph, min-sum dynamic programming instead of local search.
//
// Reference arithmetic, kept bit for bit (oracle: orc_encoding_viterbi):
//   forward  V_1 = U_1;  mincost_i[j] = min_k ( V_i[k] + bb_i[k,j] ), first strict minimum over k
//            (encode_chain.jl:50-68, one fp32 add per (k, j));  V_{i+1} = U_{i+1} + mincost_i (:45-47,:72-74)
//   last     first minimum of V_m (:76);  backward trace through minidx (:78-85)
//   bb_i[k,j] = 2<C_i[:,k], C_{i+1}[:,j]> (:104-106) = T[i+1][i][k][j] of the table set the ICM path
//   already builds (both orientations materialised), so row k of that table is contiguous in j.
//
// Kernel: one warp scores VPW vectors.  Lane l owns the 8 to-states j = 4l..4l+3, 128+4l..128+4l+3 of
// every vector; the forward loop runs over the from-states k.  Only the VALUES mincost_i[j] are tracked
// there — per (k, j) pair one FADD (FMA pipe) and one FMNMX (ALU pipe), no index bookkeeping — because
// the backward trace needs minidx_i[j] for a single j per stage, the one on the optimal path: it is
// recomputed there as the first strict minimum over k of V_i[k] + bb_i[k, j*] from the stored V_i and
// the 1 KB row T[i][i+1][j*][:] (the other orientation of the same table), i.e. from exactly the values
// the forward pass minimised, so the codes equal the reference's minidx-based trace bit for bit.  V_i
// overwrites U_i in place (U is scratch of this call).  Per k the warp reads one 1 KB table row
// (2 x LDG.128 per lane, shared by its VPW vectors and, through L1, by the CTA's warps) and VPW
// shared-memory broadcasts of V_i[k].  Bound: the ALU pipe (FMNMX issues every 2nd cycle per SM
// sub-partition): (m-1) * 65536 pairs per vector at 2 cycles per 32 pairs.
// Inputs must be finite (fminf drops a NaN where the reference's `<` would keep it).
#include "icm.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace lsq {

constexpr int VIT_WARPS = 8;

// first strict minimum of 256 values held 8 per lane (4l..4l+3, 128+4l..128+4l+3): lane-local ascending
// scan, then a lexicographic (value, index) butterfly — same reduction as the ICM kernel's argmin
__device__ __forceinline__ int warp_first_argmin(const float (&x)[8], int lane) {
  float bv = x[0];
  int bj = 4 * lane;
#pragma unroll
  for (int t = 1; t < 8; t++) {
    const int j = (t < 4) ? 4 * lane + t : 128 + 4 * lane + (t - 4);
    if (x[t] < bv) { bv = x[t]; bj = j; }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const float ov = __shfl_xor_sync(0xFFFFFFFFu, bv, off);
    const int oj = __shfl_xor_sync(0xFFFFFFFFu, bj, off);
    if (ov < bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
  }
  return bj;
}

template <int VPW>
__global__ void __launch_bounds__(VIT_WARPS * 32) viterbi_kernel(float* __restrict__ U, const float* __restrict__ T,
                                                                 int64_t n, int m, uint8_t* __restrict__ codes) {
  __shared__ __align__(16) float vit_v[VIT_WARPS][VPW][LSQ_H];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ngroups = (n + VPW - 1) / VPW;
  const int64_t nwarps = (int64_t)gridDim.x * VIT_WARPS;

  for (int64_t grp = (int64_t)blockIdx.x * VIT_WARPS + wib; grp < ngroups; grp += nwarps) {
    int64_t v[VPW];
    bool valid[VPW];
#pragma unroll
    for (int e = 0; e < VPW; e++) {
      valid[e] = grp * VPW + e < n;
      v[e] = valid[e] ? grp * VPW + e : n - 1;  // tail slots recompute the last vector and never store
    }
    // V_1 = U_1
    __syncwarp();
#pragma unroll
    for (int e = 0; e < VPW; e++) {
      float4* Vs = reinterpret_cast<float4*>(vit_v[wib][e]);
      const float4* up = reinterpret_cast<const float4*>(U + (size_t)v[e] * LSQ_H);
      Vs[lane] = up[lane];
      Vs[32 + lane] = up[32 + lane];
    }
    __syncwarp();

    float best[VPW][8];
    for (int i = 0; i < m - 1; i++) {
      const float4* trow = reinterpret_cast<const float4*>(T + ((size_t)(i + 1) * m + i) * LSQ_H * LSQ_H);
      {
        const float4 g0 = __ldg(trow + lane), g1 = __ldg(trow + 32 + lane);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < VPW; e++) {
          const float vk = vit_v[wib][e][0];
#pragma unroll
          for (int t = 0; t < 8; t++) best[e][t] = __fadd_rn(vk, gv[t]);
        }
      }
#pragma unroll 4
      for (int k = 1; k < LSQ_H; k++) {
        const float4 g0 = __ldg(trow + k * 64 + lane), g1 = __ldg(trow + k * 64 + 32 + lane);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < VPW; e++) {
          const float vk = vit_v[wib][e][k];
#pragma unroll
          for (int t = 0; t < 8; t++) best[e][t] = fminf(best[e][t], __fadd_rn(vk, gv[t]));
        }
      }
      // V_{i+1} = U_{i+1} + mincost_i, kept in shared memory for the next stage and in U for the trace
      __syncwarp();  // every lane is done reading V_i
#pragma unroll
      for (int e = 0; e < VPW; e++) {
        float4* up = reinterpret_cast<float4*>(U + ((size_t)(i + 1) * n + v[e]) * LSQ_H);
        const float4 u0 = up[lane], u1 = up[32 + lane];
        best[e][0] = __fadd_rn(u0.x, best[e][0]); best[e][1] = __fadd_rn(u0.y, best[e][1]);
        best[e][2] = __fadd_rn(u0.z, best[e][2]); best[e][3] = __fadd_rn(u0.w, best[e][3]);
        best[e][4] = __fadd_rn(u1.x, best[e][4]); best[e][5] = __fadd_rn(u1.y, best[e][5]);
        best[e][6] = __fadd_rn(u1.z, best[e][6]); best[e][7] = __fadd_rn(u1.w, best[e][7]);
        const float4 w0 = make_float4(best[e][0], best[e][1], best[e][2], best[e][3]);
        const float4 w1 = make_float4(best[e][4], best[e][5], best[e][6], best[e][7]);
        float4* Vs = reinterpret_cast<float4*>(vit_v[wib][e]);
        Vs[lane] = w0;
        Vs[32 + lane] = w1;
        if (valid[e]) { up[lane] = w0; up[32 + lane] = w1; }  // each lane re-reads only its own elements
      }
      __syncwarp();
    }

    // first minimum of V_m, then the backward trace: code_i = first argmin_k V_i[k] + bb_i[k, code_{i+1}]
#pragma unroll
    for (int e = 0; e < VPW; e++) {
      int cur = warp_first_argmin(best[e], lane);
      int mine = cur;  // lane i keeps the code of node i
      for (int i = m - 2; i >= 0; i--) {
        const float4* vp = reinterpret_cast<const float4*>(U + ((size_t)i * n + v[e]) * LSQ_H);
        const float4* tp = reinterpret_cast<const float4*>(T + (((size_t)i * m + (i + 1)) * LSQ_H + cur) * LSQ_H);
        const float4 a0 = vp[lane], a1 = vp[32 + lane];
        const float4 g0 = __ldg(tp + lane), g1 = __ldg(tp + 32 + lane);
        const float c[8] = {__fadd_rn(a0.x, g0.x), __fadd_rn(a0.y, g0.y), __fadd_rn(a0.z, g0.z), __fadd_rn(a0.w, g0.w),
                            __fadd_rn(a1.x, g1.x), __fadd_rn(a1.y, g1.y), __fadd_rn(a1.z, g1.z), __fadd_rn(a1.w, g1.w)};
        cur = warp_first_argmin(c, lane);
        if (lane == i) mine = cur;
      }
      if (lane < m && valid[e]) codes[v[e] * m + lane] = (uint8_t)mine;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pipelined variant for large n: one persistent CTA per SM, 16 consumer warps x 4 vectors = 64 vectors
// per pass, all walking the same (stage, from-state) sequence.  A dedicated producer warp streams the
// stage tables through a 4-slot ring of 16-row chunks (16 KB each) with TMA bulk copies
// (cp.async.bulk + mbarrier full/empty pairs), so every table row is fetched once per CTA instead of
// once per warp and the consumers' inner loop only touches shared memory: 2 x LDS.128 per row and
// lane + one broadcast of V_i[k] per vector, then 32 x (FADD, FMNMX).  Same arithmetic, same codes.
// ------------------------------------------------------------------------------------------------
constexpr int VC_WARPS = 16;    // consumer warps
constexpr int VC_VPW = 4;       // vectors per warp
constexpr int VC_ROWS = 16;     // table rows per chunk
constexpr int VC_SLOTS = 4;     // ring depth
constexpr int VC_CHUNKS = LSQ_H / VC_ROWS;
constexpr uint32_t VC_CHUNK_BYTES = VC_ROWS * LSQ_H * 4;
constexpr size_t VC_SMEM = (size_t)VC_SLOTS * VC_CHUNK_BYTES + (size_t)VC_WARPS * VC_VPW * LSQ_H * 4 + 2 * VC_SLOTS * 8;

__global__ void __launch_bounds__((VC_WARPS + 1) * 32, 1) viterbi_tma_kernel(float* __restrict__ U,
                                                                             const float* __restrict__ T, int64_t n,
                                                                             int m, uint8_t* __restrict__ codes) {
  extern __shared__ __align__(128) unsigned char vc_smem[];
  float* ring = reinterpret_cast<float*>(vc_smem);
  float* vall = reinterpret_cast<float*>(vc_smem + (size_t)VC_SLOTS * VC_CHUNK_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(vc_smem + (size_t)VC_SLOTS * VC_CHUNK_BYTES +
                                               (size_t)VC_WARPS * VC_VPW * LSQ_H * 4);
  uint64_t* empty = full + VC_SLOTS;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  constexpr int VPC = VC_WARPS * VC_VPW;  // vectors per CTA pass
  const int64_t npasses_all = (n + VPC - 1) / VPC;
  const int64_t my_passes = (npasses_all > (int64_t)blockIdx.x)
                                ? (npasses_all - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total_chunks = my_passes * (m - 1) * VC_CHUNKS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < VC_SLOTS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], VC_WARPS * 32); }
    fence_mbar_init();
  }
  __syncthreads();

  if (wib == VC_WARPS) {
    // ---- producer warp: one elected lane keeps the ring full ----
    if (lane == 0) {
      for (int64_t p = 0; p < total_chunks; p++) {
        const int slot = (int)(p % VC_SLOTS);
        if (p >= VC_SLOTS) mbar_wait(&empty[slot], (uint32_t)(((p / VC_SLOTS) - 1) & 1));
        const int i = (int)((p / VC_CHUNKS) % (m - 1));
        const int c = (int)(p % VC_CHUNKS);
        const float* src = T + ((size_t)(i + 1) * m + i) * LSQ_H * LSQ_H + (size_t)c * VC_ROWS * LSQ_H;
        mbar_expect_tx(&full[slot], VC_CHUNK_BYTES);
        bulk_g2s(ring + (size_t)slot * VC_ROWS * LSQ_H, src, VC_CHUNK_BYTES, &full[slot]);
      }
    }
    return;
  }

  // ---- consumer warps ----
  float* vmine = vall + (size_t)wib * VC_VPW * LSQ_H;
  const uint32_t ring_u32 = smem_u32(ring) + (uint32_t)lane * 16u;
  int64_t q = 0;  // chunk counter, identical in every consumer warp
  for (int64_t pass = blockIdx.x; pass < npasses_all; pass += gridDim.x) {
    int64_t v[VC_VPW];
    bool valid[VC_VPW];
#pragma unroll
    for (int e = 0; e < VC_VPW; e++) {
      const int64_t idx = pass * VPC + (int64_t)wib * VC_VPW + e;
      valid[e] = idx < n;
      v[e] = valid[e] ? idx : n - 1;  // tail slots recompute the last vector and never store
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < VC_VPW; e++) {
      float4* Vs = reinterpret_cast<float4*>(vmine + e * LSQ_H);
      const float4* up = reinterpret_cast<const float4*>(U + (size_t)v[e] * LSQ_H);
      Vs[lane] = up[lane];
      Vs[32 + lane] = up[32 + lane];
    }
    __syncwarp();

    float best[VC_VPW][8];
    for (int i = 0; i < m - 1; i++) {
#pragma unroll
      for (int e = 0; e < VC_VPW; e++)
#pragma unroll
        for (int t = 0; t < 8; t++) best[e][t] = INFINITY;  // min(+inf, c) = c: same as starting from k = 0
      for (int c = 0; c < VC_CHUNKS; c++, q++) {
        const int slot = (int)(q % VC_SLOTS);
        mbar_wait(&full[slot], (uint32_t)((q / VC_SLOTS) & 1));
        const uint32_t base = ring_u32 + (uint32_t)slot * VC_CHUNK_BYTES;
#pragma unroll 1
        for (int r4 = 0; r4 < VC_ROWS; r4 += 4) {
          float4 vk4[VC_VPW];  // V_i[k..k+3] of each vector: one broadcast LDS.128 per vector and 4 rows
#pragma unroll
          for (int e = 0; e < VC_VPW; e++)
            vk4[e] = *reinterpret_cast<const float4*>(vmine + e * LSQ_H + c * VC_ROWS + r4);
#pragma unroll
          for (int rr = 0; rr < 4; rr++) {
            const int r = r4 + rr;
            float4 g0, g1;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g0.x), "=f"(g0.y), "=f"(g0.z), "=f"(g0.w) : "r"(base + (uint32_t)r * 1024u));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g1.x), "=f"(g1.y), "=f"(g1.z), "=f"(g1.w) : "r"(base + (uint32_t)r * 1024u + 512u));
            const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int e = 0; e < VC_VPW; e++) {
              const float vk = (rr == 0) ? vk4[e].x : (rr == 1) ? vk4[e].y : (rr == 2) ? vk4[e].z : vk4[e].w;
#pragma unroll
              for (int t = 0; t < 8; t++) best[e][t] = fminf(best[e][t], __fadd_rn(vk, gv[t]));
            }
          }
        }
        mbar_arrive(&empty[slot]);  // every reading thread releases the slot itself
      }
      // V_{i+1} = U_{i+1} + mincost_i (per-warp state: only warp-level synchronisation)
      __syncwarp();  // every lane is done reading V_i
#pragma unroll
      for (int e = 0; e < VC_VPW; e++) {
        float4* up = reinterpret_cast<float4*>(U + ((size_t)(i + 1) * n + v[e]) * LSQ_H);
        const float4 u0 = up[lane], u1 = up[32 + lane];
        best[e][0] = __fadd_rn(u0.x, best[e][0]); best[e][1] = __fadd_rn(u0.y, best[e][1]);
        best[e][2] = __fadd_rn(u0.z, best[e][2]); best[e][3] = __fadd_rn(u0.w, best[e][3]);
        best[e][4] = __fadd_rn(u1.x, best[e][4]); best[e][5] = __fadd_rn(u1.y, best[e][5]);
        best[e][6] = __fadd_rn(u1.z, best[e][6]); best[e][7] = __fadd_rn(u1.w, best[e][7]);
        const float4 w0 = make_float4(best[e][0], best[e][1], best[e][2], best[e][3]);
        const float4 w1 = make_float4(best[e][4], best[e][5], best[e][6], best[e][7]);
        float4* Vs = reinterpret_cast<float4*>(vmine + e * LSQ_H);
        Vs[lane] = w0;
        Vs[32 + lane] = w1;
        if (valid[e]) { up[lane] = w0; up[32 + lane] = w1; }
      }
      __syncwarp();
    }

#pragma unroll
    for (int e = 0; e < VC_VPW; e++) {
      int cur = warp_first_argmin(best[e], lane);
      int mine = cur;
      for (int i = m - 2; i >= 0; i--) {
        const float4* vp = reinterpret_cast<const float4*>(U + ((size_t)i * n + v[e]) * LSQ_H);
        const float4* tp = reinterpret_cast<const float4*>(T + (((size_t)i * m + (i + 1)) * LSQ_H + cur) * LSQ_H);
        const float4 a0 = vp[lane], a1 = vp[32 + lane];
        const float4 g0 = __ldg(tp + lane), g1 = __ldg(tp + 32 + lane);
        const float cc[8] = {__fadd_rn(a0.x, g0.x), __fadd_rn(a0.y, g0.y), __fadd_rn(a0.z, g0.z), __fadd_rn(a0.w, g0.w),
                             __fadd_rn(a1.x, g1.x), __fadd_rn(a1.y, g1.y), __fadd_rn(a1.z, g1.z), __fadd_rn(a1.w, g1.w)};
        cur = warp_first_argmin(cc, lane);
        if (lane == i) mine = cur;
      }
      if (lane < m && valid[e]) codes[v[e] * m + lane] = (uint8_t)mine;
    }
  }
}

int launch_viterbi(float* dU, int64_t n, int m, const float* dT, uint8_t* dcodes, cudaStream_t st) {
  if (n == 0) return LSQ_OK;
  LSQ_CHECK_ARG(m >= 2 && m <= LSQ_MAXM, "viterbi: m must be in 2..16 (a chain needs two nodes)");
  constexpr int VPW = 4;
  int dev = 0, sms = LSQ_NUM_SMS_HINT;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // large inputs: the TMA-pipelined kernel (64 vectors per CTA pass); LSQ_B200_VITERBI=simple|tma overrides
  const char* ve = getenv("LSQ_B200_VITERBI");
  const bool want_tma = ve ? (strcmp(ve, "tma") == 0) : (n >= (int64_t)sms * VC_WARPS * VC_VPW);
  if (want_tma) {
    LSQ_CUDA(cudaFuncSetAttribute(viterbi_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VC_SMEM));
    const int64_t passes = ceil_div(n, VC_WARPS * VC_VPW);
    const unsigned grid = (unsigned)std::min<int64_t>(passes, sms);
    note_launch();
    viterbi_tma_kernel<<<grid, (VC_WARPS + 1) * 32, VC_SMEM, st>>>(dU, dT, n, m, dcodes);
    LSQ_CUDA(cudaGetLastError());
    return LSQ_OK;
  }
  int per_sm = 1;  // whole waves only: a partial last wave of resident CTAs costs a full pass
  LSQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, viterbi_kernel<VPW>, VIT_WARPS * 32, 0));
  if (per_sm < 1) per_sm = 1;
  const int64_t need = ceil_div(ceil_div(n, VPW), VIT_WARPS);
  const unsigned grid = (unsigned)std::min<int64_t>(need, (int64_t)sms * per_sm);
  note_launch();
  viterbi_kernel<VPW><<<grid, VIT_WARPS * 32, 0, st>>>(dU, dT, n, m, dcodes);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

}  // namespace lsq

using namespace lsq;

extern "C" {

int lsq_dev_viterbi(float* dU, int64_t n, int m, const float* dT, uint8_t* dcodes, void* stream) {
  LSQ_CHECK_ARG(n >= 0, "n must be >= 0");
  return launch_viterbi(dU, n, m, dT, dcodes, (cudaStream_t)stream);
}

int lsq_encoding_viterbi(const float* X, int d, int64_t n, const float* C, int m, int h, int16_t* B, int verbose) {
  LSQ_CHECK_ARG(d >= 1 && n >= 0, "bad sizes");
  LSQ_CHECK_ARG(m >= 2 && m <= LSQ_MAXM, "viterbi: m must be in 2..16 (a chain needs two nodes)");
  LSQ_CHECK_ARG(h == LSQ_H, "h must be 256");
  if (n == 0) return LSQ_OK;
  cudaStream_t st;
  LSQ_TRY(host_ctx(&st));
  DevBuf<float> dX, dC, dnorms, dT, dU;
  DevBuf<uint8_t> dcodes;
  DevBuf<int16_t> d16;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
  int64_t chunk = (int64_t)(0.5 * (double)free_b / ((double)m * LSQ_H * 4 + (double)d * 4 + 3.0 * m));
  chunk = std::max<int64_t>(1024, std::min<int64_t>(chunk, n));
  if (const char* ce = getenv("LSQ_B200_CHUNK_VECTORS")) chunk = std::max<int64_t>(1, std::min<int64_t>(atoll(ce), n));  // tests
  LSQ_CUDA(dX.alloc((size_t)chunk * d));
  LSQ_CUDA(dC.alloc((size_t)m * LSQ_H * d));
  LSQ_CUDA(dnorms.alloc((size_t)m * LSQ_H));
  LSQ_CUDA(dT.alloc((size_t)m * m * LSQ_H * LSQ_H));
  LSQ_CUDA(dU.alloc((size_t)chunk * m * LSQ_H));
  LSQ_CUDA(dcodes.alloc((size_t)chunk * m));
  LSQ_CUDA(d16.alloc((size_t)chunk * m));
  LSQ_CUDA(cudaMemcpyAsync(dC.p, C, (size_t)m * LSQ_H * d * 4, cudaMemcpyHostToDevice, st));
  LSQ_TRY(build_norms(dC.p, d, m, dnorms.p, st));
  LSQ_TRY(build_tables(dC.p, d, m, dT.p, st));
  for (int64_t lo = 0; lo < n; lo += chunk) {
    const int64_t nc = std::min<int64_t>(chunk, n - lo);
    LSQ_CUDA(cudaMemcpyAsync(dX.p, X + (size_t)lo * d, (size_t)nc * d * 4, cudaMemcpyHostToDevice, st));
    LSQ_TRY(build_unaries(dX.p, d, nc, dC.p, m, dnorms.p, dU.p, 0, st));
    LSQ_TRY(launch_viterbi(dU.p, nc, m, dT.p, dcodes.p, st));
    LSQ_TRY(launch_codes_u8_to_i16(dcodes.p, d16.p, nc * m, st));
    LSQ_CUDA(cudaMemcpyAsync(B + (size_t)lo * m, d16.p, (size_t)nc * m * 2, cudaMemcpyDeviceToHost, st));
    LSQ_CUDA(cudaStreamSynchronize(st));
    if (verbose) fprintf(stderr, "[lsq_b200] viterbi: %lld / %lld vectors\n", (long long)(lo + nc), (long long)n);
  }
  return LSQ_OK;
}

}  // extern "C"
