// icm_slice.cu — ICM / ILS encoding with shared-memory resident pair-table slices (m <= 8, large n).
//
// Same algorithm, same bits as icm_ils_warp_kernel (encode_icm.jl:4-189); different data movement.
// The warp-per-vector kernel gathers the (m-1) x 1 KB conditioning columns of every node visit from L2
// and saturates it (lts throughput 83 %).  Here the columns come from SHARED MEMORY:
//
//   * persistent grid, one CTA per SM; a CTA owns a contiguous range of vectors for the whole launch;
//   * for node j the CTA walks the 8 candidate slices of 32: slice (j, s) of all m-1 tables is
//     (m-1)*256 rows x 128 B = 224 KB at m = 8, contiguous in the pre-sliced layout Ts, and is staged
//     by TMA bulk copies (cp.async.bulk + mbarrier) — one staging serves every active vector of the CTA;
//   * a quarter-warp (8 lanes x float4) handles one vector: each LDS.128 quarter-wavefront reads one
//     128-byte table row -> conflict-free, 32 candidates per wavefront, 4 vectors per warp instruction;
//   * the vector's 32 unary values of the slice are one 128-byte line of the sliced unary layout
//     U[j][s][v][32]: with the tables on chip, the unary stream is what HBM carries;
//   * the running (value, index) minimum across slices lives in an L2-resident scratch word per
//     vector; the last slice writes the new code and updates the clean mask.
// Node visits whose conditioning codes did not change are skipped exactly as in the warp kernel
// (per-vector clean masks); a CTA-wide ordered compaction builds the active list of every visit.
#include "icm.cuh"

#include <stdlib.h>

namespace lsq {

constexpr int SLICE_THREADS = 512;
constexpr int DEPTH = 4;  // prefetched vector groups per warp

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// unary line: streamed once, keep it out of L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

template <int M>
__device__ __forceinline__ float warp_veccost_packed(const float* __restrict__ x, const float* __restrict__ C,
                                                     int d, unsigned long long codes, int lane) {
  float p = 0.0f;
  for (int t = lane; t < d; t += 32) {
    float r = 0.0f;
#pragma unroll
    for (int k = 0; k < M; k++)
      r = __fadd_rn(r, __ldg(C + ((size_t)k * LSQ_H + ((uint32_t)(codes >> (8 * k)) & 0xFFu)) * d + t));
    const float df = __fsub_rn(r, __ldg(x + t));
    p = __fadd_rn(p, __fmul_rn(df, df));
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(0xFFFFFFFFu, p, off));
  return p;
}

// veccost of one vector by a QUARTER-warp (8 lanes), bit-identical to warp_veccost_packed: lane c8 owns
// the lane-strided partial sums l = 4*c8 .. 4*c8+3 (elements t = l, l+32, ...), so the xor-butterfly
// offsets 16, 8, 4 are quarter shuffles (4, 2, 1) and offsets 2, 1 stay inside the lane.  Needs d % 4 == 0.
template <int M>
__device__ __forceinline__ float quarter_veccost_packed(const float* __restrict__ x, const float* __restrict__ C,
                                                        int d, unsigned long long codes, int c8) {
  float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
  for (int t0 = 4 * c8; t0 < d; t0 += 32) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < M; k++) {
      const float4 cv = __ldg(reinterpret_cast<const float4*>(
          C + ((size_t)k * LSQ_H + ((uint32_t)(codes >> (8 * k)) & 0xFFu)) * d + t0));
      r.x = __fadd_rn(r.x, cv.x); r.y = __fadd_rn(r.y, cv.y); r.z = __fadd_rn(r.z, cv.z); r.w = __fadd_rn(r.w, cv.w);
    }
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + t0));
    const float d0 = __fsub_rn(r.x, xv.x), d1 = __fsub_rn(r.y, xv.y), d2 = __fsub_rn(r.z, xv.z), d3 = __fsub_rn(r.w, xv.w);
    p0 = __fadd_rn(p0, __fmul_rn(d0, d0)); p1 = __fadd_rn(p1, __fmul_rn(d1, d1));
    p2 = __fadd_rn(p2, __fmul_rn(d2, d2)); p3 = __fadd_rn(p3, __fmul_rn(d3, d3));
  }
#pragma unroll
  for (int off = 4; off >= 1; off >>= 1) {
    p0 = __fadd_rn(p0, __shfl_xor_sync(0xFFFFFFFFu, p0, off));
    p1 = __fadd_rn(p1, __shfl_xor_sync(0xFFFFFFFFu, p1, off));
    p2 = __fadd_rn(p2, __shfl_xor_sync(0xFFFFFFFFu, p2, off));
    p3 = __fadd_rn(p3, __shfl_xor_sync(0xFFFFFFFFu, p3, off));
  }
  return __fadd_rn(__fadd_rn(p0, p2), __fadd_rn(p1, p3));
}

template <int M>
__global__ void __launch_bounds__(SLICE_THREADS, 1) icm_ils_slice_kernel(const __grid_constant__ IcmParams p) {
  constexpr int NT = SLICE_THREADS, NW = NT / 32;
  constexpr uint32_t TAB_BYTES = (uint32_t)(M - 1) * LSQ_H * ICM_SLICE_W * 4;
  constexpr uint32_t ALL_CLEAN = (1u << M) - 1u;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ int warp_cnt[NW];
  __shared__ int s_total;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane >> 3, c4 = lane & 7;  // quarter (vector within the group of 4), float4 within the slice
  const uint32_t tab_base = smem_u32(smem_raw);

  // this CTA's vector range (splitarray rule)
  int64_t v0, v1;
  {
    const int64_t per = p.n / gridDim.x, xtra = p.n % gridDim.x;
    const int64_t b = blockIdx.x;
    v0 = (b < xtra) ? b * (per + 1) : xtra * (per + 1) + (b - xtra) * per;
    v1 = v0 + per + ((b < xtra) ? 1 : 0);
  }
  const int nv = (int)(v1 - v0);
  int* const actp = p.act + v0;
  unsigned long long* const wcp = p.wcodes + v0;
  unsigned long long* const rbp = p.rbest + v0;
  uint16_t* const wclp = p.wclean + v0;
  const uint32_t tab_lane = tab_base + (uint32_t)c4 * 16u;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < nv; i += NT) p.clean[v0 + i] = 0;
  uint32_t phase = 0;
  __syncthreads();

  for (int it = 0; it < p.niters; it++) {
    // ---- (A) perturbation of every vector of the range (encode_icm.jl:56-70) ----
    for (int i = tid; i < nv; i += NT) {
      const int64_t v = v0 + i;
      unsigned long long c = 0;
      if (M == 8) {
        c = *reinterpret_cast<const unsigned long long*>(p.codes + v * M);
      } else {
#pragma unroll
        for (int k = 0; k < M; k++) c |= (unsigned long long)p.codes[v * M + k] << (8 * k);
      }
      uint32_t wc = p.clean[v];
      if (p.slots != nullptr) {
        const size_t base = ((size_t)it * p.n + v) * p.npert;
        for (int e = 0; e < p.npert; e++) {
          const int s = p.slots[base + e];
          const unsigned long long x = p.vals[base + e];
          if (((c >> (8 * s)) & 0xFFull) != x) { c = (c & ~(0xFFull << (8 * s))) | (x << (8 * s)); wc = 0; }
        }
      } else if (p.npert > 0) {
        uint8_t s[LSQ_MAXM], x[LSQ_MAXM];
        make_perturb_one(p.seed, p.ils_iter0 + it, p.g0 + (uint64_t)v, M, LSQ_H, p.npert, s, x);
        for (int e = 0; e < p.npert; e++)
          if (((c >> (8 * s[e])) & 0xFFull) != x[e]) {
            c = (c & ~(0xFFull << (8 * s[e]))) | ((unsigned long long)x[e] << (8 * s[e]));
            wc = 0;
          }
      }
      p.wcodes[v] = c;
      p.wclean[v] = (uint16_t)wc;
    }
    __syncthreads();

    // ---- (B) block-ICM sweeps (encode_icm.jl:72-125) ----
    for (int sweep = 0; sweep < p.icmiter; sweep++) {
      int visited = 0;
      for (int jj = 0; jj < M; jj++) {
        const int j = p.orders[it][jj];
        // B1. ordered compaction of the vectors whose node j is dirty -> act[v0 ..]; 4 vectors per thread
        int n_act = 0;
        for (int base = 0; base < nv; base += NT * 4) {
          const int i0 = base + tid * 4;
          uint32_t am = 0;
#pragma unroll
          for (int e = 0; e < 4; e++)
            if (i0 + e < nv && !((p.wclean[v0 + i0 + e] >> j) & 1u)) am |= 1u << e;
          const int cnt = __popc(am);
          int incl = cnt;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= off) incl += t;
          }
          if (lane == 31) warp_cnt[warp] = incl;
          __syncthreads();
          if (warp == 0) {
            const int wc = (lane < NW) ? warp_cnt[lane] : 0;
            int wincl = wc;
#pragma unroll
            for (int off = 1; off < NW; off <<= 1) {
              const int t = __shfl_up_sync(0xFFFFFFFFu, wincl, off);
              if (lane >= off) wincl += t;
            }
            if (lane < NW) warp_cnt[lane] = wincl - wc;  // exclusive offsets
            if (lane == NW - 1) s_total = wincl;
          }
          __syncthreads();
          int pos = n_act + warp_cnt[warp] + incl - cnt;
#pragma unroll
          for (int e = 0; e < 4; e++)
            if (am & (1u << e)) p.act[v0 + pos++] = i0 + e;
          n_act += s_total;
          __syncthreads();
        }
        if (n_act == 0) continue;
        visited = 1;
        const int ngroups = (n_act + 3) >> 2;
        // conditioning codes with byte j squeezed out: byte kk = code of the kk-th codebook != j
        const unsigned long long lowmask = (1ull << (8 * j)) - 1ull;

        for (int s = 0; s < ICM_SLICES; s++) {
          // B2. stage slice (j, s) of the m-1 tables; everyone has left the previous slice
          __syncthreads();
          if (tid == 0)
            bulk_load_issue(smem_raw, p.Ts + ((size_t)(j * ICM_SLICES + s) * (M - 1)) * LSQ_H * ICM_SLICE_W, TAB_BYTES,
                            &bar);
          const float4* uslice = reinterpret_cast<const float4*>(p.U + ((size_t)(j * ICM_SLICES + s) * p.n + v0) * ICM_SLICE_W);

          // Two-stage register pipeline per warp.  Stage 1 loads the active-list entry of the group
          // 2*DEPTH rounds ahead; stage 2 turns the entry loaded DEPTH rounds ago into the data loads
          // (codes, unary line, running best) of the group DEPTH rounds ahead.  No load result is used
          // in the round it was issued, and DEPTH x NW x 4 vectors are in flight per SM.
          const float4* ulane = uslice + c4;
          int n_i[DEPTH], r_i[DEPTH];
          unsigned long long r_c[DEPTH], r_b[DEPTH];
          float4 r_a[DEPTH];
          auto load_idx = [&](int gg) -> int {
            const int idx = gg * 4 + q;
            return (gg < ngroups) ? actp[idx < n_act ? idx : n_act - 1] : 0;
          };
          auto load_data = [&](int gg, int u) {
            const int i = n_i[u];
            r_i[u] = i;
            if (gg < ngroups) {
              r_c[u] = wcp[i];
              r_a[u] = ldg_stream(ulane + (size_t)(uint32_t)i * 8);
              if (s > 0) r_b[u] = rbp[i];
            }
          };
#pragma unroll
          for (int u = 0; u < DEPTH; u++) { n_i[u] = load_idx(warp + u * NW); r_c[u] = 0ull; r_b[u] = 0ull; r_a[u] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
          for (int u = 0; u < DEPTH; u++) { load_data(warp + u * NW, u); n_i[u] = load_idx(warp + (u + DEPTH) * NW); }
          mbar_wait(&bar, phase);
          phase ^= 1u;

          for (int g = warp; g < ngroups; g += DEPTH * NW) {
            // DEPTH independent groups per round, phase by phase and branch-free until the stores, so the
            // scheduler can overlap their LDS -> FADD -> shuffle chains (only 4 warps per scheduler).
            int ci[DEPTH], bi[DEPTH];
            unsigned long long cc[DEPTH], cb[DEPTH];
            float4 a[DEPTH];
            float best[DEPTH];
#pragma unroll
            for (int u = 0; u < DEPTH; u++) { ci[u] = r_i[u]; cc[u] = r_c[u]; cb[u] = r_b[u]; a[u] = r_a[u]; }
#pragma unroll
            for (int u = 0; u < DEPTH; u++) {
              load_data(g + (u + DEPTH) * NW, u);
              n_i[u] = load_idx(g + (u + 2 * DEPTH) * NW);
            }
#pragma unroll
            for (int u = 0; u < DEPTH; u++) {
              const unsigned long long pk = (cc[u] & lowmask) | ((cc[u] >> 8) & ~lowmask);
              const uint32_t pk0 = (uint32_t)pk, pk1 = (uint32_t)(pk >> 32);
#pragma unroll
              for (int kk = 0; kk < M - 1; kk++) {
                const uint32_t c = __byte_perm(kk < 4 ? pk0 : pk1, 0u, 0x4440u | (uint32_t)(kk & 3));
                const float4 t4 = lds128(tab_lane + (uint32_t)(kk * LSQ_H * ICM_SLICE_W * 4) + c * (ICM_SLICE_W * 4));
                a[u].x = __fadd_rn(a[u].x, t4.x); a[u].y = __fadd_rn(a[u].y, t4.y);
                a[u].z = __fadd_rn(a[u].z, t4.z); a[u].w = __fadd_rn(a[u].w, t4.w);
              }
            }
#pragma unroll
            for (int u = 0; u < DEPTH; u++) {
              // first strict minimum of this lane's 4 candidates ...
              float bv = a[u].x;
              int bx = s * ICM_SLICE_W + c4 * 4;
              if (a[u].y < bv) { bv = a[u].y; bx = s * ICM_SLICE_W + c4 * 4 + 1; }
              if (a[u].z < bv) { bv = a[u].z; bx = s * ICM_SLICE_W + c4 * 4 + 2; }
              if (a[u].w < bv) { bv = a[u].w; bx = s * ICM_SLICE_W + c4 * 4 + 3; }
              // ... then of the quarter's 32: value-only min butterfly, lowest lane holding the minimum wins
              // (lanes hold ascending candidate ranges, so this is the first strict minimum)
              float mn = bv;
              mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, 4));
              mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, 2));
              mn = fminf(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, 1));
              const uint32_t eq = __ballot_sync(0xFFFFFFFFu, bv == mn);
              const int src = (q << 3) + __ffs((eq >> (q << 3)) & 0xFFu) - 1;
              bi[u] = __shfl_sync(0xFFFFFFFFu, bx, src & 31);
              best[u] = mn;
            }
#pragma unroll
            for (int u = 0; u < DEPTH; u++) {
              const int gg = g + u * NW;
              if (c4 == 0 && gg < ngroups && (gg * 4 + q) < n_act) {
                float bv = best[u];
                int bx = bi[u];
                if (s > 0) {  // earlier slices hold lower candidate indices: they win ties
                  const float pv = __uint_as_float((uint32_t)(cb[u] >> 32));
                  if (!(bv < pv)) { bv = pv; bx = (int)(uint32_t)cb[u]; }
                }
                if (s < ICM_SLICES - 1) {
                  rbp[ci[u]] = ((unsigned long long)__float_as_uint(bv) << 32) | (uint32_t)bx;
                } else {
                  const uint32_t old = (uint32_t)(cc[u] >> (8 * j)) & 0xFFu;
                  if ((uint32_t)bx != old) {
                    const unsigned long long nc = (cc[u] & ~(0xFFull << (8 * j))) | ((unsigned long long)bx << (8 * j));
                    wcp[ci[u]] = nc;
                    // back at the accepted codes: what is known about that state applies again
                    unsigned long long ac;
                    if (M == 8) {
                      ac = *reinterpret_cast<const unsigned long long*>(p.codes + (v0 + ci[u]) * M);
                    } else {
                      ac = 0;
#pragma unroll
                      for (int k = 0; k < M; k++) ac |= (unsigned long long)p.codes[(v0 + ci[u]) * M + k] << (8 * k);
                    }
                    wclp[ci[u]] = (uint16_t)(((nc == ac) ? p.clean[v0 + ci[u]] : 0u) | (1u << j));
                  } else {
                    wclp[ci[u]] = (uint16_t)(wclp[ci[u]] | (1u << j));
                  }
                }
              }
            }
          }
        }
        __syncthreads();  // codes / clean masks of this visit visible before the next compaction
      }
      if (!visited) break;  // every vector of the range is at its ICM fixed point
    }
    __syncthreads();

    // ---- (C) accept iff strictly better (encode_icm.jl:178-186); a quarter-warp per vector ----
    const int sn = p.snap_of_iter[it];
    const bool quarter_ok = (p.d % 4 == 0);
    for (int i0 = warp * 4; i0 < nv; i0 += NW * 4) {
      const int i = i0 + q;
      const bool valid = i < nv;
      const int64_t v = v0 + (valid ? i : nv - 1);
      const unsigned long long wc = p.wcodes[v];
      unsigned long long c;
      if (M == 8) {
        c = *reinterpret_cast<const unsigned long long*>(p.codes + v * M);
      } else {
        c = 0;
#pragma unroll
        for (int k = 0; k < M; k++) c |= (unsigned long long)p.codes[v * M + k] << (8 * k);
      }
      float prev = p.cost[v];
      const bool differs = valid && (wc != c);  // identical codes cannot be strictly better
      float newc = prev;
      if (quarter_ok) {
        if (__any_sync(0xFFFFFFFFu, differs)) newc = quarter_veccost_packed<M>(p.X + (size_t)v * p.d, p.C, p.d, wc, c4);
      } else {
        // generic d: one warp per vector, quarter by quarter
        for (int qq = 0; qq < 4; qq++) {
          const unsigned long long wq = __shfl_sync(0xFFFFFFFFu, wc, qq * 8);
          const int64_t vq = __shfl_sync(0xFFFFFFFFu, v, qq * 8);
          const float t = warp_veccost_packed<M>(p.X + (size_t)vq * p.d, p.C, p.d, wq, lane);
          if (q == qq) newc = t;
        }
      }
      if (differs && newc < prev) {
        prev = newc;
        c = wc;
        if (c4 < M) p.codes[v * M + c4] = (uint8_t)(wc >> (8 * c4));
        if (c4 == 0) { p.cost[v] = newc; p.clean[v] = p.wclean[v]; }
      } else if (valid && !differs && c4 == 0) {
        p.clean[v] = (uint16_t)(p.clean[v] | p.wclean[v]);  // same codes: keep what the sweeps learned
      }
      if (sn >= 0 && valid) {
        if (c4 < M) p.snap[((size_t)sn * p.n + v) * M + c4] = (uint8_t)(c >> (8 * c4));
        if (c4 == 0 && p.snapcost != nullptr) p.snapcost[(size_t)sn * p.n + v] = prev;
      }
    }
    __syncthreads();
  }
  (void)ALL_CLEAN;
}

int icm_use_slices(int m, int64_t n) {
  if (m < 2 || m > ICM_SLICE_MAX_M) return 0;
  const char* e = getenv("LSQ_B200_ICM_KERNEL");  // testing override
  if (e != nullptr && strcmp(e, "warp") == 0) return 0;
  if (e != nullptr && strcmp(e, "slice") == 0) return 1;
  // Measured on B200 (profiles/README.md, round 1): the slice kernel moves the table traffic from L2 to
  // shared memory as designed, but pays the per-vector bookkeeping once per 32-candidate slice and is
  // instruction-issue bound at 183 ms vs 136 ms for the warp kernel (1 M x 16 iterations).  Opt-in only.
  (void)n;
  return 0;
}

template <int M>
static int launch_icm_slice_m(const IcmParams& p, cudaStream_t st) {
  int dev = 0, sms = LSQ_NUM_SMS_HINT;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = (size_t)(M - 1) * LSQ_H * ICM_SLICE_W * 4;
  LSQ_CUDA(cudaFuncSetAttribute(icm_ils_slice_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  note_launch();
  icm_ils_slice_kernel<M><<<sms, SLICE_THREADS, smem, st>>>(p);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

int launch_icm_slice(IcmParams p, cudaStream_t st) {
  if (p.n == 0 || p.niters == 0) return LSQ_OK;
  set_alloc_stream(st);
  DevBuf<unsigned long long> wcodes, rbest;
  DevBuf<uint16_t> clean, wclean;
  DevBuf<int> act;
  LSQ_CUDA(wcodes.alloc(p.n));
  LSQ_CUDA(rbest.alloc(p.n));
  LSQ_CUDA(clean.alloc(p.n));
  LSQ_CUDA(wclean.alloc(p.n));
  LSQ_CUDA(act.alloc(p.n));
  p.wcodes = wcodes.p; p.rbest = rbest.p; p.clean = clean.p; p.wclean = wclean.p; p.act = act.p;
  switch (p.m) {
#define LSQ_CASE(MM) case MM: return launch_icm_slice_m<MM>(p, st);
    LSQ_CASE(2) LSQ_CASE(3) LSQ_CASE(4) LSQ_CASE(5) LSQ_CASE(6) LSQ_CASE(7) LSQ_CASE(8)
#undef LSQ_CASE
  }
  set_error("slice kernel: m must be in 2..8");
  return LSQ_ERR_ARG;
}

}  // namespace lsq
