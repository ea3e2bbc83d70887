// unary_tc.cu — tensor-core (tcgen05 / TMEM) build of the unary tables: the one GEMM-shaped piece of
// the path (get_unaries, utils.jl:94-122: U_i = -2 * C_i' * X + ||c||^2; cuBLAS sgemm in the
// reference GPU path, encode_icm_cuda.jl:92).
//
// FAST MODE, not parity mode: a tensor core cannot reproduce the sequential fp32 FMA chain the oracle
// freezes, so this kernel is opt-in and held to the tolerance criterion (quantisation error within
// 1e-5 relative; in practice the codes come out identical).  Precision: 3xTF32 — every fp32 operand is
// split into hi = its top 19 bits and lo = the (truncated) remainder, and
//     D = lo(C).hi(X) + hi(C).lo(X) + hi(C).hi(X)
// is accumulated in fp32 in TMEM; measured |dU|/|U| ~ 7e-7 (single-pass TF32: 6e-4, which moves
// 0.08 % of the codes and the error by 3e-5 — not good enough).
//
// Structure (one CTA per SM, 256 threads):
//   * a CTA owns one 128-candidate tile (codebook j, half ah) for its whole life; the tile's hi and lo
//     operands, pre-split and pre-arranged in the UMMA no-swizzle K-major core-matrix layout by
//     split_codebooks_kernel, arrive with TMA bulk copies (cp.async.bulk + mbarrier);
//   * per 64-vector tile of X: all threads stage hi/lo of X into shared memory in the same layout
//     (row-coalesced global reads; the K-direction core-matrix stride is padded by 16 B so the 16-byte
//     shared stores of a warp spread over all banks), fence.proxy.async, __syncthreads;
//   * one thread issues 3 x K/8 tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=64, K=8) and a
//     tcgen05.commit onto an mbarrier;
//   * 8 warps drain the 128 x 64 fp32 accumulator from TMEM (tcgen05.ld 32x32b.x32), apply
//     -2*acc + ||c||^2 and store 128-byte rows of U[j][v][a0 .. a0+127].
// CTAs that share a vector tile (the 2m candidate tiles) walk the vector tiles in the same order, so X
// is read from HBM once and from L2 otherwise.
#include "icm.cuh"

namespace lsq {

constexpr int TC_M = 128;      // candidates per tile (UMMA M)
constexpr int TC_N = 64;       // vectors per tile (UMMA N): hi+lo of both operands must fit 227 KB
constexpr int TC_KMAX = 128;   // max d
constexpr int TC_THREADS = 256;

// K-major, no swizzle: core matrix = 8 rows x 16 B, contiguous (128 B).  Row groups are SBO apart,
// K-adjacent core matrices LBO apart; LBO carries 16 B of padding (bank spreading for the staging stores).
__host__ __device__ constexpr uint32_t tc_sbo() { return 128u; }
__host__ __device__ constexpr uint32_t tc_lbo(int rows) { return (uint32_t)(rows / 8) * 128u + 16u; }
__host__ __device__ constexpr uint32_t tc_operand_bytes(int rows, int k) { return (uint32_t)(k / 4) * tc_lbo(rows); }
// byte offset of element (r, k) inside an operand
__host__ __device__ inline uint32_t tc_offset(int rows, int r, int k) {
  return (uint32_t)(k >> 2) * tc_lbo(rows) + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 3) * 4u;
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Csplit[jah][part][operand layout]: part 0 = hi, 1 = lo; jah = j*2 + half
__global__ void split_codebooks_kernel(const float* __restrict__ C, int m, int d, float* __restrict__ Csplit) {
  const int jah = blockIdx.x;
  const int j = jah >> 1, a0 = (jah & 1) * TC_M;
  const uint32_t opb = tc_operand_bytes(TC_M, d);
  char* base = reinterpret_cast<char*>(Csplit) + (size_t)jah * 2 * opb;
  for (int e = threadIdx.x; e < TC_M * d; e += blockDim.x) {
    const int r = e / d, k = e % d;
    const float x = C[((size_t)j * LSQ_H + a0 + r) * d + k];
    const float hi = tf32_hi(x);
    const float lo = tf32_hi(x - hi);
    const uint32_t off = tc_offset(TC_M, r, k);
    *reinterpret_cast<float*>(base + off) = hi;
    *reinterpret_cast<float*>(base + opb + off) = lo;
  }
  // the 16-byte pads between K chunks are never read by the tensor core
}

__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // cute::UMMA::SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1 layout[61,64)=0 (no swizzle)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

__global__ void __launch_bounds__(TC_THREADS, 1) unary_tc_kernel(const float* __restrict__ X, int d, int64_t n,
                                                                 const float* __restrict__ Csplit,
                                                                 const float* __restrict__ norms, int m,
                                                                 float* __restrict__ U, int ctas_per_tile) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar_a, bar_mma;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jah = blockIdx.x / ctas_per_tile;       // which 128-candidate tile
  const int sub = blockIdx.x % ctas_per_tile;       // which stride class of vector tiles
  const int j = jah >> 1, a0 = (jah & 1) * TC_M;
  const uint32_t a_bytes = tc_operand_bytes(TC_M, d), b_bytes = tc_operand_bytes(TC_N, d);
  unsigned char* sA = smem_raw;                      // hi then lo
  unsigned char* sB = smem_raw + 2 * a_bytes;        // hi then lo
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);

  if (tid == 0) {
    mbar_init(&bar_a, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {  // TMEM: TC_N fp32 accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)TC_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  // the CTA's candidate tile: hi + lo operands, already in operand layout -> one bulk load
  if (tid == 0) bulk_load_issue(sA, reinterpret_cast<const char*>(Csplit) + (size_t)jah * 2 * a_bytes, 2 * a_bytes, &bar_a);
  const float nrm = norms[j * LSQ_H + a0 + (warp & 3) * 32 + lane];  // this thread's candidate in the epilogue

  // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  const uint32_t lboA = tc_lbo(TC_M), lboB = tc_lbo(TC_N);
  const int ksteps = d / 8;
  const int64_t ntiles = (n + TC_N - 1) / TC_N;
  uint32_t phase = 0;
  mbar_wait(&bar_a, 0);

  for (int64_t vt = sub; vt < ntiles; vt += ctas_per_tile) {
    const int64_t v0 = vt * TC_N;
    // ---- stage X tile: hi / lo in operand layout; a warp reads one 512-byte row (d = 128) at a time ----
    const int chunks = d / 4;  // float4 chunks per row
    for (int e = tid; e < TC_N * chunks; e += TC_THREADS) {
      const int r = e / chunks, kc = e % chunks;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v0 + r < n) x = __ldg(reinterpret_cast<const float4*>(X + (size_t)(v0 + r) * d) + kc);
      float4 hi, lo;
      hi.x = tf32_hi(x.x); hi.y = tf32_hi(x.y); hi.z = tf32_hi(x.z); hi.w = tf32_hi(x.w);
      lo.x = tf32_hi(x.x - hi.x); lo.y = tf32_hi(x.y - hi.y); lo.z = tf32_hi(x.z - hi.z); lo.w = tf32_hi(x.w - hi.w);
      const uint32_t off = (uint32_t)kc * lboB + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
      *reinterpret_cast<float4*>(sB + off) = hi;
      *reinterpret_cast<float4*>(sB + b_bytes + off) = lo;
    }
    fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
    __syncthreads();

    // ---- one thread issues the MMAs: lo.hi + hi.lo + hi.hi ----
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t acc = 0;
#pragma unroll 1
      for (int pass = 0; pass < 3; pass++) {
        const uint32_t aoff = (pass == 0) ? a_bytes : 0u;   // pass 0: lo(C)
        const uint32_t boff = (pass == 1) ? b_bytes : 0u;   // pass 1: lo(X)
        for (int t = 0; t < ksteps; t++) {
          const uint64_t ad = tc_smem_desc(sA_u + aoff + (uint32_t)(2 * t) * lboA, lboA, tc_sbo());
          const uint64_t bd = tc_smem_desc(sB_u + boff + (uint32_t)(2 * t) * lboB, lboB, tc_sbo());
          tc_mma_tf32(tmem_d, ad, bd, idesc, acc);
          acc = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    mbar_wait(&bar_mma, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w drains lanes 32*(w%4) .. +31, columns 32*(w/4) .. +31 ----
    const int quad = warp & 3, col0 = (warp >> 2) * 32;
    float* ubase = U + ((size_t)j * n + v0) * LSQ_H + a0 + quad * 32 + lane;
    {
      uint32_t r[32];
      tc_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const int64_t v = v0 + col0 + i;
        if (v < n) ubase[(size_t)(col0 + i) * LSQ_H] = __fadd_rn(__fmul_rn(-2.0f, __uint_as_float(r[i])), nrm);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // accumulator drained and sB consumed: next tile may overwrite both
  }

  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)TC_N) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Pipelined, warp-specialised version (the default).  Same arithmetic as the serial kernel above, but:
//   * the stationary operand (the CTA's 128-candidate codebook tile, hi and lo) lives in TENSOR MEMORY for
//     the whole kernel (tcgen05.st once; tcgen05.mma takes A from TMEM), so shared memory only holds the
//     streaming operand: 3 stages of a 64-vector X tile (hi + lo, 64 KB each);
//   * one thread streams the raw X tiles (64 vectors = one contiguous 32 KB block of X) into a 3-slot ring
//     with TMA bulk copies (cp.async.bulk + mbarrier expect_tx), running up to 3 tiles ahead, so no warp
//     ever waits on a global load; 4 producer warps read a raw slot, split it into hi / lo and write the two
//     operand images of a stage (generic-proxy stores + fence.proxy.async + an mbarrier arrive per thread);
//     one thread issues the 3 x K/8 MMAs per tile and commits twice — onto the stage's "empty" barrier
//     (releases shared memory) and onto the accumulator's "full" barrier;
//   * two TMEM accumulators (2 x 64 columns) alternate, so the MMAs of tile t+1 run while 4 epilogue warps
//     drain tile t (tcgen05.ld 32x32b.x32 -> -2*acc + ||c||^2 -> 128-byte row segments of U).
// TMEM columns: [0, 128) hi(C), [128, 256) lo(C), [256, 320) and [320, 384) accumulators (512 allocated).
// ------------------------------------------------------------------------------------------------
constexpr int TCP_STAGES = 2;     // operand-image stages
constexpr int TCP_RAW = 3;        // raw X-tile ring slots
constexpr int TCP_PRODUCERS = 256; // threads that split raw tiles into operand images (ncu: with 128 they were busy 90 % of the time)
constexpr int TCP_THREADS = 128 + TCP_PRODUCERS + 64;  // warps 0-3 epilogue (TMEM lane quadrant = warp id), producers, MMA issuer, TMA
constexpr uint32_t TCP_COL_AHI = 0, TCP_COL_ALO = 128, TCP_COL_ACC = 256;

__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
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

// D[tmem] (+)= A[tmem] * B[smem].  Called by ALL lanes of the issuing warp in uniform control flow; the
// instruction itself runs on one elected lane (elect.sync inside the asm block).  Keeping the election out of
// the C++ control flow lets ptxas hold the descriptors in uniform registers — with `if (lane == 0)` around the
// loop it wraps every UTCHMMA in an ELECT / BRA.U.ANY uniformising loop (~12 SASS instructions per MMA, the
// issuing thread became the bottleneck of the pipeline).
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tc_commit_elected(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(TCP_THREADS, 1) unary_tc_pipe_kernel(const float* __restrict__ X, int d, int64_t n,
                                                                       const float* __restrict__ C,
                                                                       const float* __restrict__ norms, int m,
                                                                       float* __restrict__ U, int ctas_per_tile) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar_full[TCP_STAGES], bar_empty[TCP_STAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint64_t bar_raw_full[TCP_RAW], bar_raw_empty[TCP_RAW];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jah = blockIdx.x / ctas_per_tile;       // which 128-candidate tile
  const int sub = blockIdx.x % ctas_per_tile;       // which stride class of vector tiles
  const int j = jah >> 1, a0 = (jah & 1) * TC_M;
  const uint32_t b_bytes = tc_operand_bytes(TC_N, d);            // one of hi / lo of a stage
  const uint32_t stage_bytes = 2 * b_bytes;
  const uint32_t lboB = tc_lbo(TC_N);

  if (tid == 0) {
    for (int s = 0; s < TCP_STAGES; s++) { mbar_init(&bar_full[s], TCP_PRODUCERS); mbar_init(&bar_empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_empty[b], 4); }
    for (int s = 0; s < TCP_RAW; s++) { mbar_init(&bar_raw_full[s], 1); mbar_init(&bar_raw_empty[s], TCP_PRODUCERS); }
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

  // ---- stationary operand -> TMEM: thread = candidate row (lane of its warp's quadrant), 32 k per store ----
  if (warp < 4) {
    const float* crow = C + ((size_t)j * LSQ_H + a0 + warp * 32 + lane) * d;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < d; k0 += 32) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const float x = (k0 + i < d) ? __ldg(crow + k0 + i) : 0.0f;
        const float h = tf32_hi(x);
        hi[i] = __float_as_uint(h);
        lo[i] = __float_as_uint(tf32_hi(x - h));
      }
      tc_st32(lane_base + TCP_COL_AHI + (uint32_t)k0, hi);
      tc_st32(lane_base + TCP_COL_ALO + (uint32_t)k0, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const int64_t ntiles = (n + TC_N - 1) / TC_N;
  const int64_t my_tiles = (ntiles > sub) ? (ntiles - sub + ctas_per_tile - 1) / ctas_per_tile : 0;

  const uint32_t raw_bytes = (uint32_t)TC_N * (uint32_t)d * 4u;               // one raw tile: 64 rows of X, contiguous
  unsigned char* raw_base = smem_raw + (size_t)TCP_STAGES * stage_bytes;

  constexpr int W_MMA = (128 + TCP_PRODUCERS) / 32, W_TMA = W_MMA + 1;
  if (warp == W_TMA) {
    // ===================== TMA: raw X tiles, up to TCP_RAW tiles ahead =====================
    if (lane == 0) {
      for (int64_t t = 0; t < my_tiles; t++) {
        const int slot = (int)(t % TCP_RAW);
        mbar_wait(&bar_raw_empty[slot], (uint32_t)((t / TCP_RAW) & 1) ^ 1u);
        const int64_t v0 = (sub + t * ctas_per_tile) * TC_N;
        const int64_t valid = (n - v0 < TC_N) ? (n - v0) : TC_N;
        bulk_load_issue(raw_base + (size_t)slot * raw_bytes, X + (size_t)v0 * d, (uint32_t)valid * (uint32_t)d * 4u,
                        &bar_raw_full[slot]);
      }
    }
  } else if (warp >= 4 && warp < W_MMA) {
    // ===================== producers: raw X tile -> hi / lo operand images =====================
    const int ptid = tid - 128;
    const int chunks = d / 4;
    for (int64_t t = 0; t < my_tiles; t++) {
      const int s = (int)(t % TCP_STAGES);
      const int slot = (int)(t % TCP_RAW);
      const int64_t v0 = (sub + t * ctas_per_tile) * TC_N;
      const int valid = (int)((n - v0 < TC_N) ? (n - v0) : TC_N);
      mbar_wait(&bar_raw_full[slot], (uint32_t)((t / TCP_RAW) & 1));
      mbar_wait(&bar_empty[s], (uint32_t)((t / TCP_STAGES) & 1) ^ 1u);
      unsigned char* sB = smem_raw + (size_t)s * stage_bytes;
      const float4* raw = reinterpret_cast<const float4*>(raw_base + (size_t)slot * raw_bytes);
      auto put = [&](int r, int kc) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < valid) x = raw[r * chunks + kc];
        float4 hi, lo;
        hi.x = tf32_hi(x.x); hi.y = tf32_hi(x.y); hi.z = tf32_hi(x.z); hi.w = tf32_hi(x.w);
        lo.x = tf32_hi(x.x - hi.x); lo.y = tf32_hi(x.y - hi.y); lo.z = tf32_hi(x.z - hi.z); lo.w = tf32_hi(x.w - hi.w);
        const uint32_t off = (uint32_t)kc * lboB + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
        *reinterpret_cast<float4*>(sB + off) = hi;
        *reinterpret_cast<float4*>(sB + b_bytes + off) = lo;
      };
      if (chunks == 32) {  // d = 128: a warp walks one 512-byte row, no division
#pragma unroll
        for (int i = 0; i < TC_N * 32 / TCP_PRODUCERS; i++) put((ptid >> 5) + (TCP_PRODUCERS / 32) * i, ptid & 31);
      } else {
        for (int e = ptid; e < TC_N * chunks; e += TCP_PRODUCERS) put(e / chunks, e % chunks);
      }
      fence_proxy_async();       // this thread's generic-proxy stores -> visible to the tensor core
      mbar_arrive(&bar_full[s]);
      mbar_arrive(&bar_raw_empty[slot]);
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (whole warp, uniform; one elected lane issues) =====================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const int ksteps = d / 8;
    const uint32_t sB_u = smem_u32(smem_raw);
    // shared-memory descriptor = constant high word (SBO, version) + low word (start >> 4 | LBO << 16); the
    // start field advances by 2 K-chunks per MMA and never carries out of its 14 bits (addresses < 256 KB)
    const uint64_t desc_hi = (uint64_t)((tc_sbo() >> 4) | (1u << 14)) << 32;
    const uint32_t desc_lbo = (lboB >> 4) << 16;
    const uint32_t kstep_enc = (2u * lboB) >> 4;
    for (int64_t t = 0; t < my_tiles; t++) {
      const int s = (int)(t % TCP_STAGES);
      const int b = (int)(t & 1);
      mbar_wait(&bar_acc_empty[b], (uint32_t)((t >> 1) & 1) ^ 1u);
      mbar_wait(&bar_full[s], (uint32_t)((t / TCP_STAGES) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t dcol = tmem + TCP_COL_ACC + (uint32_t)b * TC_N;
      const uint32_t stage_u = sB_u + (uint32_t)s * stage_bytes;
      const uint32_t lo_hiX = desc_lbo | (stage_u >> 4), lo_loX = desc_lbo | ((stage_u + b_bytes) >> 4);
      // pass 0: lo(C).hi(X)   pass 1: hi(C).lo(X)   pass 2: hi(C).hi(X)   (small terms first)
#pragma unroll 4
      for (int k = 0; k < ksteps; k++)
        tc_mma_tf32_ts(dcol, tmem + TCP_COL_ALO + (uint32_t)(8 * k), desc_hi | (lo_hiX + (uint32_t)k * kstep_enc), idesc, k > 0);
#pragma unroll 4
      for (int k = 0; k < ksteps; k++)
        tc_mma_tf32_ts(dcol, tmem + TCP_COL_AHI + (uint32_t)(8 * k), desc_hi | (lo_loX + (uint32_t)k * kstep_enc), idesc, 1u);
#pragma unroll 4
      for (int k = 0; k < ksteps; k++)
        tc_mma_tf32_ts(dcol, tmem + TCP_COL_AHI + (uint32_t)(8 * k), desc_hi | (lo_hiX + (uint32_t)k * kstep_enc), idesc, 1u);
      tc_commit_elected(&bar_empty[s]);      // shared-memory stage may be refilled once these MMAs have read it
      tc_commit_elected(&bar_acc_full[b]);   // accumulator complete
    }
  } else {
    // ===================== epilogue: warp w drains TMEM lanes 32w .. 32w+31 =====================
    const float nrm = norms[j * LSQ_H + a0 + warp * 32 + lane];
    for (int64_t t = 0; t < my_tiles; t++) {
      const int b = (int)(t & 1);
      mbar_wait(&bar_acc_full[b], (uint32_t)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t v0 = (sub + t * ctas_per_tile) * TC_N;
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + TCP_COL_ACC + (uint32_t)b * TC_N;
      float* ubase = U + ((size_t)j * n + v0) * LSQ_H + a0 + warp * 32 + lane;
      uint32_t r0[32], r1[32];
      tc_ld32(taddr, r0);
      tc_ld32(taddr + 32u, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[b]);   // the values are in registers: the accumulator is free
#pragma unroll
      for (int i = 0; i < 32; i++)
        if (v0 + i < n) ubase[(size_t)i * LSQ_H] = __fadd_rn(__fmul_rn(-2.0f, __uint_as_float(r0[i])), nrm);
#pragma unroll
      for (int i = 0; i < 32; i++)
        if (v0 + 32 + i < n) ubase[(size_t)(32 + i) * LSQ_H] = __fadd_rn(__fmul_rn(-2.0f, __uint_as_float(r1[i])), nrm);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// U[m][n][256] (plain layout) with the tensor-core kernels.  d must be a multiple of 8 and <= 128.
// LSQ_B200_UNARY_TC=serial selects the unpipelined kernel (A/B comparison).
int build_unaries_tc(const float* dX, int d, int64_t n, const float* dC, int m, const float* dnorms, float* dU,
                     cudaStream_t st) {
  if (n == 0) return LSQ_OK;
  LSQ_CHECK_ARG(d % 8 == 0 && d <= TC_KMAX, "tensor-core unary build needs d % 8 == 0 and d <= 128");
  int dev = 0, sms = LSQ_NUM_SMS_HINT;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntile = 2 * m;                                   // 128-candidate tiles
  const int per = sms / ntile > 0 ? sms / ntile : 1;          // CTAs per candidate tile
  const uint32_t a_bytes = tc_operand_bytes(TC_M, d), b_bytes = tc_operand_bytes(TC_N, d);
  const char* mode = getenv("LSQ_B200_UNARY_TC");
  if (mode == nullptr || strcmp(mode, "serial") != 0) {
    const size_t smem = (size_t)TCP_STAGES * 2 * b_bytes + (size_t)TCP_RAW * TC_N * d * 4;
    LSQ_CUDA(cudaFuncSetAttribute(unary_tc_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    note_launch();
    unary_tc_pipe_kernel<<<ntile * per, TCP_THREADS, smem, st>>>(dX, d, n, dC, dnorms, m, dU, per);
    LSQ_CUDA(cudaGetLastError());
    return LSQ_OK;
  }
  set_alloc_stream(st);
  DevBuf<float> csplit;
  LSQ_CUDA(csplit.alloc((size_t)ntile * 2 * a_bytes / 4));
  note_launch();
  split_codebooks_kernel<<<ntile, 256, 0, st>>>(dC, m, d, csplit.p);
  const size_t smem = 2 * (size_t)a_bytes + 2 * (size_t)b_bytes;
  LSQ_CUDA(cudaFuncSetAttribute(unary_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  note_launch();
  unary_tc_kernel<<<ntile * per, TC_THREADS, smem, st>>>(dX, d, n, csplit.p, dnorms, m, dU, per);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

}  // namespace lsq
