// common.cuh — shared device/host helpers of liblsq_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/lsq_b200.h"

#define LSQ_H 256      // codebook size; fixed like the reference GPU path (cudautils.cu:245)
#define LSQ_MAXM 16    // cudautils.cu:38
#define LSQ_NUM_SMS_HINT 148

namespace lsq {

void set_error(const std::string& msg);
void note_launch();  // counts the kernel launches this library issues (lsq_launch_count)
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define LSQ_CUDA(call)                                                     \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) return ::lsq::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define LSQ_CHECK_ARG(cond, msg)           \
  do {                                     \
    if (!(cond)) {                         \
      ::lsq::set_error(std::string("invalid argument: ") + msg); \
      return LSQ_ERR_ARG;                  \
    }                                      \
  } while (0)

#define LSQ_TRY(call)            \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != LSQ_OK) return rc__; \
  } while (0)

// RAII device buffer.  Stream-ordered allocation from the device's default memory pool, whose release
// threshold lsq_init raises to "never", so repeated calls reuse the same memory instead of paying
// cudaMalloc/cudaFree (milliseconds per GiB) every time.  Freed in stream order on destruction.
cudaStream_t alloc_stream();
void set_alloc_stream(cudaStream_t st);
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t count = 0;
  cudaStream_t st = nullptr;
  DevBuf() : st(alloc_stream()) {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFreeAsync(p, st); }
  cudaError_t alloc(size_t n) {
    if (p) { cudaFreeAsync(p, st); p = nullptr; }
    count = n;
    return cudaMallocAsync((void**)&p, (n ? n : 1) * sizeof(T), st);
  }
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10: the canonical schedule generator.  Must match oracle/lsq_oracle.c bit for bit.
// ------------------------------------------------------------------------------------------------
#define LSQ_STREAM_PERTURB 0u
#define LSQ_STREAM_ORDER 1u

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// word i of stream (seed, ils_iter, g, stream); callers that need several words of one block should
// use sched_block instead.
__host__ __device__ inline void sched_block(uint64_t seed, uint32_t ils_iter, uint64_t g, uint32_t stream,
                                            uint32_t block, uint32_t out[4]) {
  philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), ils_iter, (stream << 24) | block, (uint32_t)seed,
                (uint32_t)(seed >> 32), out);
}

// Perturbation of vector g: npert distinct slots (ascending) + values.  Returns the slots/values in
// slot-sorted order.  W = number of 4-word blocks needed = ceil(2*npert/4) <= 8.
__host__ __device__ inline void make_perturb_one(uint64_t seed, uint32_t ils_iter, uint64_t g, int m, int h,
                                                 int npert, uint8_t* slots, uint8_t* vals) {
  uint32_t w[2 * LSQ_MAXM];
  const int nblocks = (2 * npert + 3) >> 2;
  for (int b = 0; b < nblocks; b++) sched_block(seed, ils_iter, g, LSQ_STREAM_PERTURB, b, w + 4 * b);
  // partial Fisher-Yates over a packed nibble array (16 slots x 4 bits), register-only on the device
  uint64_t a = 0xFEDCBA9876543210ull;
  for (int i = 0; i < npert; i++) {
    const int j = i + (int)(w[i] % (uint32_t)(m - i));
    const uint64_t ai = (a >> (4 * i)) & 0xF, aj = (a >> (4 * j)) & 0xF;
    a &= ~((0xFull << (4 * i)) | (0xFull << (4 * j)));
    a |= (aj << (4 * i));
    if (j != i) a |= (ai << (4 * j));
  }
  // chosen set as a bitmask; ascending iteration gives the sorted order
  uint32_t mask = 0;
  for (int i = 0; i < npert; i++) mask |= 1u << ((a >> (4 * i)) & 0xF);
  int t = 0;
  for (int s = 0; s < m; s++)
    if (mask & (1u << s)) {
      slots[t] = (uint8_t)s;
      vals[t] = (uint8_t)(w[npert + t] % (uint32_t)h);
      t++;
    }
}

inline void make_to_look_host(uint64_t seed, uint32_t ils_iter, int m, int randord, int32_t* to_look) {
  for (int i = 0; i < m; i++) to_look[i] = i;
  if (!randord) return;
  uint32_t w[LSQ_MAXM];
  for (int b = 0; b < 4; b++) sched_block(seed, ils_iter, 0, LSQ_STREAM_ORDER, b, w + 4 * b);
  for (int i = m - 1, k = 0; i >= 1; i--, k++) {
    const int j = (int)(w[k] % (uint32_t)(i + 1));
    const int32_t t = to_look[i]; to_look[i] = to_look[j]; to_look[j] = t;
  }
}

// order-preserving float <-> uint32 map (total order equals float order for non-NaN; -0 < +0)
__host__ __device__ inline uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  const uint32_t u = __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ordered_to_float(uint32_t o) {
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// binds the device (lazily, like lsq_init) and returns the library's private stream used by the
// host-pointer API
int host_ctx(cudaStream_t* st);

// ------------------------------------------------------------------------------------------------
// TMA (1-D bulk async copy) + mbarrier wrappers.  SASS: UBLKCP / SYNCS.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy, completion signalled on `bar` (bytes % 16 == 0, both addresses 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  while (!mbar_try_wait(bar, phase)) {}
}
// whole-CTA helper: thread 0 arms `bar` and issues the copy in <= 32 KB pieces
__device__ __forceinline__ void bulk_load_issue(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  mbar_expect_tx(bar, bytes);
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t sz = (bytes - off < 32768u) ? (bytes - off) : 32768u;
    bulk_g2s((char*)dst + off, (const char*)src + off, sz, bar);
  }
}
#endif

}  // namespace lsq
