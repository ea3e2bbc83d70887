// cbupdate.cu — update_codebooks (codebook_update.jl:52-86) in normal-equation form.
//
// The reference solves min_K ||X - K*onehot(B)'||_F with d independent LSQR runs on the n-by-(m*h)
// one-hot matrix A (updatecb!, codebook_update.jl:8-46; sparsify_codes, utils.jl:50-69).  LSQR from
// x0 = 0 is, analytically, conjugate gradients on the normal equations A'A k = A'x and converges to
// the minimum-norm least-squares solution.  Here:
//   cb_accumulate : ONE pass over the shard: Gram = A'A (integer co-occurrence counts) and Rhs = A'X'
//              (64-bit fixed-point sums), both EXACT integers in one int64 buffer.  HBM-bound: reads
//              4d + m bytes per vector.  That buffer is the only thing a multi-GPU run has to all-reduce
//              (one collective per outer iteration), and because integer sums do not depend on the
//              order, the codebooks are bit-identical for any number of shards / GPUs.
//   cb_finalize : the summed integers -> float64 Gram / Rhs.
//   cb_solve : CG on Gram*K = Rhs from K0 = 0 in float64 for all d right-hand sides in lock-step
//              (per-column step sizes), i.e. the same Krylov iteration LSQR performs, run to a far
//              tighter tolerance than LSQR's sqrt(eps(Float32)).  Unused codes stay exactly zero, as
//              in the reference (Appendix A.13 of SURVEY.md).
#include "cbupdate.cuh"
#include "icm.cuh"
#include "runtime.cuh"

#include <math.h>

#include <algorithm>
#include <atomic>
#include <vector>

namespace lsq {

// ---- statistics: exact integer accumulation -------------------------------------------------------
// Layout of the statistics buffer S (int64): S[0 .. mh*mh) = co-occurrence counts (the Gram matrix A'A),
// S[mh*mh .. mh*(mh+d)) = per-code sums of x in FIXED POINT: every x is rounded once to q = rint(x * 2^e)
// (e = scale exponent, chosen from max|x| and the total number of vectors so that no sum can overflow) and
// the q are added as 64-bit integers.  Integer addition is associative, so the result is bit-identical for
// any thread schedule, any chunking, and any number of shards / GPUs — which float64 atomics are not (their
// last bit depends on the arrival order).  The one rounding per element is 2^-(62 - log2 n) relative to
// max|x| (2^-42 at n = 1 M), far below the float32 codebooks the solve produces.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ X, int64_t count, unsigned* __restrict__ out) {
  unsigned mx = 0;  // |x| as raw bits: for non-negative floats the integer order is the float order (NaN sorts last)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    mx = max(mx, __float_as_uint(X[i]) & 0x7FFFFFFFu);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, off));
  if ((threadIdx.x & 31) == 0 && mx != 0) atomicMax(out, mx);
}

int cb_absmax(const float* dX, int64_t count, float* dmax, cudaStream_t st) {
  if (count == 0) return LSQ_OK;
  const int64_t blocks = std::min<int64_t>(ceil_div(count, 256 * 8), (int64_t)LSQ_NUM_SMS_HINT * 8);
  note_launch();
  absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(dX, count, reinterpret_cast<unsigned*>(dmax));
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

// e with  n_total * max|x| * 2^e < 2^62
int cb_scale_exp(float absmax, int64_t n_total) {
  if (!(absmax > 0.0f) || !(absmax <= 3.4e38f)) return 0;  // all-zero, NaN or Inf data: any scale is as good
  int ex = 0;
  frexpf(absmax, &ex);  // absmax < 2^ex
  int nb = 0;
  while (((int64_t)1 << nb) < n_total && nb < 62) nb++;
  return 62 - nb - ex;
}

__global__ void __launch_bounds__(256) cb_accumulate_kernel(const float* __restrict__ X, int d, int64_t n,
                                                            const uint8_t* __restrict__ codes, int m, int scale_exp,
                                                            unsigned long long* __restrict__ S) {
  const int lane = threadIdx.x & 31;
  const int mh = m * LSQ_H;
  unsigned long long* cnt = S;
  unsigned long long* acc = S + (size_t)mh * mh;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t v = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); v < n; v += nwarps) {
    const int mycode = (lane < m) ? (int)codes[v * m + lane] : 0;
    // co-occurrence counts: all ordered pairs (i, j), including i == j (the diagonal = code counts)
    for (int p0 = 0; p0 < m * m; p0 += 32) {
      const int p = p0 + lane;
      const int i = (p < m * m) ? p / m : 0, j = (p < m * m) ? p % m : 0;
      const int bi = __shfl_sync(0xFFFFFFFFu, mycode, i);
      const int bj = __shfl_sync(0xFFFFFFFFu, mycode, j);
      if (p < m * m) atomicAdd(&cnt[(size_t)(i * LSQ_H + bi) * mh + (j * LSQ_H + bj)], 1ull);
    }
    // per-code fixed-point sums of x
    const float* x = X + (size_t)v * d;
    for (int t = lane; t < d; t += 32) {
      const unsigned long long q = (unsigned long long)__double2ll_rn(scalbn((double)__ldg(x + t), scale_exp));
      for (int i = 0; i < m; i++) {
        const int row = i * LSQ_H + __shfl_sync(0xFFFFFFFFu, mycode, i);
        atomicAdd(acc + (size_t)row * d + t, q);  // two's complement: unsigned add == signed add
      }
    }
  }
}

// Accumulates one shard into S (which the caller zeroes once per update).
int cb_accumulate(const float* dX, int d, int64_t n, const uint8_t* dcodes, int m, int scale_exp, int64_t* dS,
                  cudaStream_t st) {
  if (n == 0) return LSQ_OK;
  const int64_t blocks = std::min<int64_t>(ceil_div(n, 8), (int64_t)LSQ_NUM_SMS_HINT * 8);
  note_launch();
  cb_accumulate_kernel<<<(unsigned)blocks, 256, 0, st>>>(dX, d, n, dcodes, m, scale_exp,
                                                         reinterpret_cast<unsigned long long*>(dS));
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

__global__ void cb_finalize_kernel(const int64_t* __restrict__ S, int64_t ngram, int64_t total, int scale_exp,
                                   double* __restrict__ gram, double* __restrict__ rhs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (i < ngram) gram[i] = (double)S[i];
  else rhs[i - ngram] = scalbn((double)S[i], -scale_exp);
}

// Gram[mh][mh] / Rhs[mh][d] (float64) from the (summed) statistics buffer.
int cb_finalize(const int64_t* dS, int m, int d, int scale_exp, double* dGram, double* dRhs, cudaStream_t st) {
  const int64_t mh = (int64_t)m * LSQ_H;
  const int64_t total = mh * (mh + d);
  note_launch();
  cb_finalize_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(dS, mh * mh, total, scale_exp, dGram, dRhs);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

__global__ void __launch_bounds__(256) cb_finalize_peers_kernel(PeerPtrs peers, int k, int64_t ngram, int64_t total, int scale_exp,
                                                               double* __restrict__ gram, double* __restrict__ rhs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; 2 * i < total; i += stride) {
    // two int64 per thread and peer: 16-byte loads over NVLink (total = mh*(mh+d) is even: mh is a multiple of 256)
    longlong2 acc = make_longlong2(0, 0);
    for (int j = 0; j < k; j++) {
      const longlong2 v = *reinterpret_cast<const longlong2*>(peers.p[j] + 2 * i);
      acc.x += v.x; acc.y += v.y;
    }
    const int64_t e0 = 2 * i, e1 = 2 * i + 1;
    if (e0 < ngram) gram[e0] = (double)acc.x; else rhs[e0 - ngram] = scalbn((double)acc.x, -scale_exp);
    if (e1 < ngram) gram[e1] = (double)acc.y; else rhs[e1 - ngram] = scalbn((double)acc.y, -scale_exp);
  }
}

int cb_finalize_peers(const PeerPtrs& peers, int k, int m, int d, int scale_exp, double* dGram, double* dRhs, cudaStream_t st) {
  const int64_t mh = (int64_t)m * LSQ_H;
  const int64_t total = mh * (mh + d);
  note_launch();
  cb_finalize_peers_kernel<<<LSQ_NUM_SMS_HINT * 8, 256, 0, st>>>(peers, k, mh * mh, total, scale_exp, dGram, dRhs);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

// single-shard convenience: scale from this X alone
int cb_stats(const float* dX, int d, int64_t n, const uint8_t* dcodes, int m, double* dGram, double* dRhs,
             cudaStream_t st) {
  const int64_t mh = (int64_t)m * LSQ_H;
  DevBuf<int64_t> dS;
  DevBuf<float> dmax;
  dS.st = dmax.st = st;
  LSQ_CUDA(dS.alloc((size_t)(mh * (mh + d))));
  LSQ_CUDA(dmax.alloc(1));
  LSQ_CUDA(cudaMemsetAsync(dS.p, 0, (size_t)(mh * (mh + d)) * sizeof(int64_t), st));
  LSQ_CUDA(cudaMemsetAsync(dmax.p, 0, sizeof(float), st));
  LSQ_TRY(cb_absmax(dX, n * d, dmax.p, st));
  float hmax = 0.0f;
  LSQ_CUDA(cudaMemcpyAsync(&hmax, dmax.p, sizeof(float), cudaMemcpyDeviceToHost, st));
  LSQ_CUDA(cudaStreamSynchronize(st));
  const int e = cb_scale_exp(hmax, n);
  LSQ_TRY(cb_accumulate(dX, d, n, dcodes, m, e, dS.p, st));
  LSQ_TRY(cb_finalize(dS.p, m, d, e, dGram, dRhs, st));
  return LSQ_OK;
}

// ------------------------------------------------------------------------------------------------
// CG.  Work vectors are stored column-major per right-hand side: V[c][r], c < d, r < mh.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double v, double* sm) {
  const int tid = threadIdx.x;
  sm[tid] = v;
  __syncthreads();
  for (int s = 128; s >= 1; s >>= 1) {
    if (tid < s) sm[tid] += sm[tid + s];
    __syncthreads();
  }
  const double out = sm[0];
  __syncthreads();
  return out;
}

__global__ void __launch_bounds__(256) cg_init_kernel(const double* __restrict__ rhs, int mh, int d,
                                                      double* __restrict__ Xt, double* __restrict__ Rt,
                                                      double* __restrict__ Pt, double* __restrict__ rs,
                                                      double* __restrict__ rs0, int* __restrict__ done) {
  __shared__ double sm[256];
  const int c = blockIdx.x;
  double acc = 0.0;
  for (int r = threadIdx.x; r < mh; r += 256) {
    const double v = rhs[(size_t)r * d + c];
    Xt[(size_t)c * mh + r] = 0.0;
    Rt[(size_t)c * mh + r] = v;
    Pt[(size_t)c * mh + r] = v;
    acc += v * v;
  }
  const double tot = block_sum_256(acc, sm);
  if (threadIdx.x == 0) { rs[c] = tot; rs0[c] = tot; done[c] = (tot == 0.0) ? 1 : 0; }
}

// APt[z][c][r] = sum_{k in slice z} Pt[c][k] * G[k][r]   (G symmetric).  Tile 32 (c) x 64 (r), K step 16.
// The K range is split over gridDim.z CTAs (the 2048 x 128 output alone gives only 128 tiles, fewer than
// SMs, each looping over all of K: latency-bound); cg_step_kernel adds the partial products in slice
// order, so the result does not depend on scheduling.
constexpr int CG_KSPLIT = 4;
__global__ void __launch_bounds__(256) cg_gemm_kernel(const double* __restrict__ G, const double* __restrict__ Pt,
                                                      int mh, int d, double* __restrict__ APt_all) {
  __shared__ double Ps[16][32 + 1];
  __shared__ double Gs[16][64];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx: 4 r's, ty: 2 c's
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
  double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  const int kslice = mh / (int)gridDim.z;
  double* APt = APt_all + (size_t)blockIdx.z * mh * d;
  for (int k0 = (int)blockIdx.z * kslice; k0 < ((int)blockIdx.z + 1) * kslice; k0 += 16) {
    {  // Pt tile: 32 c x 16 k
      const int cc = tid >> 4, kk = tid & 15;
      for (int h = 0; h < 2; h++) {
        const int c = c0 + cc + 16 * h;
        Ps[kk][cc + 16 * h] = (c < d) ? Pt[(size_t)c * mh + k0 + kk] : 0.0;
      }
      // G tile: 16 k x 64 r
      for (int e = tid; e < 16 * 64; e += 256) {
        const int kk2 = e >> 6, rr = e & 63;
        Gs[kk2][rr] = G[(size_t)(k0 + kk2) * mh + r0 + rr];
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      const double p0 = Ps[kk][ty * 2], p1 = Ps[kk][ty * 2 + 1];
      const double g0 = Gs[kk][tx * 4], g1 = Gs[kk][tx * 4 + 1], g2 = Gs[kk][tx * 4 + 2], g3 = Gs[kk][tx * 4 + 3];
      acc[0][0] = fma(p0, g0, acc[0][0]); acc[0][1] = fma(p0, g1, acc[0][1]);
      acc[0][2] = fma(p0, g2, acc[0][2]); acc[0][3] = fma(p0, g3, acc[0][3]);
      acc[1][0] = fma(p1, g0, acc[1][0]); acc[1][1] = fma(p1, g1, acc[1][1]);
      acc[1][2] = fma(p1, g2, acc[1][2]); acc[1][3] = fma(p1, g3, acc[1][3]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int c = c0 + ty * 2 + i;
    if (c >= d) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) APt[(size_t)c * mh + r0 + tx * 4 + j] = acc[i][j];
  }
}

// one CTA per right-hand side: step sizes, updates, convergence flag
__global__ void __launch_bounds__(256) cg_step_kernel(int mh, double tol2, double* __restrict__ Xt,
                                                      double* __restrict__ Rt, double* __restrict__ Pt,
                                                      double* __restrict__ APt, int ksplit, int d,
                                                      double* __restrict__ rs,
                                                      const double* __restrict__ rs0, int* __restrict__ done) {
  __shared__ double sm[256];
  const int c = blockIdx.x;
  if (done[c]) return;
  double* x = Xt + (size_t)c * mh;
  double* r = Rt + (size_t)c * mh;
  double* p = Pt + (size_t)c * mh;
  double* ap = APt + (size_t)c * mh;
  // fold the K slices of the product in slice order (deterministic)
  for (int i = threadIdx.x; i < mh; i += 256) {
    double t = ap[i];
    for (int z = 1; z < ksplit; z++) t += ap[(size_t)z * mh * d + i];
    ap[i] = t;
  }
  double acc = 0.0;
  for (int i = threadIdx.x; i < mh; i += 256) acc += p[i] * ap[i];
  const double pAp = block_sum_256(acc, sm);
  if (!(pAp > 0.0)) {
    if (threadIdx.x == 0) done[c] = 1;
    return;
  }
  const double rsold = rs[c];
  const double alpha = rsold / pAp;
  acc = 0.0;
  for (int i = threadIdx.x; i < mh; i += 256) {
    x[i] += alpha * p[i];
    const double rn = r[i] - alpha * ap[i];
    r[i] = rn;
    acc += rn * rn;
  }
  const double rsnew = block_sum_256(acc, sm);
  const double beta = rsnew / rsold;
  for (int i = threadIdx.x; i < mh; i += 256) p[i] = r[i] + beta * p[i];
  if (threadIdx.x == 0) {
    rs[c] = rsnew;
    if (rsnew <= tol2 * rs0[c]) done[c] = 1;
  }
}

__global__ void cg_finish_kernel(const double* __restrict__ Xt, int mh, int d, float* __restrict__ Cout) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)mh * d) return;
  const int r = (int)(e / d), c = (int)(e % d);
  Cout[e] = (float)Xt[(size_t)c * mh + r];  // K2vec (utils.jl:72-87): C[i][:, b] = K[:, i*h + b]
}

int cb_solve(const double* dGram, const double* dRhs, int m, int d, float* dCout, int max_iter, double tol,
             int* iters_out, cudaStream_t st) {
  const int mh = m * LSQ_H;
  if (max_iter <= 0) max_iter = 4 * mh;
  if (!(tol > 0.0)) tol = 1e-9;
  DevBuf<double> Xt, Rt, Pt, APt, rs, rs0;
  DevBuf<int> done;
  LSQ_CUDA(Xt.alloc((size_t)mh * d));
  LSQ_CUDA(Rt.alloc((size_t)mh * d));
  LSQ_CUDA(Pt.alloc((size_t)mh * d));
  const int ksplit = (mh % (16 * CG_KSPLIT) == 0) ? CG_KSPLIT : 1;
  LSQ_CUDA(APt.alloc((size_t)ksplit * mh * d));
  LSQ_CUDA(rs.alloc(d));
  LSQ_CUDA(rs0.alloc(d));
  LSQ_CUDA(done.alloc(d));
  note_launch();
  cg_init_kernel<<<d, 256, 0, st>>>(dRhs, mh, d, Xt.p, Rt.p, Pt.p, rs.p, rs0.p, done.p);
  LSQ_CUDA(cudaGetLastError());
  std::vector<int> hdone(d);
  dim3 ggrid(mh / 64, (unsigned)ceil_div(d, 32), ksplit);
  int it = 0;
  for (; it < max_iter; it++) {
    note_launch();
    cg_gemm_kernel<<<ggrid, 256, 0, st>>>(dGram, Pt.p, mh, d, APt.p);
    note_launch();
    cg_step_kernel<<<d, 256, 0, st>>>(mh, tol * tol, Xt.p, Rt.p, Pt.p, APt.p, ksplit, d, rs.p, rs0.p, done.p);
    if ((it & 15) == 15) {
      LSQ_CUDA(cudaMemcpyAsync(hdone.data(), done.p, (size_t)d * sizeof(int), cudaMemcpyDeviceToHost, st));
      LSQ_CUDA(cudaStreamSynchronize(st));
      bool all = true;
      for (int c = 0; c < d; c++) all &= (hdone[c] != 0);
      if (all) { it++; break; }
    }
  }
  LSQ_CUDA(cudaGetLastError());
  note_launch();
  cg_finish_kernel<<<(unsigned)ceil_div((int64_t)mh * d, 256), 256, 0, st>>>(Xt.p, mh, d, dCout);
  LSQ_CUDA(cudaGetLastError());
  LSQ_CUDA(cudaStreamSynchronize(st));
  if (iters_out) *iters_out = it;
  return LSQ_OK;
}

// update_codebooks over the bound devices: splitarray shards of (X, B), local statistics, ONE all-reduce of
// the integer statistics, solve on the primary device.
static int update_codebooks_host(const float* X, int d, int64_t n, const int16_t* B, int m, int h, float* Cout,
                                 int verbose) {
  LSQ_TRY(rt_ensure_init());
  const int64_t mh = (int64_t)m * h;
  const size_t slen = (size_t)(mh * (mh + d));
  const int k = rt_devices_for(n, 4096);
  AllReduceGroup* grp = nullptr;
  if (k > 1) {
    grp = rt_allreduce_group(k);
    if (grp == nullptr) return LSQ_ERR_CUDA;
  }
  PhaseSync sync(k);
  HostBarrier& bar = sync.bar;
  std::vector<float> hmax(k, 0.0f);
  int scale_exp = 0;
  auto phase_ok = [&](int rc) { return sync.ok(rc); };
  return rt_parallel(k, [&](int r) -> int {
    int rc = rt_bind(r);
    cudaStream_t st = rt_ctx(r).st;
    int64_t lo = 0, hi = n;
    lsq_splitarray(n, k, r, &lo, &hi);
    const int64_t nl = hi - lo;
    DevBuf<float> dX, dC, dmax;
    DevBuf<int16_t> d16;
    DevBuf<uint8_t> dcodes;
    DevBuf<int64_t> dS, dS2;
    DevBuf<double> dG, dR;
    DevBuf<int> derr;
    int herr = 0;
    auto stage1 = [&]() -> int {
      LSQ_TRY(rc);
      LSQ_CUDA(dX.alloc((size_t)nl * d));
      LSQ_CUDA(d16.alloc((size_t)nl * m));
      LSQ_CUDA(dcodes.alloc((size_t)nl * m));
      LSQ_CUDA(dS.alloc(slen));
      if (k > 1 && !rt_allreduce_is_p2p(grp)) LSQ_CUDA(dS2.alloc(slen));
      LSQ_CUDA(dmax.alloc(1));
      LSQ_CUDA(derr.alloc(1));
      LSQ_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), st));
      LSQ_CUDA(cudaMemsetAsync(dmax.p, 0, sizeof(float), st));
      LSQ_CUDA(cudaMemsetAsync(dS.p, 0, slen * sizeof(int64_t), st));
      LSQ_TRY(rt_h2d(dX.p, X + (size_t)lo * d, (size_t)nl * d * 4, st));
      LSQ_TRY(rt_h2d(d16.p, B + (size_t)lo * m, (size_t)nl * m * 2, st));
      LSQ_TRY(launch_codes_i16_to_u8(d16.p, dcodes.p, nl * m, derr.p, st));
      LSQ_TRY(cb_absmax(dX.p, nl * d, dmax.p, st));
      LSQ_CUDA(cudaMemcpyAsync(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      LSQ_CUDA(cudaMemcpyAsync(&hmax[r], dmax.p, sizeof(float), cudaMemcpyDeviceToHost, st));
      LSQ_CUDA(cudaStreamSynchronize(st));
      LSQ_CHECK_ARG(herr == 0, "codes must be 1-based in 1..256");
      return LSQ_OK;
    };
    rc = stage1();
    if (!phase_ok(rc)) return rc != LSQ_OK ? rc : LSQ_ERR_CUDA;
    if (r == 0) {
      float mx = 0.0f;
      for (int i = 0; i < k; i++) mx = (hmax[i] > mx || hmax[i] != hmax[i]) ? hmax[i] : mx;
      scale_exp = cb_scale_exp(mx, n);
    }
    bar.wait();
    rc = cb_accumulate(dX.p, d, nl, dcodes.p, m, scale_exp, dS.p, st);
    if (!phase_ok(rc)) return rc != LSQ_OK ? rc : LSQ_ERR_CUDA;
    const bool fused = (k > 1) && rt_allreduce_is_p2p(grp);
    PeerPtrs peers;
    if (fused) rc = rt_peer_begin(grp, r, dS.p, st, &bar, &peers);
    else if (k > 1) rc = rt_allreduce_sum_i64(grp, r, dS.p, dS2.p, slen, st, &bar);
    auto stage3 = [&]() -> int {
      LSQ_TRY(rc);
      if (r == 0) {
        LSQ_CUDA(dG.alloc((size_t)mh * mh));
        LSQ_CUDA(dR.alloc((size_t)mh * d));
        LSQ_CUDA(dC.alloc((size_t)mh * d));
        if (fused) LSQ_TRY(cb_finalize_peers(peers, k, m, d, scale_exp, dG.p, dR.p, st));
        else LSQ_TRY(cb_finalize(dS.p, m, d, scale_exp, dG.p, dR.p, st));
        int iters = 0;
        LSQ_TRY(cb_solve(dG.p, dR.p, m, d, dC.p, 0, 0.0, &iters, st));
        if (verbose) fprintf(stderr, "[lsq_b200] codebook update: CG converged in %d iterations (%d device%s)\n", iters, k, k > 1 ? "s" : "");
        LSQ_CUDA(cudaMemcpyAsync(Cout, dC.p, (size_t)mh * d * 4, cudaMemcpyDeviceToHost, st));
      }
      LSQ_CUDA(cudaStreamSynchronize(st));
      return LSQ_OK;
    };
    rc = stage3();
    if (fused) { const int rc2 = rt_peer_end(grp, r, st, &bar); if (rc == LSQ_OK) rc = rc2; cudaStreamSynchronize(st); }
    phase_ok(rc);  // nobody frees its statistics while a peer may still be reading them
    return rc;
  });
}

}  // namespace lsq

using namespace lsq;

extern "C" {

int64_t lsq_cb_stats_len(int m, int d) {
  const int64_t mh = (int64_t)m * LSQ_H;
  return mh * (mh + d);
}

int lsq_cb_scale_exp(float absmax, int64_t n_total) { return cb_scale_exp(absmax, n_total); }

int lsq_dev_absmax(const float* dX, int64_t count, float* dmax, void* stream) {
  LSQ_CHECK_ARG(count >= 0 && dmax != nullptr, "absmax: bad arguments");
  return cb_absmax(dX, count, dmax, (cudaStream_t)stream);
}

int lsq_dev_cb_accumulate(const float* dX, int d, int64_t n, const uint8_t* dcodes, int m, int scale_exp,
                          int64_t* dstats, void* stream) {
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM && d >= 1 && n >= 0, "cb_accumulate: bad sizes");
  return cb_accumulate(dX, d, n, dcodes, m, scale_exp, dstats, (cudaStream_t)stream);
}

int lsq_dev_cb_finalize(const int64_t* dstats, int m, int d, int scale_exp, double* dGram, double* dRhs,
                        void* stream) {
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM && d >= 1, "cb_finalize: bad sizes");
  return cb_finalize(dstats, m, d, scale_exp, dGram, dRhs, (cudaStream_t)stream);
}

int lsq_dev_cb_stats(const float* dX, int d, int64_t n, const uint8_t* dcodes, int m, double* dGram, double* dRhs,
                     void* stream) {
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM && d >= 1 && n >= 0, "cb_stats: bad sizes");
  return cb_stats(dX, d, n, dcodes, m, dGram, dRhs, (cudaStream_t)stream);
}

int lsq_dev_cb_solve(const double* dGram, const double* dRhs, int m, int d, float* dCout, int max_iter, double tol,
                     int* iters_out, void* stream) {
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM && d >= 1, "cb_solve: bad sizes");
  set_alloc_stream((cudaStream_t)stream);
  return cb_solve(dGram, dRhs, m, d, dCout, max_iter, tol, iters_out, (cudaStream_t)stream);
}

int lsq_update_codebooks(const float* X, int d, int64_t n, const int16_t* B, int m, int h, float* Cout,
                         const char* method, int verbose) {
  // "Codebook update method unknown" (codebook_update.jl:59)
  LSQ_CHECK_ARG(method != nullptr && (strcmp(method, "lsqr") == 0 || strcmp(method, "lsmr") == 0),
                "Codebook update method unknown");
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM, "m must be in 1..16");
  LSQ_CHECK_ARG(h == LSQ_H, "h must be 256");
  LSQ_CHECK_ARG(d >= 1 && n >= 0, "bad sizes");
  return update_codebooks_host(X, d, n, B, m, h, Cout, verbose);
}

}  // extern "C"
