// train.cu — SURVEY.md §8(f) rows 1-3 on the device: the train_lsq alternation (src/lsq/LSQ.jl:10-88)
// with X, codes, tables and unaries resident on the GPU, the norm codebook that closes it
// (LSQ.jl:68-84), and eval_recall (src/linscan/Linscan.jl:76-117).
//
//   lsq_train_lsq   RX = R'X -> update_codebooks(RX, B) -> C_i = R*C_i -> ilsiter x encoding_icm ->
//                   niter x { obj = qerror; update_codebooks(X, B); ilsiter x encoding_icm } ->
//                   norms of the decoded vectors -> 1-D k-means (h centres) -> (C, B, cbnorms, B_norms, obj)
//                   Equals looping lsq_update_codebooks / lsq_encoding_icm with ils_iter = 0, 1, 2, ...
//                   (tests/test_gpu_parity.py::test_train_lsq_equals_public_calls), without the
//                   512 B/vector re-upload of X every call (the reference re-sends X to its workers on
//                   each encoding_icm, encode_icm.jl:165-172).
//   k-means         Clustering.jl's kmeans (k-means++ seeding from Julia's global RNG) is third-party
//                   and unpinned; the stand-in is a deterministic 1-D Lloyd iteration: values sorted
//                   once, centres seeded at the (2j+1)/(2h) quantiles, clusters are contiguous runs of
//                   the sorted array delimited by the centre midpoints, means summed in float64 in a
//                   fixed order.  B_norms is then assigned with the quantize_norms rule (utils.jl:6-31),
//                   so train's B_norms == quantize_norms(B, C, cbnorms) exactly.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "cbupdate.cuh"
#include "icm.cuh"
#include "runtime.cuh"

namespace lsq {

// out[r][j] = sum_i in[r][i] * Mx[j*sj + i*si]   (FMA chain, ascending i).  One warp per row r.
//   RX = R'X      : rx_j = sum_i R(i,j) x_i, Julia column-major R(i,j) = Rm[j*d + i]  -> sj = d, si = 1
//   C_i = R * C_i : c'_j = sum_i R(j,i) c_i = Rm[i*d + j]                             -> sj = 1, si = d
__global__ void __launch_bounds__(256) rows_times_matrix_kernel(const float* __restrict__ in, int64_t rows, int d,
                                                                const float* __restrict__ Mx, int sj, int si,
                                                                float* __restrict__ out) {
  extern __shared__ float rowbuf[];  // 8 rows of d floats
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* xr = rowbuf + (size_t)wib * d;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t r = (int64_t)blockIdx.x * 8 + wib; r < rows; r += nwarps) {
    __syncwarp();
    for (int i = lane; i < d; i += 32) xr[i] = in[(size_t)r * d + i];
    __syncwarp();
    for (int j = lane; j < d; j += 32) {
      float acc = 0.0f;
      for (int i = 0; i < d; i++) acc = __fmaf_rn(xr[i], __ldg(Mx + (size_t)j * sj + (size_t)i * si), acc);
      out[(size_t)r * d + j] = acc;
    }
  }
}

static int launch_rows_times_matrix(const float* in, int64_t rows, int d, const float* Mx, int sj, int si, float* out,
                                    cudaStream_t st) {
  if (rows == 0) return LSQ_OK;
  const int64_t b = ceil_div(rows, 8);
  const unsigned grid = (unsigned)std::min<int64_t>(b, (int64_t)LSQ_NUM_SMS_HINT * 8);
  note_launch();
  rows_times_matrix_kernel<<<grid, 256, (size_t)8 * d * sizeof(float), st>>>(in, rows, d, Mx, sj, si, out);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

// dbnorms[i] = sum_j CB[j,i]^2, j ascending, fp32 (LSQ.jl:72-77; same arithmetic as quantize_norms)
__global__ void __launch_bounds__(128) decoded_norms_kernel(const uint8_t* __restrict__ codes, int64_t n,
                                                            const float* __restrict__ C, int d, int m,
                                                            float* __restrict__ norms) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const uint8_t* cp = codes + v * m;
  float nrm = 0.0f;
  for (int t = 0; t < d; t++) {
    float r = 0.0f;
    for (int k = 0; k < m; k++) r = __fadd_rn(r, __ldg(C + ((size_t)k * LSQ_H + cp[k]) * d + t));
    nrm = __fadd_rn(nrm, __fmul_rn(r, r));
  }
  norms[v] = nrm;
}

// ---- deterministic 1-D k-means over SORTED values -------------------------------------------------
// bounds[j] = first sorted index that belongs to cluster >= j  (bounds[0] = 0, bounds[h] = n): value v
// leaves cluster j-1 for cluster j when v > (c[j-1] + c[j]) / 2 (ties stay with the lower centre, like
// the first-minimum rule).  One thread per boundary, binary search.
__global__ void kmeans1d_bounds_kernel(const float* __restrict__ sorted, int64_t n, const float* __restrict__ cent,
                                       int h, int64_t* __restrict__ bounds, int* __restrict__ changed) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > h) return;
  int64_t b;
  if (j == 0) b = 0;
  else if (j == h) b = n;
  else {
    const float mid = __fmul_rn(0.5f, __fadd_rn(cent[j - 1], cent[j]));
    int64_t lo = 0, hi = n;  // first index with sorted[i] > mid
    while (lo < hi) {
      const int64_t md = (lo + hi) >> 1;
      if (sorted[md] > mid) hi = md; else lo = md + 1;
    }
    b = lo;
  }
  if (bounds[j] != b) { bounds[j] = b; *changed = 1; }
}

// cent[j] = mean of sorted[bounds[j] .. bounds[j+1]) in float64 (fixed order: thread-strided partials,
// then a tree); an empty cluster keeps its centre.  One CTA per cluster.
__global__ void __launch_bounds__(256) kmeans1d_means_kernel(const float* __restrict__ sorted,
                                                             const int64_t* __restrict__ bounds,
                                                             float* __restrict__ cent) {
  __shared__ double sm[256];
  const int j = blockIdx.x;
  const int64_t lo = bounds[j], hi = bounds[j + 1];
  double acc = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc += (double)sorted[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s >= 1; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && hi > lo) cent[j] = (float)(sm[0] / (double)(hi - lo));
}

__global__ void kmeans1d_seed_kernel(const float* __restrict__ sorted, int64_t n, int h, float* __restrict__ cent,
                                     int64_t* __restrict__ bounds) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j <= h) bounds[j] = -1;
  if (j < h) cent[j] = sorted[(int64_t)(((2 * (int64_t)j + 1) * n) / (2 * (int64_t)h))];
}

// values [n] (device, unsorted) -> centres [h] ascending (device).  maxiter = 100 like Clustering.jl.
static int kmeans1d_device(const float* dvals, int64_t n, int h, float* dcent, int maxiter, int* iters_out,
                           cudaStream_t st) {
  LSQ_CHECK_ARG(n >= 1 && h >= 1 && n < ((int64_t)1 << 31), "kmeans1d: need 1 <= n < 2^31");
  DevBuf<float> dsorted;
  DevBuf<int64_t> dbounds;
  DevBuf<int> dchanged;
  DevBuf<unsigned char> dtmp;
  LSQ_CUDA(dsorted.alloc(n));
  LSQ_CUDA(dbounds.alloc(h + 1));
  LSQ_CUDA(dchanged.alloc(1));
  size_t tmp_bytes = 0;
  LSQ_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, dvals, dsorted.p, (int)n, 0, 32, st));
  LSQ_CUDA(dtmp.alloc(tmp_bytes));
  LSQ_CUDA(cub::DeviceRadixSort::SortKeys(dtmp.p, tmp_bytes, dvals, dsorted.p, (int)n, 0, 32, st));
  const unsigned gb = (unsigned)ceil_div(h + 1, 128);
  note_launch();
  kmeans1d_seed_kernel<<<gb, 128, 0, st>>>(dsorted.p, n, h, dcent, dbounds.p);
  LSQ_CUDA(cudaGetLastError());
  int it = 0;
  for (; it < maxiter; it++) {
    LSQ_CUDA(cudaMemsetAsync(dchanged.p, 0, sizeof(int), st));
    note_launch();
    kmeans1d_bounds_kernel<<<gb, 128, 0, st>>>(dsorted.p, n, dcent, h, dbounds.p, dchanged.p);
    int hchanged = 0;
    LSQ_CUDA(cudaMemcpyAsync(&hchanged, dchanged.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LSQ_CUDA(cudaStreamSynchronize(st));
    if (!hchanged) break;  // assignments are a fixed point: the means would not move either
    note_launch();
    kmeans1d_means_kernel<<<h, 256, 0, st>>>(dsorted.p, dbounds.p, dcent);
    LSQ_CUDA(cudaGetLastError());
  }
  if (iters_out) *iters_out = it;
  return LSQ_OK;
}

// ---- eval_recall (Linscan.jl:76-117) --------------------------------------------------------------
// rank[q] = 1-based position of gnd[q] in query q's WHOLE list pred[q][0..ld) if it occurs exactly once there
// (`find` over the full column, Linscan.jl:91-98), else k+1; positions beyond k are misses (:112); hist[r]++.
__global__ void __launch_bounds__(256) recall_rank_kernel(const int32_t* __restrict__ gnd,
                                                          const int32_t* __restrict__ pred, int nq, int ld, int k,
                                                          int* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= nq) return;
  const int32_t g = gnd[q];
  int count = 0, first = ld;
  for (int i = lane; i < ld; i += 32)
    if (pred[(size_t)q * ld + i] == g) { count++; first = min(first, i); }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    count += __shfl_xor_sync(0xFFFFFFFFu, count, off);
    first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, off));
  }
  if (lane == 0) atomicAdd(&hist[(count == 1 && first < k) ? first : k], 1);  // integer counts: order-independent
}

// recall[i] = #{rank <= i+1} / nq  (float64 division of integer counts, like the reference's `./ nquery`)
__global__ void recall_curve_kernel(const int* __restrict__ hist, int k, int nq, double* __restrict__ recall) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  long long run = 0;
  for (int i = 0; i < k; i++) {
    run += hist[i];
    recall[i] = (double)run / (double)nq;
  }
}

// ---- the resident training state -------------------------------------------------------------------
struct TrainCtx {
  int d, m;
  int64_t n;
  cudaStream_t st;
  const float* dX;
  uint8_t* dcodes;
  float* dcost;
  float* dC;
  float* dnorms;   // [m][256] codeword norms
  float* dT;       // pair tables
  float* dU;       // unaries of one chunk
  int64_t chunk;   // vectors per chunk (U holds m*chunk*256 floats)
  int icmiter, npert, randord;
  uint64_t seed;
  uint64_t g0;     // global index of this shard's first vector
};

// `niters` ILS iterations (ils_iter0, ils_iter0+1, ...) over the whole set with the current codebooks
static int train_encode(const TrainCtx& T, uint32_t ils_iter0, int niters) {
  LSQ_TRY(build_norms(T.dC, T.d, T.m, T.dnorms, T.st));
  LSQ_TRY(build_tables(T.dC, T.d, T.m, T.dT, T.st));
  LSQ_TRY(launch_veccost(T.dX, T.d, T.n, T.dcodes, T.dC, T.m, T.dcost, T.st));  // prevcost, encode_icm.jl:147
  for (int64_t lo = 0; lo < T.n; lo += T.chunk) {
    const int64_t nc = std::min<int64_t>(T.chunk, T.n - lo);
    // LSQ_B200_UNARY=tc: tensor-core unaries (fast mode), exactly as the host encode path does (capi.cu)
    const char* um = getenv("LSQ_B200_UNARY");
    if (um != nullptr && strcmp(um, "tc") == 0 && T.d % 8 == 0 && T.d <= 128)
      LSQ_TRY(build_unaries_tc(T.dX + (size_t)lo * T.d, T.d, nc, T.dC, T.m, T.dnorms, T.dU, T.st));
    else
      LSQ_TRY(build_unaries(T.dX + (size_t)lo * T.d, T.d, nc, T.dC, T.m, T.dnorms, T.dU, 0, T.st));
    for (int it0 = 0; it0 < niters; it0 += ICM_MAX_ITERS_PER_LAUNCH) {
      const int nit = std::min(ICM_MAX_ITERS_PER_LAUNCH, niters - it0);
      IcmParams p;
      memset(&p, 0, sizeof(p));
      p.X = T.dX + (size_t)lo * T.d; p.C = T.dC; p.U = T.dU; p.T = T.dT;
      p.codes = T.dcodes + (size_t)lo * T.m; p.cost = T.dcost + lo;
      p.n = nc; p.seed = T.seed; p.g0 = T.g0 + (uint64_t)lo; p.ils_iter0 = ils_iter0 + (uint32_t)it0;
      p.d = T.d; p.m = T.m; p.icmiter = T.icmiter; p.npert = T.npert; p.niters = nit;
      for (int i = 0; i < nit; i++) {
        p.snap_of_iter[i] = -1;
        int32_t order[LSQ_MAXM];
        make_to_look_host(T.seed, ils_iter0 + (uint32_t)(it0 + i), T.m, T.randord, order);
        for (int k = 0; k < T.m; k++) p.orders[i][k] = (int8_t)order[k];
      }
      LSQ_TRY(launch_icm_warp(p, T.st));
    }
  }
  return LSQ_OK;
}

// float64 sum of the costs of this shard's current codes
static int train_cost_sum(const TrainCtx& T, double* dsum, double* out) {
  LSQ_TRY(launch_veccost(T.dX, T.d, T.n, T.dcodes, T.dC, T.m, T.dcost, T.st));
  LSQ_TRY(launch_sum_f32_to_f64(T.dcost, T.n, dsum, T.st));
  LSQ_CUDA(cudaMemcpyAsync(out, dsum, sizeof(double), cudaMemcpyDeviceToHost, T.st));
  LSQ_CUDA(cudaStreamSynchronize(T.st));
  return LSQ_OK;
}

}  // namespace lsq

using namespace lsq;

extern "C" {

// The training set is sharded over the bound devices (splitarray, contiguous): every device keeps its shard of
// X, the codes, the unaries and a replica of the tables resident for the whole alternation.  Per outer
// iteration the devices exchange exactly one buffer — the integer codebook-update statistics (one all-reduce)
// — and one float64 per device for the objective; the solve is replicated (deterministic kernels on identical
// inputs give identical codebooks, no broadcast).  The statistics are exact integers and the schedule is
// keyed by the global vector index, so the result is bit-identical for any number of devices.
int lsq_train_lsq(const float* X, int d, int64_t n, int m, int h, const float* R, int16_t* B, float* C, int niter,
                  int ilsiter, int icmiter, int randord, int npert, uint64_t seed, float* cbnorms,
                  int16_t* B_norms, float* obj, int verbose) {
  LSQ_CHECK_ARG(d >= 1 && n >= 1, "train_lsq: need d >= 1, n >= 1");
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM, "m must be in 1..16 (cudautils.cu:38)");
  LSQ_CHECK_ARG(h == LSQ_H, "h must be 256 (cudautils.cu:245)");
  LSQ_CHECK_ARG(niter >= 0 && ilsiter >= 0 && icmiter >= 0, "iteration counts must be >= 0");
  LSQ_CHECK_ARG(npert >= 0 && npert <= m, "npert must be in 0..m (sample without replacement, encode_icm.jl:58)");
  LSQ_CHECK_ARG(X != nullptr && B != nullptr && C != nullptr, "X, B and C are required");
  LSQ_TRY(rt_ensure_init());
  const int64_t mh = (int64_t)m * h;
  const size_t slen = (size_t)(mh * (mh + d));
  if (verbose) {
    // LSQ.jl:25-29 prints this banner unconditionally; here it follows V
    fprintf(stderr, "Doing local search with %d codebooks, %d perturbations, %d icm iterations and random order = %s\n",
            m, npert, icmiter, randord ? "true" : "false");
  }
  const int k = rt_devices_for(n, 4096);
  AllReduceGroup* grp = nullptr;
  if (k > 1) {
    grp = rt_allreduce_group(k);
    if (grp == nullptr) return LSQ_ERR_CUDA;
    if (verbose) fprintf(stderr, "[lsq_b200] train_lsq on %d devices, statistics all-reduce: %s\n", k, rt_allreduce_backend(grp));
  }
  PhaseSync sync(k);
  std::vector<float> hmax(k, 0.0f);
  std::vector<double> hsum(k, 0.0);
  std::vector<float> hnorms, hcent(h, 0.0f);
  if (cbnorms != nullptr || B_norms != nullptr) hnorms.resize(n);
  const bool want_norms = !hnorms.empty();

  return rt_parallel(k, [&](int r) -> int {
    int rc = rt_bind(r);
    const cudaStream_t st = rt_ctx(r).st;
    int64_t lo = 0, hi = n;
    lsq_splitarray(n, k, r, &lo, &hi);
    const int64_t nl = hi - lo;
#define STEP(expr) do { if (rc == LSQ_OK) rc = (expr); } while (0)
#define STEP_CUDA(expr) do { if (rc == LSQ_OK) { cudaError_t e__ = (expr); if (e__ != cudaSuccess) rc = cuda_fail(e__, #expr, __FILE__, __LINE__); } } while (0)
#define SYNC() do { if (!sync.ok(rc)) return rc != LSQ_OK ? rc : LSQ_ERR_CUDA; } while (0)

    DevBuf<float> dX, dRX, dRm, dC, dC2, dnorms, dT, dU, dcost, dvn, dcent, dmax;
    DevBuf<int16_t> d16;
    DevBuf<uint8_t> dcodes;
    DevBuf<int64_t> dS, dS2;
    DevBuf<double> dG, dRhs, dsum;
    DevBuf<int> derr;
    int herr = 0;
    STEP_CUDA(dX.alloc((size_t)nl * d));
    STEP_CUDA(dC.alloc((size_t)mh * d));
    STEP_CUDA(dnorms.alloc(mh));
    STEP_CUDA(dT.alloc((size_t)m * m * LSQ_H * LSQ_H));
    STEP_CUDA(dcost.alloc(nl));
    STEP_CUDA(d16.alloc((size_t)nl * m));
    STEP_CUDA(dcodes.alloc((size_t)nl * m));
    STEP_CUDA(dS.alloc(slen));
    if (k > 1 && !rt_allreduce_is_p2p(grp)) STEP_CUDA(dS2.alloc(slen));
    STEP_CUDA(dG.alloc((size_t)mh * mh));
    STEP_CUDA(dRhs.alloc((size_t)mh * d));
    STEP_CUDA(dsum.alloc(1025));
    STEP_CUDA(derr.alloc(1));
    STEP_CUDA(dmax.alloc(1));
    STEP_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), st));
    STEP(rt_h2d(dX.p, X + (size_t)lo * d, (size_t)nl * d * 4, st));
    STEP(rt_h2d(d16.p, B + (size_t)lo * m, (size_t)nl * m * 2, st));
    STEP(launch_codes_i16_to_u8(d16.p, dcodes.p, nl * m, derr.p, st));
    STEP_CUDA(cudaMemcpyAsync(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    STEP_CUDA(cudaStreamSynchronize(st));
    if (rc == LSQ_OK && herr != 0) { set_error("invalid argument: codes must be 1-based in 1..256"); rc = LSQ_ERR_ARG; }

    // unaries: as many vectors per chunk as fit comfortably (all of them for the usual training sets)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
    int64_t chunk = (int64_t)(0.6 * (double)free_b / ((double)m * LSQ_H * 4));
    chunk = std::max<int64_t>(1024, std::min<int64_t>(chunk, nl));
    if (const char* ce = getenv("LSQ_B200_CHUNK_VECTORS")) chunk = std::max<int64_t>(1, std::min<int64_t>(atoll(ce), nl));  // tests
    STEP_CUDA(dU.alloc((size_t)chunk * mh));

    TrainCtx T;
    T.d = d; T.m = m; T.n = nl; T.g0 = (uint64_t)lo; T.st = st; T.dX = dX.p; T.dcodes = dcodes.p; T.dcost = dcost.p;
    T.dC = dC.p; T.dnorms = dnorms.p; T.dT = dT.p; T.dU = dU.p; T.chunk = chunk;
    T.icmiter = icmiter; T.npert = npert; T.randord = randord; T.seed = seed;

    const bool fused = (k > 1) && rt_allreduce_is_p2p(grp);
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;   // device time of the exchange + finalize on the primary device
    struct EvGuard { cudaEvent_t& a; cudaEvent_t& b; ~EvGuard() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evg{ev_a, ev_b};
    if (r == 0) { STEP_CUDA(cudaEventCreate(&ev_a)); STEP_CUDA(cudaEventCreate(&ev_b)); }
    // scale exponent of the fixed-point statistics for data source `src` (global max|x| over the shards)
    int scale_exp = 0;
    auto exchange_scale = [&](const float* src) -> int {
      STEP_CUDA(cudaMemsetAsync(dmax.p, 0, sizeof(float), st));
      STEP(cb_absmax(src, nl * d, dmax.p, st));
      STEP_CUDA(cudaMemcpyAsync(&hmax[r], dmax.p, sizeof(float), cudaMemcpyDeviceToHost, st));
      STEP_CUDA(cudaStreamSynchronize(st));
      SYNC();
      float mx = 0.0f;
      for (int i = 0; i < k; i++) mx = (hmax[i] > mx || hmax[i] != hmax[i]) ? hmax[i] : mx;
      scale_exp = cb_scale_exp(mx, n);
      return LSQ_OK;
    };
    // update_codebooks(src, codes): local statistics -> ONE all-reduce -> replicated solve into dC
    auto update = [&](const float* src) -> int {
      STEP_CUDA(cudaMemsetAsync(dS.p, 0, slen * sizeof(int64_t), st));
      STEP(cb_accumulate(src, d, nl, dcodes.p, m, scale_exp, dS.p, st));
      SYNC();
      // the one exchange per outer iteration: NCCL all-reduce + finalize, or the fused peer-memory form in which
      // the finalize kernel itself reads and sums every device's statistics (no intermediate buffer or copy)
      if (r == 0) STEP_CUDA(cudaEventRecord(ev_a, st));
      if (fused) {
        PeerPtrs peers;
        rc = rt_peer_begin(grp, r, dS.p, st, &sync.bar, &peers);
        STEP(cb_finalize_peers(peers, k, m, d, scale_exp, dG.p, dRhs.p, st));
        const int rc2 = rt_peer_end(grp, r, st, &sync.bar);
        if (rc == LSQ_OK) rc = rc2;
      } else {
        if (k > 1) rc = rt_allreduce_sum_i64(grp, r, dS.p, dS2.p, slen, st, &sync.bar);
        STEP(cb_finalize(dS.p, m, d, scale_exp, dG.p, dRhs.p, st));
      }
      if (r == 0) STEP_CUDA(cudaEventRecord(ev_b, st));
      int iters = 0;
      STEP(cb_solve(dG.p, dRhs.p, m, d, dC.p, 0, 0.0, &iters, st));
      if (verbose && r == 0 && rc == LSQ_OK) fprintf(stderr, "[lsq_b200] codebook update: CG converged in %d iterations\n", iters);
      if (r == 0 && rc == LSQ_OK) rt_note_collective(ev_a, ev_b);   // (cb_solve synchronised the stream)
      SYNC();
      return LSQ_OK;
    };
    // objective over all shards: one float64 per device, added in shard order
    float q = 0.0f;
    auto qerror_all = [&]() -> int {
      STEP(train_cost_sum(T, dsum.p, &hsum[r]));
      SYNC();
      double tot = 0.0;
      for (int i = 0; i < k; i++) tot += hsum[i];
      q = (float)(tot / (double)n);
      return LSQ_OK;
    };
#define RUN(call) do { const int rr__ = (call); if (rr__ != LSQ_OK) return rr__; } while (0)

    // C = update_codebooks(R'X, B); C_i = R*C_i  (LSQ.jl:31-41)
    if (R != nullptr) {
      STEP_CUDA(dRX.alloc((size_t)nl * d));
      STEP_CUDA(dRm.alloc((size_t)d * d));
      STEP_CUDA(dC2.alloc((size_t)mh * d));
      STEP_CUDA(cudaMemcpyAsync(dRm.p, R, (size_t)d * d * 4, cudaMemcpyHostToDevice, st));
      STEP(launch_rows_times_matrix(dX.p, nl, d, dRm.p, d, 1, dRX.p, st));
      RUN(exchange_scale(dRX.p));
      RUN(update(dRX.p));
      STEP(launch_rows_times_matrix(dC.p, mh, d, dRm.p, 1, d, dC2.p, st));
      STEP_CUDA(cudaMemcpyAsync(dC.p, dC2.p, (size_t)mh * d * 4, cudaMemcpyDeviceToDevice, st));
      RUN(exchange_scale(dX.p));
    } else {
      RUN(exchange_scale(dX.p));
      RUN(update(dX.p));
    }
    if (verbose) { RUN(qerror_all()); if (r == 0) fprintf(stderr, "%3d %e \n", -2, q); }

    // Initialize B (LSQ.jl:44-49)
    uint32_t ils_count = 0;
    STEP(train_encode(T, ils_count, ilsiter));
    ils_count += (uint32_t)ilsiter;
    if (verbose) { RUN(qerror_all()); if (r == 0) fprintf(stderr, "%3d %e \n", -1, q); }

    for (int iter = 0; iter < niter; iter++) {
      RUN(qerror_all());  // LSQ.jl:55
      if (r == 0 && obj) obj[iter] = q;
      if (verbose && r == 0) fprintf(stderr, "%3d %e \n", iter + 1, q);
      RUN(update(dX.p));
      STEP(train_encode(T, ils_count, ilsiter));
      ils_count += (uint32_t)ilsiter;
    }

    // norm codebook (LSQ.jl:68-84): norms of all shards -> one 1-D k-means on the primary device -> every
    // shard quantises its own norms with the shared centres
    if (want_norms) {
      STEP_CUDA(dvn.alloc(nl));
      STEP_CUDA(dcent.alloc(h));
      if (rc == LSQ_OK) {
        note_launch();
        decoded_norms_kernel<<<(unsigned)ceil_div(nl, 128), 128, 0, st>>>(dcodes.p, nl, dC.p, d, m, dvn.p);
        STEP_CUDA(cudaGetLastError());
      }
      STEP_CUDA(cudaMemcpyAsync(hnorms.data() + lo, dvn.p, (size_t)nl * 4, cudaMemcpyDeviceToHost, st));
      STEP_CUDA(cudaStreamSynchronize(st));
      SYNC();
      if (r == 0) {
        DevBuf<float> dall;
        const float* src = dvn.p;
        if (k > 1) {
          STEP_CUDA(dall.alloc(n));
          STEP_CUDA(cudaMemcpyAsync(dall.p, hnorms.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
          src = dall.p;
        }
        int kit = 0;
        STEP(kmeans1d_device(src, n, h, dcent.p, 100, &kit, st));
        if (verbose && rc == LSQ_OK) fprintf(stderr, "[lsq_b200] norm codebook: 1-D k-means stopped after %d iterations\n", kit);
        STEP_CUDA(cudaMemcpyAsync(hcent.data(), dcent.p, (size_t)h * 4, cudaMemcpyDeviceToHost, st));
        STEP_CUDA(cudaStreamSynchronize(st));
        if (cbnorms && rc == LSQ_OK) memcpy(cbnorms, hcent.data(), (size_t)h * 4);
      }
      SYNC();
      if (B_norms) {
        if (r != 0) STEP_CUDA(cudaMemcpyAsync(dcent.p, hcent.data(), (size_t)h * 4, cudaMemcpyHostToDevice, st));
        STEP(launch_quantize_norms(dcodes.p, nl, dC.p, d, m, dcent.p, h, d16.p, st));
        STEP_CUDA(cudaMemcpyAsync(B_norms + lo, d16.p, (size_t)nl * 2, cudaMemcpyDeviceToHost, st));
        STEP_CUDA(cudaStreamSynchronize(st));
      }
    }

    STEP(launch_codes_u8_to_i16(dcodes.p, d16.p, nl * m, st));
    STEP_CUDA(cudaMemcpyAsync(B + (size_t)lo * m, d16.p, (size_t)nl * m * 2, cudaMemcpyDeviceToHost, st));
    if (r == 0) STEP_CUDA(cudaMemcpyAsync(C, dC.p, (size_t)mh * d * 4, cudaMemcpyDeviceToHost, st));
    STEP_CUDA(cudaStreamSynchronize(st));
    SYNC();  // no device frees its buffers while a peer might still be inside the last collective
    return LSQ_OK;
#undef STEP
#undef STEP_CUDA
#undef SYNC
#undef RUN
  });
}

int lsq_kmeans1d(const float* values, int64_t n, int h, int maxiter, float* centers, int* iters_out) {
  LSQ_CHECK_ARG(values != nullptr && centers != nullptr, "values and centers are required");
  LSQ_CHECK_ARG(n >= 1 && h >= 1 && maxiter >= 0, "kmeans1d: bad sizes");
  cudaStream_t st;
  LSQ_TRY(host_ctx(&st));
  DevBuf<float> dv, dc;
  LSQ_CUDA(dv.alloc(n));
  LSQ_CUDA(dc.alloc(h));
  LSQ_CUDA(cudaMemcpyAsync(dv.p, values, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  LSQ_TRY(kmeans1d_device(dv.p, n, h, dc.p, maxiter, iters_out, st));
  LSQ_CUDA(cudaMemcpyAsync(centers, dc.p, (size_t)h * 4, cudaMemcpyDeviceToHost, st));
  LSQ_CUDA(cudaStreamSynchronize(st));
  return LSQ_OK;
}

int lsq_dev_eval_recall(const int32_t* dgnd, const int32_t* dpred, int nq, int ld, int k, double* drecall,
                        void* stream) {
  LSQ_CHECK_ARG(nq >= 1 && k >= 1 && ld >= k, "eval_recall: need nq >= 1, 1 <= k <= row length");
  cudaStream_t st = (cudaStream_t)stream;
  set_alloc_stream(st);
  DevBuf<int> dhist;
  LSQ_CUDA(dhist.alloc(k + 1));
  LSQ_CUDA(cudaMemsetAsync(dhist.p, 0, (size_t)(k + 1) * sizeof(int), st));
  note_launch();
  recall_rank_kernel<<<(unsigned)ceil_div(nq, 8), 256, 0, st>>>(dgnd, dpred, nq, ld, k, dhist.p);
  note_launch();
  recall_curve_kernel<<<1, 32, 0, st>>>(dhist.p, k, nq, drecall);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

int lsq_eval_recall(const int32_t* ids_gnd, const int32_t* ids_predicted, int nq, int ld, int k, double* recall) {
  LSQ_CHECK_ARG(ids_gnd != nullptr && ids_predicted != nullptr && recall != nullptr, "null argument");
  LSQ_CHECK_ARG(nq >= 1 && k >= 1 && ld >= k, "eval_recall: need nq >= 1, 1 <= k <= row length");
  cudaStream_t st;
  LSQ_TRY(host_ctx(&st));
  DevBuf<int32_t> dg, dp;
  DevBuf<double> dr;
  LSQ_CUDA(dg.alloc(nq));
  LSQ_CUDA(dp.alloc((size_t)nq * ld));
  LSQ_CUDA(dr.alloc(k));
  LSQ_CUDA(cudaMemcpyAsync(dg.p, ids_gnd, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
  LSQ_CUDA(cudaMemcpyAsync(dp.p, ids_predicted, (size_t)nq * ld * 4, cudaMemcpyHostToDevice, st));
  LSQ_TRY(lsq_dev_eval_recall(dg.p, dp.p, nq, ld, k, dr.p, st));
  set_alloc_stream(st);
  LSQ_CUDA(cudaMemcpyAsync(recall, dr.p, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
  LSQ_CUDA(cudaStreamSynchronize(st));
  return LSQ_OK;
}

}  // extern "C"
