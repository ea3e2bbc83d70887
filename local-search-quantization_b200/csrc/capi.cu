// capi.cu — the extern "C" boundary (include/lsq_b200.h): argument checks, host<->device staging, and
// the orchestration of the kernels.  No compute happens on the host; without a GPU every compute call
// fails with LSQ_ERR_CUDA.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "icm.cuh"
#include "linscan.cuh"
#include "cbupdate.cuh"
#include "runtime.cuh"

namespace lsq {

static int check_encode_args(int d, int64_t n, int m, int h, int niter, int npert) {
  LSQ_CHECK_ARG(d >= 1, "d must be >= 1");
  LSQ_CHECK_ARG(n >= 0, "n must be >= 0");
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM, "m must be in 1..16 (cudautils.cu:38)");
  LSQ_CHECK_ARG(h == LSQ_H, "h must be 256 (cudautils.cu:245)");
  LSQ_CHECK_ARG(niter >= 0, "niter must be >= 0");
  LSQ_CHECK_ARG(npert >= 0 && npert <= m, "npert must be in 0..m (sample without replacement, encode_icm.jl:58)");
  return LSQ_OK;
}

// Number of vectors whose unaries (m KB each) fit comfortably in free device memory.
static int64_t unary_chunk_capacity(int m, int d) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
  const double budget = 0.6 * (double)free_b;
  const double per_vec = (double)m * LSQ_H * 4 + (double)d * 4 + 64;
  int64_t cap = (int64_t)(budget / per_vec);
  return cap < 1024 ? 1024 : cap;
}

// upload Int16 1-based codes and convert to uint8 0-based on the device
static int upload_codes(const int16_t* hB, int64_t count, uint8_t* d8, cudaStream_t st) {
  DevBuf<int16_t> d16;
  DevBuf<int> derr;
  LSQ_CUDA(d16.alloc(count));
  LSQ_CUDA(derr.alloc(1));
  LSQ_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), st));
  LSQ_CUDA(cudaMemcpyAsync(d16.p, hB, count * sizeof(int16_t), cudaMemcpyHostToDevice, st));
  LSQ_TRY(launch_codes_i16_to_u8(d16.p, d8, count, derr.p, st));
  int herr = 0;
  LSQ_CUDA(cudaMemcpyAsync(&herr, derr.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  LSQ_CUDA(cudaStreamSynchronize(st));
  LSQ_CHECK_ARG(herr == 0, "codes must be 1-based in 1..256");
  return LSQ_OK;
}

// The shared driver of lsq_encoding_icm[_sched] and lsq_encode_icm_cuda: uploads, builds tables,
// walks the base set in memory-bounded chunks, runs `total_iters` ILS iterations per chunk.
struct EncodeJob {
  const float* X; int d; int64_t n;
  const int16_t* B_in;
  const float* C; int m;
  int icmiter, npert, randord;
  uint64_t seed, g0; uint32_t ils_iter0; int total_iters;
  // explicit schedule (single iteration only) or null
  const int32_t* to_look; const uint8_t* slots; const int16_t* vals;
  // outputs
  int16_t* B_out;                 // final codes (may be null)
  const int64_t* ilsiters; int nr; int16_t* Bs; float* objs;  // snapshots (may be null/0)
  int nsplits; int verbose;
  // sharding over the bound devices: this job covers vectors [off, off + n) of a set of n_total
  int64_t n_total, off;
};

// Per-chunk device state of the encode pipeline.  Two slots alternate: while the GPU runs the ILS
// kernel of one chunk, the host stages the next chunk's X and codes into the other slot, so the
// host->device copies (which the reference pays up front, encode_icm_cuda.jl:79) hide behind compute.
struct EncodeSlot {
  cudaStream_t st = nullptr;
  DevBuf<float> dX, dU, dcost, dsnapcost;
  DevBuf<uint8_t> dcodes, dsnap, dslots, dvals;
  DevBuf<int16_t> d16;
  DevBuf<double> dsum;
  DevBuf<int> derr;
  int64_t lo = 0, nc = 0;
  bool busy = false;
};

static int slot_alloc(EncodeSlot& S, cudaStream_t st, int64_t maxchunk, int d, int m, int nr) {
  S.st = st;
  set_alloc_stream(st);
  // (DevBuf captures the allocation stream at construction: rebind each member to this slot's stream)
  S.dX.st = S.dU.st = S.dcost.st = S.dsnapcost.st = st;
  S.dcodes.st = S.dsnap.st = S.dslots.st = S.dvals.st = st;
  S.d16.st = st; S.dsum.st = st; S.derr.st = st;
  LSQ_CUDA(S.dX.alloc((size_t)maxchunk * d));
  LSQ_CUDA(S.dU.alloc((size_t)maxchunk * m * LSQ_H));
  LSQ_CUDA(S.dcost.alloc(maxchunk));
  LSQ_CUDA(S.dcodes.alloc((size_t)maxchunk * m));
  LSQ_CUDA(S.d16.alloc((size_t)maxchunk * m));
  LSQ_CUDA(S.dsum.alloc((size_t)(nr > 0 ? nr : 1) * 1025));
  LSQ_CUDA(S.derr.alloc(1));
  if (nr > 0) {
    LSQ_CUDA(S.dsnap.alloc((size_t)nr * maxchunk * m));
    LSQ_CUDA(S.dsnapcost.alloc((size_t)nr * maxchunk));
  }
  return LSQ_OK;
}

// One device's share of an encode call.  obj_sum[r] receives the float64 sum of the snapshot-r costs of this
// shard (the caller adds the shards in order and divides by n_total).
static int run_encode_job(const EncodeJob& J, int devidx, double* obj_sum) {
  LSQ_TRY(rt_bind(devidx));
  const cudaStream_t st0 = rt_ctx(devidx).st, st1 = rt_ctx(devidx).st2;
  const int d = J.d, m = J.m;
  const int64_t n = J.n;

  DevBuf<float> dC, dnorms, dT, dTs;
  LSQ_CUDA(dC.alloc((size_t)m * LSQ_H * d));
  LSQ_CUDA(dnorms.alloc((size_t)m * LSQ_H));
  LSQ_CUDA(dT.alloc((size_t)m * m * LSQ_H * LSQ_H));
  LSQ_CUDA(cudaMemcpyAsync(dC.p, J.C, (size_t)m * LSQ_H * d * sizeof(float), cudaMemcpyHostToDevice, st0));
  LSQ_TRY(build_norms(dC.p, d, m, dnorms.p, st0));
  LSQ_TRY(build_tables(dC.p, d, m, dT.p, st0));

  // snapshot map: ILS iteration i (1-based) -> first r with ilsiters[r] == i (encode_icm_cuda.jl:211-213)
  std::vector<int> snap_of(J.total_iters, -1);
  for (int i = 1; i <= J.total_iters; i++)
    for (int r = 0; r < J.nr; r++)
      if (J.ilsiters[r] == i) { snap_of[i - 1] = r; break; }
  std::vector<char> snap_used(J.nr > 0 ? J.nr : 1, 0);
  for (int i = 0; i < J.total_iters; i++)
    if (snap_of[i] >= 0) snap_used[snap_of[i]] = 1;
  for (int r = 0; r < J.nr; r++) obj_sum[r] = 0.0;

  // chunking: at least nsplits splitarray parts (encode_icm_cuda.jl:272), more if memory demands, and
  // a few parts for large inputs so that the copies of part i+1 overlap the kernels of part i
  const int64_t cap = unary_chunk_capacity(m, d) / 2;  // two slots
  int nparts = J.nsplits > 1 ? J.nsplits : 1;
  // Pipeline schedule.  Equal parts (measured on B200, 1 M vectors, static ICM split: 1 part 165.6 ms,
  // 2: 158.0, 4: 151.3, 8: 149.7, resident-data step 148.4) expose the first part's H2D copy and pay one
  // kernel tail + launch gap per part; with the dynamic ICM split the resident step dropped to 117 ms while
  // 8 equal parts stayed at 128 ms.  A GEOMETRIC schedule (each part twice the previous one) exposes only
  // the copy of a small first part and needs half as many launches: part i+1's copy still hides behind
  // part i's kernels because copying a vector is ~12x cheaper than encoding it.
  // LSQ_B200_PIPELINE_PARTS=k forces k equal parts (k = 1: no pipelining).
  std::vector<int64_t> cuts;  // part p = [cuts[p], cuts[p+1])
  const char* pe = getenv("LSQ_B200_PIPELINE_PARTS");
  bool geometric = (pe == nullptr) && n >= 65536 && nparts <= 4 && (n * 8 / 15 + 1) <= cap;
  if (geometric) {
    cuts = {0, n / 15, n / 15 + 2 * n / 15, n / 15 + 2 * n / 15 + 4 * n / 15, n};
    nparts = 4;
  } else {
    int pipe_parts = (int)std::min<int64_t>(8, n / 32768);
    if (pe) pipe_parts = std::max(1, atoi(pe));
    if (n >= 262144 && nparts < pipe_parts) nparts = pipe_parts;
    while (ceil_div(n, nparts) > cap) nparts++;
    if (n == 0) nparts = 1;
    for (int part = 0; part <= nparts; part++) {
      int64_t lo = n, hi = n;
      if (part < nparts) lsq_splitarray(n, nparts, part, &lo, &hi);
      cuts.push_back(lo);
    }
  }
  int64_t maxchunk = 1;
  for (int part = 0; part < nparts; part++) maxchunk = std::max(maxchunk, cuts[part + 1] - cuts[part]);
  maxchunk += 1;
  const int nslots = nparts > 1 ? 2 : 1;

  // the sliced layout is decided once for the job (all chunks have the same size up to one vector)
  const int sliced = icm_use_slices(m, maxchunk - 1);
  if (sliced) {
    LSQ_CUDA(dTs.alloc((size_t)m * (m - 1) * LSQ_H * LSQ_H));
    LSQ_TRY(build_sliced_tables(dT.p, m, dTs.p, st0));
  }
  const char* um = getenv("LSQ_B200_UNARY");  // "tc": tensor-core unaries (fast mode, tolerance-checked)
  const bool unary_tc = (um != nullptr && strcmp(um, "tc") == 0 && !sliced && d % 8 == 0 && d <= 128);

  int fin_err = 0;                                       // D2H targets of finish(): they must outlive `drain`
  std::vector<double> fin_sums(J.nr > 0 ? J.nr : 1, 0.0);
  EncodeSlot slots[2];
  // Declared after every device buffer of the job, so it is destroyed first: on ANY exit (also the error
  // returns below) both streams are drained before a buffer is freed or a host output buffer is left behind
  // with a copy still in flight.
  struct Drain {
    cudaStream_t a, b; cudaEvent_t ev = nullptr;
    ~Drain() { cudaStreamSynchronize(a); cudaStreamSynchronize(b); if (ev) cudaEventDestroy(ev); cudaGetLastError(); }
  } drain{st0, st1};
  LSQ_CUDA(cudaEventCreateWithFlags(&drain.ev, cudaEventDisableTiming));
  LSQ_CUDA(cudaEventRecord(drain.ev, st0));
  LSQ_TRY(slot_alloc(slots[0], st0, maxchunk, d, m, J.nr));
  if (nslots == 2) {
    LSQ_TRY(slot_alloc(slots[1], st1, maxchunk, d, m, J.nr));
    LSQ_CUDA(cudaStreamWaitEvent(st1, drain.ev, 0));
  }
  set_alloc_stream(st0);

  // copy results of a finished chunk back (D2H into pageable memory blocks the host: call it only
  // after the next chunk's work has been queued on the other stream)
  auto finish = [&](EncodeSlot& S) -> int {
    const int64_t lo = S.lo, nc = S.nc;
    int& herr = fin_err;
    std::vector<double>& sums = fin_sums;
    herr = 0;
    std::fill(sums.begin(), sums.end(), 0.0);
    LSQ_CUDA(cudaMemcpyAsync(&herr, S.derr.p, sizeof(int), cudaMemcpyDeviceToHost, S.st));
    if (J.B_out) {
      LSQ_TRY(launch_codes_u8_to_i16(S.dcodes.p, S.d16.p, nc * m, S.st));
      LSQ_CUDA(cudaMemcpyAsync(J.B_out + (size_t)lo * m, S.d16.p, (size_t)nc * m * sizeof(int16_t), cudaMemcpyDeviceToHost, S.st));
    }
    for (int r = 0; r < J.nr; r++) {
      if (!snap_used[r]) continue;
      LSQ_TRY(launch_codes_u8_to_i16(S.dsnap.p + (size_t)r * nc * m, S.d16.p, nc * m, S.st));
      LSQ_CUDA(cudaMemcpyAsync(J.Bs + ((size_t)r * J.n_total + J.off + lo) * m, S.d16.p, (size_t)nc * m * sizeof(int16_t), cudaMemcpyDeviceToHost, S.st));
      LSQ_TRY(launch_sum_f32_to_f64(S.dsnapcost.p + (size_t)r * nc, nc, S.dsum.p + (size_t)r * 1025, S.st));
      LSQ_CUDA(cudaMemcpyAsync(&sums[r], S.dsum.p + (size_t)r * 1025, sizeof(double), cudaMemcpyDeviceToHost, S.st));
    }
    LSQ_CUDA(cudaStreamSynchronize(S.st));
    S.busy = false;
    LSQ_CHECK_ARG(herr == 0, "codes must be 1-based in 1..256");
    for (int r = 0; r < J.nr; r++) obj_sum[r] += sums[r];
    return LSQ_OK;
  };

  for (int part = 0; part < nparts; part++) {
    const int64_t lo = cuts[part], hi = cuts[part + 1];
    const int64_t nc = hi - lo;
    if (nc <= 0) continue;
    EncodeSlot& S = slots[part % nslots];
    if (S.busy) LSQ_TRY(finish(S));
    cudaStream_t st = S.st;
    S.lo = lo; S.nc = nc;
    LSQ_CUDA(cudaMemsetAsync(S.derr.p, 0, sizeof(int), st));
    LSQ_TRY(rt_h2d(S.dX.p, J.X + (size_t)lo * d, (size_t)nc * d * sizeof(float), st));
    LSQ_TRY(rt_h2d(S.d16.p, J.B_in + (size_t)lo * m, (size_t)nc * m * sizeof(int16_t), st));
    LSQ_TRY(launch_codes_i16_to_u8(S.d16.p, S.dcodes.p, nc * m, S.derr.p, st));
    if (unary_tc) LSQ_TRY(build_unaries_tc(S.dX.p, d, nc, dC.p, m, dnorms.p, S.dU.p, st));
    else LSQ_TRY(build_unaries(S.dX.p, d, nc, dC.p, m, dnorms.p, S.dU.p, sliced, st));
    set_alloc_stream(st0);
    LSQ_TRY(launch_veccost(S.dX.p, d, nc, S.dcodes.p, dC.p, m, S.dcost.p, st));
    if (J.nr > 0) LSQ_CUDA(cudaMemsetAsync(S.dsnap.p, 0, (size_t)J.nr * nc * m, st));

    if (J.slots != nullptr && J.npert > 0) {
      // explicit perturbations: [n][npert] uint8 slots, int16 0-based values -> uint8
      // (range-checked by lsq_encoding_icm_sched before anything was queued)
      std::vector<uint8_t> v8((size_t)nc * J.npert);
      for (size_t i = 0; i < v8.size(); i++) v8[i] = (uint8_t)J.vals[(size_t)lo * J.npert + i];
      S.dslots.st = S.dvals.st = st;
      LSQ_CUDA(S.dslots.alloc(v8.size()));
      LSQ_CUDA(S.dvals.alloc(v8.size()));
      LSQ_CUDA(cudaMemcpyAsync(S.dslots.p, J.slots + (size_t)lo * J.npert, v8.size(), cudaMemcpyHostToDevice, st));
      LSQ_CUDA(cudaMemcpyAsync(S.dvals.p, v8.data(), v8.size(), cudaMemcpyHostToDevice, st));
      LSQ_CUDA(cudaStreamSynchronize(st));  // v8 dies at the end of this scope
    }

    for (int it0 = 0; it0 < J.total_iters; it0 += ICM_MAX_ITERS_PER_LAUNCH) {
      const int nit = std::min(ICM_MAX_ITERS_PER_LAUNCH, J.total_iters - it0);
      IcmParams p;
      memset(&p, 0, sizeof(p));
      p.X = S.dX.p; p.C = dC.p; p.U = S.dU.p; p.T = dT.p; p.Ts = dTs.p;
      p.codes = S.dcodes.p; p.cost = S.dcost.p;
      p.slots = (J.slots && J.npert > 0) ? S.dslots.p : nullptr;
      p.vals = (J.slots && J.npert > 0) ? S.dvals.p : nullptr;
      p.snap = J.nr > 0 ? S.dsnap.p : nullptr;
      p.snapcost = J.nr > 0 ? S.dsnapcost.p : nullptr;
      p.n = nc; p.seed = J.seed; p.g0 = J.g0 + (uint64_t)lo;
      p.ils_iter0 = J.ils_iter0 + (uint32_t)it0;
      p.d = d; p.m = m; p.icmiter = J.icmiter; p.npert = J.npert; p.niters = nit;
      for (int i = 0; i < nit; i++) {
        p.snap_of_iter[i] = (int16_t)snap_of[it0 + i];
        int32_t order[LSQ_MAXM];
        if (J.to_look) memcpy(order, J.to_look, sizeof(int32_t) * m);
        else make_to_look_host(J.seed, J.ils_iter0 + (uint32_t)(it0 + i), m, J.randord, order);
        for (int k = 0; k < m; k++) p.orders[i][k] = (int8_t)order[k];
      }
      LSQ_TRY(sliced ? launch_icm_slice(p, st) : launch_icm_warp(p, st));
      set_alloc_stream(st0);
    }
    S.busy = true;
    // now that this chunk is queued, collect the previous one from the other slot
    if (nslots == 2 && slots[(part + 1) % 2].busy) LSQ_TRY(finish(slots[(part + 1) % 2]));
    if (J.verbose) fprintf(stderr, "[lsq_b200] queued part %d/%d (%lld vectors)\n", part + 1, nparts, (long long)nc);
  }
  for (int s = 0; s < nslots; s++)
    if (slots[s].busy) LSQ_TRY(finish(slots[s]));
  LSQ_CUDA(cudaStreamSynchronize(st0));
  if (nslots == 2) LSQ_CUDA(cudaStreamSynchronize(st1));
  return LSQ_OK;
}

// Shards the job over the bound devices by the reference's splitarray rule (utils.jl:152-177), one worker
// thread per device, no communication: the schedule is keyed by the global vector index, so the codes do
// not depend on the number of devices.
static int run_encode(EncodeJob J) {
  LSQ_TRY(rt_ensure_init());
  J.n_total = J.n; J.off = 0;
  const int nr = J.nr > 0 ? J.nr : 1;
  std::vector<char> snap_used(nr, 0);
  for (int i = 1; i <= J.total_iters; i++)
    for (int r = 0; r < J.nr; r++)
      if (J.ilsiters[r] == i) { snap_used[r] = 1; break; }
  // a snapshot slot no iteration maps to (duplicate or zero iteration count) stays all-zero, like the
  // reference's preallocated Bs entries; every other slot is fully overwritten, so it is not touched here
  // (zeroing 16 MB of fresh host pages costs ~4 ms per million vectors)
  for (int r = 0; r < J.nr; r++)
    if (!snap_used[r] && J.Bs) memset(J.Bs + (size_t)r * J.n * J.m, 0, (size_t)J.n * J.m * sizeof(int16_t));
  const int k = rt_devices_for(J.n, 4096);
  std::vector<double> sums((size_t)k * nr, 0.0);
  LSQ_TRY(rt_parallel(k, [&](int r) -> int {
    EncodeJob S = J;
    int64_t lo = 0, hi = J.n;
    lsq_splitarray(J.n, k, r, &lo, &hi);
    S.n = hi - lo; S.off = lo; S.g0 = J.g0 + (uint64_t)lo;
    S.X = J.X + (size_t)lo * J.d;
    S.B_in = J.B_in + (size_t)lo * J.m;
    if (J.B_out) S.B_out = J.B_out + (size_t)lo * J.m;
    if (J.slots) { S.slots = J.slots + (size_t)lo * J.npert; S.vals = J.vals + (size_t)lo * J.npert; }
    return run_encode_job(S, r, sums.data() + (size_t)r * nr);
  }));
  for (int r = 0; r < J.nr; r++) {
    double tot = 0.0;
    for (int i = 0; i < k; i++) tot += sums[(size_t)i * nr + r];  // shard order: deterministic
    if (J.objs) J.objs[r] = snap_used[r] ? (float)(tot / (double)(J.n ? J.n : 1)) : 0.0f;
  }
  return LSQ_OK;
}

}  // namespace lsq

using namespace lsq;

extern "C" {

int lsq_splitarray(int64_t n, int nparts, int p, int64_t* lo, int64_t* hi) {
  LSQ_CHECK_ARG(nparts >= 1 && p >= 0 && p < nparts && n >= 0, "splitarray: need 0 <= p < nparts, n >= 0");
  const int64_t per = n / nparts, xtra = n % nparts;
  if (p < xtra) { *lo = p * (per + 1); *hi = *lo + per + 1; }
  else { *lo = xtra * (per + 1) + (p - xtra) * per; *hi = *lo + per; }
  return LSQ_OK;
}

int lsq_make_to_look(uint64_t seed, uint32_t ils_iter, int m, int randord, int32_t* to_look) {
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM, "m must be in 1..16");
  make_to_look_host(seed, ils_iter, m, randord, to_look);
  return LSQ_OK;
}

int lsq_make_perturb(uint64_t seed, uint32_t ils_iter, uint64_t g0, int64_t n, int m, int h, int npert,
                     uint8_t* slots, int16_t* vals) {
  LSQ_CHECK_ARG(m >= 1 && m <= LSQ_MAXM, "m must be in 1..16");
  LSQ_CHECK_ARG(npert >= 0 && npert <= m, "npert must be in 0..m");
  LSQ_CHECK_ARG(h >= 1 && h <= 256, "h must be in 1..256");
  for (int64_t v = 0; v < n; v++) {
    uint8_t s[LSQ_MAXM], x[LSQ_MAXM];
    make_perturb_one(seed, ils_iter, g0 + (uint64_t)v, m, h, npert, s, x);
    for (int i = 0; i < npert; i++) { slots[v * npert + i] = s[i]; vals[v * npert + i] = x[i]; }
  }
  return LSQ_OK;
}

int lsq_get_unaries(const float* X, int d, int64_t n, const float* C, int m, int h, float* U) {
  LSQ_TRY(check_encode_args(d, n, m, h, 0, 0));
  cudaStream_t st;
  LSQ_TRY(host_ctx(&st));
  DevBuf<float> dX, dC, dn, dU;
  LSQ_CUDA(dX.alloc((size_t)n * d));
  LSQ_CUDA(dC.alloc((size_t)m * h * d));
  LSQ_CUDA(dn.alloc((size_t)m * h));
  LSQ_CUDA(dU.alloc((size_t)m * n * h));
  LSQ_CUDA(cudaMemcpyAsync(dX.p, X, (size_t)n * d * 4, cudaMemcpyHostToDevice, st));
  LSQ_CUDA(cudaMemcpyAsync(dC.p, C, (size_t)m * h * d * 4, cudaMemcpyHostToDevice, st));
  LSQ_TRY(build_norms(dC.p, d, m, dn.p, st));
  LSQ_TRY(build_unaries(dX.p, d, n, dC.p, m, dn.p, dU.p, 0, st));
  LSQ_CUDA(cudaMemcpyAsync(U, dU.p, (size_t)m * n * h * 4, cudaMemcpyDeviceToHost, st));
  LSQ_CUDA(cudaStreamSynchronize(st));
  return LSQ_OK;
}

int lsq_get_binaries(const float* C, int d, int m, int h, float* G, int32_t* cbi) {
  LSQ_TRY(check_encode_args(d, 0, m, h, 0, 0));
  cudaStream_t st;
  LSQ_TRY(host_ctx(&st));
  DevBuf<float> dC, dT;
  LSQ_CUDA(dC.alloc((size_t)m * h * d));
  LSQ_CUDA(dT.alloc((size_t)m * m * h * h));
  LSQ_CUDA(cudaMemcpyAsync(dC.p, C, (size_t)m * h * d * 4, cudaMemcpyHostToDevice, st));
  LSQ_TRY(build_tables(dC.p, d, m, dT.p, st));
  int idx = 0;
  for (int i = 0; i < m; i++)
    for (int j = i + 1; j < m; j++, idx++) {
      cbi[2 * idx] = i + 1; cbi[2 * idx + 1] = j + 1;
      // binaries[idx][b][a] = 2<C_i[:,a], C_j[:,b]> = T[i][j][b][a]
      LSQ_CUDA(cudaMemcpyAsync(G + (size_t)idx * h * h, dT.p + (size_t)(i * m + j) * h * h, (size_t)h * h * 4,
                               cudaMemcpyDeviceToHost, st));
    }
  LSQ_CUDA(cudaStreamSynchronize(st));
  return LSQ_OK;
}

static int cost_common(const float* X, int d, int64_t n, const int16_t* B, const float* C, int m, int h,
                       float* cost, float* mean_out) {
  LSQ_TRY(check_encode_args(d, n, m, h, 0, 0));
  cudaStream_t st;
  LSQ_TRY(host_ctx(&st));
  DevBuf<float> dX, dC, dcost;
  DevBuf<uint8_t> dcodes;
  DevBuf<double> dsum;
  LSQ_CUDA(dX.alloc((size_t)n * d));
  LSQ_CUDA(dC.alloc((size_t)m * h * d));
  LSQ_CUDA(dcost.alloc(n));
  LSQ_CUDA(dcodes.alloc((size_t)n * m));
  LSQ_CUDA(dsum.alloc(1025));
  LSQ_CUDA(cudaMemcpyAsync(dX.p, X, (size_t)n * d * 4, cudaMemcpyHostToDevice, st));
  LSQ_CUDA(cudaMemcpyAsync(dC.p, C, (size_t)m * h * d * 4, cudaMemcpyHostToDevice, st));
  LSQ_TRY(upload_codes(B, n * m, dcodes.p, st));
  LSQ_TRY(launch_veccost(dX.p, d, n, dcodes.p, dC.p, m, dcost.p, st));
  if (cost) LSQ_CUDA(cudaMemcpyAsync(cost, dcost.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (mean_out) {
    LSQ_TRY(launch_sum_f32_to_f64(dcost.p, n, dsum.p, st));
    double s = 0;
    LSQ_CUDA(cudaMemcpyAsync(&s, dsum.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    LSQ_CUDA(cudaStreamSynchronize(st));
    *mean_out = (float)(n ? s / (double)n : 0.0);
  }
  LSQ_CUDA(cudaStreamSynchronize(st));
  return LSQ_OK;
}

int lsq_veccost(const float* X, int d, int64_t n, const int16_t* B, const float* C, int m, int h, float* cost) {
  return cost_common(X, d, n, B, C, m, h, cost, nullptr);
}

int lsq_qerror(const float* X, int d, int64_t n, const int16_t* B, const float* C, int m, int h, float* out) {
  return cost_common(X, d, n, B, C, m, h, nullptr, out);
}

int lsq_reconstruct(const int16_t* B, int64_t n, const float* C, int d, int m, int h, float* CB) {
  LSQ_TRY(check_encode_args(d, n, m, h, 0, 0));
  cudaStream_t st;
  LSQ_TRY(host_ctx(&st));
  DevBuf<float> dC, dCB;
  DevBuf<uint8_t> dcodes;
  LSQ_CUDA(dC.alloc((size_t)m * h * d));
  LSQ_CUDA(dCB.alloc((size_t)n * d));
  LSQ_CUDA(dcodes.alloc((size_t)n * m));
  LSQ_CUDA(cudaMemcpyAsync(dC.p, C, (size_t)m * h * d * 4, cudaMemcpyHostToDevice, st));
  LSQ_TRY(upload_codes(B, n * m, dcodes.p, st));
  LSQ_TRY(launch_reconstruct(dcodes.p, n, dC.p, d, m, dCB.p, st));
  LSQ_CUDA(cudaMemcpyAsync(CB, dCB.p, (size_t)n * d * 4, cudaMemcpyDeviceToHost, st));
  LSQ_CUDA(cudaStreamSynchronize(st));
  return LSQ_OK;
}

int lsq_quantize_norms(const int16_t* B, int64_t n, const float* C, int d, int m, int h, const float* cbnorms,
                       int hn, int16_t* out) {
  LSQ_TRY(check_encode_args(d, n, m, h, 0, 0));
  LSQ_CHECK_ARG(hn >= 1, "norm codebook must be non-empty");
  cudaStream_t st;
  LSQ_TRY(host_ctx(&st));
  DevBuf<float> dC, dcb;
  DevBuf<uint8_t> dcodes;
  DevBuf<int16_t> dout;
  LSQ_CUDA(dC.alloc((size_t)m * h * d));
  LSQ_CUDA(dcb.alloc(hn));
  LSQ_CUDA(dcodes.alloc((size_t)n * m));
  LSQ_CUDA(dout.alloc(n));
  LSQ_CUDA(cudaMemcpyAsync(dC.p, C, (size_t)m * h * d * 4, cudaMemcpyHostToDevice, st));
  LSQ_CUDA(cudaMemcpyAsync(dcb.p, cbnorms, (size_t)hn * 4, cudaMemcpyHostToDevice, st));
  LSQ_TRY(upload_codes(B, n * m, dcodes.p, st));
  LSQ_TRY(launch_quantize_norms(dcodes.p, n, dC.p, d, m, dcb.p, hn, dout.p, st));
  LSQ_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)n * 2, cudaMemcpyDeviceToHost, st));
  LSQ_CUDA(cudaStreamSynchronize(st));
  return LSQ_OK;
}

int lsq_encoding_icm(const float* X, int d, int64_t n, const int16_t* oldB, int16_t* newB, const float* C,
                     int m, int h, int niter, int randord, int npert, uint64_t seed, uint32_t ils_iter,
                     uint64_t g0, int verbose) {
  LSQ_TRY(check_encode_args(d, n, m, h, niter, npert));
  EncodeJob J;
  memset(&J, 0, sizeof(J));
  J.X = X; J.d = d; J.n = n; J.B_in = oldB; J.C = C; J.m = m;
  J.icmiter = niter; J.npert = npert; J.randord = randord;
  J.seed = seed; J.g0 = g0; J.ils_iter0 = ils_iter; J.total_iters = 1;
  J.B_out = newB; J.nsplits = 1; J.verbose = verbose;
  return run_encode(J);
}

int lsq_encoding_icm_sched(const float* X, int d, int64_t n, const int16_t* oldB, int16_t* newB,
                           const float* C, int m, int h, int niter, const int32_t* to_look, int npert,
                           const uint8_t* slots, const int16_t* vals, int verbose) {
  LSQ_TRY(check_encode_args(d, n, m, h, niter, npert));
  LSQ_CHECK_ARG(to_look != nullptr, "to_look is required");
  LSQ_CHECK_ARG(npert == 0 || (slots != nullptr && vals != nullptr), "slots/vals are required when npert > 0");
  uint32_t seen = 0;
  for (int i = 0; i < m; i++) {
    LSQ_CHECK_ARG(to_look[i] >= 0 && to_look[i] < m, "to_look must be a 0-based permutation of 0..m-1");
    seen |= 1u << to_look[i];
  }
  LSQ_CHECK_ARG(seen == (m == 32 ? 0xFFFFFFFFu : ((1u << m) - 1)), "to_look must be a permutation");
  for (int64_t i = 0; i < n * npert; i++) {  // before any work is queued (nothing to unwind on failure)
    LSQ_CHECK_ARG(vals[i] >= 0 && vals[i] < LSQ_H, "perturbation values must be 0-based in 0..255");
    LSQ_CHECK_ARG(slots[i] < m, "perturbation slots must be < m");
  }
  EncodeJob J;
  memset(&J, 0, sizeof(J));
  J.X = X; J.d = d; J.n = n; J.B_in = oldB; J.C = C; J.m = m;
  J.icmiter = niter; J.npert = npert; J.randord = 0;
  J.total_iters = 1; J.to_look = to_look; J.slots = slots; J.vals = vals;
  J.B_out = newB; J.nsplits = 1; J.verbose = verbose;
  return run_encode(J);
}

int lsq_encode_icm_cuda(const float* RX, int d, int64_t n, const int16_t* B, const float* C, int m, int h,
                        const int64_t* ilsiters, int nr, int icmiter, int npert, int randord, int nsplits,
                        uint64_t seed, uint64_t g0, int16_t* Bs, float* objs, int verbose) {
  LSQ_TRY(check_encode_args(d, n, m, h, icmiter, npert));
  LSQ_CHECK_ARG(nr >= 1 && ilsiters != nullptr, "ilsiters must hold at least one iteration count");
  LSQ_CHECK_ARG(nr < 32767, "too many snapshots");
  LSQ_CHECK_ARG(nsplits >= 1, "nsplits must be >= 1");
  int64_t maxit = 0;
  for (int r = 0; r < nr; r++) maxit = std::max(maxit, ilsiters[r]);
  LSQ_CHECK_ARG(maxit >= 0 && maxit < (1 << 30), "ilsiters out of range");
  EncodeJob J;
  memset(&J, 0, sizeof(J));
  J.X = RX; J.d = d; J.n = n; J.B_in = B; J.C = C; J.m = m;
  J.icmiter = icmiter; J.npert = npert; J.randord = randord;
  J.seed = seed; J.g0 = g0; J.ils_iter0 = 0; J.total_iters = (int)maxit;
  J.ilsiters = ilsiters; J.nr = nr; J.Bs = Bs; J.objs = objs;
  J.nsplits = nsplits; J.verbose = verbose;
  // (snapshots that no iteration fills are zeroed in run_encode_job; the others are fully overwritten)
  return run_encode(J);
}

// ---- device-pointer API ----------------------------------------------------------------------
int64_t lsq_dev_tables_bytes(int m) { return (int64_t)m * m * LSQ_H * LSQ_H * (int64_t)sizeof(float); }

int64_t lsq_dev_sliced_tables_bytes(int m) {
  return (int64_t)m * (m > 1 ? m - 1 : 1) * LSQ_H * LSQ_H * (int64_t)sizeof(float);
}

int lsq_dev_icm_layout(int m, int64_t n) { return icm_use_slices(m, n); }

int lsq_dev_build_tables(const float* dC, int d, int m, float* dT, float* dTs, void* stream) {
  LSQ_TRY(check_encode_args(d, 0, m, LSQ_H, 0, 0));
  LSQ_TRY(build_tables(dC, d, m, dT, (cudaStream_t)stream));
  if (dTs != nullptr) LSQ_TRY(build_sliced_tables(dT, m, dTs, (cudaStream_t)stream));
  return LSQ_OK;
}

int lsq_dev_build_unaries(const float* dX, int d, int64_t n, const float* dC, int m, float* dU, int sliced,
                          void* stream) {
  LSQ_TRY(check_encode_args(d, n, m, LSQ_H, 0, 0));
  cudaStream_t st = (cudaStream_t)stream;
  float* dn = nullptr;
  LSQ_CUDA(cudaMallocAsync((void**)&dn, (size_t)m * LSQ_H * sizeof(float), st));
  int rc = build_norms(dC, d, m, dn, st);
  if (rc == LSQ_OK) rc = build_unaries(dX, d, n, dC, m, dn, dU, sliced, st);
  cudaFreeAsync(dn, st);
  return rc;
}

int lsq_dev_build_unaries_tc(const float* dX, int d, int64_t n, const float* dC, int m, float* dU, void* stream) {
  LSQ_TRY(check_encode_args(d, n, m, LSQ_H, 0, 0));
  cudaStream_t st = (cudaStream_t)stream;
  float* dn = nullptr;
  LSQ_CUDA(cudaMallocAsync((void**)&dn, (size_t)m * LSQ_H * sizeof(float), st));
  int rc = build_norms(dC, d, m, dn, st);
  if (rc == LSQ_OK) rc = build_unaries_tc(dX, d, n, dC, m, dn, dU, st);
  cudaFreeAsync(dn, st);
  return rc;
}

int lsq_dev_veccost(const float* dX, int d, int64_t n, const uint8_t* dcodes, const float* dC, int m,
                    float* dcost, void* stream) {
  LSQ_TRY(check_encode_args(d, n, m, LSQ_H, 0, 0));
  return launch_veccost(dX, d, n, dcodes, dC, m, dcost, (cudaStream_t)stream);
}

int lsq_dev_icm_visit_counter(unsigned long long* dcounter) {
  set_icm_visit_counter(dcounter);
  return LSQ_OK;
}

int lsq_dev_icm_ils(const float* dX, int d, int64_t n, const float* dC, int m, const float* dU, const float* dT,
                    const float* dTs, int sliced, uint8_t* dcodes, float* dcost, int icmiter, int npert, const int8_t* orders,
                    const uint8_t* dslots, const uint8_t* dvals, uint64_t seed, uint32_t ils_iter0, int niters,
                    uint64_t g0, uint8_t* dsnap, float* dsnapcost, const int32_t* snap_of_iter, void* stream) {
  LSQ_TRY(check_encode_args(d, n, m, LSQ_H, icmiter, npert));
  LSQ_CHECK_ARG(niters >= 0 && orders != nullptr, "orders (host int8 [niters][m]) is required");
  LSQ_CHECK_ARG(!sliced || (dTs != nullptr && m >= 2 && m <= ICM_SLICE_MAX_M), "sliced layout needs dTs and 2 <= m <= 8");
  for (int it0 = 0; it0 < niters; it0 += ICM_MAX_ITERS_PER_LAUNCH) {
    const int nit = std::min(ICM_MAX_ITERS_PER_LAUNCH, niters - it0);
    IcmParams p;
    memset(&p, 0, sizeof(p));
    p.X = dX; p.C = dC; p.U = dU; p.T = dT; p.Ts = dTs; p.codes = dcodes; p.cost = dcost;
    p.slots = dslots ? dslots + (size_t)it0 * n * npert : nullptr;
    p.vals = dvals ? dvals + (size_t)it0 * n * npert : nullptr;
    p.snap = dsnap; p.snapcost = dsnapcost;
    p.n = n; p.seed = seed; p.g0 = g0; p.ils_iter0 = ils_iter0 + (uint32_t)it0;
    p.d = d; p.m = m; p.icmiter = icmiter; p.npert = npert; p.niters = nit;
    for (int i = 0; i < nit; i++) {
      p.snap_of_iter[i] = (dsnap && snap_of_iter) ? (int16_t)snap_of_iter[it0 + i] : (int16_t)-1;
      for (int k = 0; k < m; k++) {
        const int8_t o = orders[(size_t)(it0 + i) * m + k];
        LSQ_CHECK_ARG(o >= 0 && o < m, "orders entries must be in 0..m-1");
        p.orders[i][k] = o;
      }
    }
    LSQ_TRY(sliced ? launch_icm_slice(p, (cudaStream_t)stream) : launch_icm_warp(p, (cudaStream_t)stream));
  }
  return LSQ_OK;
}

}  // extern "C"
