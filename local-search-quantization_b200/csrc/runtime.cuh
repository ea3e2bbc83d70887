// runtime.cuh — the device runtime behind the host-pointer API: the set of bound GPUs (lsq_init /
// lsq_init_devices, replacing CuDevice(0)/CuContext of encode_icm_cuda.jl:59-64, where the device is
// hard-coded), one worker thread per bound GPU for the calls that shard, staged host->device copies for
// pageable callers (Julia arrays are pageable), and the one collective of the path: the sum of the
// codebook-update statistics over the bound GPUs (NCCL all-reduce, or a peer-memory reduction kernel).
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>

#include "common.cuh"

namespace lsq {

struct DevCtx {
  int dev = -1;
  cudaStream_t st = nullptr;   // main stream of the host-pointer API on this device
  cudaStream_t st2 = nullptr;  // second slot of the encode pipeline
};

// Binds the device set lazily (the current device alone) if lsq_init / lsq_init_devices was never called.
int rt_ensure_init();
int rt_num_devices();            // bound GPUs (>= 1 after rt_ensure_init)
const DevCtx& rt_ctx(int i);     // i-th bound GPU
int rt_bind(int i);              // cudaSetDevice + allocation stream of the calling thread -> ctx i
int rt_init_devices(const int* devs, int n);
int rt_finalize();

// Runs fn(i) for i = 0..k-1, one thread per bound device (the caller's thread takes i = 0), joins them, and
// returns the first non-OK status in index order (its message becomes the caller's lsq_last_error()).
int rt_parallel(int k, const std::function<int(int)>& fn);

// Reusable barrier for the k worker threads of one rt_parallel call.
class HostBarrier {
 public:
  explicit HostBarrier(int k) : k_(k) {}
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    const unsigned gen = gen_;
    if (++count_ == k_) { count_ = 0; gen_++; cv_.notify_all(); }
    else cv_.wait(lk, [&] { return gen_ != gen; });
  }
 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int k_, count_ = 0;
  unsigned gen_ = 0;
};

// Phase boundary for k workers that must all leave together when any of them failed: every worker calls
// ok(rc) at the same points of the same sequence; it returns false on ALL of them if any passed an error,
// so nobody is left waiting at a later barrier or inside a collective.
class PhaseSync {
 public:
  explicit PhaseSync(int k) : bar(k) {}
  bool ok(int rc) {
    if (rc != LSQ_OK) failed_.store(1);
    bar.wait();
    const bool good = failed_.load() == 0;
    bar.wait();  // nobody races ahead and flags the NEXT phase before everyone has read this one
    return good;
  }
  HostBarrier bar;
 private:
  std::atomic<int> failed_{0};
};

// devices a call over n items spreads over: all bound ones, fewer when a shard would drop below `min_per_device`
int rt_devices_for(int64_t n, int64_t min_per_device);

// Host -> device copy of caller memory, ordered on `st`.  Pinned / registered sources go straight to
// cudaMemcpyAsync.  Pageable sources (Julia arrays) are cut into 2 MB chunks; the calling thread and a few
// persistent helper threads of the device each grab the next chunk, memcpy it into one of their own two pinned
// buffers and queue its DMA, so the DMA of one chunk overlaps the host copies of the others and the copy is not
// limited to one core's memcpy rate (LSQ_B200_H2D=direct|staged, default staged; LSQ_B200_COPY_THREADS=k).
int rt_h2d(void* ddst, const void* hsrc, size_t bytes, cudaStream_t st);

// ---- the collective: in-place sum of `count` int64 over the k bound devices ---------------------------
// Called concurrently by worker i = 0..k-1 of one rt_parallel call with its own device buffer, on its own
// stream.  Integer sums are exact, so the result does not depend on the reduction order or on k.
//   backend "nccl": ncclAllReduce(ncclInt64, ncclSum) on a clique built by ncclCommInitAll (libnccl.so.2 is
//                   bound at run time, so single-GPU deployments do not need it);
//   backend "p2p" : every device reads its peers' buffers over NVLink peer memory in rank order and writes
//                   the sum locally (one kernel per device + event hand-shakes); used when NCCL is absent or
//                   LSQ_B200_ALLREDUCE=p2p.
struct AllReduceGroup;
AllReduceGroup* rt_allreduce_group(int k);   // nullptr + lsq_last_error() on failure; owned by the runtime
// p2p backend, fused form: instead of materialising the sum, the CONSUMER kernel reads every peer's buffer itself
// (cb_finalize_peers: sum over the peers in rank order + conversion to float64 in one pass, no intermediate copy).
//   rt_peer_begin : publish `dbuf`, wait (stream-side) until every peer's buffer is complete, return their pointers
//   rt_peer_end   : after the consumer kernel was queued: nobody overwrites its buffer while a peer still reads it
struct PeerPtrs { const int64_t* p[16]; };
bool rt_allreduce_is_p2p(const AllReduceGroup* g);
int rt_peer_begin(AllReduceGroup* g, int rank, const int64_t* dbuf, cudaStream_t st, HostBarrier* bar, PeerPtrs* peers);
int rt_peer_end(AllReduceGroup* g, int rank, cudaStream_t st, HostBarrier* bar);
// device-time of the most recent statistics exchange on the primary device (measurement aid, lsq_last_collective_ms)
void rt_note_collective(cudaEvent_t a, cudaEvent_t b);
int rt_allreduce_sum_i64(AllReduceGroup* g, int rank, int64_t* dbuf, int64_t* dscratch, size_t count,
                         cudaStream_t st, HostBarrier* bar);
const char* rt_allreduce_backend(const AllReduceGroup* g);

}  // namespace lsq
