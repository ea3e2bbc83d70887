/*
 * lsq_b200.h — C ABI of the B200-native LSQ hot path (liblsq_b200.so).
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and returns an int status
 * (0 = ok, non-zero = error; text via lsq_last_error()) — except the two linscan symbols, which keep
 * the reference's `void` signatures byte for byte so that src/linscan/Linscan.jl works unchanged.
 * Citations `file:line` are relative to the reference repository (una-dinosauria/local-search-quantization).
 *
 * Array conventions = the bytes Julia's ccall passes for its column-major arrays:
 *   X  (d-by-n  Matrix{Float32})      -> const float*   [n][d]
 *   B  (m-by-n  Matrix{Int16})        -> const int16_t* [n][m]   1-BASED codes, as everywhere in Julia
 *   C  (cat(3, C...), d-by-h-by-m)    -> const float*   [m][h][d]
 *   uint8 codes (UInt8(B-1))          -> const uint8_t* [n][m]   0-based (linscan, device API)
 * h must be 256 and m <= 16 on the encode path (the reference GPU path hard-codes both,
 * src/encodings/cuda/cudautils.cu:38,57,94,155,245).
 *
 * Host-pointer functions (lsq_*) copy to the device, run, copy back, and are what the Julia overlay
 * binds.  Device-pointer functions (lsq_dev_*) take device buffers and a CUDA stream and never touch
 * the host; the host-pointer functions are thin wrappers over them.  There is no CPU fallback: with no
 * usable GPU every compute entry point returns LSQ_ERR_CUDA.
 */
#ifndef LSQ_B200_H_
#define LSQ_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSQ_OK 0
#define LSQ_ERR_ARG 1   /* invalid argument (m > 16, h != 256, npert > m, ...) */
#define LSQ_ERR_CUDA 2  /* CUDA runtime / no device */
#define LSQ_ERR_LIMIT 3 /* documented capacity limit exceeded */

/* ---- runtime (replaces CudaUtilsModule.init / finit, cudaUtilsModule.jl:37-43, and the
 *      CuDevice(0)/CuContext of encode_icm_cuda.jl:59-64) ------------------------------------------ */
int lsq_init(int device);
/* Binds SEVERAL GPUs of one box (devices == NULL or n <= 0: all visible ones).  The reference is hard-wired to
 * device 0 (encode_icm_cuda.jl:59-64); here the calls that shard naturally — lsq_encoding_icm[_sched],
 * lsq_encode_icm_cuda, lsq_update_codebooks, lsq_train_lsq, and the linscan symbols (by queries) — split their
 * input over the bound devices by the splitarray rule (utils.jl:152-177), one internal worker thread per
 * device, from the single calling thread; results are bit-identical for any number of devices.  The only
 * collective is the all-reduce of the codebook-update statistics (NCCL, bound with dlopen at first use; a
 * peer-memory reduction kernel when NCCL is absent or LSQ_B200_ALLREDUCE=p2p).  devices[0] is the primary
 * device: every other host-pointer call runs there. */
int lsq_init_devices(const int* devices, int n);
int lsq_num_bound_devices(void);
int lsq_finalize(void);
const char* lsq_last_error(void);
int lsq_device_count(void);
/* number of CUDA kernels this library has launched so far in the process (measurement aid: bench.py reports
 * the difference over its timed region as gpu_launches) */
unsigned long long lsq_launch_count(void);
/* device time (ms, primary device) of the most recent statistics exchange + finalize inside lsq_train_lsq
 * (NCCL all-reduce, or the fused peer-memory kernel); -1 if none happened yet.  Measurement aid. */
float lsq_last_collective_ms(void);
const char* lsq_version(void);

/* ---- a10: splitarray (utils.jl:152-177): part p of nparts over 0..n-1 -> [lo, hi) ------------- */
int lsq_splitarray(int64_t n, int nparts, int p, int64_t* lo, int64_t* hi);

/* ---- canonical schedule (replaces randperm/sample/rand of encode_icm.jl:46-64 and the curand
 *      setup_kernel/perturb of cudautils.cu:14-80): Philox4x32-10 keyed by seed, counter = (global
 *      vector index, ILS iteration); independent of sharding ------------------------------------- */
int lsq_make_to_look(uint64_t seed, uint32_t ils_iter, int m, int randord, int32_t* to_look);
int lsq_make_perturb(uint64_t seed, uint32_t ils_iter, uint64_t g0, int64_t n, int m, int h, int npert,
                     uint8_t* slots /*[n][npert]*/, int16_t* vals /*[n][npert] 0-based*/);

/* ---- a3 get_unaries (utils.jl:94-122): U[m][n][h] = -2*C_i'*X + ||c||^2 ------------------------ */
int lsq_get_unaries(const float* X, int d, int64_t n, const float* C, int m, int h, float* U);
/* ---- a4 get_binaries (utils.jl:125-144): G[ncbi][h][h] (G[idx][b][a] = 2<C_i[:,a],C_j[:,b]>, i<j),
 *      cbi[ncbi][2] 1-based pairs like the reference's `cbi` -------------------------------------- */
int lsq_get_binaries(const float* C, int d, int m, int h, float* G, int32_t* cbi);
/* ---- a5 veccost / qerror / reconstruct (utils.jl:225-254, 257-285, 203-223) ------------------- */
int lsq_veccost(const float* X, int d, int64_t n, const int16_t* B, const float* C, int m, int h,
                float* cost);
int lsq_qerror(const float* X, int d, int64_t n, const int16_t* B, const float* C, int m, int h,
               float* out);
int lsq_reconstruct(const int16_t* B, int64_t n, const float* C, int d, int m, int h, float* CB);

/* ---- a1/a2 encoding_icm (encode_icm.jl:131-189): ONE ILS iteration.  newB may alias oldB.
 *      g0 = global index of X's first vector (0 unless the caller shards). ----------------------- */
int lsq_encoding_icm(const float* X, int d, int64_t n, const int16_t* oldB, int16_t* newB,
                     const float* C, int m, int h, int niter, int randord, int npert, uint64_t seed,
                     uint32_t ils_iter, uint64_t g0, int verbose);
/* same, explicit schedule: to_look[m] 0-based; slots[n][npert] 0-based ascending; vals[n][npert]
 * 0-based new codes.  This is the entry point the parity tests drive. */
int lsq_encoding_icm_sched(const float* X, int d, int64_t n, const int16_t* oldB, int16_t* newB,
                           const float* C, int m, int h, int niter, const int32_t* to_look, int npert,
                           const uint8_t* slots, const int16_t* vals, int verbose);

/* ---- a6 encode_icm_cuda (encode_icm_cuda.jl:253-296): all ILS iterations in one call; snapshots of
 *      the codes and objective at the iteration counts in ilsiters[nr].  Bs: [nr][n][m] 1-based.
 *      nsplits only bounds device memory (as in the reference); results do not depend on it. -------- */
int lsq_encode_icm_cuda(const float* RX, int d, int64_t n, const int16_t* B, const float* C, int m,
                        int h, const int64_t* ilsiters, int nr, int icmiter, int npert, int randord,
                        int nsplits, uint64_t seed, uint64_t g0, int16_t* Bs, float* objs, int verbose);

/* ---- a7 update_codebooks (codebook_update.jl:52-86): Cout[m][h][d] = min-norm least-squares
 *      codebooks for fixed codes.  method: "lsqr" or "lsmr" (anything else -> LSQ_ERR_ARG, as the
 *      reference's error at :59). --------------------------------------------------------------- */
int lsq_update_codebooks(const float* X, int d, int64_t n, const int16_t* B, int m, int h, float* Cout,
                         const char* method, int verbose);

/* ---- a8 linscan_lsq: EXACT reference symbol + signature (linscan_aqd_pairwise_byte.cpp:97-104;
 *      bound at Linscan.jl:63-69).  idx is 1-based; rows ascending by (distance, id). ------------- */
void linscan_aqd_query_extra_byte(float* dists, int* idx, unsigned char* codes, float* queries,
                                  float* codebooks, float* dbnorms, int nqueries, int ncodes, int m,
                                  int h, int d, int nn);
/* ---- a9 linscan_pq / linscan_opq: EXACT reference symbol + signature (linscan_aqd.cpp:107-113;
 *      bound at Linscan.jl:19-23).  res is 0-based (Julia adds 1, Linscan.jl:25). ----------------- */
void linscan_aqd_query(float* dists, unsigned int* res, unsigned char* codes, float* centers,
                       float* queries, int N, unsigned int NQ, int B, int K, int dim1codes,
                       int dim1queries, int subdim);
/* which main pass linscan_lsq runs for n base vectors and nq queries: 1 = tensor-core filter (tcgen05 bf16 GEMM over the
 * decoded base vectors) + exact rescoring of the survivors (csrc/adc_tc.cu; n >= 64 K, nq * m >= 8000, d a multiple
 * of 16 up to 128), 0 = lookup-table scan (csrc/linscan.cu).  Both give the reference's results bit for bit.
 * linscan_pq / linscan_opq follow the same rule with d = dim1codes * subdim.  Environment override: LSQ_B200_ADC=scan|tc. */
int lsq_linscan_path(int64_t n, int64_t nq, int m, int d);
/* Measurement aid: with LSQ_B200_ADC_TIMING set in the environment every linscan call times its phases with CUDA
 * events (and prints them to stderr); this returns the device time (ms) and name of phase i of the calling thread's
 * most recent call, and the number of phases; i = -2 returns the number of products the tensor-core filter ran with in
 * that call (1 or 2; 0 = lookup scan). */
int lsq_linscan_last_phases(int i, float* ms, const char** name);
/* status-returning twins of the two above (same arguments) */
int lsq_linscan_lsq(float* dists, int* idx, const unsigned char* codes, const float* queries,
                    const float* codebooks, const float* dbnorms, int nqueries, int ncodes, int m, int h,
                    int d, int nn);
int lsq_linscan_pq(float* dists, unsigned int* res, const unsigned char* codes, const float* centers,
                   const float* queries, int N, unsigned int NQ, int B, int K, int dim1codes,
                   int dim1queries, int subdim);

/* ---- f2 quantize_norms (utils.jl:6-31): out[n] 1-based index of the nearest norm-codebook entry - */
int lsq_quantize_norms(const int16_t* B, int64_t n, const float* C, int d, int m, int h,
                       const float* cbnorms, int hn, int16_t* out);

/* ---- f1 train_lsq (src/lsq/LSQ.jl:10-88) with X, codes, unaries and tables resident on the GPU for
 *      the whole alternation: C = update_codebooks(R'X, B); C_i = R*C_i; ilsiter x encoding_icm; then
 *      niter x { obj[iter] = qerror; C = update_codebooks(X, B); ilsiter x encoding_icm }; finally the
 *      norm codebook of LSQ.jl:68-84.  Equal, bit for bit, to looping lsq_update_codebooks /
 *      lsq_encoding_icm with ils_iter = 0, 1, 2, ... and the same seed.
 *        R        d-by-d Matrix{Float32} (column-major) or NULL for the identity
 *        B        in: initial codes, out: final codes   [n][m] 1-based
 *        C        out: codebooks [m][h][d]   (the C argument of train_lsq is overwritten before use, :34)
 *        cbnorms  out [h] or NULL;  B_norms out [n] 1-based or NULL;  obj out [niter] or NULL
 *      The norm codebook is a deterministic 1-D k-means (Clustering.jl's seeded kmeans is third-party
 *      and unpinned): see lsq_kmeans1d. ----------------------------------------------------------- */
int lsq_train_lsq(const float* X, int d, int64_t n, int m, int h, const float* R, int16_t* B, float* C,
                  int niter, int ilsiter, int icmiter, int randord, int npert, uint64_t seed,
                  float* cbnorms, int16_t* B_norms, float* obj, int verbose);
/* ---- f2 norm codebook (stand-in for `kmeans(dbnorms, h)`, LSQ.jl:80): Lloyd iterations on scalars,
 *      centres seeded at the (2j+1)/(2h) quantiles, ties to the lower centre, float64 means in a fixed
 *      order; stops at a fixed point or after maxiter (Clustering.jl default: 100).  centers[h] come
 *      back ascending. -------------------------------------------------------------------------- */
int lsq_kmeans1d(const float* values, int64_t n, int h, int maxiter, float* centers, int* iters_out);
/* ---- f3 eval_recall (Linscan.jl:76-117): recall[i-1] = fraction of queries whose ground-truth id is
 *      found (exactly once) among the first i predictions, i = 1..k.  ids_predicted is [nq][ld] with
 *      ld >= k (a row = one query's ranked list, i.e. a column of the Julia matrix). --------------- */
int lsq_eval_recall(const int32_t* ids_gnd, const int32_t* ids_predicted, int nq, int ld, int k,
                    double* recall);

/* ---- f4 encoding_viterbi (src/encodings/encode_chain.jl:95-127): exact MAP codes on the chain
 *      1-2-...-m by min-sum dynamic programming (the ChainQ encoder), same unary / pair tables as the
 *      ICM path.  B out [n][m] 1-based.  Needs 2 <= m <= 16, h == 256.  No randomness. ------------- */
int lsq_encoding_viterbi(const float* X, int d, int64_t n, const float* C, int m, int h, int16_t* B,
                         int verbose);

/* =====================================================================================================
 * Device-pointer API.  All pointers are device pointers; `stream` is a cudaStream_t (0 = legacy
 * default stream).  Codes are uint8 0-based [n][m].  Nothing here synchronises the host.
 * =================================================================================================== */
/* bytes of table storage for m codebooks: T[m][m][256][256] floats (both orientations materialised,
 * cf. binaries + binaries_t, encode_icm.jl:25-28) */
int64_t lsq_dev_tables_bytes(int m);
/* bytes of the sliced copy Ts[m][8][m-1][256][32] the shared-memory-slice ICM kernel stages by TMA */
int64_t lsq_dev_sliced_tables_bytes(int m);
/* which ICM kernel / unary layout serves (m, n): 0 = warp-per-vector kernel, tables gathered from L2,
 * U[m][n][256]; 1 = slice kernel (m <= 8, large n), tables in shared memory, U[m][8][n][32].
 * Environment override for testing: LSQ_B200_ICM_KERNEL=warp|slice. */
int lsq_dev_icm_layout(int m, int64_t n);
/* T[j][k][b][a] = 2<C_j[:,a], C_k[:,b]> for j != k; dTs (may be NULL) receives the sliced copy */
int lsq_dev_build_tables(const float* dC, int d, int m, float* dT, float* dTs, void* stream);
/* sliced = 0: U[m][n][256];  sliced = 1: U[m][8][n][32] */
int lsq_dev_build_unaries(const float* dX, int d, int64_t n, const float* dC, int m, float* dU,
                          int sliced, void* stream);
/* FAST MODE: the same table (plain layout) on the tensor cores — tcgen05.mma kind::tf32 with a 3xTF32
 * operand split, fp32 accumulation in TMEM.  Not bit-exact with the sequential fp32 chain of the
 * parity path (|dU|/|U| ~ 1e-6; quantisation error within 1e-5 relative).  Needs d % 8 == 0, d <= 128.
 * The host-pointer encode calls and lsq_train_lsq use it when the environment has LSQ_B200_UNARY=tc. */
int lsq_dev_build_unaries_tc(const float* dX, int d, int64_t n, const float* dC, int m, float* dU,
                             void* stream);
int lsq_dev_veccost(const float* dX, int d, int64_t n, const uint8_t* dcodes, const float* dC, int m,
                    float* dcost, void* stream);
/* `niters` ILS iterations, numbered ils_iter0 .. ils_iter0+niters-1, in ONE launch.
 *   orders   : HOST int8 [niters][m] visit orders (travel in the kernel parameter block)
 *   dslots/dvals: explicit perturbations [niters][n][npert] (uint8 / uint8) or NULL -> Philox(seed)
 *   dcodes   : in/out accepted codes; dcost: in/out cost of dcodes (must be valid on entry)
 *   dsnap    : NULL or uint8 [nsnap][n][m]; dsnapcost: NULL or float [nsnap][n];
 *              snap_of_iter (HOST int32[niters], -1 = none) says which snapshot slot receives the
 *              accepted codes (and their costs) after each iteration. */
/* Measurement hook: every later ICM launch of the calling thread adds the number of node visits it actually
 * executed to *dcounter (device memory; NULL switches it off).  The kernel skips visits whose outcome is
 * already known, so the count is data dependent (nominal: n * niters * icmiter * m). */
int lsq_dev_icm_visit_counter(unsigned long long* dcounter);
int lsq_dev_icm_ils(const float* dX, int d, int64_t n, const float* dC, int m, const float* dU,
                    const float* dT, const float* dTs, int sliced, uint8_t* dcodes, float* dcost, int icmiter, int npert,
                    const int8_t* orders, const uint8_t* dslots, const uint8_t* dvals, uint64_t seed,
                    uint32_t ils_iter0, int niters, uint64_t g0, uint8_t* dsnap, float* dsnapcost,
                    const int32_t* snap_of_iter, void* stream);
/* Codebook-update statistics (codebook_update.jl:8-46 in normal-equation form), EXACT and order-independent:
 * one int64 buffer S[mh*mh + mh*d] (mh = m*256): co-occurrence counts, then per-code sums of x in fixed
 * point, x rounded once to rint(x * 2^scale_exp).  Integer sums do not depend on thread order, chunking or
 * sharding, so summing the S of several shards (ONE all-reduce of int64) gives bit-identical codebooks for
 * any number of GPUs.  scale_exp = lsq_cb_scale_exp(max|x| over ALL shards, total vector count) keeps every
 * sum below 2^62; max|x| comes from lsq_dev_absmax (it only changes when X changes, not per iteration).
 *   lsq_dev_absmax        *dmax = max(*dmax, max|x|) over `count` floats (zero *dmax first)
 *   lsq_dev_cb_accumulate S += this shard (zero S once per update)
 *   lsq_dev_cb_finalize   summed S -> Gram[mh][mh], Rhs[mh][d] (float64) for lsq_dev_cb_solve
 *   lsq_dev_cb_stats      single-shard convenience: absmax + accumulate + finalize; OVERWRITES Gram / Rhs
 *                         (synchronises the stream once to read max|x|) */
int64_t lsq_cb_stats_len(int m, int d);
int lsq_cb_scale_exp(float absmax, int64_t n_total);
int lsq_dev_absmax(const float* dX, int64_t count, float* dmax, void* stream);
int lsq_dev_cb_accumulate(const float* dX, int d, int64_t n, const uint8_t* dcodes, int m, int scale_exp,
                          int64_t* dstats, void* stream);
int lsq_dev_cb_finalize(const int64_t* dstats, int m, int d, int scale_exp, double* dGram, double* dRhs,
                        void* stream);
int lsq_dev_cb_stats(const float* dX, int d, int64_t n, const uint8_t* dcodes, int m, double* dGram,
                     double* dRhs, void* stream);
/* min-norm solve of Gram * K = Rhs by conjugate gradients from K0 = 0; dCout float [m][256][d]. */
int lsq_dev_cb_solve(const double* dGram, const double* dRhs, int m, int d, float* dCout, int max_iter,
                     double tol, int* iters_out, void* stream);
/* ADC scan with device buffers: lut_kind 0 = LSQ (-2<q,c>, + dbnorms, 1-based ids), 1 = PQ. */
int lsq_dev_linscan(const uint8_t* dcodes, int64_t n, int m, const float* dqueries, int nq, int d,
                    const float* dcodebooks, const float* dbnorms, int lut_kind, int subdim, int nn,
                    float* ddists, int32_t* dids, void* stream);

/* Test hook of the tensor-core ADC prefilter (csrc/adc_tc.cu): the filter values dbnorm[v] - 2<q, xhat_v> as the
 * bf16 hi/lo tcgen05 GEMM computes them, for every (query, base vector) pair.  dout: float [nq][ld],
 * ld >= 128*ceil(n/128); columns >= n are padding (> 1e37).  Needs d % 16 == 0, 16 <= d <= 128. */
int lsq_dev_adc_filter_values(const uint8_t* dcodes, int64_t n, int m, const float* dqueries, int nq, int d,
                              const float* dcodebooks, const float* dbnorms, float* dout, int64_t ld, void* stream);

/* chain encoder on device buffers: dU = U[m][n][256] (lsq_dev_build_unaries, plain layout) is CONSUMED
 * (the forward messages overwrite it in place); dT from lsq_dev_build_tables; dcodes uint8 [n][m] out. */
int lsq_dev_viterbi(float* dU, int64_t n, int m, const float* dT, uint8_t* dcodes, void* stream);
/* eval_recall on device buffers: dgnd int32[nq], dpred int32[nq][ld], drecall double[k]. */
int lsq_dev_eval_recall(const int32_t* dgnd, const int32_t* dpred, int nq, int ld, int k, double* drecall,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSQ_B200_H_ */
