/*
 * lsq_oracle.c — CPU restatement of the reference LSQ hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The shipped path is the CUDA
 * library (local-search-quantization_b200/csrc), which never links or calls anything in here.
 *
 * PARITY STATUS: "parity unpinned" for the Julia functions.  The reference ships no tests, no golden
 * vectors, and Julia is not installable in this image, so the Julia code cannot be executed.  This
 * restatement follows the reference line by line (citations below, relative to /root/reference) and
 * is cross-checked against an independent NumPy twin (oracle/np_twin.py).  Three things the reference
 * leaves unspecified are frozen here as the CANONICAL definition that the CUDA kernels must match
 * bit for bit:
 *   (1) dot products (BLAS sgemm in the reference, utils.jl:108,137) = sequential-k fp32 FMA chain;
 *   (2) the @simd reduction of veccost (utils.jl:241-249) = 32 lane-strided partial sums followed by
 *       an xor-butterfly (offsets 16,8,4,2,1);
 *   (3) the random schedule (randperm / StatsBase.sample / rand on MersenneTwister,
 *       encode_icm.jl:47,58,63) = Philox4x32-10 keyed by (seed), counter (global vector idx, ILS iter).
 * The linear-scan functions ARE pinned: tests run them against the reference's own C++ compiled
 * unmodified into oracle/_ref/ (see oracle/Makefile).
 *
 * Layout conventions (Julia column-major arrays seen from C):
 *   X  d-by-n            -> float  X[n][d]
 *   B  m-by-n Int16      -> int16  B[n][m]          (1-based at the Julia boundary, 0-based inside)
 *   C  m x (d-by-h)      -> float  C[m][h][d]       (cat(3, C...))
 *   unaries[i] h-by-n    -> float  U[m][n][h]
 *   binaries[idx] h-by-h -> float  G[idx][b][a],  G(a,b) = 2<C_i[:,a], C_j[:,b]>, i<j
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXM 16

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11) — the counter-based generator of the canonical schedule.
 * ---------------------------------------------------------------------------------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; r++) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(c));
}

/* word i of the stream (seed, ils_iter, g, stream): block = i/4 */
#define ORC_STREAM_PERTURB 0u
#define ORC_STREAM_ORDER 1u
static uint32_t sched_word(uint64_t seed, uint32_t ils_iter, uint64_t g, uint32_t stream, int i) {
  uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), ils_iter, (stream << 24) | (uint32_t)(i >> 2)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  orc_philox4x32_10(ctr, key, out);
  return out[i & 3];
}

/* Visit order of one ILS iteration: randperm(m) (encode_icm.jl:46-49; encode_icm_cuda.jl:141-144),
 * or 0..m-1 when randord is false.  Fisher-Yates, descending. */
void orc_make_to_look(uint64_t seed, uint32_t ils_iter, int m, int randord, int32_t* to_look) {
  for (int i = 0; i < m; i++) to_look[i] = i;
  if (!randord) return;
  for (int i = m - 1, w = 0; i >= 1; i--, w++) {
    uint32_t r = sched_word(seed, ils_iter, 0, ORC_STREAM_ORDER, w);
    int j = (int)(r % (uint32_t)(i + 1));
    int32_t t = to_look[i]; to_look[i] = to_look[j]; to_look[j] = t;
  }
}

/* Perturbation of vector g: npert DISTINCT slots, sorted ascending (sample(1:m,npert,replace=false,
 * ordered=true), encode_icm.jl:58), each paired with an independent uniform value in 0..h-1
 * (rand(1:h,npert,n), encode_icm.jl:63) which may equal the old code. */
void orc_make_perturb_one(uint64_t seed, uint32_t ils_iter, uint64_t g, int m, int h, int npert,
                          uint8_t* slots, int16_t* vals) {
  int a[ORC_MAXM];
  for (int i = 0; i < m; i++) a[i] = i;
  for (int i = 0; i < npert; i++) {
    uint32_t r = sched_word(seed, ils_iter, g, ORC_STREAM_PERTURB, i);
    int j = i + (int)(r % (uint32_t)(m - i));
    int t = a[i]; a[i] = a[j]; a[j] = t;
  }
  /* insertion sort of the chosen prefix */
  for (int i = 1; i < npert; i++) {
    int v = a[i], j = i - 1;
    while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
    a[j + 1] = v;
  }
  for (int i = 0; i < npert; i++) {
    slots[i] = (uint8_t)a[i];
    vals[i] = (int16_t)(sched_word(seed, ils_iter, g, ORC_STREAM_PERTURB, npert + i) % (uint32_t)h);
  }
}

void orc_make_perturb(uint64_t seed, uint32_t ils_iter, uint64_t g0, int64_t n, int m, int h,
                      int npert, uint8_t* slots, int16_t* vals) {
  for (int64_t v = 0; v < n; v++)
    orc_make_perturb_one(seed, ils_iter, g0 + (uint64_t)v, m, h, npert, slots + v * npert,
                         vals + v * npert);
}

/* ------------------------------------------------------------------------------------------------
 * Canonical dot product: acc = fmaf(a[k], b[k], acc), k ascending, acc0 = 0.
 * target_clones gives an inlined vfmadd on FMA hosts and exact libm fmaf elsewhere.
 * ---------------------------------------------------------------------------------------------- */
__attribute__((target_clones("fma", "default"))) static float dot_fma(const float* a, const float* b,
                                                                      int d) {
  float acc = 0.0f;
  for (int k = 0; k < d; k++) acc = fmaf(a[k], b[k], acc);
  return acc;
}

/* many-against-one: out[a] = dot(C[a], x) for a<h.  Same arithmetic per output as dot_fma. */
__attribute__((target_clones("fma", "default"))) static void dots_fma(const float* C, const float* x,
                                                                      int h, int d, float* out) {
  for (int a = 0; a < h; a++) {
    const float* c = C + (size_t)a * d;
    float acc = 0.0f;
    for (int k = 0; k < d; k++) acc = fmaf(c[k], x[k], acc);
    out[a] = acc;
  }
}

/* diag(C_i' * C_i) (utils.jl:109) */
void orc_get_norms(const float* C, int m, int h, int d, float* norms /*[m][h]*/) {
  for (int i = 0; i < m * h; i++) norms[i] = dot_fma(C + (size_t)i * d, C + (size_t)i * d, d);
}

/* The sgemm part of get_unaries alone: G_i = -2*C_i'*X (utils.jl:108; on the reference GPU path this is the
 * cuBLAS gemm of encode_icm_cuda.jl:92, to which the reference's own vec_add kernel then adds the norms). */
void orc_get_unary_gemm(const float* X, const float* C, int64_t n, int m, int h, int d,
                        float* G /*[m][n][h]*/) {
#pragma omp parallel
  {
    float* tmp = (float*)malloc(sizeof(float) * h);
#pragma omp for collapse(2) schedule(static)
    for (int i = 0; i < m; i++)
      for (int64_t v = 0; v < n; v++) {
        dots_fma(C + (size_t)i * h * d, X + (size_t)v * d, h, d, tmp);
        float* g = G + ((size_t)i * n + v) * h;
        for (int a = 0; a < h; a++) g[a] = -2.0f * tmp[a];
      }
    free(tmp);
  }
}

/* get_unaries (utils.jl:94-122): U_i = -2*C_i'*X, then += ||c||^2 per row. */
void orc_get_unaries(const float* X, const float* C, int64_t n, int m, int h, int d,
                     float* U /*[m][n][h]*/) {
  float* norms = (float*)malloc(sizeof(float) * m * h);
  orc_get_norms(C, m, h, d, norms);
#pragma omp parallel
  {
    float* tmp = (float*)malloc(sizeof(float) * h);
#pragma omp for collapse(2) schedule(static)
    for (int i = 0; i < m; i++)
      for (int64_t v = 0; v < n; v++) {
        dots_fma(C + (size_t)i * h * d, X + (size_t)v * d, h, d, tmp);
        float* u = U + ((size_t)i * n + v) * h;
        for (int a = 0; a < h; a++) u[a] = -2.0f * tmp[a] + norms[i * h + a];
      }
    free(tmp);
  }
  free(norms);
}

/* get_binaries (utils.jl:125-144): binaries[idx] = 2*C_i'*C_j for i<j; cbi[:,idx] = (i,j) (0-based
 * here).  Memory: G[idx][b][a]. */
void orc_get_binaries(const float* C, int m, int h, int d, float* G, int32_t* cbi /*[ncbi][2]*/) {
  int idx = 0;
  for (int i = 0; i < m; i++)
    for (int j = i + 1; j < m; j++, idx++) {
      cbi[2 * idx] = i; cbi[2 * idx + 1] = j;
      float* g = G + (size_t)idx * h * h;
#pragma omp parallel for schedule(static)
      for (int b = 0; b < h; b++)
        for (int a = 0; a < h; a++)
          g[(size_t)b * h + a] =
              2.0f * dot_fma(C + ((size_t)i * h + a) * d, C + ((size_t)j * h + b) * d, d);
    }
}

/* ------------------------------------------------------------------------------------------------
 * veccost (utils.jl:225-254).  Reconstruction is accumulated codebook by codebook in order
 * (CBi[j] += Ci[j,code], 238-243); the squared-error reduction uses canonical order (2).
 * B is 0-based here.
 * ---------------------------------------------------------------------------------------------- */
static float veccost_one(const float* x, const int16_t* b, const float* C, int m, int h, int d) {
  float p[32];
  for (int l = 0; l < 32; l++) p[l] = 0.0f;
  for (int t = 0; t < d; t++) {
    float r = 0.0f;
    for (int k = 0; k < m; k++) r = r + C[((size_t)k * h + b[k]) * d + t];
    float df = r - x[t];
    float sq = df * df;
    p[t & 31] = p[t & 31] + sq;
  }
  for (int off = 16; off >= 1; off >>= 1) {
    float q[32];
    for (int l = 0; l < 32; l++) q[l] = p[l] + p[l ^ off];
    memcpy(p, q, sizeof(p));
  }
  return p[0];
}

void orc_veccost(const float* X, const int16_t* B0, const float* C, int64_t n, int m, int h, int d,
                 float* cost) {
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < n; v++) cost[v] = veccost_one(X + (size_t)v * d, B0 + (size_t)v * m, C, m, h, d);
}

/* qerror (utils.jl:257-285): mean of the per-vector costs.  The reference accumulates everything in
 * one Float32 @simd accumulator (order unspecified); here: float64 sum of the canonical per-vector
 * float32 costs.  Compared with a relative tolerance, never bit-wise. */
double orc_qerror(const float* X, const int16_t* B0, const float* C, int64_t n, int m, int h, int d) {
  double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
  for (int64_t v = 0; v < n; v++)
    acc += (double)veccost_one(X + (size_t)v * d, B0 + (size_t)v * m, C, m, h, d);
  return n ? acc / (double)n : 0.0;
}

/* reconstruct (utils.jl:203-223) */
void orc_reconstruct(const int16_t* B0, const float* C, int64_t n, int m, int h, int d, float* CB) {
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < n; v++)
    for (int t = 0; t < d; t++) {
      float r = 0.0f;
      for (int k = 0; k < m; k++) r = r + C[((size_t)k * h + B0[(size_t)v * m + k]) * d + t];
      CB[(size_t)v * d + t] = r;
    }
}

/* splitarray (utils.jl:152-177): part p of nparts over 0..n-1 -> [*lo, *hi) */
void orc_splitarray(int64_t n, int nparts, int p, int64_t* lo, int64_t* hi) {
  int64_t per = n / nparts, xtra = n % nparts;
  if (p < xtra) { *lo = p * (per + 1); *hi = *lo + per + 1; }
  else { *lo = xtra * (per + 1) + (p - xtra) * per; *hi = *lo + per; }
}

/* ------------------------------------------------------------------------------------------------
 * encode_icm_fully! (encode_icm.jl:4-127) on vectors [0,n) of one worker, given precomputed unaries
 * and binaries, an explicit visit order and explicit perturbations.  Keeps the reference's loop
 * structure and memory behaviour (h-by-n `ub` buffer, node-major passes), so it doubles as the CPU
 * baseline.  B is 0-based [n][m], modified in place.
 * ---------------------------------------------------------------------------------------------- */
void orc_icm_fully(int16_t* B, const float* U /*[m][n][h]*/, const float* G, const float* Gt,
                   int64_t n, int m, int h, int niter, const int32_t* to_look, int npert,
                   const uint8_t* slots, const int16_t* vals) {
  /* cbpair2binaryidx (encode_icm.jl:31-34) */
  int pair2idx[ORC_MAXM][ORC_MAXM];
  int idx = 0;
  for (int i = 0; i < m; i++)
    for (int j = i + 1; j < m; j++) pair2idx[i][j] = idx++;

  /* perturb (encode_icm.jl:66-70) */
  for (int64_t v = 0; v < n; v++)
    for (int j = 0; j < npert; j++) B[v * m + slots[v * npert + j]] = vals[v * npert + j];

  float* ub = (float*)malloc(sizeof(float) * (size_t)h * n);
  for (int it = 0; it < niter; it++) {
    for (int jj = 0; jj < m; jj++) {
      const int j = to_look[jj];
      /* ub = unaries[j] (78-81) */
      memcpy(ub, U + (size_t)j * n * h, sizeof(float) * (size_t)h * n);
      /* condition on every other node, ascending k (84-102) */
      for (int k = 0; k < m; k++) {
        if (k == j) continue;
        /* j<k: binaries[(j,k)] column code_k.  j>k: transpose of binaries[(k,j)], column code_k.
         * Either way 2<C_j[:,a], C_k[:,code_k]> for a=0..h-1, contiguous (87-93). */
        const float* bb = (j < k) ? G + (size_t)pair2idx[j][k] * h * h : Gt + (size_t)pair2idx[k][j] * h * h;
        for (int64_t l = 0; l < n; l++) {
          const float* col = bb + (size_t)B[l * m + k] * h;
          float* u = ub + (size_t)l * h;
          for (int a = 0; a < h; a++) u[a] += col[a];
        }
      }
      /* first strict minimum (105-119) */
      for (int64_t l = 0; l < n; l++) {
        const float* u = ub + (size_t)l * h;
        float minv = u[0];
        int mini = 0;
        for (int a = 1; a < h; a++)
          if (u[a] < minv) { minv = u[a]; mini = a; }
        B[l * m + j] = (int16_t)mini;
      }
    }
  }
  free(ub);
}

static void transpose_tables(const float* G, int ncbi, int h, float* Gt) {
  for (int t = 0; t < ncbi; t++)
    for (int b = 0; b < h; b++)
      for (int a = 0; a < h; a++)
        Gt[((size_t)t * h + a) * h + b] = G[((size_t)t * h + b) * h + a];
}

/* ------------------------------------------------------------------------------------------------
 * encoding_icm (encode_icm.jl:131-189): ONE ILS iteration over the whole set with an explicit
 * schedule.  Worker fan-out (165-173) = `nworkers` contiguous splitarray parts run as OpenMP
 * threads; every part uses the same visit order (the GPU-path convention, encode_icm_cuda.jl:141).
 * oldB/newB are 0-based [n][m].  cost_out (optional) receives the cost of the returned codes.
 * ---------------------------------------------------------------------------------------------- */
void orc_encoding_icm_sched(const float* X, const int16_t* oldB, int16_t* newB, const float* C,
                            int64_t n, int m, int h, int d, int niter, const int32_t* to_look,
                            int npert, const uint8_t* slots, const int16_t* vals, int nworkers,
                            float* cost_out) {
  const int ncbi = m * (m - 1) / 2;
  float* G = (float*)malloc(sizeof(float) * (size_t)(ncbi ? ncbi : 1) * h * h);
  float* Gt = (float*)malloc(sizeof(float) * (size_t)(ncbi ? ncbi : 1) * h * h);
  int32_t* cbi = (int32_t*)malloc(sizeof(int32_t) * 2 * (ncbi ? ncbi : 1));
  orc_get_binaries(C, m, h, d, G, cbi);
  transpose_tables(G, ncbi, h, Gt);

  float* prevcost = (float*)malloc(sizeof(float) * (size_t)(n ? n : 1));
  float* newcost = (float*)malloc(sizeof(float) * (size_t)(n ? n : 1));
  orc_veccost(X, oldB, C, n, m, h, d, prevcost);
  memcpy(newB, oldB, sizeof(int16_t) * (size_t)n * m);

  if (nworkers < 1) nworkers = 1;
#pragma omp parallel for schedule(static, 1) num_threads(nworkers)
  for (int p = 0; p < nworkers; p++) {
    int64_t lo, hi;
    orc_splitarray(n, nworkers, p, &lo, &hi);
    const int64_t np = hi - lo;
    if (np <= 0) continue;
    /* each worker computes its own unaries (encode_icm.jl:16) */
    float* U = (float*)malloc(sizeof(float) * (size_t)m * np * h);
    {
      float* norms = (float*)malloc(sizeof(float) * m * h);
      float* tmp = (float*)malloc(sizeof(float) * h);
      orc_get_norms(C, m, h, d, norms);
      for (int i = 0; i < m; i++)
        for (int64_t v = 0; v < np; v++) {
          dots_fma(C + (size_t)i * h * d, X + (size_t)(lo + v) * d, h, d, tmp);
          float* u = U + ((size_t)i * np + v) * h;
          for (int a = 0; a < h; a++) u[a] = -2.0f * tmp[a] + norms[i * h + a];
        }
      free(norms); free(tmp);
    }
    orc_icm_fully(newB + lo * m, U, G, Gt, np, m, h, niter, to_look, npert, slots + lo * npert,
                  vals + lo * npert);
    free(U);
  }

  /* keep only strictly better codes (encode_icm.jl:178-186) */
  orc_veccost(X, newB, C, n, m, h, d, newcost);
  for (int64_t v = 0; v < n; v++) {
    if (!(newcost[v] < prevcost[v])) {
      memcpy(newB + v * m, oldB + v * m, sizeof(int16_t) * m);
      newcost[v] = prevcost[v];
    }
  }
  if (cost_out) memcpy(cost_out, newcost, sizeof(float) * (size_t)n);
  free(G); free(Gt); free(cbi); free(prevcost); free(newcost);
}

/* encoding_icm with the canonical Philox schedule.  g0 = global index of vector 0 (sharding). */
void orc_encoding_icm(const float* X, const int16_t* oldB, int16_t* newB, const float* C, int64_t n,
                      int m, int h, int d, int niter, int randord, int npert, uint64_t seed,
                      uint32_t ils_iter, uint64_t g0, int nworkers, float* cost_out) {
  int32_t to_look[ORC_MAXM];
  orc_make_to_look(seed, ils_iter, m, randord, to_look);
  uint8_t* slots = (uint8_t*)malloc((size_t)(n * npert + 1));
  int16_t* vals = (int16_t*)malloc(sizeof(int16_t) * (size_t)(n * npert + 1));
  orc_make_perturb(seed, ils_iter, g0, n, m, h, npert, slots, vals);
  orc_encoding_icm_sched(X, oldB, newB, C, n, m, h, d, niter, to_look, npert, slots, vals, nworkers,
                         cost_out);
  free(slots); free(vals);
}

/* encode_icm_cuda_single (encode_icm_cuda.jl:22-234): all ILS iterations, snapshots + objectives at
 * the requested iteration counts.  ILS iteration i (1-based in the reference) uses ils_iter = i-1.
 * Bs is [nr][n][m] 0-based. */
void orc_encode_icm_ils(const float* X, const int16_t* B, const float* C, int64_t n, int m, int h,
                        int d, const int64_t* ilsiters, int nr, int icmiter, int npert, int randord,
                        uint64_t seed, uint64_t g0, int nworkers, int16_t* Bs, float* objs) {
  int64_t maxit = 0;
  for (int r = 0; r < nr; r++) if (ilsiters[r] > maxit) maxit = ilsiters[r];
  int16_t* cur = (int16_t*)malloc(sizeof(int16_t) * (size_t)(n * m + 1));
  int16_t* nxt = (int16_t*)malloc(sizeof(int16_t) * (size_t)(n * m + 1));
  memcpy(cur, B, sizeof(int16_t) * (size_t)n * m);
  for (int64_t i = 1; i <= maxit; i++) {
    orc_encoding_icm(X, cur, nxt, C, n, m, h, d, icmiter, randord, npert, seed, (uint32_t)(i - 1), g0,
                     nworkers, NULL);
    int16_t* t = cur; cur = nxt; nxt = t;
    for (int r = 0; r < nr; r++)
      if (ilsiters[r] == i) { /* find(i .== ilsiters)[1]: first match only (213) */
        memcpy(Bs + (size_t)r * n * m, cur, sizeof(int16_t) * (size_t)n * m);
        objs[r] = (float)orc_qerror(X, cur, C, n, m, h, d);
        break;
      }
  }
  free(cur); free(nxt);
}

/* ------------------------------------------------------------------------------------------------
 * Linear scan restatements.  These are checked against the reference's own C++ (oracle/_ref).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float d; int32_t i; } orc_pair;
static int pair_cmp(const void* a, const void* b) {
  const orc_pair* x = (const orc_pair*)a; const orc_pair* y = (const orc_pair*)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return (x->i > y->i) - (x->i < y->i);
}

/* LUT of linscan_aqd_pairwise_byte.cpp:42-48: tentry[j] -= 2*query[k]*centry[k], k ascending. */
void orc_lut_lsq(const float* query, const float* codebooks, int mh, int d, float* lut) {
  for (int j = 0; j < mh; j++) {
    const float* c = codebooks + (size_t)j * d;
    float t = 0.0f;
    for (int k = 0; k < d; k++) t -= 2 * query[k] * c[k];
    lut[j] = t;
  }
}

/* LUT of linscan_aqd.cpp:66-74: dis[t] += sqr(centers[t*subdim+s] - q[k*subdim+s]). */
void orc_lut_pq(const float* query, const float* centers, int m, int h, int subdim, float* lut) {
  for (int k = 0; k < m; k++)
    for (int r = 0; r < h; r++) {
      int t = k * h + r;
      float acc = 0.0f;
      for (int s = 0; s < subdim; s++) {
        float df = centers[(size_t)t * subdim + s] - query[k * subdim + s];
        acc += df * df;
      }
      lut[t] = acc;
    }
}

/* _linscan_aqd_query_extra_byte (linscan_aqd_pairwise_byte.cpp:14-93): ids 1-based, ascending by
 * (distance, id) (partial_sort on pair<float,int>, :81).  A full sort of all pairs followed by a
 * prefix gives the same first nn entries as the reference's chunked partial_sort. */
void orc_linscan_lsq(float* dists, int32_t* idx, const uint8_t* codes, const float* queries,
                     const float* codebooks, const float* dbnorms, int nq, int n, int m, int h, int d,
                     int nn) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int q = 0; q < nq; q++) {
    float* lut = (float*)malloc(sizeof(float) * m * h);
    orc_pair* pairs = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)(n ? n : 1));
    orc_lut_lsq(queries + (size_t)q * d, codebooks, m * h, d, lut);
    for (int i = 0; i < n; i++) {
      float s = 0;
      for (int k = 0; k < m; k++) s += lut[h * k + codes[(size_t)i * m + k]];
      s += dbnorms[i];
      pairs[i].d = s; pairs[i].i = i + 1;
    }
    qsort(pairs, n, sizeof(orc_pair), pair_cmp);
    for (int j = 0; j < nn; j++) { dists[(size_t)q * nn + j] = pairs[j].d; idx[(size_t)q * nn + j] = pairs[j].i; }
    free(lut); free(pairs);
  }
}

/* _linscan_aqd_query (linscan_aqd.cpp:37-102): ids 0-based, no norm term. */
void orc_linscan_pq(float* dists, uint32_t* res, const uint8_t* codes, const float* centers,
                    const float* queries, int n, int nq, int m, int h, int K, int dim1queries,
                    int subdim) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int q = 0; q < nq; q++) {
    float* lut = (float*)malloc(sizeof(float) * m * h);
    orc_pair* pairs = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)(n ? n : 1));
    orc_lut_pq(queries + (size_t)q * dim1queries, centers, m, h, subdim, lut);
    for (int i = 0; i < n; i++) {
      float s = 0;
      for (int k = 0; k < m; k++) s += lut[h * k + codes[(size_t)i * m + k]];
      pairs[i].d = s; pairs[i].i = i;
    }
    qsort(pairs, n, sizeof(orc_pair), pair_cmp);
    for (int j = 0; j < K; j++) { dists[(size_t)q * K + j] = pairs[j].d; res[(size_t)q * K + j] = (uint32_t)pairs[j].i; }
    free(lut); free(pairs);
  }
}

/* quantize_norms (utils.jl:6-31): nearest norm-codebook entry to ||reconstruction||^2, first min. */
void orc_quantize_norms(const int16_t* B0, const float* C, const float* cbnorms, int64_t n, int m,
                        int h, int d, int hn, int16_t* out0) {
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < n; v++) {
    float nrm = 0.0f;
    for (int t = 0; t < d; t++) {
      float r = 0.0f;
      for (int k = 0; k < m; k++) r = r + C[((size_t)k * h + B0[(size_t)v * m + k]) * d + t];
      nrm = nrm + r * r;
    }
    float best = (nrm - cbnorms[0]) * (nrm - cbnorms[0]);
    int bi = 0;
    for (int j = 1; j < hn; j++) {
      float dd = (nrm - cbnorms[j]) * (nrm - cbnorms[j]);
      if (dd < best) { best = dd; bi = j; }
    }
    out0[v] = (int16_t)bi;
  }
}

/* ------------------------------------------------------------------------------------------------
 * encoding_viterbi (encode_chain.jl:95-127) + encode_viterbi! (encode_chain.jl:1-92): exact MAP on the
 * chain 1-2-...-m by dynamic programming (the ChainQ encoder).
 *   binaries[i] = 2*C[i]'*C[i+1]  (:104-106; canonical dot = sequential-k FMA chain, then *2)
 *   forward (:41-70): U[:,i] += mincost[:,i-1] (i>1); mincost[j,i] = min_k U[k,i] + bb[k,j], first
 *   strict minimum over k (:56-65); last node: U[:,m] += mincost[:,m-1], findmin (:72-76);
 *   backward trace (:78-85).  Codes out 0-based B0[n][m].  Needs m >= 2.
 * ---------------------------------------------------------------------------------------------- */
void orc_encoding_viterbi(const float* X, const float* C, int64_t n, int m, int h, int d, int16_t* B0) {
  float* norms = (float*)malloc(sizeof(float) * m * h);
  orc_get_norms(C, m, h, d, norms);
  /* bb[i][j][k] = 2<C_i[:,k], C_{i+1}[:,j]>  (Julia bb[k,j], column-major: k contiguous) */
  float* bb = (float*)malloc(sizeof(float) * (size_t)(m - 1) * h * h);
  for (int i = 0; i < m - 1; i++) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < h; j++)
      for (int k = 0; k < h; k++)
        bb[((size_t)i * h + j) * h + k] =
            2.0f * dot_fma(C + ((size_t)i * h + k) * d, C + ((size_t)(i + 1) * h + j) * d, d);
  }
#pragma omp parallel
  {
    float* U = (float*)malloc(sizeof(float) * m * h);
    float* tmp = (float*)malloc(sizeof(float) * h);
    float* mincost = (float*)malloc(sizeof(float) * m * h);
    int* minidx = (int*)malloc(sizeof(int) * m * h);
#pragma omp for schedule(static)
    for (int64_t v = 0; v < n; v++) {
      for (int i = 0; i < m; i++) { /* get_unaries (utils.jl:94-122) for this vector */
        dots_fma(C + (size_t)i * h * d, X + (size_t)v * d, h, d, tmp);
        for (int a = 0; a < h; a++) U[i * h + a] = -2.0f * tmp[a] + norms[i * h + a];
      }
      for (int i = 0; i < m - 1; i++) {
        if (i > 0)
          for (int j = 0; j < h; j++) U[i * h + j] += mincost[(i - 1) * h + j];
        const float* b = bb + (size_t)i * h * h;
        for (int j = 0; j < h; j++) {
          float minv = U[i * h + 0] + b[(size_t)j * h + 0];
          int mini = 0;
          for (int k = 1; k < h; k++) {
            const float c = U[i * h + k] + b[(size_t)j * h + k];
            if (c < minv) { minv = c; mini = k; }
          }
          mincost[i * h + j] = minv;
          minidx[i * h + j] = mini;
        }
      }
      for (int j = 0; j < h; j++) U[(m - 1) * h + j] += mincost[(m - 2) * h + j];
      float minv = U[(m - 1) * h];
      int mini = 0;
      for (int j = 1; j < h; j++)
        if (U[(m - 1) * h + j] < minv) { minv = U[(m - 1) * h + j]; mini = j; }
      B0[v * m + (m - 1)] = (int16_t)mini;
      for (int i = m - 2; i >= 0; i--) {
        mini = minidx[i * h + mini];
        B0[v * m + i] = (int16_t)mini;
      }
    }
    free(U); free(tmp); free(mincost); free(minidx);
  }
  free(bb);
  free(norms);
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline legs of bench.py ask for all host
 * cores explicitly so that the reference arm is not handicapped under N > 1 launches. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
