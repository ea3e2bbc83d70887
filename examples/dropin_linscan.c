/*
 * dropin_linscan.c — the drop-in claim at the C level, without Python or Julia in between.
 *
 * Loads TWO shared objects that export the reference's symbol `linscan_aqd_query_extra_byte`
 * (src/linscan/cpp/linscan_aqd_pairwise_byte.cpp:97-104, bound by Linscan.jl:63-69):
 *   argv[1]  liblsq_b200.so                         (this repository: CUDA, sm_100a)
 *   argv[2]  linscan_aqd_pairwise_byte.so           (the reference's own C++, e.g. oracle/_ref/)
 * calls both exactly as the Julia ccall does — same argument list, caller-allocated outputs — on the same
 * pseudo-random problem, and compares ids and distances bit for bit.  Exit code 0 = identical.
 *
 *   gcc -O2 -o dropin_linscan examples/dropin_linscan.c -ldl
 *   ./dropin_linscan local-search-quantization_b200/liblsq_b200.so oracle/_ref/linscan_aqd_pairwise_byte.so
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void (*scan_fn)(float* dists, int* idx, unsigned char* codes, float* queries, float* codebooks,
                        float* dbnorms, int nqueries, int ncodes, int m, int h, int d, int nn);

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) {  /* xorshift64* */
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return (uint32_t)((rng_state * 0x2545F4914F6CDD1Dull) >> 32);
}

static scan_fn load(const char* path) {
  void* hnd = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!hnd) { fprintf(stderr, "dlopen(%s): %s\n", path, dlerror()); exit(2); }
  scan_fn f = (scan_fn)dlsym(hnd, "linscan_aqd_query_extra_byte");
  if (!f) { fprintf(stderr, "%s does not export linscan_aqd_query_extra_byte\n", path); exit(2); }
  return f;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s liblsq_b200.so reference_linscan.so [n nq]\n", argv[0]); return 2; }
  const int n = argc > 3 ? atoi(argv[3]) : 200000, nq = argc > 4 ? atoi(argv[4]) : 40;
  const int m = 8, h = 256, d = 64, nn = 100;
  unsigned char* codes = malloc((size_t)n * m);
  float* queries = malloc(sizeof(float) * (size_t)nq * d);
  float* codebooks = malloc(sizeof(float) * (size_t)m * h * d);
  float* norms = malloc(sizeof(float) * (size_t)n);
  for (size_t i = 0; i < (size_t)n * m; i++) codes[i] = (unsigned char)(rnd() & 0xFF);
  for (size_t i = 0; i < (size_t)nq * d; i++) queries[i] = (float)(rnd() % 200);
  for (size_t i = 0; i < (size_t)m * h * d; i++) codebooks[i] = (float)(rnd() % 64) * 0.25f;
  for (int i = 0; i < n; i++) norms[i] = (float)(rnd() % 100000) * 0.01f;
  /* duplicates make exact ties: the order must still agree (distance, then lower id) */
  memcpy(codes + (size_t)(n / 2) * m, codes, (size_t)1000 * m);
  memcpy(norms + n / 2, norms, sizeof(float) * 1000);

  float* d1 = calloc((size_t)nq * nn, sizeof(float)); int* i1 = calloc((size_t)nq * nn, sizeof(int));
  float* d2 = calloc((size_t)nq * nn, sizeof(float)); int* i2 = calloc((size_t)nq * nn, sizeof(int));
  scan_fn ours = load(argv[1]), ref = load(argv[2]);
  ours(d1, i1, codes, queries, codebooks, norms, nq, n, m, h, d, nn);
  ref(d2, i2, codes, queries, codebooks, norms, nq, n, m, h, d, nn);
  const int same_ids = memcmp(i1, i2, sizeof(int) * (size_t)nq * nn) == 0;
  const int same_d = memcmp(d1, d2, sizeof(float) * (size_t)nq * nn) == 0;
  printf("n=%d nq=%d m=%d nn=%d: ids %s, distances %s (first hit: id %d, dist %.4f)\n", n, nq, m, nn,
         same_ids ? "identical" : "DIFFER", same_d ? "identical" : "DIFFER", i1[0], d1[0]);
  return (same_ids && same_d) ? 0 : 1;
}
