// adc_tc.cuh — tensor-core prefilter of the exact ADC scan (see adc_tc.cu).
#pragma once
#include "linscan.cuh"

namespace lsq {

// bit patterns of non-negative floats (ordered like unsigned integers): maxima gathered with atomicMax
struct AdcStats {
  unsigned int xmax2_bits;  // max_v ||xhat_v||^2
  unsigned int nmax_bits;   // max_v |dbnorm_v|
  unsigned int cmax2_bits;  // max codeword ||c||^2
  unsigned int xlo2_bits;   // max_v ||xhat_v - bf16(xhat_v)||^2
};

// per-call image of the base set: bf16 hi/lo UMMA operand tiles of the decoded vectors (norms folded in), maxima;
// the same for the strided sample the thresholds are estimated on
struct AdcTcBase {
  DevBuf<unsigned char> img, simg, s1img;
  DevBuf<AdcStats> stats;
  int64_t ntiles = 0, stiles = 0, scount = 0, sstride = 1, s1tiles = 0, s1count = 0;
  int subdim = 0;    // 0: LSQ tables (-2<q,c> + dbnorm); > 0: PQ / OPQ tables (squared differences per sub-space)
  int qstride = 0;   // floats between query rows
};

bool adc_tc_shape_ok(int64_t n, int64_t nq, int m, int d);
bool adc_tc_applicable(const uint8_t* dcodes, int64_t n, int64_t nq, int m, int d, const float* dqueries,
                       const float* dcodebooks);
// base image + image of the sample {i * sstride : i < scount}
// PQ / OPQ: subdim > 0, d = m * subdim, dcodebooks = centers [m][256][subdim], dbnorms = nullptr
int adc_tc_prepare(const uint8_t* dcodes, int64_t n, int m, const float* dcodebooks, int d, const float* dbnorms,
                   int64_t scount, int64_t sstride, int subdim, int qstride, cudaStream_t st, AdcTcBase& B);
// filter values of the sample, ordered, in threshold_kernel's layout with 32-query tiles: [(q/32 * scount + t) * 32 + q%32]
// subsample: the 1/8 sub-sample (B.s1count steps) instead of the sample (B.scount steps)
int adc_tc_sample(const AdcTcBase& B, bool subsample, const float* dq, int nb, int d, int m, uint32_t* dsbuf,
                  cudaStream_t st);
// thresholds without a sample buffer: sample positions with filter value <= dbound[q] (+ margin) -> dlist / dlcnt
// (lcap per query), scored exactly, r-th smallest -> dtau[q] (+inf if the list overflowed or is shorter than r);
// dinfl / dnpass (optional): per-query estimate of how much a one-product filter would lengthen the survivor lists,
// and the number of products chosen from its mean (1 or 2), read by the main pass on the device
int adc_tc_sample_tau(const AdcTcBase& B, const uint8_t* dcodes, int m, const float* dq, int nb, int d,
                      const float* dbnorms, const float* dlutq, const float* dbound, uint32_t* dlist, int* dlcnt,
                      int lcap, int r, float* dtau, float* dinfl, int* dnpass, cudaStream_t st);
// exact LUT rows lutq[q][m*256] (the reference's fp32 chain) for the rescoring
int adc_tc_lut_rows(const AdcTcBase& B, const float* dq, int nb, int d, const float* dcodebooks, int m, float* dlutq,
                    cudaStream_t st);
// filter (-> dcandidx / dccnt) + exact rescoring (-> dcand / dcnt, the buffers the top-k kernels read); dtau[q].
// dcand == nullptr: filter only.  ddbg (optional): [nb][dbg_ld] filter values.
int adc_tc_main_pass(const AdcTcBase& B, const uint8_t* dcodes, int64_t n, int m, const float* dq, int nb, int d,
                     const float* dbnorms, const float* dlutq, const float* dtau, uint32_t* dcandidx, int* dccnt,
                     int64_t ccap, unsigned long long* dcand, int* dcnt, int64_t cap, int id_base, float* ddbg,
                     int64_t dbg_ld, const int* dnpass, cudaStream_t st);

// exact rescoring of the filter's survivors: dcandidx / dccnt -> keys in dcand / dcnt (the buffers the top-k kernels read)
int adc_tc_rescore(const uint8_t* dcodes, int64_t n, int m, int nb, const float* dbnorms, const float* dlutq,
                   const float* dtau, const uint32_t* dcandidx, const int* dccnt, int64_t ccap, unsigned long long* dcand,
                   int* dcnt, int64_t cap, int id_base, cudaStream_t st);

}  // namespace lsq
