// adc_tc.cuh — tensor-core prefilter of the exact ADC scan (see adc_tc.cu).
#pragma once
#include "linscan.cuh"

namespace lsq {

// bit patterns of non-negative floats (ordered like unsigned integers): maxima gathered with atomicMax
struct AdcStats {
  unsigned int xmax2_bits;  // max_v ||xhat_v||^2
  unsigned int nmax_bits;   // max_v |dbnorm_v|
  unsigned int cmax2_bits;  // max codeword ||c||^2
  unsigned int pad;
};

// per-call image of the base set: bf16 hi/lo UMMA operand tiles of the decoded vectors, padded norms, maxima
struct AdcTcBase {
  DevBuf<unsigned char> img;
  DevBuf<float> normpad;
  DevBuf<AdcStats> stats;
  int64_t ntiles = 0;
};

bool adc_tc_applicable(int64_t n, int m, int d, const float* dqueries, const float* dcodebooks, const float* dbnorms);
int adc_tc_prepare(const uint8_t* dcodes, int64_t n, int m, const float* dcodebooks, int d, const float* dbnorms,
                   cudaStream_t st, AdcTcBase& B);
// filter (-> dcandidx / dccnt) + exact rescoring (-> dcand / dcnt, the buffers the top-k kernels read).
// dcand == nullptr: filter only.  ddbg (optional): [nb][dbg_ld] filter values.
int adc_tc_main_pass(const AdcTcBase& B, const uint8_t* dcodes, int64_t n, int m, const float* dq, int nb, int d,
                     const float* dbnorms, const float* dlut, int QT, const float* dtau, uint32_t* dcandidx,
                     int* dccnt, int64_t ccap, unsigned long long* dcand, int* dcnt, int64_t cap, int id_base,
                     float* ddbg, int64_t dbg_ld, cudaStream_t st);

}  // namespace lsq
