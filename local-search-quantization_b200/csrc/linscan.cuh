// linscan.cuh — ADC linear scan launcher (linscan_aqd_pairwise_byte.cpp / linscan_aqd.cpp replacement).
#pragma once
#include "common.cuh"

namespace lsq {

enum { LUT_LSQ = 0, LUT_PQ = 1 };

constexpr int LINSCAN_MAX_NN = 16384;  // top-k capacity of the shared-memory sorter

// All pointers are device pointers.  LSQ: codebooks float[m*256][d], dbnorms float[n], ids 1-based.
// PQ: codebooks = centers float[m][256][subdim], queries row stride = d, ids 0-based.
int linscan_device(const uint8_t* dcodes, int64_t n, int m, const float* dqueries, int nq, int d,
                   const float* dcodebooks, const float* dbnorms, int lut_kind, int subdim, int nn,
                   float* ddists, int32_t* dids, cudaStream_t st);

}  // namespace lsq
