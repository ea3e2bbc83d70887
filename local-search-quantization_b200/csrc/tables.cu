// tables.cu — unary / binary table builders (get_unaries utils.jl:94-122, get_binaries utils.jl:125-144).
//
// Canonical arithmetic (frozen with the oracle, oracle/lsq_oracle.c dot_fma): every dot product is a
// sequential-k fp32 FMA chain from 0, k ascending.  A classic shared-memory tiled SIMT GEMM keeps
// exactly that order per output element as long as the K chunks are visited in ascending order, so
// the kernel below is bit-identical to the oracle.  (A tcgen05 tensor-core version cannot reproduce a
// sequential fp32 chain; it belongs to the tolerance-checked "fast" mode, not to parity mode.)
#include "icm.cuh"

namespace lsq {

constexpr int TM = 128;  // candidates per block tile
constexpr int TN = 128;  // rows (vectors) per block tile
constexpr int TK = 16;
constexpr int TPAD = 4;

enum { EPI_UNARY = 0, EPI_BINARY = 1 };

// out[z][r][a] = epi( sum_k A_z[a][k] * B_z[r][k] ),  a < 256, r < R
//   A_z = Abase + az(z) * 256*d        B_z = Bbase + bz(z) * R*d (binary) / Bbase (unary)
template <int EPI>
__global__ void __launch_bounds__(256, 2) gemm_tables_kernel(const float* __restrict__ Abase,
                                                          const float* __restrict__ Bbase,
                                                          const float* __restrict__ norms, float* __restrict__ out,
                                                          int64_t R, int d, int m, int sliced, int64_t Rs) {
  __shared__ __align__(16) float As[TK][TM + TPAD];
  __shared__ __align__(16) float Bs[TK][TN + TPAD];

  const int z = blockIdx.z;
  const float* A;
  const float* B;
  float* O;
  if (EPI == EPI_UNARY) {
    A = Abase + (size_t)z * LSQ_H * d;
    B = Bbase;
    O = out + (size_t)z * R * LSQ_H;
  } else {
    const int j = z / m, k = z % m;
    if (j == k) return;
    A = Abase + (size_t)j * LSQ_H * d;
    B = Bbase + (size_t)k * LSQ_H * d;
    O = out + (size_t)z * LSQ_H * LSQ_H;
  }
  const int a0 = blockIdx.x * TM;
  const int64_t r0 = (int64_t)blockIdx.y * TN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;

  const bool vec_ok = (d % 4 == 0);
  // register-staged prefetch (vector path): the global loads of K tile t+1 are issued before the FMAs of
  // tile t, so their latency hides behind 16 x 64 FMAs per thread instead of sitting between two barriers
  float4 pa[2], pb[2];
  auto fetch = [&](int kk) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int row = (tid >> 2) + 64 * i;
      const int k4 = (tid & 3) * 4;
      pa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      pb[i] = pa[i];
      if (kk + k4 < d) {
        pa[i] = *reinterpret_cast<const float4*>(A + (size_t)(a0 + row) * d + kk + k4);
        if (r0 + row < R) pb[i] = *reinterpret_cast<const float4*>(B + (size_t)(r0 + row) * d + kk + k4);
      }
    }
  };
  if (vec_ok) fetch(0);
  for (int kk = 0; kk < d; kk += TK) {
    // stage A[a0..a0+127][kk..kk+15] and B[r0..r0+127][kk..kk+15], transposed to k-major
    if (vec_ok) {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int row = (tid >> 2) + 64 * i;
        const int k4 = (tid & 3) * 4;
        As[k4 + 0][row] = pa[i].x; As[k4 + 1][row] = pa[i].y; As[k4 + 2][row] = pa[i].z; As[k4 + 3][row] = pa[i].w;
        Bs[k4 + 0][row] = pb[i].x; Bs[k4 + 1][row] = pb[i].y; Bs[k4 + 2][row] = pb[i].z; Bs[k4 + 3][row] = pb[i].w;
      }
    } else {
      for (int e = tid; e < TM * TK; e += 256) {
        const int row = e / TK, k = e % TK;
        float va = 0.f, vb = 0.f;
        if (kk + k < d) {
          va = A[(size_t)(a0 + row) * d + kk + k];
          if (r0 + row < R) vb = B[(size_t)(r0 + row) * d + kk + k];
        }
        As[k][row] = va;
        Bs[k][row] = vb;
      }
    }
    __syncthreads();
    if (vec_ok && kk + TK < d) fetch(kk + TK);
#pragma unroll
    for (int k = 0; k < TK; k++) {
      const float4 a_lo = *reinterpret_cast<const float4*>(&As[k][tx * 4]);
      const float4 a_hi = *reinterpret_cast<const float4*>(&As[k][64 + tx * 4]);
      const float4 b_lo = *reinterpret_cast<const float4*>(&Bs[k][ty * 4]);
      const float4 b_hi = *reinterpret_cast<const float4*>(&Bs[k][64 + ty * 4]);
      const float av[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
      const float bv[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[j], bv[i], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue: rows r = r0 + ty*4 + {0..3} and r0 + 64 + ty*4 + {0..3}; cols a0 + tx*4.., a0+64+tx*4..
  float nlo[4] = {0, 0, 0, 0}, nhi[4] = {0, 0, 0, 0};
  if (EPI == EPI_UNARY) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      nlo[j] = norms[z * LSQ_H + a0 + tx * 4 + j];
      nhi[j] = norms[z * LSQ_H + a0 + 64 + tx * 4 + j];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int64_t r = r0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= R) continue;
    float4 lo, hi;
    if (EPI == EPI_UNARY) {
      // fl(-2*dot) + ||c||^2  (utils.jl:108,116) — separate multiply and add, never contracted
      lo.x = __fadd_rn(__fmul_rn(-2.0f, acc[i][0]), nlo[0]);
      lo.y = __fadd_rn(__fmul_rn(-2.0f, acc[i][1]), nlo[1]);
      lo.z = __fadd_rn(__fmul_rn(-2.0f, acc[i][2]), nlo[2]);
      lo.w = __fadd_rn(__fmul_rn(-2.0f, acc[i][3]), nlo[3]);
      hi.x = __fadd_rn(__fmul_rn(-2.0f, acc[i][4]), nhi[0]);
      hi.y = __fadd_rn(__fmul_rn(-2.0f, acc[i][5]), nhi[1]);
      hi.z = __fadd_rn(__fmul_rn(-2.0f, acc[i][6]), nhi[2]);
      hi.w = __fadd_rn(__fmul_rn(-2.0f, acc[i][7]), nhi[3]);
    } else {
      lo = make_float4(2.0f * acc[i][0], 2.0f * acc[i][1], 2.0f * acc[i][2], 2.0f * acc[i][3]);
      hi = make_float4(2.0f * acc[i][4], 2.0f * acc[i][5], 2.0f * acc[i][6], 2.0f * acc[i][7]);
    }
    if (!sliced) {
      float* orow = O + (size_t)r * LSQ_H + a0;
      *reinterpret_cast<float4*>(orow + tx * 4) = lo;
      *reinterpret_cast<float4*>(orow + 64 + tx * 4) = hi;
    } else {
      // U[slice][r][32]: candidate a lives in slice a/32 at offset a%32
      const int alo = a0 + tx * 4, ahi = a0 + 64 + tx * 4;
      *reinterpret_cast<float4*>(O + ((size_t)(alo >> 5) * Rs + r) * ICM_SLICE_W + (alo & 31)) = lo;
      *reinterpret_cast<float4*>(O + ((size_t)(ahi >> 5) * Rs + r) * ICM_SLICE_W + (ahi & 31)) = hi;
    }
  }
}

// norms[i] = sum_k c[k]^2 as an FMA chain (diag(C'C), utils.jl:109)
__global__ void norms_kernel(const float* __restrict__ C, int rows, int d, float* __restrict__ norms) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* c = C + (size_t)i * d;
  float acc = 0.0f;
  for (int k = 0; k < d; k++) acc = fmaf(c[k], c[k], acc);
  norms[i] = acc;
}

int build_norms(const float* dC, int d, int m, float* dnorms, cudaStream_t st) {
  const int rows = m * LSQ_H;
  note_launch();
  norms_kernel<<<(rows + 127) / 128, 128, 0, st>>>(dC, rows, d, dnorms);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

int build_unaries(const float* dX, int d, int64_t n, const float* dC, int m, const float* dnorms, float* dU,
                  int sliced, cudaStream_t st) {
  if (n == 0) return LSQ_OK;
  // gridDim.y is limited to 65535 tiles of 128 rows (8.3 M vectors) per launch
  const int64_t max_rows = (int64_t)65535 * TN;
  for (int64_t r0 = 0; r0 < n; r0 += max_rows) {
    const int64_t rows = (n - r0 < max_rows) ? (n - r0) : max_rows;
    if (r0 == 0 && rows == n) {
      dim3 grid(LSQ_H / TM, (unsigned)ceil_div(rows, TN), m);
      note_launch();
      gemm_tables_kernel<EPI_UNARY><<<grid, 256, 0, st>>>(dC, dX, dnorms, dU, n, d, m, sliced, n);
    } else {
      // one launch per codebook when the launch is chunked, since U's z-stride is n*256
      for (int j = 0; j < m; j++) {
        dim3 grid(LSQ_H / TM, (unsigned)ceil_div(rows, TN), 1);
        float* o = dU + (size_t)j * n * LSQ_H + (sliced ? (size_t)r0 * ICM_SLICE_W : (size_t)r0 * LSQ_H);
        note_launch();
        gemm_tables_kernel<EPI_UNARY><<<grid, 256, 0, st>>>(dC + (size_t)j * LSQ_H * d, dX + (size_t)r0 * d,
                                                            dnorms + j * LSQ_H, o, rows, d, m, sliced, n);
      }
    }
    LSQ_CUDA(cudaGetLastError());
  }
  return LSQ_OK;
}

int build_tables(const float* dC, int d, int m, float* dT, cudaStream_t st) {
  dim3 grid(LSQ_H / TM, LSQ_H / TN, m * m);
  note_launch();
  gemm_tables_kernel<EPI_BINARY><<<grid, 256, 0, st>>>(dC, dC, nullptr, dT, LSQ_H, d, m, 0, LSQ_H);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

// Ts[j][s][kk][b][w] = T[j][k(kk)][b][32 s + w], k(kk) = kk-th codebook != j in ascending order: the
// (m-1)*256 rows of 128 bytes that node j / candidate slice s needs, contiguous for one TMA bulk load.
__global__ void slice_tables_kernel(const float* __restrict__ T, int m, float* __restrict__ Ts) {
  const int64_t total = (int64_t)m * ICM_SLICES * (m - 1) * LSQ_H * ICM_SLICE_W;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(e % ICM_SLICE_W);
    int64_t r = e / ICM_SLICE_W;
    const int b = (int)(r % LSQ_H); r /= LSQ_H;
    const int kk = (int)(r % (m - 1)); r /= (m - 1);
    const int s = (int)(r % ICM_SLICES);
    const int j = (int)(r / ICM_SLICES);
    const int k = kk + (kk >= j ? 1 : 0);
    Ts[e] = T[(((size_t)j * m + k) * LSQ_H + b) * LSQ_H + s * ICM_SLICE_W + w];
  }
}

int build_sliced_tables(const float* dT, int m, float* dTs, cudaStream_t st) {
  if (m < 2) return LSQ_OK;
  note_launch();
  slice_tables_kernel<<<LSQ_NUM_SMS_HINT * 8, 256, 0, st>>>(dT, m, dTs);
  LSQ_CUDA(cudaGetLastError());
  return LSQ_OK;
}

}  // namespace lsq
