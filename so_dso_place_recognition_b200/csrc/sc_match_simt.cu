// Plain fp32 CUDA-core implementation of processSC.m:12-45.  NOT the product path: it is the
// on-GPU cross-check for the tcgen05 kernel (sc_match_tc.cu) at sizes the CPU oracle cannot
// reach in test time, selectable with sodso_ctx_set_match_algo(SODSO_ALGO_SIMT).
#include "../../include/sodso_pr.h"
#include "common.cuh"

namespace sodso {
namespace {

constexpr int PREP_THREADS = 128;

// processSC.m:15-20: row / norm(row), per channel.  out_t: [ch][1200][ld] fp32 (K-major).
__global__ void __launch_bounds__(PREP_THREADS)
sc_prep_simt_kernel(const double *__restrict__ hist, int rows, float *__restrict__ out_t, int ld) {
  __shared__ double red[2][PREP_THREADS / 32];
  const int row = blockIdx.x;
  const double *h = hist + (size_t)row * 2 * SC_SIZE;
  double ss[2] = {0.0, 0.0};
  for (int ch = 0; ch < 2; ch++)
    for (int k = threadIdx.x; k < SC_SIZE; k += PREP_THREADS) {
      double v = h[ch * SC_SIZE + k];
      ss[ch] += v * v;
    }
  for (int ch = 0; ch < 2; ch++) {
    double s = ss[ch];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[ch][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  double nrm[2];
  for (int ch = 0; ch < 2; ch++) {
    double s = 0.0;
    for (int w = 0; w < PREP_THREADS / 32; w++) s += red[ch][w];
    nrm[ch] = sqrt(s);
  }
  for (int ch = 0; ch < 2; ch++)
    for (int k = threadIdx.x; k < SC_SIZE; k += PREP_THREADS)
      out_t[((size_t)ch * SC_SIZE + k) * ld + row] = (float)(h[ch * SC_SIZE + k] / nrm[ch]);
}

constexpr int SIMT_THREADS = 128;
constexpr int SIMT_ST = 10;  // shifts per sweep (per base vector)

// One CTA: one query x 128 DB rows.  All 120 variants (processSC.m:24-28): 60 forward shifts
// of the query image x and 60 forward shifts of its sector-reversed image y
// (reverse shift k of x == forward shift (60-k+1) mod 60 of y).
__global__ void __launch_bounds__(SIMT_THREADS)
sc_match_simt_kernel(const float *__restrict__ q_t, int m, int ldq, const float *__restrict__ h_t,
                     int n, int ldh, float *__restrict__ d_p, float *__restrict__ d_i, int ldd) {
  __shared__ float x2[2 * SC_SIZE], y2[2 * SC_SIZE];
  const int qi = blockIdx.y;
  const int j = blockIdx.x * SIMT_THREADS + threadIdx.x;
  const int jc = j < n ? j : n - 1;
  for (int ch = 0; ch < 2; ch++) {
    __syncthreads();
    const float *qv = q_t + (size_t)ch * SC_SIZE * ldq + qi;
    for (int k = threadIdx.x; k < SC_SIZE; k += SIMT_THREADS) {
      float v = qv[(size_t)k * ldq];
      int c = k / SC_NUM_R, r = k - c * SC_NUM_R;
      int cr = (SC_NUM_S - c) % SC_NUM_S;
      x2[k] = v;
      x2[k + SC_SIZE] = v;
      y2[cr * SC_NUM_R + r] = v;
      y2[cr * SC_NUM_R + r + SC_SIZE] = v;
    }
    __syncthreads();
    const float *hv = h_t + (size_t)ch * SC_SIZE * ldh + jc;
    float best = __int_as_float(0x7fc00000);  // NaN: MATLAB min ignores NaN (processSC.m:31)
    for (int s0 = 0; s0 < SC_NUM_S; s0 += SIMT_ST) {
      float ax[SIMT_ST], ay[SIMT_ST];
#pragma unroll
      for (int i = 0; i < SIMT_ST; i++) ax[i] = ay[i] = 0.0f;
      const float *xb = x2 + SC_NUM_R * s0, *yb = y2 + SC_NUM_R * s0;
#pragma unroll 4
      for (int k = 0; k < SC_SIZE; k++) {
        const float h = hv[(size_t)k * ldh];
#pragma unroll
        for (int i = 0; i < SIMT_ST; i++) {
          ax[i] = fmaf(xb[k + SC_NUM_R * i], h, ax[i]);
          ay[i] = fmaf(yb[k + SC_NUM_R * i], h, ay[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < SIMT_ST; i++) {
        best = fminf(best, (1.0f - ax[i]) * 0.5f);  // processSC.m:30
        best = fminf(best, (1.0f - ay[i]) * 0.5f);
      }
    }
    float *out = ch == 0 ? d_p : d_i;
    if (j < n && out) out[(size_t)qi * ldd + j] = best;
  }
}

}  // namespace

cudaError_t launch_sc_prep_simt(const double *hist, int rows, float *out_t, int ld, cudaStream_t st,
                                int64_t *launches) {
  if (rows <= 0) return cudaSuccess;
  sc_prep_simt_kernel<<<rows, PREP_THREADS, 0, st>>>(hist, rows, out_t, ld);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_sc_match_simt(const float *q_t, int m, int ldq, const float *h_t, int n, int ldh,
                                 float *d_p, float *d_i, int ldd, cudaStream_t st, int64_t *launches) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  // gridDim.y is limited to 65535 queries per launch
  for (int q0 = 0; q0 < m; q0 += 65535) {
    int mq = m - q0 < 65535 ? m - q0 : 65535;
    dim3 grid((n + SIMT_THREADS - 1) / SIMT_THREADS, mq);
    sc_match_simt_kernel<<<grid, SIMT_THREADS, 0, st>>>(q_t + q0, mq, ldq, h_t, n, ldh,
                                                        d_p ? d_p + (size_t)q0 * ldd : nullptr,
                                                        d_i ? d_i + (size_t)q0 * ldd : nullptr, ldd);
    if (launches) ++*launches;
  }
  return cudaGetLastError();
}

}  // namespace sodso
