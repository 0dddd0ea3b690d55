// M2DP all-pairs match: processM2DP.m:12-22.  diff_full = (1 - hist1 * hist2') / 2 on the 4m x 4n
// variant rows (NO normalisation, processM2DP.m:15), then the minimum of every 4 x 4 block.
//
// This fp32 CUDA-core kernel is the on-GPU cross-check (SODSO_ALGO_SIMT) of the tensor-core matcher in
// m2dp_match_tc.cu, which is the product path.  Tiled contraction: a 16 x 16 thread block owns a 64 x 64 tile of variant rows
// (16 queries x 16 DB entries), every thread accumulates the 4 x 4 block of ONE (query, DB) pair
// in registers over K = 192 and reduces it with the NaN-ignoring min (MATLAB min).  12 288 FLOP per
// pair: ~0.3 TFLOP for 5k x 5k, three orders of magnitude below the Scan Context matcher, and the
// signatures are unit vectors, so fp32 accumulation is ~1e-7 from the fp64 reference (bar 1e-5).
#include "../../include/sodso_pr.h"
#include "common.cuh"

namespace sodso {
namespace {

constexpr int MT = 16;   // pairs per block edge
constexpr int KC = 32;   // K chunk

__global__ void __launch_bounds__(MT * MT)
m2dp_match_kernel(const double *__restrict__ h1, int m, const double *__restrict__ h2, int n,
                  float *__restrict__ d_p, float *__restrict__ d_i, int ldd) {
  __shared__ float sa[KC][4 * MT + 1], sb[KC][4 * MT + 1];
  const int tx = threadIdx.x & (MT - 1), ty = threadIdx.x / MT;
  const int q0 = blockIdx.y * MT, j0 = blockIdx.x * MT;
  for (int ch = 0; ch < 2; ch++) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = 0.0f;
    for (int k0 = 0; k0 < M2DP_SIG; k0 += KC) {
      __syncthreads();
      // 64 rows x KC columns of each operand; consecutive threads read consecutive k (coalesced fp64)
      for (int e = threadIdx.x; e < 4 * MT * KC; e += MT * MT) {
        const int r = e / KC, k = e - r * KC;
        const int ra = q0 * 4 + r, rb = j0 * 4 + r;
        sa[k][r] = ra < 4 * m ? (float)h1[(size_t)ra * 2 * M2DP_SIG + ch * M2DP_SIG + k0 + k] : 0.0f;
        sb[k][r] = rb < 4 * n ? (float)h2[(size_t)rb * 2 * M2DP_SIG + ch * M2DP_SIG + k0 + k] : 0.0f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < KC; k++) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          a[i] = sa[k][ty * 4 + i];
          b[i] = sb[k][tx * 4 + i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    float best = __int_as_float(0x7fc00000);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) best = fminf(best, (1.0f - acc[i][j]) * 0.5f);  // processM2DP.m:15,19
    const int qi = q0 + ty, dj = j0 + tx;
    float *out = ch == 0 ? d_p : d_i;
    if (qi < m && dj < n && out) out[(size_t)qi * ldd + dj] = best;
  }
}

}  // namespace

size_t m2dp_match_workspace_bytes(int, int) { return 256; }

cudaError_t launch_m2dp_match(const double *hist1, int m, const double *hist2, int n, float *d_p, float *d_i,
                              int ldd, void *, cudaStream_t st, int64_t *launches) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  for (int qb = 0; qb < m; qb += 65535 * MT) {
    const int mq = m - qb < 65535 * MT ? m - qb : 65535 * MT;
    dim3 grid((n + MT - 1) / MT, (mq + MT - 1) / MT);
    m2dp_match_kernel<<<grid, MT * MT, 0, st>>>(hist1 + (size_t)qb * 4 * 2 * M2DP_SIG, mq, hist2, n,
                                                d_p ? d_p + (size_t)qb * ldd : nullptr,
                                                d_i ? d_i + (size_t)qb * ldd : nullptr, ldd);
    if (launches) ++*launches;
  }
  return cudaGetLastError();
}

}  // namespace sodso
