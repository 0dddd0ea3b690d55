// DELIGHT: DELIGHT::getSignature (DELIGHT.cpp:6-24) for a batch of scans, the chi-square matcher
// processDELIGHT.m:1-38, and the single-matrix decision of run_test.m:47-57.
//
// Generation: one CTA per scan.  Moments pass (one-pass sums about the first point, as in sc_generate.cu) -> 3x3
// eigen-solve on one thread -> second pass from L2: rotate (fp64), cast to float like the reference, octant + inside /
// outside RADIUS -> one of 16 histograms, bin int(intensity); shared-memory u32 histogram, 16 x 256 doubles out.
//
// Matching is not a GEMM: per pair min over 4 row permutations of  mean over {a + b > 0} of 2 (a - b)^2 / (a + b).
// A CTA owns a 16 x 16 tile of pairs (one pair per thread); per histogram row r it stages the 16 query rows r and,
// for each of the 4 permutations, the 16 DB rows Mut[k][r] in shared memory as fp32 counts (exact below 2^24) and
// accumulates the terms in fp32 per row (256 terms), the row sums in fp64; the number of contributing bins comes from
// the popcount of the OR of 256-bit non-zero masks.  HBM traffic is the signatures (16 KB
// each, L2 resident across tiles); the kernel is bound by the ~16 k divide-accumulate terms per pair.
// Tried (tools/experiments/delight_match_sparse_masks.patch, parity green): only the bins where both counts are non-zero
// need the reciprocal (one in seven at 37 % density; the others contribute 2 x the non-zero count, in closed form from row
// sums), found from the AND of 256-bit non-zero masks.  38 ms instead of 32 for 2000 x 2000: the per-lane bit loops
// diverge and their shared-memory reads hit random banks, where the dense loop broadcasts.
// Compiled with -fmad=false (generation restates fp64 arithmetic operation by operation, see pca.cuh).
#include <climits>

#include "../../include/sodso_pr.h"
#include "pca.cuh"

namespace sodso {
namespace {

constexpr int DL_BINS = 256;       // DELIGHT.h:10
constexpr int DL_ROWS = 16;
constexpr int DL_SIZE = DL_ROWS * DL_BINS;
constexpr double DL_RADIUS = 10.0;  // DELIGHT.h:9
constexpr int DL_THREADS = 256;

struct DlSmem {
  double scratch[9 * 32];
  double bc[16];
  unsigned hist[DL_SIZE];
};

__global__ void __launch_bounds__(DL_THREADS)
delight_generate_kernel(const double *__restrict__ xyz, const float *__restrict__ inten,
                        const int64_t *__restrict__ off, int nscan, double *__restrict__ out) {
  __shared__ DlSmem S;
  for (int scan = blockIdx.x; scan < nscan; scan += gridDim.x) {
    const int64_t p0 = off[scan];
    const int n = (int)(off[scan + 1] - p0);
    const double *g = xyz + 3 * p0;
    const float *gi = inten + p0;
    // ---- moments about the first point: mean and scatter matrix (pts_align.h:10-30)
    double s9[9];
#pragma unroll
    for (int k = 0; k < 9; k++) s9[k] = 0.0;
    double ox = 0.0, oy = 0.0, oz = 0.0;
    if (n > 0) {
      ox = g[0];
      oy = g[1];
      oz = g[2];
    }
#pragma unroll 2
    for (int i = threadIdx.x; i < n; i += DL_THREADS) {
      const double x = g[3 * (size_t)i + 0] - ox, y = g[3 * (size_t)i + 1] - oy, z = g[3 * (size_t)i + 2] - oz;
      s9[0] += x;
      s9[1] += y;
      s9[2] += z;
      s9[3] = fma(x, x, s9[3]);
      s9[4] = fma(x, y, s9[4]);
      s9[5] = fma(x, z, s9[5]);
      s9[6] = fma(y, y, s9[6]);
      s9[7] = fma(y, z, s9[7]);
      s9[8] = fma(z, z, s9[8]);
    }
    for (int b = threadIdx.x; b < DL_SIZE; b += DL_THREADS) S.hist[b] = 0u;
    __syncthreads();
    block_sum<9>(s9, S.scratch);
    if (threadIdx.x == 0 && n > 0) {
      const double dn = (double)n;
      const double mx = s9[0] / dn, my = s9[1] / dn, mz = s9[2] / dn;
      double c6[6] = {s9[3] - s9[0] * mx, s9[4] - s9[0] * my, s9[5] - s9[0] * mz,
                      s9[6] - s9[1] * my, s9[7] - s9[1] * mz, s9[8] - s9[2] * mz};
      S.bc[0] = ox + mx;
      S.bc[1] = oy + my;
      S.bc[2] = oz + mz;
      sym_eig3_fast(c6, S.bc);   // pts_align.h:31-34
    }
    __syncthreads();
    // ---- DELIGHT.cpp:14-23
    for (int i = threadIdx.x; i < n; i += DL_THREADS) {
      double ax, ay, az;
      pca_rotate(S.bc, g[3 * (size_t)i + 0], g[3 * (size_t)i + 1], g[3 * (size_t)i + 2], ax, ay, az);
      const float x = (float)ax, y = (float)ay, z = (float)az;
      const float d = (float)sqrt((ax * ax + ay * ay) + az * az);
      const float clr = gi[i];
      const int h = 8 * ((double)d > DL_RADIUS) + 4 * (z > 0.0f) + 2 * (y > 0.0f) + 1 * (x > 0.0f);
      if (!(clr > -1.0f && clr < 256.0f)) continue;   // out of the 256 columns in the reference: dropped
      atomicAdd(&S.hist[h * DL_BINS + (int)clr], 1u);
    }
    __syncthreads();
    double *o = out + (size_t)scan * DL_SIZE;
    for (int b = threadIdx.x; b < DL_SIZE; b += DL_THREADS) o[b] = (double)S.hist[b];
    __syncthreads();
  }
}

// fp64 signature rows -> fp32 counts [scan][16][256] and their non-zero masks [scan][16][8]; one warp per (scan, row)
__global__ void __launch_bounds__(256)
delight_prep_kernel(const double *__restrict__ hist, size_t rows, float *__restrict__ out, unsigned *__restrict__ mask) {
  const int lane = threadIdx.x & 31;
  for (size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (size_t)gridDim.x * 8) {
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const float v = (float)hist[row * 256 + 32 * w + lane];
      out[row * 256 + 32 * w + lane] = v;
      const unsigned bits = __ballot_sync(0xffffffffu, v > 0.0f);
      if (lane == 0) mask[row * 8 + w] = bits;
    }
  }
}

__constant__ int c_mut[4][16] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},           // processDELIGHT.m:2-5
                                 {5, 4, 7, 6, 1, 0, 3, 2, 13, 12, 15, 14, 9, 8, 11, 10},
                                 {6, 7, 4, 5, 2, 3, 0, 1, 14, 15, 12, 13, 10, 11, 8, 9},
                                 {3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12}};

constexpr int DT = 16;              // pairs per tile edge
constexpr int DPAD = DL_BINS + 1;   // row pitch in shared memory (bank-conflict free for the per-thread DB rows)

constexpr int DW = DL_BINS / 32;    // mask words per row

__global__ void __launch_bounds__(DT * DT)
delight_match_kernel(const float *__restrict__ h1, const unsigned *__restrict__ m1, int m, const float *__restrict__ h2,
                     const unsigned *__restrict__ m2, int n, double *__restrict__ dist) {
  extern __shared__ float sm[];
  float *sa = sm;                      // [DT][DPAD]     query rows r
  float *sb = sm + DT * DPAD;          // [4][DT][DPAD]  DB rows Mut[k][r]
  unsigned *ma = reinterpret_cast<unsigned *>(sm + 5 * DT * DPAD);   // [DT][DW]     non-zero masks of the query rows
  unsigned *mb = ma + DT * DW;                                       // [4][DT][DW]  ... of the DB rows
  const int tx = threadIdx.x & (DT - 1), ty = threadIdx.x / DT;
  const int q0 = blockIdx.y * DT, j0 = blockIdx.x * DT;
  double ts[4] = {0.0, 0.0, 0.0, 0.0};
  int tc[4] = {0, 0, 0, 0};
  for (int r = 0; r < DL_ROWS; r++) {
    __syncthreads();
    for (int e = threadIdx.x; e < DT * DL_BINS; e += DT * DT) {
      const int s = e >> 8, c = e & 255;
      sa[s * DPAD + c] = q0 + s < m ? h1[((size_t)(q0 + s) * DL_ROWS + r) * DL_BINS + c] : 0.0f;
#pragma unroll
      for (int k = 0; k < 4; k++)
        sb[(k * DT + s) * DPAD + c] = j0 + s < n ? h2[((size_t)(j0 + s) * DL_ROWS + c_mut[k][r]) * DL_BINS + c] : 0.0f;
    }
    if (threadIdx.x < DT * DW) {
      const int s = threadIdx.x / DW, w = threadIdx.x % DW;
      ma[s * DW + w] = q0 + s < m ? m1[((size_t)(q0 + s) * DL_ROWS + r) * DW + w] : 0u;
#pragma unroll
      for (int k = 0; k < 4; k++)
        mb[(k * DT + s) * DW + w] = j0 + s < n ? m2[((size_t)(j0 + s) * DL_ROWS + c_mut[k][r]) * DW + w] : 0u;
    }
    __syncthreads();
    // processDELIGHT.m:25: the number of bins with a + b > 0 (counts are >= 0) from the masks, not term by term
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int w = 0; w < DW; w++) tc[k] += __popc(ma[ty * DW + w] | mb[(k * DT + tx) * DW + w]);
    float rs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
    for (int c = 0; c < DL_BINS; c++) {
      const float a = sa[ty * DPAD + c];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float b = sb[(k * DT + tx) * DPAD + c];
        // processDELIGHT.m:25-29, branch-free: a + b == 0 means a == b == 0 and the term vanishes; the factor 2 of
        // :26 is applied to the row sum (exact)
        const float ab = a + b, df = a - b;
        rs[k] = __fmaf_rn(df * df, __fdividef(1.0f, fmaxf(ab, 1e-30f)), rs[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) ts[k] += (double)(2.0f * rs[k]);
  }
  const int qi = q0 + ty, dj = j0 + tx;
  if (qi < m && dj < n) {
    double best = INFINITY;                                // :16
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double v = ts[k] / (double)tc[k];              // :31 (0/0 -> NaN, never the minimum)
      if (best > v) best = v;                              // :32-34
    }
    dist[(size_t)qi * n + dj] = best;
  }
}

// run_test.m:47-57 for one distance matrix: mask, first minimum, NaN skipped
__global__ void __launch_bounds__(256)
top1_single_kernel(const double *__restrict__ d, int n, int mask_width, int32_t *__restrict__ idx, double *__restrict__ score) {
  __shared__ double sv[256];
  __shared__ int si[256];
  const int row = blockIdx.x;
  double bv = 0.0;
  int bi = -1;
  for (int j = threadIdx.x; j < n; j += 256) {
    int dd = row - j;
    if (dd < 0) dd = -dd;
    const double v = dd < mask_width ? INFINITY : d[(size_t)row * n + j];
    if (v != v) continue;
    if (bi < 0 || v < bv) {
      bv = v;
      bi = j;
    }
  }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double v2 = sv[threadIdx.x + o];
      const int i2 = si[threadIdx.x + o];
      const int i1 = si[threadIdx.x];
      if (i2 >= 0 && (i1 < 0 || v2 < sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < i1))) {
        sv[threadIdx.x] = v2;
        si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    idx[row] = si[0] < 0 ? 0 : si[0];
    score[row] = si[0] < 0 ? NAN : sv[0];
  }
}

}  // namespace

cudaError_t launch_delight_generate(const double *xyz, const float *inten, const int64_t *off, int nscan, double *hist,
                                    int num_sms, cudaStream_t st, int64_t *launches) {
  if (nscan <= 0) return cudaSuccess;
  const int grid = nscan < 8 * num_sms ? nscan : 8 * num_sms;
  delight_generate_kernel<<<grid, DL_THREADS, 0, st>>>(xyz, inten, off, nscan, hist);
  if (launches) ++*launches;
  return cudaGetLastError();
}

static size_t dl_align(size_t x) { return (x + 255) & ~(size_t)255; }
size_t delight_match_workspace_bytes(int m, int n) {
  const size_t rows = ((size_t)m + (size_t)n) * DL_ROWS;
  return dl_align(rows * DL_BINS * sizeof(float)) + dl_align(rows * DW * sizeof(unsigned)) + 256;
}

cudaError_t launch_delight_match(const double *hist1, int m, const double *hist2, int n, double *dist, void *workspace,
                                 cudaStream_t st, int64_t *launches) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  const size_t rows1 = (size_t)m * DL_ROWS, rows2 = (size_t)n * DL_ROWS, rows = rows1 + rows2;
  float *f1 = reinterpret_cast<float *>(workspace), *f2 = f1 + rows1 * DL_BINS;
  unsigned *k1 = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(workspace) + dl_align(rows * DL_BINS * sizeof(float)));
  unsigned *k2 = k1 + rows1 * DW;
  delight_prep_kernel<<<1024, 256, 0, st>>>(hist1, rows1, f1, k1);
  delight_prep_kernel<<<1024, 256, 0, st>>>(hist2, rows2, f2, k2);
  const size_t smem = (size_t)5 * DT * DPAD * sizeof(float) + (size_t)5 * DT * DW * sizeof(unsigned);
  cudaError_t e = cudaFuncSetAttribute(delight_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  for (int qb = 0; qb < m; qb += 65535 * DT) {
    const int mq = m - qb < 65535 * DT ? m - qb : 65535 * DT;
    dim3 grid((n + DT - 1) / DT, (mq + DT - 1) / DT);
    delight_match_kernel<<<grid, DT * DT, smem, st>>>(f1 + (size_t)qb * DL_SIZE, k1 + (size_t)qb * DL_ROWS * DW, mq, f2, k2, n,
                                                      dist + (size_t)qb * n);
  }
  if (launches) *launches += 3;
  return cudaGetLastError();
}

cudaError_t launch_top1_single(const double *d, int m, int n, int mask_width, int32_t *idx, double *score, cudaStream_t st,
                               int64_t *launches) {
  if (m <= 0) return cudaSuccess;
  top1_single_kernel<<<m, 256, 0, st>>>(d, n, mask_width, idx, score);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
