// M2DP signature generation: test_m2dp.cpp:41-67 (PCA once, 4 sign variants) around
// M2DP::getSignature (M2DP.cpp:38-109), one persistent CTA per SM (1024 threads), the whole batch in
// one launch.
//
// Per scan: moments pass (HBM read) -> mean + 3x3 eigenproblem (pts_align.h) -> staging pass (L2 read):
// the PCA-aligned points go to shared memory as fp32.  Per variant: 64 planes x n points are projected
// and binned by (rho, theta) into a 64 x 128 count histogram (u32) and intensity-sum histogram (int32 when
// the sums are provably exact, fp64 otherwise) with shared-memory atomics; the sums are binarised against
// the float average intensity; the signature is the dominant left/right singular vector pair of each
// 64 x 128 matrix.
//
// Binning: the bin of a (point, plane) pair is proposed in fp32 (6 FMAs, fast_turns, rsqrt) and accepted
// only if both polar coordinates are further from a bin edge than 3x the fp32 error bound (which grows
// like 1/rho for the angle); everything else takes the reference's fp64 expression (M2DP.cpp:56-63) on the
// fp64 point re-read from L2.  The degenerate plane p=2,q=0, whose projection vectors are exactly zero
// (SURVEY F8), puts every point into one of two bins depending on the signs of the zeros; it is handled
// with one warp-aggregated update per 32 points.
//
// SVD: only the dominant pair is needed (M2DP.cpp:96-103).  G = A A^T (64 x 64, exact in integers) is
// formed by the whole CTA, then four warps per matrix run the power iteration u <- G u / |G u| in fp64 from
// the all-ones vector until the update falls below 1e-14 (both matrices concurrently), and
// v = A^T u / sigma.  Both matrices are entrywise non-negative, so the iteration converges to the Perron
// pair, which fixes the sign the same way as the oracle (sum(u) >= 0; Eigen's own sign is unobservable,
// SURVEY §8c).
//
// The projection table (xProj / yProj, M2DP.cpp:4-34) is computed on the host with float cosf/sinf exactly
// like the reference constructor and passed in constant memory.
// Compiled with -fmad=false (see pca.cuh); fused multiply-adds are written explicitly where wanted.
#include <climits>
#include <cmath>

#include "../../include/sodso_pr.h"
#include "pca.cuh"

namespace sodso {
namespace {

constexpr int M2_THREADS = 1024;
constexpr int M2_CAP = 4096;            // aligned points of a scan staged in shared memory (fp32)
constexpr int HB = M2DP_PQ * M2DP_SR;   // 8192 histogram bins
constexpr float M2_GUARD_R = 3e-5f;      // ring coordinate guard (see the binning loop)
constexpr int SVD_WARPS = 4;            // warps per matrix in the power iteration

__constant__ double c_xproj[3 * M2DP_PQ];
__constant__ double c_yproj[3 * M2DP_PQ];
__constant__ float c_xproj32[3 * M2DP_PQ];
__constant__ float c_yproj32[3 * M2DP_PQ];

struct M2Smem {
  double hsum[HB];             // fp64 intensity sums (int32 in exact mode) -> binarised matrix as u32 in the
                               // first half, G of the binarised matrix in the second half
  double G0[M2DP_PQ * M2DP_PQ];  // Gram matrix of the count matrix
  double T[M2DP_PQ * M2DP_PQ];   // squaring workspace
  double scratch[11 * 32];
  double uvec[2][M2DP_PQ], yv[2][M2DP_SR], red[2][SVD_WARPS], red2[2][SVD_WARPS], sig[2];
  double bc[16];
  float4 tab[2 * M2DP_PQ];      // fp32 projection table: (x0, x1, x2, y0), (y1, y2, 0, 0) per plane
  float ax[M2_CAP], ay[M2_CAP], az[M2_CAP];
  unsigned hcnt[HB];
  int ibc[4];
};
static_assert(sizeof(M2Smem) <= 227 * 1024, "shared memory");

struct ScanRef {
  const double *g;
  const float *gi;
  const double *bc;   // mean + eigenvectors (identity for pre-aligned input)
  int n, nst;
  int identity;       // class contract (M2DP.h:18-20): the input is used as it is (signed zeros included)
  double dx, dy, dz;  // sign variant
  double S_res_inv, R_res_inv;
};

// aligned + sign-flipped fp64 point i (pts_align.h:37-45, test_m2dp.cpp:49-53)
__device__ __forceinline__ void aligned_point64(const ScanRef &R, int i, double &px, double &py, double &pz) {
  double ax, ay, az;
  if (R.identity) {
    ax = R.g[3 * (size_t)i + 0];
    ay = R.g[3 * (size_t)i + 1];
    az = R.g[3 * (size_t)i + 2];
    px = ax;   // (dx = dy = dz = 1; no multiplication either)
    py = ay;
    pz = az;
    return;
  }
  pca_rotate(R.bc, R.g[3 * (size_t)i + 0], R.g[3 * (size_t)i + 1], R.g[3 * (size_t)i + 2], ax, ay, az);
  px = R.dx * ax;
  py = R.dy * ay;
  pz = R.dz * az;
}

// M2DP.cpp:56-68 in fp64, exactly the reference's expression: flat histogram index or -1
__device__ __noinline__ int m2dp_bin_exact(const ScanRef &R, int i, int pq) {
  const double PI = 3.14159265358979323846;
  double px, py, pz;
  aligned_point64(R, i, px, py, pz);
  const double xp = (c_xproj[3 * pq] * px + c_xproj[3 * pq + 1] * py) + c_xproj[3 * pq + 2] * pz;
  const double yp = (c_yproj[3 * pq] * px + c_yproj[3 * pq + 1] * py) + c_yproj[3 * pq + 2] * pz;
  const double ang = (atan2(yp, xp) + PI) * R.S_res_inv;
  const double rad = sqrt(xp * xp + yp * yp) * R.R_res_inv;
  if (!(rad < (double)M2DP_SR) || !(ang < 32.0)) return -1;
  const int si = (int)floor(ang), ri = (int)floor(rad);
  const int sr = ri * M2DP_NUM_S + si;  // M2DP.cpp:63
  return sr < M2DP_SR ? pq * M2DP_SR + sr : -1;  // M2DP.cpp:66 (si == 16 aliases into ring ri+1)
}

// Y = X X for a symmetric 64 x 64 matrix (all threads): Y[i][j] = sum_k X[k][i] X[k][j]
__device__ __forceinline__ void sym_square64(const double *X, double *Y) {
  const int j = threadIdx.x & 63, i0 = (threadIdx.x >> 6) * 4;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
  for (int k = 0; k < M2DP_PQ; k++) {
    const double xj = X[k * M2DP_PQ + j];
    const double2 p = reinterpret_cast<const double2 *>(X + k * M2DP_PQ + i0)[0];
    const double2 q = reinterpret_cast<const double2 *>(X + k * M2DP_PQ + i0)[1];
    a0 = fma(p.x, xj, a0);
    a1 = fma(p.y, xj, a1);
    a2 = fma(q.x, xj, a2);
    a3 = fma(q.y, xj, a3);
  }
  Y[(i0 + 0) * M2DP_PQ + j] = a0;
  Y[(i0 + 1) * M2DP_PQ + j] = a1;
  Y[(i0 + 2) * M2DP_PQ + j] = a2;
  Y[(i0 + 3) * M2DP_PQ + j] = a3;
}

// dominant singular pairs of the two 64 x 128 matrices A0 (counts) and A1 (binarised), both u32 in shared
// memory.  G0 / G1 / T: 64 x 64 fp64 workspaces.  Writes [u (64), v (128)] of each to out0 / out1.
//   G = A A^T exactly (64-bit integers), scaled by a power of two to trace ~ 1;  G^16 by four squarings;
//   power iteration with G^16 from the all-ones vector (Perron pair: entrywise non-negative);
//   sigma = |A^T u|, v = A^T u / sigma.
__device__ void dominant_pairs(const unsigned *A0, const unsigned *A1, double *G0, double *G1, double *T, M2Smem &S,
                               double *out0, double *out1) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- Gram matrices: one (i, j <= i) entry per warp step, lanes split k (conflict-free LDS.128)
  // warp w takes rows w and 63 - w of both matrices: 65 entries each, no index decoding
  for (int m = 0; m < 2; m++) {
    const unsigned *A = m ? A1 : A0;
    double *G = m ? G1 : G0;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      const int i = h ? M2DP_PQ - 1 - warp : warp;
      const uint4 a = reinterpret_cast<const uint4 *>(A + i * M2DP_SR)[lane];
#pragma unroll 2
      for (int j = 0; j <= i; j++) {
        const uint4 b = reinterpret_cast<const uint4 *>(A + j * M2DP_SR)[lane];
        unsigned long long acc = (unsigned long long)a.x * b.x + (unsigned long long)a.y * b.y +
                                 (unsigned long long)a.z * b.z + (unsigned long long)a.w * b.w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
          G[i * M2DP_PQ + j] = (double)acc;
          G[j * M2DP_PQ + i] = (double)acc;
        }
      }
    }
  }
  __syncthreads();
  // ---- scale to trace in [1, 2) (exact, power of two), so that G^16 neither overflows nor underflows
  if (warp < 2) {
    double *G = warp ? G1 : G0;
    double tr = G[lane * (M2DP_PQ + 1)] + G[(lane + 32) * (M2DP_PQ + 1)];
    tr = warp_sum(tr);
    tr = __shfl_sync(0xffffffffu, tr, 0);
    if (lane == 0) S.sig[warp] = tr > 0.0 ? scalbn(1.0, -ilogb(tr)) : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < 2 * M2DP_PQ * M2DP_PQ; e += M2_THREADS) {
    double *G = (e >> 12) ? G1 : G0;
    G[e & 4095] *= S.sig[e >> 12];
  }
  __syncthreads();
  for (int m = 0; m < 2; m++) {
    double *G = m ? G1 : G0;
    sym_square64(G, T);   // G^2
    __syncthreads();
    sym_square64(T, G);   // G^4
    __syncthreads();
    sym_square64(G, T);   // G^8
    __syncthreads();
    sym_square64(T, G);   // G^16
    __syncthreads();
  }
  if (tid < 2 * M2DP_PQ) S.uvec[tid >> 6][tid & 63] = 0.125;  // all-ones / |.| (Perron start)
  __syncthreads();
  // ---- power iteration with G^16: SVD_WARPS warps per matrix, 2 threads per row, named barrier per matrix
  if (warp < 2 * SVD_WARPS) {
    const int m = warp / SVD_WARPS, wl = warp % SVD_WARPS;
    const double *G = m ? G1 : G0;
    const int t = wl * 32 + lane;          // 0..127
    const int row = t >> 1, half = t & 1;
    const int bar_id = 1 + m, bar_n = SVD_WARPS * 32;
    for (int iter = 0; iter < 2000; iter++) {
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
      for (int k = 0; k < 32; k += 2) {
        const int c = half * 32 + k;
        acc0 = fma(G[c * M2DP_PQ + row], S.uvec[m][c], acc0);   // G symmetric: column access is conflict-free
        acc1 = fma(G[(c + 1) * M2DP_PQ + row], S.uvec[m][c + 1], acc1);
      }
      double acc = acc0 + acc1;
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      double sq = half == 0 ? acc * acc : 0.0;
      sq = warp_sum(sq);
      if (lane == 0) S.red[m][wl] = sq;
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");   // all reads of uvec done
      double nn = 0.0;
#pragma unroll
      for (int w = 0; w < SVD_WARPS; w++) nn += S.red[m][w];
      if (nn == 0.0) break;  // zero matrix (uniform over the group)
      const double inv = 1.0 / sqrt(nn);
      double d2 = 0.0;
      if (half == 0) {
        const double nv = acc * inv;
        const double d = nv - S.uvec[m][row];
        d2 = d * d;
        S.uvec[m][row] = nv;
      }
      d2 = warp_sum(d2);
      if (lane == 0) S.red2[m][wl] = d2;
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");   // uvec / red2 complete
      double dd = 0.0;
#pragma unroll
      for (int w = 0; w < SVD_WARPS; w++) dd += S.red2[m][w];
      if (dd < 1e-29) break;
    }
  }
  __syncthreads();
  // ---- y = A^T u (128 per matrix), sigma = |y|
  if (tid < 2 * M2DP_SR) {
    const int m = tid >> 7, col = tid & 127;
    const unsigned *A = m ? A1 : A0;
    double acc = 0.0;
#pragma unroll 8
    for (int r = 0; r < M2DP_PQ; r++) acc = fma((double)A[r * M2DP_SR + col], S.uvec[m][r], acc);
    S.yv[m][col] = acc;
    const double sq = warp_sum(acc * acc);
    if (lane == 0) S.red[m][warp & 3] = sq;
  }
  __syncthreads();
  // ---- outputs: u, v = y / sigma  (zero matrix: u = e_0, v = 0, the oracle's convention)
  for (int e = tid; e < 2 * M2DP_SIG; e += M2_THREADS) {
    const int m = e / M2DP_SIG, k = e % M2DP_SIG;
    double *out = m ? out1 : out0;
    const double sigma = sqrt((S.red[m][0] + S.red[m][1]) + (S.red[m][2] + S.red[m][3]));
    double v;
    if (k < M2DP_PQ)
      v = sigma == 0.0 ? (k == 0 ? 1.0 : 0.0) : S.uvec[m][k];
    else
      v = sigma == 0.0 ? 0.0 : S.yv[m][k - M2DP_PQ] / sigma;
    out[k] = v;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(M2_THREADS, 1)
m2dp_generate_kernel(const double *__restrict__ xyz, const float *__restrict__ inten,
                     const int64_t *__restrict__ off, int nscan, double S_res_inv, double R_res_inv,
                     int variants, double *__restrict__ hist, unsigned long long degen_mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  M2Smem &S = *reinterpret_cast<M2Smem *>(smem_raw);
  const int lane = threadIdx.x & 31;
  const unsigned dm_lo = (unsigned)degen_mask, dm_hi = (unsigned)(degen_mask >> 32);
  if (threadIdx.x < M2DP_PQ) {
    const int k = threadIdx.x;
    S.tab[2 * k] = make_float4(c_xproj32[3 * k], c_xproj32[3 * k + 1], c_xproj32[3 * k + 2], c_yproj32[3 * k]);
    S.tab[2 * k + 1] = make_float4(c_yproj32[3 * k + 1], c_yproj32[3 * k + 2], 0.0f, 0.0f);
  }
  // the two bins of a plane with zero projection vectors: M2DP.cpp:59-63 evaluated at (+0, +0) and (-0, -0)
  int degen_sr_pos, degen_sr_neg;
  {
    const double PI = 3.14159265358979323846;
    const double ap = (atan2(0.0, 0.0) + PI) * S_res_inv, an = (atan2(-0.0, -0.0) + PI) * S_res_inv;
    const int sp = (int)floor(ap), sn = (int)floor(an);   // ring 0
    degen_sr_pos = sp < M2DP_SR ? sp : -1;
    degen_sr_neg = sn < M2DP_SR ? sn : -1;
  }
  const float S_f = (float)M2DP_NUM_S, R_f = (float)R_res_inv;
  int *h_isum = reinterpret_cast<int *>(S.hsum);
  unsigned *bin_mat = reinterpret_cast<unsigned *>(S.hsum);            // first 32 KB of hsum
  double *G1 = S.hsum + HB / 2;                                       // second 32 KB of hsum

  for (int scan = blockIdx.x; scan < nscan; scan += gridDim.x) {
    const int64_t p0 = off[scan];
    const int n = (int)(off[scan + 1] - p0);
    const double *g = xyz + 3 * p0;
    const float *gi = inten + p0;
    const int nst = n < M2_CAP ? n : M2_CAP;

    // ---- pass 1 (HBM): moments about the first point (mean, scatter matrix) and intensity sums
    double s11[11];
#pragma unroll
    for (int k = 0; k < 11; k++) s11[k] = 0.0;
    int emin = 1 << 20;
    double ox = 0.0, oy = 0.0, oz = 0.0;
    if (n > 0) {
      ox = g[0];
      oy = g[1];
      oz = g[2];
    }
    if (threadIdx.x == 0) S.ibc[0] = 1 << 20;
#pragma unroll 2
    for (int i = threadIdx.x; i < n; i += M2_THREADS) {
      const float v = gi[i];
      const double x = g[3 * (size_t)i + 0] - ox, y = g[3 * (size_t)i + 1] - oy, z = g[3 * (size_t)i + 2] - oz;
      s11[0] += x;
      s11[1] += y;
      s11[2] += z;
      s11[3] = fma(x, x, s11[3]);
      s11[4] = fma(x, y, s11[4]);
      s11[5] = fma(x, z, s11[5]);
      s11[6] = fma(y, y, s11[6]);
      s11[7] = fma(y, z, s11[7]);
      s11[8] = fma(z, z, s11[8]);
      s11[9] += (double)v;
      s11[10] += fabs((double)v);
      const unsigned b = __float_as_uint(v);
      const int e = (int)((b >> 23) & 0xffu) - 151 + __ffs(b | 0x800000u);
      emin = (v != 0.0f && e < emin) ? e : emin;
    }
    emin = __reduce_min_sync(0xffffffffu, emin);
    __syncthreads();
    if (lane == 0) atomicMin(&S.ibc[0], emin);
    block_sum<11>(s11, S.scratch);
    emin = S.ibc[0];
    // every partial sum of the sequential float loop (M2DP.cpp:77-80) is exact -> the float sum is the exact sum
    const bool exact = emin == (1 << 20) ? s11[10] == 0.0 : s11[10] < ldexp(1.0, 24 + emin);
    if (threadIdx.x == 0) {
      if (variants && n > 0) {  // test_m2dp.cpp:44-45
        const double dn = (double)n;
        const double mx = s11[0] / dn, my = s11[1] / dn, mz = s11[2] / dn;
        double c6[6] = {s11[3] - s11[0] * mx, s11[4] - s11[0] * my, s11[5] - s11[0] * mz,
                        s11[6] - s11[1] * my, s11[7] - s11[1] * mz, s11[8] - s11[2] * mz};
        S.bc[0] = ox + mx;
        S.bc[1] = oy + my;
        S.bc[2] = oz + mz;
        sym_eig3_fast(c6, S.bc);
      } else {  // the class contract (M2DP.h:18-20): input already aligned -> identity transform
        for (int k = 0; k < 16; k++) S.bc[k] = (k == 3 || k == 7 || k == 11) ? 1.0 : 0.0;
      }
    } else if (threadIdx.x == 32 && !exact) {
      float a = 0.0f;
#pragma unroll 16
      for (int i = 0; i < n; i++) a += gi[i];
      S.ibc[2] = __float_as_int(a);
    }
    __syncthreads();
    const float ave = (exact ? (float)s11[9] : __int_as_float(S.ibc[2])) / (float)n;  // M2DP.cpp:81
    const float iscale = exact && emin != (1 << 20) ? (float)ldexp(1.0, -emin) : 0.0f;
    const double unscale = exact && emin != (1 << 20) ? ldexp(1.0, emin) : 0.0;

    // ---- pass 2 (L2): aligned points -> shared memory as fp32 (the proposal path only needs fp32)
    for (int i = threadIdx.x; i < nst; i += M2_THREADS) {
      double ax, ay, az;
      if (variants) {
        pca_rotate(S.bc, g[3 * (size_t)i + 0], g[3 * (size_t)i + 1], g[3 * (size_t)i + 2], ax, ay, az);
      } else {
        ax = g[3 * (size_t)i + 0];
        ay = g[3 * (size_t)i + 1];
        az = g[3 * (size_t)i + 2];
      }
      S.ax[i] = (float)ax;
      S.ay[i] = (float)ay;
      S.az[i] = (float)az;
    }

    const int nvar = variants ? 4 : 1;
    for (int var = 0; var < nvar; var++) {
      // test_m2dp.cpp:47-57: dx outer, dy inner, both in {-1, +1}
      ScanRef R;
      R.g = g;
      R.gi = gi;
      R.bc = S.bc;
      R.n = n;
      R.nst = nst;
      R.identity = variants ? 0 : 1;
      R.dx = variants ? ((var >> 1) ? 1.0 : -1.0) : 1.0;
      R.dy = variants ? ((var & 1) ? 1.0 : -1.0) : 1.0;
      R.dz = R.dx * R.dy;
      R.S_res_inv = S_res_inv;
      R.R_res_inv = R_res_inv;
      const float fdx = (float)R.dx, fdy = (float)R.dy, fdz = (float)R.dz;
      for (int b = threadIdx.x; b < HB; b += M2_THREADS) {
        S.hcnt[b] = 0u;
        S.hsum[b] = 0.0;
      }
      __syncthreads();
      // ---- M2DP.cpp:47-74
      for (int i0 = 0; i0 < n; i0 += M2_THREADS) {
        const int i = i0 + threadIdx.x;
        const bool act = i < n;
        float px = 0.0f, py = 0.0f, pz = 0.0f, it = 0.0f;
        if (act) {
          it = gi[i];
          if (i < nst) {
            px = fdx * S.ax[i];
            py = fdy * S.ay[i];
            pz = fdz * S.az[i];
          } else {
            double ax, ay, az;
            aligned_point64(R, i, ax, ay, az);
            px = (float)ax;
            py = (float)ay;
            pz = (float)az;
          }
        }
        const int iv = (int)(it * iscale);
        // the fp32 error bound below assumes coordinates below 128 m (the staging crops at 45 m)
        const float r2_lim = fmaxf(fabsf(px), fmaxf(fabsf(py), fabsf(pz))) < 128.0f ? 1e10f : -1.0f;
        // planes whose projection vectors are exactly zero (p=2, q=0): xp = yp = -0 if all three
        // coordinates are negative, +0 otherwise, and every point lands in one of two bins -> warp-aggregated
        for (unsigned long long dm = degen_mask; dm; dm &= dm - 1) {
          const int pq = __ffsll((long long)dm) - 1;
          bool neg;
          if (i < nst || !act) {
            neg = px < 0.0f && py < 0.0f && pz < 0.0f;
            // an fp32 coordinate that rounded to zero: decide on the fp64 point
            if (act && (px == 0.0f || py == 0.0f || pz == 0.0f)) {
              double ax, ay, az;
              aligned_point64(R, i, ax, ay, az);
              neg = signbit(ax) && signbit(ay) && signbit(az);
            }
          } else {
            double ax, ay, az;
            aligned_point64(R, i, ax, ay, az);
            neg = signbit(ax) && signbit(ay) && signbit(az);
          }
#pragma unroll
          for (int cls = 0; cls < 2; cls++) {
            const bool mine = act && (neg == (cls == 1));
            const unsigned ball = __ballot_sync(0xffffffffu, mine);
            const int sr = cls ? degen_sr_neg : degen_sr_pos;
            if (ball == 0u || sr < 0) continue;
            if (exact) {
              const int sum = __reduce_add_sync(0xffffffffu, mine ? iv : 0);
              if (lane == 0) {
                atomicAdd(&S.hcnt[pq * M2DP_SR + sr], (unsigned)__popc(ball));
                atomicAdd(&h_isum[2 * (pq * M2DP_SR + sr)], sum);
              }
            } else {
              // the reference adds the intensities point by point (M2DP.cpp:71); fp64 sums of floats in another
              // order differ only when the exact sum needs more than 53 bits
              const double sum = warp_sum(mine ? (double)it : 0.0);
              if (lane == 0) {
                atomicAdd(&S.hcnt[pq * M2DP_SR + sr], (unsigned)__popc(ball));
                atomicAdd(&S.hsum[pq * M2DP_SR + sr], sum);
              }
            }
          }
        }
#pragma unroll 2
        for (int pq = 0; pq < M2DP_PQ; pq++) {
          if (((pq < 32 ? dm_lo : dm_hi) >> (pq & 31)) & 1u) continue;
          const float4 ta = S.tab[2 * pq], tb = S.tab[2 * pq + 1];   // (x0, x1, x2, y0), (y1, y2, -, -)
          const float xp = __fmaf_rn(ta.z, pz, __fmaf_rn(ta.y, py, ta.x * px));
          const float yp = __fmaf_rn(tb.y, pz, __fmaf_rn(tb.x, py, ta.w * px));
          const float r2 = __fmaf_rn(xp, xp, yp * yp);
          const float rinv = rsqrtf(r2);
          const float tf = fast_turns(yp, xp) * S_f;
          const float rf = r2 * rinv * R_f;
          const float ft = floorf(tf), fr = floorf(rf);
          // distance to the nearest bin edge = 0.5 - |frac - 0.5|
          const float et = fabsf((tf - ft) - 0.5f), er = fabsf((rf - fr) - 0.5f);
          // fp32 error of the proposal: the projections are good to 2e-5 m (|p| < 128 m), i.e. 5e-5 / rho sectors
          // and 4e-6 rings, plus 4e-6 sectors from fast_turns and 3e-6 rings from rsqrt; accepted with a 3x margin
          const bool safe = et < 0.5f - 2e-5f - 1.5e-4f * rinv && er < 0.5f - M2_GUARD_R && tf < (float)M2DP_NUM_S &&
                            rf < (float)M2DP_NUM_R && r2 > 1e-2f && r2 < r2_lim;
          int idx = safe ? pq * M2DP_SR + __float2int_rz(__fmaf_rn(fr, (float)M2DP_NUM_S, ft)) : -1;
          if (__any_sync(0xffffffffu, act && !safe)) {   // rare: the reference's fp64 expression
            if (act && !safe) idx = m2dp_bin_exact(R, i, pq);
          }
          if (act && idx >= 0) {
            atomicAdd(&S.hcnt[idx], 1u);
            if (exact)
              atomicAdd(&h_isum[2 * idx], iv);  // exact integer multiple of 2^emin, |sum| < 2^24
            else
              atomicAdd(&S.hsum[idx], (double)it);
          }
        }
      }
      __syncthreads();
      // ---- M2DP.cpp:84-91: binarise (registers first: the u32 matrix overwrites the sums in place)
      unsigned bv[HB / M2_THREADS];
#pragma unroll
      for (int k = 0; k < HB / M2_THREADS; k++) {
        const int b = threadIdx.x + k * M2_THREADS;
        const unsigned c = S.hcnt[b];
        unsigned v = 0u;
        if (c) {
          const double sum = exact ? (double)h_isum[2 * b] * unscale : S.hsum[b];
          v = (sum / (double)c) > (double)ave ? 1u : 0u;
        }
        bv[k] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < HB / M2_THREADS; k++) bin_mat[threadIdx.x + k * M2_THREADS] = bv[k];
      __syncthreads();
      double *row = hist + ((size_t)scan * nvar + var) * 2 * M2DP_SIG;
      dominant_pairs(S.hcnt, bin_mat, S.G0, G1, S.T, S, row, row + M2DP_SIG);   // M2DP.cpp:94-108
    }
    __syncthreads();
  }
}

// M2DP::M2DP (M2DP.cpp:4-34) -- float azimuth / elevation and float cos/sin products, like the reference
void build_tables(double *xproj, double *yproj) {
  for (int p = 0; p < M2DP_NUM_P; p++) {
    float azm = -M_PI / 2.0 + (M_PI / M2DP_NUM_P) * p;
    for (int q = 0; q < M2DP_NUM_Q; q++) {
      float elv = (M_PI / 2.0 / M2DP_NUM_Q) * q;
      double n0 = std::cos(elv) * std::cos(azm);
      double n1 = std::cos(elv) * std::sin(azm);
      double n2 = std::sin(elv);
      double d = (1.0 * n0 + 0.0 * n1) + 0.0 * n2;
      double x0 = 1.0 - d * n0, x1 = 0.0 - d * n1, x2 = 0.0 - d * n2;
      double y0 = n1 * x2 - n2 * x1;
      double y1 = n2 * x0 - n0 * x2;
      double y2 = n0 * x1 - n1 * x0;
      int k = p * M2DP_NUM_Q + q;
      xproj[3 * k + 0] = x0;
      xproj[3 * k + 1] = x1;
      xproj[3 * k + 2] = x2;
      yproj[3 * k + 0] = y0;
      yproj[3 * k + 1] = y1;
      yproj[3 * k + 2] = y2;
    }
  }
}

}  // namespace

size_t m2dp_generate_workspace_bytes(int, bool) { return 256; }

cudaError_t launch_m2dp_generate(const double *xyz, const float *inten, const int64_t *off, int nscan,
                                 double max_rho, bool do_align_and_variants, double *hist, void *, size_t,
                                 int num_sms, cudaStream_t st, int64_t *launches) {
  if (nscan <= 0) return cudaSuccess;
  double xp[3 * M2DP_PQ], yp[3 * M2DP_PQ];
  build_tables(xp, yp);
  cudaError_t e = cudaMemcpyToSymbolAsync(c_xproj, xp, sizeof(xp), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(c_yproj, yp, sizeof(yp), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  float xp32[3 * M2DP_PQ], yp32[3 * M2DP_PQ];
  for (int k = 0; k < 3 * M2DP_PQ; k++) {
    xp32[k] = (float)xp[k];
    yp32[k] = (float)yp[k];
  }
  e = cudaMemcpyToSymbolAsync(c_xproj32, xp32, sizeof(xp32), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(c_yproj32, yp32, sizeof(yp32), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(m2dp_generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(M2Smem));
  if (e != cudaSuccess) return e;
  const double S_res_inv = M2DP_NUM_S / (2.0 * 3.14159265358979323846);  // M2DP.cpp:32
  const double R_res_inv = M2DP_NUM_R / max_rho;                          // M2DP.cpp:33
  int grid = nscan < num_sms ? nscan : num_sms;
  unsigned long long degen_mask = 0;   // planes with exactly zero projection vectors (M2DP.cpp:21-25 at p=2, q=0)
  for (int k = 0; k < M2DP_PQ; k++) {
    bool zero = true;
    for (int c = 0; c < 3; c++) zero = zero && xp[3 * k + c] == 0.0 && yp[3 * k + c] == 0.0;
    // the sign rule used by the kernel needs +0 entries
    for (int c = 0; c < 3; c++) zero = zero && !std::signbit(xp[3 * k + c]) && !std::signbit(yp[3 * k + c]);
    if (zero) degen_mask |= 1ull << k;
  }
  m2dp_generate_kernel<<<grid, M2_THREADS, sizeof(M2Smem), st>>>(xyz, inten, off, nscan, S_res_inv, R_res_inv,
                                                                 do_align_and_variants ? 1 : 0, hist, degen_mask);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
