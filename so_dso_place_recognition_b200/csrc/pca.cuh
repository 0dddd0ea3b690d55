// Per-scan PCA alignment on one CTA (pts_align.h:7-46) and small block-level helpers shared by
// the generation kernels.  Translation units including this file are compiled with
// -fmad=false: every fp64 operation is a separately rounded IEEE operation, like the
// reference built without contraction and like the CPU oracle.
#pragma once
#include "common.cuh"

namespace sodso {

// order-preserving map double -> int64 (involution), so that smem atomicMin/Max on int64
// implement min/max on doubles.
__device__ __forceinline__ long long f64_key(double v) {
  long long b = __double_as_longlong(v);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double f64_unkey(long long k) {
  return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Fixed-tree (deterministic) block sum of NV doubles per thread; result valid in ALL threads.
// scratch: NV * 32 doubles.  Contains two __syncthreads().
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double s = warp_sum(v[k]);
    if (lane == 0) scratch[k * 32 + warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double s = lane < nwarp ? scratch[k * 32 + lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) scratch[k * 32] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = scratch[k * 32];
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi, eigenvalues ascending, standing in
// for Eigen::SelfAdjointEigenSolver (pts_align.h:31-34).  Sign convention (Eigen's is
// implementation-defined): the largest-magnitude component of every eigenvector is
// positive (ties -> lowest index).  Same operation sequence as the CPU oracle.
// a: row-major 3x3.  v: row-major, column k = k-th eigenvector.
__device__ inline void sym_eig3(const double a_in[9], double w[3], double v[9]) {
  double a[3][3], q[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      a[i][j] = a_in[i * 3 + j];
      q[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; p++) {
#pragma unroll
      for (int r = p + 1; r < 3; r++) {
        double apq = a[p][r];
        if (apq == 0.0) continue;
        double theta = (a[r][r] - a[p][p]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        double c = 1.0 / sqrt(t * t + 1.0);
        double s = t * c;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double akp = a[k][p], akr = a[k][r];
          a[k][p] = c * akp - s * akr;
          a[k][r] = s * akp + c * akr;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double apk = a[p][k], ark = a[r][k];
          a[p][k] = c * apk - s * ark;
          a[r][k] = s * apk + c * ark;
        }
        a[p][r] = 0.0;
        a[r][p] = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double qkp = q[k][p], qkr = q[k][r];
          q[k][p] = c * qkp - s * qkr;
          q[k][r] = s * qkp + c * qkr;
        }
      }
    }
  }
  int order[3] = {0, 1, 2};
  double d[3] = {a[0][0], a[1][1], a[2][2]};
  for (int i = 1; i < 3; i++) {  // stable insertion sort, ascending
    int oi = order[i];
    int j = i - 1;
    while (j >= 0 && d[order[j]] > d[oi]) {
      order[j + 1] = order[j];
      j--;
    }
    order[j + 1] = oi;
  }
  for (int k = 0; k < 3; k++) {
    int src = order[k];
    w[k] = d[src];
    double col[3] = {q[0][src], q[1][src], q[2][src]};
    int big = 0;
    for (int i = 1; i < 3; i++)
      if (fabs(col[i]) > fabs(col[big])) big = i;
    double sgn = (col[big] < 0.0) ? -1.0 : 1.0;
    for (int i = 0; i < 3; i++) v[i * 3 + k] = sgn * col[i];
  }
}

// Symmetric 3x3 eigen-decomposition for the kernels (ascending eigenvalues, sign convention of
// sym_eig3).  Cyclic Jacobi like the oracle's, arranged for a short dependency chain on one
// thread: per rotation  h = sqrt(d^2 + b^2),  t = +-b / (|d| + h)  and  c = sqrt((h + |d|) / 2h)  are
// algebraically the oracle's  t = sgn/(|theta| + sqrt(theta^2 + 1)),  c = 1/sqrt(t^2 + 1)  with theta =
// d / b, and a rotation is skipped once |a_pq| <= 2^-54 (|a_pp| + |a_qq|) (a backward error below
// half an ulp of the diagonal) instead of iterating to exact zeros.  Eigenvectors agree with the
// oracle's to ~1e-15 / gap.
__device__ __forceinline__ bool jacobi_rot(double &app, double &arr, double &apr, double &akp, double &akr,
                                           double &q0p, double &q0r, double &q1p, double &q1r, double &q2p,
                                           double &q2r) {
  if (fabs(apr) <= 0x1p-54 * (fabs(app) + fabs(arr))) {
    apr = 0.0;
    return false;
  }
  const double d = arr - app, b = 2.0 * apr;
  const double h = sqrt(d * d + b * b);
  const double ad = fabs(d);
  double t = b / (ad + h);
  if (d < 0.0) t = -t;
  const double c = sqrt((h + ad) / (2.0 * h));
  const double s = t * c;
  app = app - t * apr;
  arr = arr + t * apr;
  apr = 0.0;
  const double kp = akp, kr = akr;
  akp = c * kp - s * kr;
  akr = s * kp + c * kr;
  double a, bb;
  a = q0p, bb = q0r, q0p = c * a - s * bb, q0r = s * a + c * bb;
  a = q1p, bb = q1r, q1p = c * a - s * bb, q1r = s * a + c * bb;
  a = q2p, bb = q2r, q2p = c * a - s * bb, q2r = s * a + c * bb;
  return true;
}

static __device__ __noinline__ void sym_eig3_fast(const double *cov6, double *bc) {
  // cov6 = {xx, xy, xz, yy, yz, zz}
  double a00 = cov6[0], a01 = cov6[1], a02 = cov6[2], a11 = cov6[3], a12 = cov6[4], a22 = cov6[5];
  double q00 = 1, q01 = 0, q02 = 0, q10 = 0, q11 = 1, q12 = 0, q20 = 0, q21 = 0, q22 = 1;
  for (int sweep = 0; sweep < 32; sweep++) {
    bool any = false;
    any |= jacobi_rot(a00, a11, a01, a02, a12, q00, q01, q10, q11, q20, q21);  // (0,1), k = 2
    any |= jacobi_rot(a00, a22, a02, a01, a12, q00, q02, q10, q12, q20, q22);  // (0,2), k = 1
    any |= jacobi_rot(a11, a22, a12, a01, a02, q01, q02, q11, q12, q21, q22);  // (1,2), k = 0
    if (!any) break;
  }
  double w0 = a00, w1 = a11, w2 = a22;
  double v0[3] = {q00, q10, q20}, v1[3] = {q01, q11, q21}, v2[3] = {q02, q12, q22};
  // stable ascending sort of three (same result as the oracle's insertion sort)
#define SODSO_SWAP(wa, va, wb, vb)              \
  if (wa > wb) {                                \
    double tw = wa;                             \
    wa = wb;                                    \
    wb = tw;                                    \
    for (int i_ = 0; i_ < 3; i_++) {            \
      double tv = va[i_];                       \
      va[i_] = vb[i_];                          \
      vb[i_] = tv;                              \
    }                                           \
  }
  SODSO_SWAP(w0, v0, w1, v1)
  SODSO_SWAP(w1, v1, w2, v2)
  SODSO_SWAP(w0, v0, w1, v1)
#undef SODSO_SWAP
  double *vs[3] = {v0, v1, v2};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double *col = vs[k];
    int big = 0;
    if (fabs(col[1]) > fabs(col[big])) big = 1;
    if (fabs(col[2]) > fabs(col[big])) big = 2;
    const double sgn = (col[big] < 0.0) ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < 3; i++) bc[3 + i * 3 + k] = sgn * col[i];
  }
  bc[12] = w0;
  bc[13] = w1;
  bc[14] = w2;
}

// Points of one scan: the first `nst` are staged in shared memory as SoA, the rest (scans
// larger than the staging capacity) are read from global memory.
struct ScanPoints {
  const double *g;   // AoS n x 3 in global memory
  const double *sx, *sy, *sz;
  int n, nst;
  __device__ __forceinline__ void get(int i, double &x, double &y, double &z) const {
    if (i < nst) {
      x = sx[i];
      y = sy[i];
      z = sz[i];
    } else {
      x = g[3 * (size_t)i + 0];
      y = g[3 * (size_t)i + 1];
      z = g[3 * (size_t)i + 2];
    }
  }
};

// Cooperative staging of up to cap points of a scan into SoA shared memory.
__device__ __forceinline__ void stage_scan(const double *g, int n, int cap, double *sx, double *sy,
                                           double *sz) {
  const int nst = n < cap ? n : cap;
  for (int f = threadIdx.x; f < 3 * nst; f += blockDim.x) {
    double v = g[f];
    int p = f / 3, c = f - 3 * p;
    double *dst = c == 0 ? sx : (c == 1 ? sy : sz);
    dst[p] = v;
  }
}

// pts_align.h:10-34 for one scan on one CTA.  On return (after the internal barriers) bc[0..2]
// = mean, bc[3..11] = eigenvectors (row-major, column k = k-th), bc[12..14] = eigenvalues, in
// shared memory, visible to all threads.  scratch: 6*32 doubles; bc: 16 doubles.
__device__ inline void scan_pca(const ScanPoints &P, double *scratch, double *bc) {
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < P.n; i += blockDim.x) {
    double x, y, z;
    P.get(i, x, y, z);
    s[0] += x;
    s[1] += y;
    s[2] += z;
  }
  block_sum<3>(s, scratch);
  const double mx = s[0] / (double)P.n, my = s[1] / (double)P.n, mz = s[2] / (double)P.n;
  double c[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < P.n; i += blockDim.x) {
    double x, y, z;
    P.get(i, x, y, z);
    x -= mx;
    y -= my;
    z -= mz;
    c[0] += x * x;
    c[1] += x * y;
    c[2] += x * z;
    c[3] += y * y;
    c[4] += y * z;
    c[5] += z * z;
  }
  __syncthreads();  // scratch reuse
  block_sum<6>(c, scratch);
  if (threadIdx.x == 0) {
    double cov[9] = {c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5]};
    double w[3], v[9];
    sym_eig3(cov, w, v);
    bc[0] = mx;
    bc[1] = my;
    bc[2] = mz;
    for (int k = 0; k < 9; k++) bc[3 + k] = v[k];
    for (int k = 0; k < 3; k++) bc[12 + k] = w[k];
  }
  __syncthreads();
}

// pts_align.h:37-45 for one point: (x,y,z) raw -> PCA frame.
__device__ __forceinline__ void pca_rotate(const double *bc, double x, double y, double z, double &ox,
                                           double &oy, double &oz) {
  x -= bc[0];
  y -= bc[1];
  z -= bc[2];
  ox = (x * bc[3] + y * bc[6]) + z * bc[9];
  oy = (x * bc[4] + y * bc[7]) + z * bc[10];
  oz = (x * bc[5] + y * bc[8]) + z * bc[11];
}

// SC.cpp:60-64 / M2DP.cpp:77-81: `float ave = 0; for (...) ave += intensity; ave = ave / n`.
// The float running sum is order dependent in general.  It is reproduced exactly:
//  - if every value is a non-negative-or-negative multiple of a common power of two u and
//    sum|v| < 2^24 u, every partial sum is exactly representable in fp32, so the sequential
//    float sum equals the exact sum (computed here in fp64, any order);   [SO-DSO
//    intensities are means of 8 uint8 pixels = multiples of 1/8, OutputWrapperSODSO.cpp:47-50]
//  - otherwise thread 0 replays the sequential fp32 loop.
// Result (float, already divided by n) is returned in all threads.  scratch: 3*32 doubles,
// ibc: 2 ints + 1 float of shared memory.
__device__ inline float scan_ave_intensity(const float *gi, const float *si, int n, int nst,
                                           double *scratch, int *ibc) {
  double s[2] = {0.0, 0.0};
  int emin = 1 << 20;
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = i < nst ? si[i] : gi[i];
    s[0] += (double)v;
    s[1] += fabs((double)v);
    unsigned b = __float_as_uint(v);
    unsigned ex = (b >> 23) & 0xffu, man = b & 0x7fffffu;
    if (ex == 0xffu) bad = 1;
    if (ex != 0 || man != 0) {
      unsigned m = ex ? (man | 0x800000u) : man;
      int e = (ex ? (int)ex - 150 : -149) + (__ffs(m) - 1);
      emin = e < emin ? e : emin;
    }
  }
  // reduce emin / bad with warp ops + shared atomics
  if (threadIdx.x == 0) {
    ibc[0] = 1 << 20;
    ibc[1] = 0;
  }
  __syncthreads();
  emin = __reduce_min_sync(0xffffffffu, emin);
  bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&ibc[0], emin);
    if (bad) atomicOr(&ibc[1], 1);
  }
  block_sum<2>(s, scratch);  // has barriers: ibc complete afterwards
  emin = ibc[0];
  bad = ibc[1];
  float ave;
  bool exact = !bad && (emin == (1 << 20) || s[1] < ldexp(1.0, 24 + emin));
  if (exact) {
    ave = (float)s[0];
  } else {
    if (threadIdx.x == 0) {
      float a = 0.0f;
      for (int i = 0; i < n; i++) a += (i < nst ? si[i] : gi[i]);
      reinterpret_cast<float *>(ibc)[2] = a;
    }
    __syncthreads();
    ave = reinterpret_cast<float *>(ibc)[2];
  }
  __syncthreads();  // ibc / scratch may be reused by the caller
  return ave / (float)n;
}

}  // namespace sodso
