// Scan Context signature generation: SC::getSignature (SC.cpp:12-76) + align_points_PCA
// (pts_align.h:7-46) for a whole batch of scans in ONE launch.
//
// Two persistent CTAs per SM (256 threads each) walk over scans; while one CTA sits in a
// latency-bound phase (bulk copy in flight, the serial 3x3 eigen-solve) the other one computes.
// Per scan:
//   cp.async.bulk (TMA, 1-D) lands the scan's raw AoS fp64 points in shared memory, no register
//   staging, next scan prefetched into L2   ->   mean + intensity sums (one fixed-tree block
//   reduction)   ->   scatter matrix (second reduction)   ->   3x3 Jacobi on one thread while the
//   other threads clear the bins   ->   point -> (sector, ring) scatter with shared-memory atomics
//   (u32 count, int32 / fp64 intensity sum, order-preserving int64 min / max of the height)   ->
//   2 x 1200 coalesced fp64 stores.
// The bin of a point is found in fp32 (atan2f / sqrtf) and accepted only when the fractional
// position is at least GUARD away from a bin edge -- 10x the worst-case fp32 error -- otherwise the
// point takes the reference's own fp64 atan2 / sqrt / floor expression (SC.cpp:37-38).  The bins
// are therefore exactly those of the fp64 expression.
// HBM traffic per scan = the algorithmic bytes: 28 B/point in, 19 200 B out.
// Compiled with -fmad=false (see pca.cuh).
#include <climits>

#include "../../include/sodso_pr.h"
#include "pca.cuh"

namespace sodso {
namespace {

constexpr int GEN_THREADS = 256;
constexpr int GEN_CAP = 3200;       // points of a scan staged in shared memory; the rest is re-read from L2
constexpr float BIN_GUARD = 2.5e-4f;  // fp32 bin coordinate error is < 2e-5 (see bin_of_point)

struct ScSmem {
  double pts[3 * GEN_CAP + 2];   // raw AoS as landed by the bulk copy; element j of the scan at pts[j + shift]
  double b_sum[SC_SIZE];         // fp64 intensity sums; aliased as int32 sums in exact mode
  long long b_lo[SC_SIZE], b_hi[SC_SIZE];
  double scratch[6 * 32];
  double bc[16];
  unsigned long long mbar;
  unsigned b_cnt[SC_SIZE];
  int ibc[4];
};
static_assert(sizeof(ScSmem) <= 113 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint32_t gen_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool gen_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Thread 0: start the bulk copy of the 16-byte aligned part of a scan's xyz block.  Element j of the
// scan (double index) lands at pts[j + shift], shift = 1 when the block starts at an address that
// is 8 mod 16; the unaligned head / tail element is copied by gen_fix_edges.
__device__ __forceinline__ void gen_issue_load(ScSmem &S, const double *g, int nst) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(g);
  const int shift = (int)((addr >> 3) & 1);
  const int ne = 3 * nst;
  const int e0 = ne > 0 ? shift : 0;
  const uint32_t bytes = ne > e0 ? (uint32_t)(((ne - e0) * 8) & ~15) : 0u;
  const uint32_t bar = gen_smem_u32(&S.mbar);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of pts vs the async write
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  if (bytes)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     gen_smem_u32(&S.pts[e0 + shift])),
                 "l"(g + e0), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void gen_fix_edges(ScSmem &S, const double *g, int nst) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(g);
  const int shift = (int)((addr >> 3) & 1);
  const int ne = 3 * nst;
  if (ne == 0) return;
  const int e0 = shift;
  const int bulk_e = ne > e0 ? (((ne - e0) * 8) & ~15) / 8 : 0;
  if (threadIdx.x == 0 && e0 == 1) S.pts[shift] = g[0];
  if (threadIdx.x == 32 && e0 + bulk_e < ne) S.pts[shift + ne - 1] = g[ne - 1];  // at most one tail element
}

__device__ __forceinline__ void gen_prefetch_l2(const void *p, int64_t bytes) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uintptr_t a0 = (a + 15) & ~(uintptr_t)15;
  const int64_t b = (bytes - (int64_t)(a0 - a)) & ~(int64_t)15;
  if (b > 0)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((uint32_t)b) : "memory");
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

__device__ __noinline__ void sym_eig3_fast(const double *cov6, double *bc) {
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

// SC.cpp:33-44 for one point already in the PCA frame: flat bin index or -1 (dropped).
// Fast path: fp32.  Error budget of the fp32 sector coordinate tf (<= 60.5): inputs rounded to fp32
// (1.2e-7 rad), atan2f (<= 3 ulp at pi = 7.2e-7), the +pi add (3.3e-7), the multiply by 60/2pi
// (x 9.55) and its rounding (1.9e-6 + 3.6e-6): < 2e-5.  Ring coordinate rf: relative 3e-7 of a value
// < 1199: accepted only below 64 where that is < 2e-5.  A point closer than BIN_GUARD to an edge, or
// anything unusual (NaN, huge), takes the fp64 expression of the reference.
__device__ __forceinline__ int bin_of_point(double yp, double zp, double S_res_inv, double R_res_inv, float S_f,
                                            float R_f) {
  const float yf = (float)yp, zf = (float)zp;
  const float tf = (atan2f(zf, yf) + 3.14159274f) * S_f;
  const float rf = sqrtf(yf * yf + zf * zf) * R_f;
  const float ft = floorf(tf), fr = floorf(rf);
  const float dt = tf - ft, dr = rf - fr;
  const bool safe = dt > BIN_GUARD && dt < 1.0f - BIN_GUARD && dr > BIN_GUARD && dr < 1.0f - BIN_GUARD &&
                    tf > 0.0f && tf < 60.0f && rf < 64.0f;
  int si, ri;
  if (safe) {
    si = (int)ft;
    ri = (int)fr;
  } else {
    const double PI = 3.14159265358979323846;  // M_PI, SC.cpp:37
    const double ang = (atan2(zp, yp) + PI) * S_res_inv;
    const double rad = sqrt(yp * yp + zp * zp) * R_res_inv;
    // `idx >= getSignatureSize()` (SC.cpp:42) is the ONLY range check: a point with ri >= 20 and
    // si < 59 aliases into the next sector.  NaN / huge values give INT_MIN on x86 and
    // are dropped there; dropped explicitly here.
    if (!(rad < (double)SC_SIZE) || !(ang < 64.0)) return -1;
    si = (int)floor(ang);
    ri = (int)floor(rad);
  }
  const int idx = si * SC_NUM_R + ri;
  return (unsigned)idx >= (unsigned)SC_SIZE ? -1 : idx;
}

// point i of the scan: staged part from shared memory, the rest (scans above GEN_CAP) from global / L2
__device__ __forceinline__ void gen_point(const double *sp, const double *g, int nst, int i, double &x, double &y,
                                          double &z) {
  if (i < nst) {
    x = sp[3 * i + 0];
    y = sp[3 * i + 1];
    z = sp[3 * i + 2];
  } else {
    x = g[3 * (size_t)i + 0];
    y = g[3 * (size_t)i + 1];
    z = g[3 * (size_t)i + 2];
  }
}

__global__ void __launch_bounds__(GEN_THREADS, 2)
sc_generate_kernel(const double *__restrict__ xyz, const float *__restrict__ inten,
                   const int64_t *__restrict__ off, int nscan, double S_res_inv, double R_res_inv,
                   double *__restrict__ hist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScSmem &S = *reinterpret_cast<ScSmem *>(smem_raw);
  const float S_f = (float)S_res_inv, R_f = (float)R_res_inv;
  const uint32_t bar = gen_smem_u32(&S.mbar);
  int *b_isum = reinterpret_cast<int *>(S.b_sum);

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if ((int)blockIdx.x < nscan) {
      const int64_t p0 = off[blockIdx.x];
      const int n0 = (int)(off[blockIdx.x + 1] - p0);
      gen_issue_load(S, xyz + 3 * p0, n0 < GEN_CAP ? n0 : GEN_CAP);
    }
  }
  __syncthreads();

  uint32_t phase = 0;
  for (int scan = blockIdx.x; scan < nscan; scan += gridDim.x) {
    const int64_t p0 = off[scan];
    const int n = (int)(off[scan + 1] - p0);
    const double *g = xyz + 3 * p0;
    const float *gi = inten + p0;
    const int nst = n < GEN_CAP ? n : GEN_CAP;
    const int shift = (int)((reinterpret_cast<uintptr_t>(g) >> 3) & 1);
    const double *sp = S.pts + shift;

    // pull the next scan of this CTA towards L2 while this one is processed
    const int nxt = scan + gridDim.x;
    if (threadIdx.x == 64 && nxt < nscan) {
      const int64_t q0 = off[nxt];
      const int64_t nn = off[nxt + 1] - q0;
      gen_prefetch_l2(xyz + 3 * q0, nn * 24);
      gen_prefetch_l2(inten + q0, nn * 4);
    }

    // ---- pass 1: mean (pts_align.h:10-18) and the intensity sums for SC.cpp:60-64
    double s5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    int emin = 1 << 20, bad = 0;
    for (int i = threadIdx.x; i < n; i += GEN_THREADS) {
      const float v = gi[i];
      s5[3] += (double)v;
      s5[4] += fabs((double)v);
      const unsigned b = __float_as_uint(v);
      const unsigned ex = (b >> 23) & 0xffu, man = b & 0x7fffffu;
      if (ex == 0xffu) bad = 1;
      if (ex != 0 || man != 0) {
        const unsigned m = ex ? (man | 0x800000u) : man;
        const int e = (ex ? (int)ex - 150 : -149) + (__ffs(m) - 1);
        emin = e < emin ? e : emin;
      }
    }
    if (threadIdx.x == 0) {
      S.ibc[0] = 1 << 20;
      S.ibc[1] = 0;
    }
    while (!gen_mbar_try(bar, phase)) {
    }
    phase ^= 1;
    gen_fix_edges(S, g, nst);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += GEN_THREADS) {
      double x, y, z;
      gen_point(sp, g, nst, i, x, y, z);
      s5[0] += x;
      s5[1] += y;
      s5[2] += z;
    }
    emin = __reduce_min_sync(0xffffffffu, emin);
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&S.ibc[0], emin);
      if (bad) atomicOr(&S.ibc[1], 1);
    }
    block_sum<5>(s5, S.scratch);
    const double mx = s5[0] / (double)n, my = s5[1] / (double)n, mz = s5[2] / (double)n;
    emin = S.ibc[0];
    // every partial sum of the sequential float loop is exact -> the float sum is the exact sum
    const bool exact = !S.ibc[1] && (emin == (1 << 20) || s5[4] < ldexp(1.0, 24 + emin));

    // ---- pass 2: scatter matrix of the centred points (pts_align.h:21-30)
    double c6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += GEN_THREADS) {
      double x, y, z;
      gen_point(sp, g, nst, i, x, y, z);
      x -= mx;
      y -= my;
      z -= mz;
      c6[0] += x * x;
      c6[1] += x * y;
      c6[2] += x * z;
      c6[3] += y * y;
      c6[4] += y * z;
      c6[5] += z * z;
    }
    __syncthreads();  // scratch reuse
    block_sum<6>(c6, S.scratch);

    // ---- eigen-solve on one thread (pts_align.h:31-34); the others clear the bins meanwhile
    if (threadIdx.x == 0) {
      S.bc[0] = mx;
      S.bc[1] = my;
      S.bc[2] = mz;
      sym_eig3_fast(c6, S.bc);
    } else if (threadIdx.x == 32) {
      if (!exact) {  // replay the sequential float loop of SC.cpp:60-63 (order dependent rounding)
        float a = 0.0f;
#pragma unroll 16
        for (int i = 0; i < n; i++) a += gi[i];
        reinterpret_cast<float *>(S.ibc)[2] = a;
      }
    } else {
      for (int b = threadIdx.x - (threadIdx.x > 32 ? 2 : 1); b < SC_SIZE; b += GEN_THREADS - 2) {
        S.b_sum[b] = 0.0;
        S.b_cnt[b] = 0u;
        S.b_lo[b] = LLONG_MAX;
        S.b_hi[b] = LLONG_MIN;
      }
    }
    __syncthreads();
    const float ave = (exact ? (float)s5[3] : reinterpret_cast<float *>(S.ibc)[2]) / (float)n;  // SC.cpp:64
    const float iscale = exact && emin != (1 << 20) ? (float)ldexp(1.0, -emin) : 0.0f;
    const double unscale = exact && emin != (1 << 20) ? ldexp(1.0, emin) : 0.0;

    // ---- SC.cpp:29-57
    for (int i = threadIdx.x; i < n; i += GEN_THREADS) {
      const float it = gi[i];
      double x, y, z, hx, yp, zp;
      gen_point(sp, g, nst, i, x, y, z);
      pca_rotate(S.bc, x, y, z, hx, yp, zp);
      const int idx = bin_of_point(yp, zp, S_res_inv, R_res_inv, S_f, R_f);
      if (idx < 0) continue;
      atomicAdd(&S.b_cnt[idx], 1u);
      if (exact)
        atomicAdd(&b_isum[2 * idx], (int)(it * iscale));  // exact integer multiple of 2^emin, |sum| < 2^24
      else
        atomicAdd(&S.b_sum[idx], (double)it);
      const long long key = f64_key(hx);
      if (key < S.b_lo[idx]) atomicMin(&S.b_lo[idx], key);
      if (key > S.b_hi[idx]) atomicMax(&S.b_hi[idx], key);
    }
    __syncthreads();

    // the point buffer is free: start the next scan's copy before writing this one's signature
    if (threadIdx.x == 0 && nxt < nscan) {
      const int64_t q0 = off[nxt];
      const int nn = (int)(off[nxt + 1] - q0);
      gen_issue_load(S, xyz + 3 * q0, nn < GEN_CAP ? nn : GEN_CAP);
    }

    // ---- SC.cpp:67-75
    double *row = hist + (size_t)scan * 2 * SC_SIZE;
    for (int b = threadIdx.x; b < SC_SIZE; b += GEN_THREADS) {
      const unsigned c = S.b_cnt[b];
      double st = 0.0, iv = 0.0;
      if (c) {
        st = f64_unkey(S.b_hi[b]) - f64_unkey(S.b_lo[b]);
        const double sum = exact ? (double)b_isum[2 * b] * unscale : S.b_sum[b];
        const double mean = sum / (double)c;
        iv = mean > (double)ave ? 1.0 : 0.0;
      }
      row[b] = st;
      row[SC_SIZE + b] = iv;
    }
    __syncthreads();
  }
}

struct PcaSmem {
  double sx[GEN_CAP], sy[GEN_CAP], sz[GEN_CAP];
  double scratch[6 * 32];
  double bc[16];
};

__global__ void __launch_bounds__(GEN_THREADS, 1)
align_pca_kernel(const double *__restrict__ xyz, const int64_t *__restrict__ off, int nscan,
                 double *__restrict__ out, double *__restrict__ evec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PcaSmem &S = *reinterpret_cast<PcaSmem *>(smem_raw);
  for (int scan = blockIdx.x; scan < nscan; scan += gridDim.x) {
    const int64_t p0 = off[scan];
    const int n = (int)(off[scan + 1] - p0);
    const double *g = xyz + 3 * p0;
    const int nst = n < GEN_CAP ? n : GEN_CAP;
    stage_scan(g, n, GEN_CAP, S.sx, S.sy, S.sz);
    __syncthreads();
    ScanPoints P{g, S.sx, S.sy, S.sz, n, nst};
    scan_pca(P, S.scratch, S.bc);
    double *o = out + 3 * p0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double x, y, z, ox, oy, oz;
      P.get(i, x, y, z);
      pca_rotate(S.bc, x, y, z, ox, oy, oz);
      o[3 * (size_t)i + 0] = ox;
      o[3 * (size_t)i + 1] = oy;
      o[3 * (size_t)i + 2] = oz;
    }
    if (evec && threadIdx.x < 9) evec[(size_t)scan * 9 + threadIdx.x] = S.bc[3 + threadIdx.x];
    __syncthreads();
  }
}

}  // namespace

cudaError_t launch_sc_generate(const double *xyz, const float *inten, const int64_t *off, int nscan,
                               double max_rho, double *hist, int num_sms, cudaStream_t st,
                               int64_t *launches) {
  if (nscan <= 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(sc_generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(ScSmem));
  if (e != cudaSuccess) return e;
  const double S_res_inv = SC_NUM_S / (2.0 * 3.14159265358979323846);  // SC.cpp:6
  const double R_res_inv = SC_NUM_R / max_rho;                          // SC.cpp:7
  int grid = nscan < 2 * num_sms ? nscan : 2 * num_sms;
  sc_generate_kernel<<<grid, GEN_THREADS, sizeof(ScSmem), st>>>(xyz, inten, off, nscan, S_res_inv,
                                                                R_res_inv, hist);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_align_pca(const double *xyz, const int64_t *off, int nscan, double *out_xyz,
                             double *evec, int num_sms, cudaStream_t st, int64_t *launches) {
  if (nscan <= 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(align_pca_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(PcaSmem));
  if (e != cudaSuccess) return e;
  int grid = nscan < num_sms ? nscan : num_sms;
  align_pca_kernel<<<grid, GEN_THREADS, sizeof(PcaSmem), st>>>(xyz, off, nscan, out_xyz, evec);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
