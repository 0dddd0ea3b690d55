// Scan Context signature generation: SC::getSignature (SC.cpp:12-76) + align_points_PCA
// (pts_align.h:7-46) for a whole batch of scans in ONE launch.
//
// One persistent CTA per SM walks over scans (grid-stride).  Per scan:
//   stage points (coalesced, AoS -> SoA shared memory)  ->  PCA (two block reductions + a
//   3x3 Jacobi on one thread, overlapped with bin clearing / intensity average on the others)
//   ->  point -> (sector, ring) scatter with shared-memory atomics (u32 count, fp64 sum,
//   order-preserving int64 min / max)  ->  2 x 1200 coalesced fp64 stores.
// HBM traffic per scan = the algorithmic bytes: 28 B/point in, 19 200 B out.
// Compiled with -fmad=false (see pca.cuh).
#include <climits>

#include "../../include/sodso_pr.h"
#include "pca.cuh"

namespace sodso {
namespace {

constexpr int GEN_THREADS = 512;
constexpr int GEN_CAP = 6144;  // staged points per scan; larger scans spill to global re-reads

struct ScSmem {
  double sx[GEN_CAP], sy[GEN_CAP], sz[GEN_CAP];
  double b_sum[SC_SIZE];
  long long b_lo[SC_SIZE], b_hi[SC_SIZE];
  double scratch[6 * 32];
  double bc[16];
  float si[GEN_CAP];
  unsigned b_cnt[SC_SIZE];
  int ibc[4];
};

__global__ void __launch_bounds__(GEN_THREADS, 1)
sc_generate_kernel(const double *__restrict__ xyz, const float *__restrict__ inten,
                   const int64_t *__restrict__ off, int nscan, double S_res_inv, double R_res_inv,
                   double *__restrict__ hist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScSmem &S = *reinterpret_cast<ScSmem *>(smem_raw);
  const double PI = 3.14159265358979323846;  // M_PI, SC.cpp:37

  for (int scan = blockIdx.x; scan < nscan; scan += gridDim.x) {
    const int64_t p0 = off[scan];
    const int n = (int)(off[scan + 1] - p0);
    const double *g = xyz + 3 * p0;
    const float *gi = inten + p0;
    const int nst = n < GEN_CAP ? n : GEN_CAP;

    // pull the next scan of this CTA towards L2 while this one is processed
    {
      int nxt = scan + gridDim.x;
      if (nxt < nscan) {
        const int64_t q0 = off[nxt];
        const int64_t nb = (off[nxt + 1] - q0) * 24;
        const char *base = reinterpret_cast<const char *>(xyz + 3 * q0);
        for (int64_t o = (int64_t)threadIdx.x * 128; o < nb; o += (int64_t)blockDim.x * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
      }
    }

    stage_scan(g, n, GEN_CAP, S.sx, S.sy, S.sz);
    for (int i = threadIdx.x; i < nst; i += blockDim.x) S.si[i] = gi[i];
    for (int b = threadIdx.x; b < SC_SIZE; b += blockDim.x) {
      S.b_sum[b] = 0.0;
      S.b_cnt[b] = 0u;
      S.b_lo[b] = LLONG_MAX;
      S.b_hi[b] = LLONG_MIN;
    }
    __syncthreads();

    ScanPoints P{g, S.sx, S.sy, S.sz, n, nst};
    const float ave = scan_ave_intensity(gi, S.si, n, nst, S.scratch, S.ibc);  // SC.cpp:60-64
    scan_pca(P, S.scratch, S.bc);                                              // SC.cpp:17

    // SC.cpp:29-57
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double x, y, z, hx, yp, zp;
      P.get(i, x, y, z);
      pca_rotate(S.bc, x, y, z, hx, yp, zp);
      const double ang = (atan2(zp, yp) + PI) * S_res_inv;
      const double rad = sqrt(yp * yp + zp * zp) * R_res_inv;
      // `idx >= getSignatureSize()` (SC.cpp:42) is the ONLY range check: a point with
      // ri >= 20 and si < 59 aliases into the next sector.  NaN / huge values give
      // INT_MIN on x86 and are dropped there; dropped explicitly here.
      if (!(rad < (double)SC_SIZE) || !(ang < 64.0)) continue;
      const int si = (int)floor(ang);
      const int ri = (int)floor(rad);
      const int idx = si * SC_NUM_R + ri;
      if ((unsigned)idx >= (unsigned)SC_SIZE) continue;
      const float it = i < nst ? S.si[i] : gi[i];
      atomicAdd(&S.b_cnt[idx], 1u);
      atomicAdd(&S.b_sum[idx], (double)it);
      const long long key = f64_key(hx);
      atomicMin(&S.b_lo[idx], key);
      atomicMax(&S.b_hi[idx], key);
    }
    __syncthreads();

    // SC.cpp:67-75
    double *row = hist + (size_t)scan * 2 * SC_SIZE;
    for (int b = threadIdx.x; b < SC_SIZE; b += blockDim.x) {
      const unsigned c = S.b_cnt[b];
      double st = 0.0, iv = 0.0;
      if (c) {
        st = f64_unkey(S.b_hi[b]) - f64_unkey(S.b_lo[b]);
        const double mean = S.b_sum[b] / (double)c;
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
  int grid = nscan < num_sms ? nscan : num_sms;
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
