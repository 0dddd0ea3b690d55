// Scan Context signature generation: SC::getSignature (SC.cpp:12-76) + align_points_PCA
// (pts_align.h:7-46) for a whole batch of scans in ONE launch.
//
// Four persistent CTAs per SM (256 threads each) walk over scans.  The only shared-memory state of a
// CTA is the 1200-bin accumulator set (37 KB), so many CTAs are resident and hide each other's
// latency-bound phases.  Per scan:
//   pass 1 (HBM read): one pass of sums about the scan's first point -> mean, 3x3 scatter matrix,
//           intensity sums; one fixed-tree block reduction
//   serial: thread 0 solves the 3x3 eigenproblem (short-chain Jacobi), thread 32 replays the order
//           dependent float intensity sum of SC.cpp:60-63 if it cannot be proven exact
//   pass 2 (L2 read, backwards so that the lines pass 1 touched last are re-read first; measured DRAM
//           read = 1.07 x the algorithmic bytes): point -> (sector, ring), shared-memory atomics (u32 count,
//           int32 / fp64 intensity sum, order-preserving int64 min / max of the height); 128-point
//           chunks are handed out by a shared counter so that all warps reach the barrier together
//   output: 2 x 1200 coalesced fp64 stores, bins cleared in the same sweep.
// The bin of a point is found in fp32 (fast_turns / rsqrt) and accepted only when the fractional
// position is at least BIN_GUARD away from a bin edge -- 10x the worst-case fp32 error -- otherwise
// the point takes the reference's own fp64 atan2 / sqrt / floor expression (SC.cpp:37-38).  The bins
// are therefore exactly those of the fp64 expression.
// HBM traffic per scan = the algorithmic bytes: 28 B/point in, 19 200 B out.
// Compiled with -fmad=false (see pca.cuh); fused multiply-adds are written explicitly where wanted.
#include <climits>
#include <cstdlib>

#include "../../include/sodso_pr.h"
#include "pca.cuh"

namespace sodso {
namespace {

constexpr int GEN_THREADS = 256;
constexpr int GEN_CTAS_PER_SM = 4;
constexpr int GEN_CHUNK = 128;        // points per work grab in the scatter pass
constexpr int GEN_CAP = 3200;         // align_pca_kernel: points of a scan staged in shared memory
constexpr float BIN_GUARD = 2.5e-4f;  // fp32 bin coordinates are accurate to < 2.5e-5 (see bin_of_point)

struct ScSmem {
  double b_sum[SC_SIZE];                    // fp64 intensity sums (int32 sums in exact mode)
  long long b_lo[SC_SIZE], b_hi[SC_SIZE];   // min / max of the height (order-preserving int64 keys)
  unsigned b_cnt[SC_SIZE];
  double scratch[11 * 32];
  double bc[16];                            // mean + eigenvectors
  int ibc[4];                               // [emin, -, float sum bits, -]
  int next_chunk;
};
static_assert((sizeof(ScSmem) + 1024) * GEN_CTAS_PER_SM <= 228 * 1024, "resident CTAs per SM");

__device__ __forceinline__ void gen_prefetch_l2(const void *p, int64_t bytes) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uintptr_t a0 = (a + 15) & ~(uintptr_t)15;
  const int64_t b = (bytes - (int64_t)(a0 - a)) & ~(int64_t)15;
  if (b > 0)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((uint32_t)b) : "memory");
}

// SC.cpp:33-44 in fp64, exactly the reference's expression: flat bin index or -1 (dropped).
__device__ __noinline__ int bin_of_point_exact(double yp, double zp, double S_res_inv, double R_res_inv) {
  const double PI = 3.14159265358979323846;  // M_PI, SC.cpp:37
  const double ang = (atan2(zp, yp) + PI) * S_res_inv;
  const double rad = sqrt(yp * yp + zp * zp) * R_res_inv;
  // `idx >= getSignatureSize()` (SC.cpp:42) is the ONLY range check: a point with ri >= 20 and
  // si < 59 aliases into the next sector.  NaN / huge values give INT_MIN on x86 and
  // are dropped there; dropped explicitly here.
  if (!(rad < (double)SC_SIZE) || !(ang < 64.0)) return -1;
  const int idx = (int)floor(ang) * SC_NUM_R + (int)floor(rad);
  return (unsigned)idx >= (unsigned)SC_SIZE ? -1 : idx;
}

// The bin of a point already in the PCA frame.  Fast path in fp32: the sector coordinate comes from
// fast_turns (common.cuh, |error| < 2e-7 turns -> 1.2e-5 sectors, plus the fp32 rounding of the inputs
// 1.1e-6 and of the final multiply 3.8e-6), the ring coordinate from r2 * rsqrt(r2) (relative 3e-7 of a
// value accepted only below 64: < 2e-5).  A point closer than BIN_GUARD = 2.5e-4 to a bin edge, or
// anything unusual (NaN, origin, huge), takes the fp64 expression of the reference, so the bins are
// exactly those of SC.cpp:37-39.
__device__ __forceinline__ int bin_of_point(double yp, double zp, double S_res_inv, double R_res_inv, float R_f) {
  const float yf = (float)yp, zf = (float)zp;
  const float tf = fast_turns(zf, yf) * (float)SC_NUM_S;
  const float r2 = __fmaf_rn(yf, yf, zf * zf);
  const float rf = r2 * rsqrtf(r2) * R_f;
  const float ft = floorf(tf), fr = floorf(rf);
  const float dt = tf - ft, dr = rf - fr;
  const bool safe = fminf(dt, dr) > BIN_GUARD && fmaxf(dt, dr) < 1.0f - BIN_GUARD && tf < (float)SC_NUM_S && rf < 64.0f;
  if (!safe) return bin_of_point_exact(yp, zp, S_res_inv, R_res_inv);
  const int idx = (int)ft * SC_NUM_R + (int)fr;
  return (unsigned)idx >= (unsigned)SC_SIZE ? -1 : idx;
}

// min / max on the order-preserving int64 keys: the value read for the pre-check seeds the CAS loop
__device__ __forceinline__ void smem_min_i64(long long *addr, long long key) {
  long long cur = *addr;
  while (key < cur) {
    const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long *>(addr),
                                                (unsigned long long)cur, (unsigned long long)key);
    if (prev == cur) break;
    cur = prev;
  }
}
__device__ __forceinline__ void smem_max_i64(long long *addr, long long key) {
  long long cur = *addr;
  while (key > cur) {
    const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long *>(addr),
                                                (unsigned long long)cur, (unsigned long long)key);
    if (prev == cur) break;
    cur = prev;
  }
}

__global__ void __launch_bounds__(GEN_THREADS, GEN_CTAS_PER_SM)
sc_generate_kernel(const double *__restrict__ xyz, const float *__restrict__ inten,
                   const int64_t *__restrict__ off, int nscan, double S_res_inv, double R_res_inv,
                   double *__restrict__ hist, int flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScSmem &S = *reinterpret_cast<ScSmem *>(smem_raw);
  const float R_f = (float)R_res_inv;
  int *b_isum = reinterpret_cast<int *>(S.b_sum);
  for (int b = threadIdx.x; b < SC_SIZE; b += GEN_THREADS) {
    S.b_sum[b] = 0.0;
    S.b_cnt[b] = 0u;
    S.b_lo[b] = LLONG_MAX;
    S.b_hi[b] = LLONG_MIN;
  }
  if (threadIdx.x == 0) S.next_chunk = 0;

  for (int scan = blockIdx.x; scan < nscan; scan += gridDim.x) {
    const int64_t p0 = off[scan];
    const int n = (int)(off[scan + 1] - p0);
    const double *g = xyz + 3 * p0;
    const float *gi = inten + p0;

    // ---- pass 1 (HBM): moments about the scan's first point o, d = p - o:
    //   s11 = [sum d (3), sum d d^T (6), sum v, sum |v|]
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
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += GEN_THREADS) {
      const float v = __ldg(gi + i);
      const double x = __ldg(g + 3 * (size_t)i + 0) - ox;
      const double y = __ldg(g + 3 * (size_t)i + 1) - oy;
      const double z = __ldg(g + 3 * (size_t)i + 2) - oz;
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
      // lowest set bit of v as a power of two (denormals come out one too small, which only makes the
      // exactness test below more conservative; inf / NaN poison s11[10])
      const unsigned b = __float_as_uint(v);
      const int e = (int)((b >> 23) & 0xffu) - 151 + __ffs(b | 0x800000u);
      emin = (v != 0.0f && e < emin) ? e : emin;
    }
    emin = __reduce_min_sync(0xffffffffu, emin);
    __syncthreads();  // ibc initialised, bins and scratch free
    if ((threadIdx.x & 31) == 0) atomicMin(&S.ibc[0], emin);
    block_sum<11>(s11, S.scratch);
    emin = S.ibc[0];
    // every partial sum of the sequential float loop is exact -> the float sum is the exact sum
    const bool exact = emin == (1 << 20) ? s11[10] == 0.0 : s11[10] < ldexp(1.0, 24 + emin);

    // ---- serial jobs: thread 0 solves the eigenproblem (pts_align.h:31-34), thread 32 replays the order
    // dependent float sum of SC.cpp:60-63 when it is not provably exact.  The other CTAs of the SM hide this.
    if (threadIdx.x == 0 && n > 0) {
      // mean = o + sum(d)/n (pts_align.h:10-18); scatter matrix = sum(d d^T) - sum(d) sum(d)^T / n (pts_align.h:21-30)
      const double dn = (double)n;
      const double mx = s11[0] / dn, my = s11[1] / dn, mz = s11[2] / dn;
      double c6[6] = {s11[3] - s11[0] * mx, s11[4] - s11[0] * my, s11[5] - s11[0] * mz,
                      s11[6] - s11[1] * my, s11[7] - s11[1] * mz, s11[8] - s11[2] * mz};
      S.bc[0] = ox + mx;
      S.bc[1] = oy + my;
      S.bc[2] = oz + mz;
      sym_eig3_fast(c6, S.bc);
    } else if (threadIdx.x == 32 && !exact) {
      float a = 0.0f;
#pragma unroll 16
      for (int i = 0; i < n; i++) a += gi[i];
      S.ibc[2] = __float_as_int(a);
    } else if (threadIdx.x == 64 && (flags & 1)) {  // pull the next scan of this CTA towards L2
      const int nxt = scan + gridDim.x;
      if (nxt < nscan) {
        const int64_t q0 = off[nxt];
        const int64_t nn = off[nxt + 1] - q0;
        gen_prefetch_l2(xyz + 3 * q0, nn * 24);
        gen_prefetch_l2(inten + q0, nn * 4);
      }
    }
    __syncthreads();

    // ---- pass 2 (L2): SC.cpp:29-57.  Backwards, so that the lines pass 1 touched last are re-read first.
    {
      const float iscale = exact && emin != (1 << 20) ? (float)ldexp(1.0, -emin) : 0.0f;
      // chunks of GEN_CHUNK points are handed out by a shared counter: warps that hit the fp64 path or
      // contended bins take fewer chunks, so all warps reach the barrier together
      const int nchunk = (n + GEN_CHUNK - 1) / GEN_CHUNK;
      const int lane = threadIdx.x & 31;
      for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&S.next_chunk, 1);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= nchunk) break;
        if (flags & 2) c = nchunk - 1 - c;
#pragma unroll
        for (int r = 0; r < GEN_CHUNK / 32; r++) {
          const int i = c * GEN_CHUNK + r * 32 + lane;
          if (i >= n) continue;
          const float it = __ldg(gi + i);
          // pts_align.h:37-45 (fused multiply-adds: the frame itself already differs from the oracle's in the last bits)
          const double x = __ldg(g + 3 * (size_t)i + 0) - S.bc[0];
          const double y = __ldg(g + 3 * (size_t)i + 1) - S.bc[1];
          const double z = __ldg(g + 3 * (size_t)i + 2) - S.bc[2];
          const double hx = fma(z, S.bc[9], fma(y, S.bc[6], x * S.bc[3]));
          const double yp = fma(z, S.bc[10], fma(y, S.bc[7], x * S.bc[4]));
          const double zp = fma(z, S.bc[11], fma(y, S.bc[8], x * S.bc[5]));
          const int idx = bin_of_point(yp, zp, S_res_inv, R_res_inv, R_f);
          if (idx >= 0) {
            atomicAdd(&S.b_cnt[idx], 1u);
            if (exact)
              atomicAdd(&b_isum[2 * idx], (int)(it * iscale));  // exact integer multiple of 2^emin, |sum| < 2^24
            else
              atomicAdd(&S.b_sum[idx], (double)it);
            const long long key = f64_key(hx);
            smem_min_i64(&S.b_lo[idx], key);
            smem_max_i64(&S.b_hi[idx], key);
          }
        }
      }
    }
    __syncthreads();

    // ---- SC.cpp:60-75: binarise against the float average, write the signature, clear the bins
    {
      const float fsum = exact ? (float)s11[9] : __int_as_float(S.ibc[2]);
      const float ave = fsum / (float)n;  // SC.cpp:64
      const double unscale = exact && emin != (1 << 20) ? ldexp(1.0, emin) : 0.0;
      double *row = hist + (size_t)scan * 2 * SC_SIZE;
      for (int b = threadIdx.x; b < SC_SIZE; b += GEN_THREADS) {
        const unsigned c = S.b_cnt[b];
        double st = 0.0, iv = 0.0;
        if (c) {
          st = f64_unkey(S.b_hi[b]) - f64_unkey(S.b_lo[b]);
          const double sum = exact ? (double)b_isum[2 * b] * unscale : S.b_sum[b];
          const double mean = sum / (double)c;
          iv = mean > (double)ave ? 1.0 : 0.0;
          S.b_sum[b] = 0.0;
          S.b_cnt[b] = 0u;
          S.b_lo[b] = LLONG_MAX;
          S.b_hi[b] = LLONG_MIN;
        }
        row[b] = st;
        row[SC_SIZE + b] = iv;
      }
      if (threadIdx.x == 0) S.next_chunk = 0;
    }
    // no barrier needed here: the next scan touches the bins / ibc / scratch only after its first barrier
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
  const int per_sm = g_debug.gen_ctas > 0 ? g_debug.gen_ctas : GEN_CTAS_PER_SM, flags = g_debug.gen_flags;
  int grid = nscan < per_sm * num_sms ? nscan : per_sm * num_sms;
  sc_generate_kernel<<<grid, GEN_THREADS, sizeof(ScSmem), st>>>(xyz, inten, off, nscan, S_res_inv,
                                                                R_res_inv, hist, flags);
  if (launches) ++*launches;
  return cudaGetLastError();
}

namespace {
__global__ void fast_turns_probe_kernel(const float *__restrict__ num, const float *__restrict__ den, int64_t n,
                                        float *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = fast_turns(num[i], den[i]);
}
}  // namespace

// test hook: the fp32 angle proposal used by the generation kernels, so that its error bound can be checked directly
cudaError_t launch_fast_turns_probe(const float *num, const float *den, int64_t n, float *out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  fast_turns_probe_kernel<<<1024, 256, 0, st>>>(num, den, n, out);
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
