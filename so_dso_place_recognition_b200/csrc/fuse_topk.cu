// Channel fusion + decision, run_test.m:38-57: row z-score of both distance matrices (std with
// N-1 over the UNMASKED row), fused = p_weight * z_p + z_i, temporal mask |i-j| < mask_width ->
// Inf, first-index argmin (generalised to the k best for the row-sharded database).
// One CTA per query row; all statistics in fp64 with a fixed reduction tree.
#include <cmath>
#include <cstring>

#include "../../include/sodso_pr.h"
#include "common.cuh"

namespace sodso {
namespace {

constexpr int FUSE_THREADS = 256;
constexpr int FUSE_UNROLL = 8;     // row elements a thread loads together in fuse_topk_kernel
constexpr int STATS_UNROLL = 10;   // ... and in row_stats_kernel

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__device__ __forceinline__ void block_sum_d(double (&v)[NV], double *scratch /* NV*32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double s = warp_sum_d(v[k]);
    if (lane == 0) scratch[k * 32 + warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double s = lane < nwarp ? scratch[k * 32 + lane] : 0.0;
      s = warp_sum_d(s);
      if (lane == 0) scratch[k * 32] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = scratch[k * 32];
  __syncthreads();
}

// (score, idx) lexicographic "less" with NaN never winning; idx < 0 means "none".
__device__ __forceinline__ bool cand_less(double sa, long long ia, double sb, long long ib) {
  if (ia < 0) return false;
  if (ib < 0) return true;
  return sa < sb || (sa == sb && ia < ib);
}

__device__ __forceinline__ void block_argmin(double &s, long long &i, double *ss, long long *si) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double s2 = __shfl_down_sync(0xffffffffu, s, o);
    long long i2 = __shfl_down_sync(0xffffffffu, i, o);
    if (cand_less(s2, i2, s, i)) {
      s = s2;
      i = i2;
    }
  }
  if (lane == 0) {
    ss[warp] = s;
    si[warp] = i;
  }
  __syncthreads();
  if (warp == 0) {
    s = lane < nwarp ? ss[lane] : 0.0;
    i = lane < nwarp ? si[lane] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double s2 = __shfl_down_sync(0xffffffffu, s, o);
      long long i2 = __shfl_down_sync(0xffffffffu, i, o);
      if (cand_less(s2, i2, s, i)) {
        s = s2;
        i = i2;
      }
    }
    if (lane == 0) {
      ss[0] = s;
      si[0] = i;
    }
  }
  __syncthreads();
  s = ss[0];
  i = si[0];
  __syncthreads();
}

// ---- peer-memory exchange helpers (PeerExchange, common.cuh) ----
__device__ __forceinline__ unsigned char *px_slot(const PeerExchange &px, int owner, int source) {
  return px.win[owner] + ((size_t)(px.epoch & 1u) * px.nranks + source) * px.slot_bytes;
}
__device__ __forceinline__ void px_store_flag(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// poll a flag in this rank's own window until it carries the batch epoch; a peer that never arrives (crashed rank)
// ends the wait after ~4 s with the error word set instead of hanging the GPU
__device__ __forceinline__ void px_wait_flag(const unsigned *p, unsigned epoch, int *err) {
  unsigned v;
  long long t0 = 0;
  unsigned spins = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v == epoch) return;
    if ((++spins & 0x3ffu) == 0) {
      if (t0 == 0)
        t0 = clock64();
      else if (clock64() - t0 > 8000000000LL) {
        *err = 1;
        return;
      }
      __nanosleep(64);
    }
  }
}

// stats[row] = [sum(dp-c), sum((dp-c)^2), count(dp), sum(di-c), sum((di-c)^2), count(di)], c = STAT_SHIFT, over the
// non-NaN entries only: MATLAB's normalize (run_test.m:40) computes its mean and std with 'omitnan', so the NaN column
// of one zero-norm DB signature (processSC.m:15-20) stays NaN and every other candidate is still ranked.
__global__ void __launch_bounds__(FUSE_THREADS)
row_stats_kernel(const float *__restrict__ d_p, const float *__restrict__ d_i, int n, int ldd,
                 double *__restrict__ stats, const PeerExchange px) {
  __shared__ double scratch[STATS_W * 32];
  const int row = blockIdx.x;
  const float *p = d_p + (size_t)row * ldd, *q = d_i + (size_t)row * ldd;
  double v[STATS_W] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  // (the loads of STATS_UNROLL elements per channel are issued together; the sums run in the same order as a plain loop)
  for (int j0 = threadIdx.x; j0 < n; j0 += FUSE_THREADS * STATS_UNROLL) {
    float pv[STATS_UNROLL], qv[STATS_UNROLL];
#pragma unroll
    for (int u = 0; u < STATS_UNROLL; u++) {
      const int j = j0 + u * FUSE_THREADS;
      pv[u] = j < n ? p[j] : 0.0f;
      qv[u] = j < n ? q[j] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < STATS_UNROLL; u++) {
      if (j0 + u * FUSE_THREADS >= n) break;
      const float pf = pv[u], qf = qv[u];
      const double a = (double)pf - STAT_SHIFT, b = (double)qf - STAT_SHIFT;
      if (pf == pf) {
        v[0] += a;
        v[1] += a * a;
        v[2] += 1.0;
      }
      if (qf == qf) {
        v[3] += b;
        v[4] += b * b;
        v[5] += 1.0;
      }
    }
  }
  block_sum_d<STATS_W>(v, scratch);
  if (threadIdx.x < STATS_W) stats[(size_t)row * STATS_W + threadIdx.x] = v[threadIdx.x];
  // sharded database: this shard's partial statistics of the row go straight into every rank's window over NVLink
  // (its own included); the consumer is fuse_topk_kernel on each rank
  if (px.win[0] && threadIdx.x < px.nranks) {
    unsigned char *slot = px_slot(px, threadIdx.x, px.rank);
    double *dst = reinterpret_cast<double *>(slot + PX_OFF_STATS) + (size_t)row * STATS_W;
#pragma unroll
    for (int t = 0; t < STATS_W; t++) dst[t] = v[t];
    __threadfence_system();
    px_store_flag(reinterpret_cast<unsigned *>(slot + PX_OFF_SFLAG) + row, px.epoch);
  }
}

// sharded database: the k candidates of this shard for `row` (written by thread 0 to the local lists [4][m][k] = idx |
// score | d_p | d_i) go into every rank's window; the consumer is topk_merge_px_kernel on each rank.  Called by ALL
// threads of the CTA: one (rank, array, entry) store per thread, fence, barrier, then one flag store per rank.
__device__ __forceinline__ void px_publish_row(const PeerExchange &px, int row, int m, int k, const int64_t *idx,
                                               const double *score, const double *dp_at, const double *di_at) {
  __syncthreads();   // thread 0's local lists are complete (block-scope visibility)
  const size_t mk = (size_t)m * k, o = (size_t)row * k;
  for (int e = threadIdx.x; e < px.nranks * 4 * k; e += blockDim.x) {
    const int r = e / (4 * k), a = (e / k) & 3, t = e % k;
    long long *dst = reinterpret_cast<long long *>(px_slot(px, r, px.rank) + PX_OFF_LISTS) + (size_t)a * mk + o + t;
    const long long v = a == 0 ? (long long)idx[o + t]
                               : __double_as_longlong(a == 1 ? score[o + t] : a == 2 ? dp_at[o + t] : di_at[o + t]);
    *dst = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < px.nranks)
    px_store_flag(reinterpret_cast<unsigned *>(px_slot(px, threadIdx.x, px.rank) + PX_OFF_LFLAG) + row, px.epoch);
}

// KL: length of the per-thread candidate list (1 for top-1, 8 for k <= 8, 0 = no lists: one row scan per selected entry)
constexpr int TOPK_CAND_CAP = 96;   // entries at or below the selection bound that the single-scan path can hold

template <int KL>
__global__ void __launch_bounds__(FUSE_THREADS)
fuse_topk_kernel(const float *__restrict__ d_p, const float *__restrict__ d_i, int n, int ldd,
                 const double *__restrict__ gstats, long long n_global, long long q_row0,
                 long long db_row0, int mask_width, double p_weight, int k, int64_t *__restrict__ idx,
                 double *__restrict__ score, double *__restrict__ dp_at, double *__restrict__ di_at,
                 const PeerExchange px) {
  __shared__ double ss[32];
  __shared__ long long si[32];
  __shared__ double cand_f[TOPK_CAND_CAP];
  __shared__ long long cand_j[TOPK_CAND_CAP];
  __shared__ int cand_n;
  __shared__ double gsum[STATS_W];
  const int row = blockIdx.x;
  const float *p = d_p + (size_t)row * ldd, *q = d_i + (size_t)row * ldd;
  const double *st = gstats + (size_t)row * STATS_W;
  if (px.win[0]) {
    // sharded database: the row statistics over the WHOLE database (run_test.m:40) = the sum of the shards' partial
    // statistics, which their row_stats_kernel wrote into this rank's window; summed in rank order, so every rank
    // gets the same bits
    if (threadIdx.x < px.nranks)
      px_wait_flag(reinterpret_cast<const unsigned *>(px_slot(px, px.rank, threadIdx.x) + PX_OFF_SFLAG) + row, px.epoch, px.err);
    __syncthreads();
    if (threadIdx.x < STATS_W) {
      double a = 0.0;
      for (int s = 0; s < px.nranks; s++)
        a += reinterpret_cast<const volatile double *>(px_slot(px, px.rank, s) + PX_OFF_STATS)[(size_t)row * STATS_W + threadIdx.x];
      gsum[threadIdx.x] = a;
    }
    __syncthreads();
    st = gsum;
  }
  const double Np = st[2], Ni = st[5];   // non-NaN entries of the whole (global) row, per channel
  const double mu_p = STAT_SHIFT + st[0] / Np, mu_i = STAT_SHIFT + st[3] / Ni;
  const double sd_p = sqrt((st[1] - st[0] * st[0] / Np) / (Np - 1.0));
  const double sd_i = sqrt((st[4] - st[3] * st[3] / Ni) / (Ni - 1.0));
  const long long qg = q_row0 + row;
  double last_s = 0.0;
  long long last_i = -1;  // nothing selected yet
  // The two fp64 divisions per element are only needed to order the few entries around the k smallest.  A linear
  // form a_p p + a_i q + c0 (two FMAs, |error| ~ 1e-13) is used to SELECT: every entry of the exact top-k lies at or
  // below the k-th smallest linear value + margin, and the exact expression of run_test.m:40-46 is evaluated only for
  // those.  Rows with fewer than k finite unmasked entries, or with degenerate statistics, take the plain loop (bound
  // = +inf admits everything, masked entries included).
  const double a_p = p_weight / sd_p, a_i = 1.0 / sd_i, c0 = -(a_p * mu_p + a_i * mu_i);
  const bool lin_ok = isfinite(a_p) && isfinite(a_i) && isfinite(c0);
  double bound = INFINITY;
  if (lin_ok && KL > 0) {
    // ONE scan: every thread keeps the KL smallest (value, index) of its elements in registers (sorted), then k rounds
    // of block arg-min over the list heads give the k-th smallest of the row
    constexpr int KLL = KL > 0 ? KL : 1;
    double lv[KLL];
    long long lj[KLL];
#pragma unroll
    for (int t = 0; t < KLL; t++) {
      lv[t] = 0.0;
      lj[t] = -1;
    }
    // (the loads of FUSE_UNROLL elements are issued together: a thread has only ~20 elements of a row, and one load
    // in flight at a time made the row scan a chain of memory latencies)
    for (int j0 = threadIdx.x; j0 < n; j0 += FUSE_THREADS * FUSE_UNROLL) {
      float pv[FUSE_UNROLL], qv[FUSE_UNROLL];
#pragma unroll
      for (int u = 0; u < FUSE_UNROLL; u++) {
        const int j = j0 + u * FUSE_THREADS;
        pv[u] = j < n ? p[j] : 0.0f;
        qv[u] = j < n ? q[j] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < FUSE_UNROLL; u++) {
        const int j = j0 + u * FUSE_THREADS;
        if (j >= n) break;
        const long long jg = db_row0 + j;
        long long dist = qg - jg;
        if (dist < 0) dist = -dist;
        if (dist < (long long)mask_width) continue;
        const double fl = fma(a_p, (double)pv[u], fma(a_i, (double)qv[u], c0));
        if (!isfinite(fl)) continue;
        if (cand_less(fl, jg, lv[KLL - 1], lj[KLL - 1])) {
          lv[KLL - 1] = fl;
          lj[KLL - 1] = jg;
#pragma unroll
          for (int t = KLL - 1; t > 0; t--)
            if (cand_less(lv[t], lj[t], lv[t - 1], lj[t - 1])) {
              const double tv = lv[t];
              const long long tj = lj[t];
              lv[t] = lv[t - 1];
              lj[t] = lj[t - 1];
              lv[t - 1] = tv;
              lj[t - 1] = tj;
            }
        }
      }
    }
    double sel_s = 0.0;
    bool enough = true;
    for (int r = 0; r < k; r++) {
      double bs = lv[0];
      long long bi = lj[0];
      block_argmin(bs, bi, ss, si);
      if (bi < 0) {
        enough = false;
        break;
      }
      if (lj[0] == bi) {   // the winner pops its head
#pragma unroll
        for (int t = 0; t < KLL - 1; t++) {
          lv[t] = lv[t + 1];
          lj[t] = lj[t + 1];
        }
        lj[KLL - 1] = -1;
      }
      sel_s = bs;
    }
    if (enough) bound = sel_s + 1e-9 * (1.0 + fabs(sel_s));
    if (enough) {
      // second scan: the (typically exactly k) entries at or below the bound, evaluated exactly and sorted by one thread
      if (threadIdx.x == 0) cand_n = 0;
      __syncthreads();
      for (int j0 = threadIdx.x; j0 < n; j0 += FUSE_THREADS * FUSE_UNROLL) {
        float pv[FUSE_UNROLL], qv[FUSE_UNROLL];
#pragma unroll
        for (int u = 0; u < FUSE_UNROLL; u++) {
          const int j = j0 + u * FUSE_THREADS;
          pv[u] = j < n ? p[j] : 0.0f;
          qv[u] = j < n ? q[j] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < FUSE_UNROLL; u++) {
          const int j = j0 + u * FUSE_THREADS;
          if (j >= n) break;
          const long long jg = db_row0 + j;
          long long dist = qg - jg;
          if (dist < 0) dist = -dist;
          if (dist < (long long)mask_width) continue;
          if (!(fma(a_p, (double)pv[u], fma(a_i, (double)qv[u], c0)) <= bound)) continue;
          const double f = p_weight * (((double)pv[u] - mu_p) / sd_p) + ((double)qv[u] - mu_i) / sd_i;   // run_test.m:40
          if (f != f) continue;
          const int slot = atomicAdd(&cand_n, 1);
          if (slot < TOPK_CAND_CAP) {
            cand_f[slot] = f;
            cand_j[slot] = jg;
          }
        }
      }
      __syncthreads();
      const int cn = cand_n;
      if (cn <= TOPK_CAND_CAP) {
        if (threadIdx.x == 0) {
          for (int a = 1; a < cn; a++) {   // insertion sort by (score, global index)
            const double f = cand_f[a];
            const long long jg = cand_j[a];
            int b = a - 1;
            while (b >= 0 && cand_less(f, jg, cand_f[b], cand_j[b])) {
              cand_f[b + 1] = cand_f[b];
              cand_j[b + 1] = cand_j[b];
              b--;
            }
            cand_f[b + 1] = f;
            cand_j[b + 1] = jg;
          }
          for (int r = 0; r < k; r++) {
            const size_t o = (size_t)row * k + r;
            const bool have = r < cn;
            const long long bi = have ? cand_j[r] : -1;
            idx[o] = bi;
            score[o] = have ? cand_f[r] : NAN;
            if (dp_at) dp_at[o] = have ? (double)p[bi - db_row0] : NAN;
            if (di_at) di_at[o] = have ? (double)q[bi - db_row0] : NAN;
          }
        }
        if (px.win[0]) px_publish_row(px, row, (int)gridDim.x, k, idx, score, dp_at, di_at);
        return;
      }
      // more near-ties than the list holds: fall through to the scan-per-entry loop below with the bound as filter
    }
  } else if (lin_ok) {
    double sel_s = 0.0;
    long long sel_i = -1;
    bool enough = true;
    for (int r = 0; r < k; r++) {
      double bs = 0.0;
      long long bi = -1;
      for (int j = threadIdx.x; j < n; j += FUSE_THREADS) {
        const long long jg = db_row0 + j;
        long long dist = qg - jg;
        if (dist < 0) dist = -dist;
        if (dist < (long long)mask_width) continue;
        const double fl = fma(a_p, (double)p[j], fma(a_i, (double)q[j], c0));
        if (!isfinite(fl)) continue;
        if (sel_i >= 0 && !cand_less(sel_s, sel_i, fl, jg)) continue;  // selected in an earlier pass
        if (cand_less(fl, jg, bs, bi)) {
          bs = fl;
          bi = jg;
        }
      }
      block_argmin(bs, bi, ss, si);
      if (bi < 0) {
        enough = false;
        break;
      }
      sel_s = bs;
      sel_i = bi;
    }
    if (enough) bound = sel_s + 1e-9 * (1.0 + fabs(sel_s));
  }
  for (int r = 0; r < k; r++) {
    double bs = 0.0;
    long long bi = -1;
    const bool filtered = bound < INFINITY;
    for (int j = threadIdx.x; j < n; j += FUSE_THREADS) {
      const long long jg = db_row0 + j;
      if (filtered) {
        // an unmasked entry above the bound cannot be among the k smallest; masked entries (+inf) lose against the
        // k finite entries that are known to exist
        long long dist0 = qg - jg;
        if (dist0 < 0) dist0 = -dist0;
        if (dist0 < (long long)mask_width) continue;
        if (!(fma(a_p, (double)p[j], fma(a_i, (double)q[j], c0)) <= bound)) continue;
      }
      double f = p_weight * (((double)p[j] - mu_p) / sd_p) + ((double)q[j] - mu_i) / sd_i;
      long long dist = qg - jg;
      if (dist < 0) dist = -dist;
      if (dist < (long long)mask_width) f = INFINITY;  // run_test.m:47-53
      if (f != f) continue;                              // min skips NaN (run_test.m:57)
      if (last_i >= 0 && !cand_less(last_s, last_i, f, jg)) continue;  // already emitted
      if (cand_less(f, jg, bs, bi)) {
        bs = f;
        bi = jg;
      }
    }
    block_argmin(bs, bi, ss, si);
    if (threadIdx.x == 0) {
      const size_t o = (size_t)row * k + r;
      idx[o] = bi;
      score[o] = bi >= 0 ? bs : NAN;
      if (dp_at) dp_at[o] = bi >= 0 ? (double)p[bi - db_row0] : NAN;
      if (di_at) di_at[o] = bi >= 0 ? (double)q[bi - db_row0] : NAN;
    }
    if (bi < 0) {  // exhausted: fill the rest
      if (threadIdx.x == 0)
        for (int r2 = r + 1; r2 < k; r2++) {
          const size_t o = (size_t)row * k + r2;
          idx[o] = -1;
          score[o] = NAN;
          if (dp_at) dp_at[o] = NAN;
          if (di_at) di_at[o] = NAN;
        }
      break;
    }
    last_s = bs;
    last_i = bi;
  }
  if (px.win[0]) px_publish_row(px, row, (int)gridDim.x, k, idx, score, dp_at, di_at);
}

// run_test.m:38-57 on caller-supplied fp64 matrices, two-pass statistics over the non-NaN entries like MATLAB normalize.
__global__ void __launch_bounds__(FUSE_THREADS)
fuse_top1_f64_kernel(const double *__restrict__ d_p, const double *__restrict__ d_i, int n,
                     int mask_width, double p_weight, int32_t *__restrict__ idx,
                     double *__restrict__ score) {
  __shared__ double scratch[4 * 32];
  __shared__ double ss[32];
  __shared__ long long si[32];
  const int row = blockIdx.x;
  const double *p = d_p + (size_t)row * n, *q = d_i + (size_t)row * n;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int j = threadIdx.x; j < n; j += FUSE_THREADS) {
    const double a = p[j], b = q[j];
    if (a == a) {
      s[0] += a;
      s[2] += 1.0;
    }
    if (b == b) {
      s[1] += b;
      s[3] += 1.0;
    }
  }
  block_sum_d<4>(s, scratch);
  const double mu_p = s[0] / s[2], mu_i = s[1] / s[3];
  double v[2] = {0.0, 0.0};
  for (int j = threadIdx.x; j < n; j += FUSE_THREADS) {
    const double a = p[j] - mu_p, b = q[j] - mu_i;
    if (a == a) v[0] += a * a;
    if (b == b) v[1] += b * b;
  }
  block_sum_d<2>(v, scratch);
  const double sd_p = sqrt(v[0] / (s[2] - 1.0)), sd_i = sqrt(v[1] / (s[3] - 1.0));
  double bs = 0.0;
  long long bi = -1;
  for (int j = threadIdx.x; j < n; j += FUSE_THREADS) {
    double f = p_weight * ((p[j] - mu_p) / sd_p) + (q[j] - mu_i) / sd_i;
    int dist = row - j;
    if (dist < 0) dist = -dist;
    if (dist < mask_width) f = INFINITY;
    if (f != f) continue;
    if (cand_less(f, j, bs, bi)) {
      bs = f;
      bi = j;
    }
  }
  block_argmin(bs, bi, ss, si);
  if (threadIdx.x == 0) {
    idx[row] = bi >= 0 ? (int32_t)bi : 0;  // MATLAB min of an all-NaN row returns index 1
    score[row] = bi >= 0 ? bs : NAN;
  }
}

__global__ void f32_to_f64_kernel(const float *__restrict__ src, int rows, int cols, int ld,
                                  double *__restrict__ dst) {
  const size_t total = (size_t)rows * cols;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    size_t r = t / cols, c = t - r * cols;
    dst[t] = (double)src[r * ld + c];
  }
}

}  // namespace

static PeerExchange no_exchange() {
  PeerExchange px;
  memset(&px, 0, sizeof(px));
  return px;
}

cudaError_t launch_row_stats(const float *d_p, const float *d_i, int m, int n, int ldd, double *stats,
                             cudaStream_t st, int64_t *launches, const PeerExchange *px) {
  if (m <= 0) return cudaSuccess;
  row_stats_kernel<<<m, FUSE_THREADS, 0, st>>>(d_p, d_i, n, ldd, stats, px ? *px : no_exchange());
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_fuse_topk(const float *d_p, const float *d_i, int m, int n, int ldd,
                             const double *global_stats, int64_t n_global, int64_t q_row0,
                             int64_t db_row0, int mask_width, double p_weight, int k, int64_t *idx,
                             double *score, double *dp_at, double *di_at, cudaStream_t st,
                             int64_t *launches, const PeerExchange *px) {
  if (m <= 0) return cudaSuccess;
  auto kern = k == 1 ? fuse_topk_kernel<1> : k <= 8 ? fuse_topk_kernel<8> : fuse_topk_kernel<0>;
  kern<<<m, FUSE_THREADS, 0, st>>>(d_p, d_i, n, ldd, global_stats, n_global, q_row0, db_row0, mask_width, p_weight, k, idx,
                                   score, dp_at, di_at, px ? *px : no_exchange());
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_fuse_top1_f64(const double *d_p, const double *d_i, int m, int n, int mask_width,
                                 double p_weight, int32_t *idx, double *score, cudaStream_t st,
                                 int64_t *launches) {
  if (m <= 0) return cudaSuccess;
  fuse_top1_f64_kernel<<<m, FUSE_THREADS, 0, st>>>(d_p, d_i, n, mask_width, p_weight, idx, score);
  if (launches) ++*launches;
  return cudaGetLastError();
}

// Ground-truth loop set of run_test.m:3-21: for every query position the nearest database position outside
// the temporal mask (first one on ties: the comparison at :13 is strict); a loop if closer than loop_diff.
// nearest[i] = 0-based j or -1, dist2[i] = squared distance.
__global__ void __launch_bounds__(FUSE_THREADS)
gt_loops_kernel(const double *__restrict__ gt1, const double *__restrict__ gt2, int n, int mask_width,
                int32_t *__restrict__ nearest, double *__restrict__ dist2) {
  __shared__ double ss[32];
  __shared__ long long si[32];
  const int i = blockIdx.x;
  const double x = gt1[3 * (size_t)i], y = gt1[3 * (size_t)i + 1], z = gt1[3 * (size_t)i + 2];
  double bs = 0.0;
  long long bi = -1;
  for (int j = threadIdx.x; j < n; j += FUSE_THREADS) {
    int dji = i - j;
    if (dji < 0) dji = -dji;
    if (dji < mask_width) continue;                              // run_test.m:8-10
    const double dx = x - gt2[3 * (size_t)j], dy = y - gt2[3 * (size_t)j + 1], dz = z - gt2[3 * (size_t)j + 2];
    const double d = (dx * dx + dy * dy) + dz * dz;              // diff*diff' (:11-12)
    if (d != d) continue;                                        // `min_diff > diff` is false for NaN
    if (cand_less(d, (long long)j, bs, bi)) {
      bs = d;
      bi = j;
    }
  }
  block_argmin(bs, bi, ss, si);
  if (threadIdx.x == 0) {
    nearest[i] = (int32_t)bi;
    dist2[i] = bi >= 0 ? bs : INFINITY;
  }
}

// Merge of R per-shard top-k lists (each ascending by (score, index)) into the global top-k of every query: one
// thread per query walks the R list heads k times.  idx < 0 marks an exhausted list.
__global__ void topk_merge_kernel(const int64_t *__restrict__ idx, const double *__restrict__ score,
                                  const double *__restrict__ d_p, const double *__restrict__ d_i, int nshards, int m,
                                  int k, size_t shard_stride, int64_t *__restrict__ out_idx,
                                  double *__restrict__ out_score, double *__restrict__ out_d_p,
                                  double *__restrict__ out_d_i) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  int pos[16];   // list heads (at most 16 shards)
  for (int s = 0; s < 16; s++) pos[s] = 0;
  for (int r = 0; r < k; r++) {
    int best = -1;
    size_t bo = 0;
    for (int s = 0; s < nshards; s++) {
      if (pos[s] >= k) continue;
      const size_t o = (size_t)s * shard_stride + (size_t)q * k + pos[s];
      if (idx[o] < 0) continue;
      if (best < 0 || score[o] < score[bo] || (score[o] == score[bo] && idx[o] < idx[bo])) {
        best = s;
        bo = o;
      }
    }
    const size_t oo = (size_t)q * k + r;
    if (best < 0) {
      out_idx[oo] = -1;
      out_score[oo] = NAN;
      if (out_d_p) out_d_p[oo] = NAN;
      if (out_d_i) out_d_i[oo] = NAN;
    } else {
      out_idx[oo] = idx[bo];
      out_score[oo] = score[bo];
      if (out_d_p) out_d_p[oo] = d_p ? d_p[bo] : NAN;
      if (out_d_i) out_d_i[oo] = d_i ? d_i[bo] : NAN;
      pos[best]++;
    }
  }
}

// The merge for the peer-memory exchange: the R lists of query q are in this rank's own window, one slot per source
// rank; a thread first waits until every source has published its row.
__global__ void topk_merge_px_kernel(const PeerExchange px, int m, int k, int64_t *__restrict__ out_idx,
                                     double *__restrict__ out_score, double *__restrict__ out_d_p,
                                     double *__restrict__ out_d_i) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const size_t mk = (size_t)m * k;
  const volatile long long *li[PX_MAX_RANKS];
  const volatile double *ls[PX_MAX_RANKS];
  int pos[PX_MAX_RANKS];
  for (int s = 0; s < px.nranks; s++) {
    const unsigned char *slot = px_slot(px, px.rank, s);
    px_wait_flag(reinterpret_cast<const unsigned *>(slot + PX_OFF_LFLAG) + q, px.epoch, px.err);
    li[s] = reinterpret_cast<const volatile long long *>(slot + PX_OFF_LISTS);
    ls[s] = reinterpret_cast<const volatile double *>(slot + PX_OFF_LISTS) + mk;
    pos[s] = 0;
  }
  for (int r = 0; r < k; r++) {
    int best = -1;
    long long bidx = -1;
    double bsc = 0.0;
    for (int s = 0; s < px.nranks; s++) {
      if (pos[s] >= k) continue;
      const size_t o = (size_t)q * k + pos[s];
      const long long id = li[s][o];
      if (id < 0) continue;
      const double sc = ls[s][o];
      if (best < 0 || sc < bsc || (sc == bsc && id < bidx)) {
        best = s;
        bidx = id;
        bsc = sc;
      }
    }
    const size_t oo = (size_t)q * k + r;
    if (best < 0) {
      out_idx[oo] = -1;
      out_score[oo] = NAN;
      if (out_d_p) out_d_p[oo] = NAN;
      if (out_d_i) out_d_i[oo] = NAN;
    } else {
      const size_t o = (size_t)q * k + pos[best];
      out_idx[oo] = bidx;
      out_score[oo] = bsc;
      if (out_d_p) out_d_p[oo] = ls[best][mk + o];
      if (out_d_i) out_d_i[oo] = ls[best][2 * mk + o];
      pos[best]++;
    }
  }
}

cudaError_t launch_topk_merge_px(const PeerExchange &px, int m, int k, int64_t *out_idx, double *out_score,
                                 double *out_d_p, double *out_d_i, cudaStream_t st, int64_t *launches) {
  if (m <= 0) return cudaSuccess;
  topk_merge_px_kernel<<<(m + 63) / 64, 64, 0, st>>>(px, m, k, out_idx, out_score, out_d_p, out_d_i);
  if (launches) ++*launches;
  return cudaGetLastError();
}

// shard_stride: elements between the lists of consecutive shards in each of the four arrays (0 = m * k, i.e. four
// separate R x m x k arrays; the packed all-gather buffer of sharded.cu uses 4 * m * k)
cudaError_t launch_topk_merge(const int64_t *idx, const double *score, const double *d_p, const double *d_i,
                              int nshards, int m, int k, int64_t *out_idx, double *out_score, double *out_d_p,
                              double *out_d_i, cudaStream_t st, int64_t *launches, size_t shard_stride) {
  if (m <= 0) return cudaSuccess;
  if (nshards > 16) return cudaErrorInvalidValue;
  if (shard_stride == 0) shard_stride = (size_t)m * k;
  topk_merge_kernel<<<(m + 127) / 128, 128, 0, st>>>(idx, score, d_p, d_i, nshards, m, k, shard_stride, out_idx, out_score,
                                                     out_d_p, out_d_i);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_gt_loops(const double *gt1, int m, const double *gt2, int n, int mask_width, int32_t *nearest,
                            double *dist2, cudaStream_t st, int64_t *launches) {
  if (m <= 0) return cudaSuccess;
  gt_loops_kernel<<<m, FUSE_THREADS, 0, st>>>(gt1, gt2, n, mask_width, nearest, dist2);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_f32_to_f64(const float *src, int rows, int cols, int ld, double *dst,
                              cudaStream_t st, int64_t *launches) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  size_t total = (size_t)rows * cols;
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  f32_to_f64_kernel<<<grid, 256, 0, st>>>(src, rows, cols, ld, dst);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
