// M2DP signature generation: test_m2dp.cpp:41-67 (PCA once, 4 sign variants) around
// M2DP::getSignature (M2DP.cpp:38-109), one persistent CTA per SM (1024 threads), the whole batch in
// one launch.
//
// Per scan: moments pass (HBM read) -> mean + 3x3 eigenproblem (pts_align.h).  Then, per sign variant, 64 planes x n
// points are projected and binned by (rho, theta) into a 64 x 128 count histogram (u32) and intensity-sum histogram
// (int32 when the sums are provably exact, fp64 otherwise) with shared-memory atomics; the sums are binarised against
// the float average intensity; the signature is the dominant left / right singular vector pair of each 64 x 128 matrix.
//
// Binning.  The bin of a (point, plane) pair is PROPOSED in fp32 and accepted only when both polar coordinates are
// further from a bin edge than 3x the fp32 error bound; everything else takes the reference's fp64 expression
// (M2DP.cpp:56-63) on the fp64 point re-read from L2.  The proposal needs no arctangent: with 16 sectors of 22.5
// degrees the sector follows from the signs of (xp, yp), from |yp| > |xp| and from min/max against tan(22.5 deg), and
// the distance to the nearest sector edge is the distance of that ratio to {0, tan 22.5, 1}; the ring comes from
// sqrt(xp^2 + yp^2).  ~45 instructions per evaluation instead of ~100 with the polynomial angle.
//
// Variant sharing.  Variant v evaluates plane (p, q) with the vectors S_v xProj, S_v yProj, S_v = diag(dx, dy, dx dy)
// (test_m2dp.cpp:47-57).  For the azimuths p = 1, 2, 3 the float table of M2DP.cpp:4-34 is exactly mirror symmetric:
// S_{v+2} xProj[4-p][q] = -S_v xProj[p][q] and S_{v+2} yProj[4-p][q] = +S_v yProj[p][q] (checked on the host at launch),
// so variant v+2's rows for those planes are variant v's with the sectors mapped s -> (7 - s) mod 16 -- not a single
// evaluation is needed.  Variants are therefore processed in pairs (0, 2) and (1, 3): 64 + 16 planes per point and
// pair instead of 128.  Evaluations inside the guard band are deferred to a small queue and replayed in fp64 for each
// variant separately, so the result is bit-identical to processing the variants one by one.
//
// The degenerate plane p=2,q=0, whose projection vectors are exactly zero, puts every point into one of
// two bins depending on the signs of the zeros; it is handled with one warp-aggregated update per 32 points.
//
// SVD: only the dominant pair is needed (M2DP.cpp:96-103).  G = A A^T (64 x 64, exact in 64-bit integers: 2 x 2 blocks
// of the lower triangle per thread, lanes skewed along k so that any row assignment is free of bank conflicts), scaled
// by a power of two, squared four times (G^16), then four warps run the power iteration u <- G u / |G u| in fp64 from
// the all-ones vector until the update falls below 1e-14, and v = A^T u / sigma.  Both matrices are entrywise
// non-negative, so the iteration converges to the Perron pair, which fixes the sign the same way as the oracle
// (sum(u) >= 0; Eigen's own sign is unobservable).
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
constexpr int HB = M2DP_PQ * M2DP_SR;   // 8192 histogram bins
constexpr float M2_GUARD_R = 3e-5f;      // ring coordinate guard (see the binning loop)
constexpr int SVD_WARPS = 4;            // warps in the power iteration
constexpr int QCAP = 4096;              // deferred (point, plane) evaluations of a pass (8 bytes each, in the T area)
// the two SVD groups of a CTA work at the same time on their own workspaces: each takes one variant of a pair (count
// matrix, then binarised matrix), so that both are in the same phase -- the tensor-core squarings of one group do not
// sit in front of the dependent fp64 chains of the other's power iteration
constexpr int SVD_G0 = 512, SVD_G1 = M2_THREADS - SVD_G0;
constexpr int WS = M2DP_PQ * M2DP_PQ;   // doubles in a 64 x 64 workspace
constexpr float TAN22 = 0.41421356237f;
// cross-pair stash (global memory, per CTA): the p = 0 rows of variants 1 and 3, derived from those of variants 0 and 2
constexpr int P0_BINS = M2DP_NUM_Q * M2DP_SR;        // 2048 bins of the 16 planes with p = 0
constexpr int STASH_U32 = 4 * P0_BINS;               // [cnt slot 0 | cnt slot 1 | isum slot 0 | isum slot 1]
constexpr int STASH_MAX_CTAS = 256;

__constant__ double c_xproj[3 * M2DP_PQ];
__constant__ double c_yproj[3 * M2DP_PQ];
// fp32 proposal table: (x0, x1, x2, y0), (y1, y2, 0, 0) per plane; the plane index is warp-uniform, so the two
// 16-byte reads are uniform constant loads, not shared-memory traffic
__constant__ float4 c_tab4[2 * M2DP_PQ];

struct M2Smem {
  unsigned cnt[2][HB];          // count histograms of the two variants of a pair (A0 of the SVD)
  int isum[2][HB];              // exact mode: intensity sums in units of 2^emin; then, in place, the binarised matrix.
                                // inexact mode (one variant at a time): the 64 KB are HB fp64 sums
  double G[WS];                 // Gram matrix -> G^16 of SVD group 0 (swizzled 64 x 64, see gi())
  double T[WS];                 // its squaring workspace; during binning: the queue of deferred evaluations
                                // (SVD group 1 takes its two workspaces from the isum area, free once binarised)
  unsigned bits[2][M2DP_PQ * 4];   // the binarised intensity matrices of the two slots: 128-bit row masks
  double scratch[11 * 32];
  // per SVD group: power iteration state, the result vectors, the Gram scale
  // per SVD group, two parities: the unnormalised iterate, partial squared norms
  double ubuf[2][2][M2DP_PQ], nrm[2][2][SVD_WARPS];
  double uvec[2][M2DP_PQ], yv[2][M2DP_SR], red[2][SVD_WARPS], sig[2];
  double bc[16];
  int ibc[4];
  int qn;
  int prof_on;                       // sodso_debug_phase_profile: thread 0 sums clock64 deltas per phase
  long long prof_t;
  unsigned long long prof[16];
};
static_assert(sizeof(M2Smem) <= 227 * 1024, "shared memory");
static_assert(2 * WS * sizeof(double) <= sizeof(int) * 2 * HB, "SVD group 1's workspaces fit in the isum area");
static_assert(QCAP * sizeof(unsigned long long) <= WS * sizeof(double), "the queue fits in T");

// element (r, c) of a 64 x 64 fp64 workspace.  The column is XOR-swizzled by the row (4-double granules), so that the
// 8 rows x 4 columns fragment loads of the fp64 tensor-core squaring hit 32 different 8-byte words of one aligned 256 B
// span (conflict free) without padding the rows -- two workspaces are exactly the 64 KB of the isum area.
__device__ __forceinline__ int gi(int r, int c) { return r * M2DP_PQ + (c ^ ((r & 7) << 2)); }
__device__ __forceinline__ void group_sync(int g) {
  asm volatile("bar.sync %0, %1;" ::"r"(8 + g), "r"(g ? SVD_G1 : SVD_G0) : "memory");
}

// phases: 10 moments pass, 0 eigen-solve, 1 main binning pass, 2 twin-row copy, 3 binning of the p = 0 planes of the twin,
// 4 queue replay, 5 binarise, 6 Gram matrices, 7 squarings, 8 power iteration, 9 v = A^T u + output
#define M2_PROF(k)                                            \
  do {                                                        \
    if (S.prof_on && threadIdx.x == 0) {                      \
      const long long t_ = clock64();                         \
      S.prof[k] += (unsigned long long)(t_ - S.prof_t);       \
      S.prof_t = t_;                                          \
    }                                                         \
  } while (0)

struct ScanRef {
  const double *g;
  const float *gi;
  const double *bc;   // mean + eigenvectors (identity for pre-aligned input)
  int n;
  int identity;       // class contract (M2DP.h:18-20): the input is used as it is (signed zeros included)
  double S_res_inv, R_res_inv;
};

// aligned fp64 point i (pts_align.h:37-45), before the sign flips of a variant
__device__ __forceinline__ void aligned_point64(const ScanRef &R, int i, double &ax, double &ay, double &az) {
  if (R.identity) {
    ax = R.g[3 * (size_t)i + 0];
    ay = R.g[3 * (size_t)i + 1];
    az = R.g[3 * (size_t)i + 2];
    return;
  }
  pca_rotate(R.bc, R.g[3 * (size_t)i + 0], R.g[3 * (size_t)i + 1], R.g[3 * (size_t)i + 2], ax, ay, az);
}

// sign pattern of variant var (test_m2dp.cpp:47-57: dx outer, dy inner, both in {-1, +1}); var < 0: identity
__device__ __forceinline__ void variant_signs(int var, double &dx, double &dy, double &dz) {
  dx = var < 0 ? 1.0 : ((var >> 1) ? 1.0 : -1.0);
  dy = var < 0 ? 1.0 : ((var & 1) ? 1.0 : -1.0);
  dz = dx * dy;
}

// M2DP.cpp:56-68 in fp64 for point i under variant var: flat histogram index or -1.  xp, yp and the ring are the
// reference's own expressions.  The sector floor((atan2(yp, xp) + pi) * 16 / 2pi) is first decided without the
// arctangent, by the comparisons of propose_bin carried out in fp64: the reference's angle is good to ~1e-15 sectors,
// so outside a 1e-13 band around the sector edges (in units of min / max) the comparisons give its floor; inside the
// band the reference's expression itself is evaluated.
__device__ __noinline__ int m2dp_bin_exact(const ScanRef &R, int i, int pq, int var) {
  const double PI = 3.14159265358979323846;
  double px, py, pz;
  aligned_point64(R, i, px, py, pz);
  if (!R.identity) {   // test_m2dp.cpp:49-53 (no multiplication at all for the class contract: signed zeros survive)
    double dx, dy, dz;
    variant_signs(var, dx, dy, dz);
    px = dx * px;
    py = dy * py;
    pz = dz * pz;
  }
  const double xp = (c_xproj[3 * pq] * px + c_xproj[3 * pq + 1] * py) + c_xproj[3 * pq + 2] * pz;
  const double yp = (c_yproj[3 * pq] * px + c_yproj[3 * pq + 1] * py) + c_yproj[3 * pq + 2] * pz;
  const double rad = sqrt(xp * xp + yp * yp) * R.R_res_inv;
  if (!(rad < (double)M2DP_SR)) return -1;     // (also NaN)
  const int ri = (int)floor(rad);
  const double ax = fabs(xp), ay = fabs(yp);
  const double mx = fmax(ax, ay), mn = fmin(ax, ay);
  const double T = 0.41421356237309503;        // tan(pi / 8)
  const double tm = T * mx, band = 1e-13 * mx;
  int si;
  if (mn > band && fabs(mn - tm) > band && mx - mn > band && mx < 1e300) {
    int k = mn > tm ? 1 : 0;
    k = ay > ax ? 3 - k : k;
    const bool sx = xp < 0.0, sy = yp < 0.0;
    si = (sy ? 0 : 8) + (k ^ (sx != sy ? 7 : 0));
  } else {
    const double ang = (atan2(yp, xp) + PI) * R.S_res_inv;
    if (!(ang < 32.0)) return -1;
    si = (int)floor(ang);
  }
  const int sr = ri * M2DP_NUM_S + si;  // M2DP.cpp:63
  return sr < M2DP_SR ? pq * M2DP_SR + sr : -1;  // M2DP.cpp:66 (si == 16 aliases into ring ri+1)
}

// the reference's bin of (point i, plane pq) under variant var, added to histogram slot `slot`
template <bool EXACT>
__device__ __forceinline__ void add_exact(M2Smem &S, const ScanRef &R, int i, int pq, int var, int slot, float iscale) {
  const int idx = m2dp_bin_exact(R, i, pq, var);
  if (idx < 0) return;
  atomicAdd(&S.cnt[slot][idx], 1u);
  if (EXACT)
    atomicAdd(&S.isum[slot][idx], (int)(R.gi[i] * iscale));
  else
    atomicAdd(reinterpret_cast<double *>(S.isum) + idx, (double)R.gi[i]);
}

// plane of variant a + 2 that is the mirror image of plane pq (p >= 1) of variant a
__device__ __forceinline__ int twin_plane(int pq) {
  return (M2DP_NUM_P - pq / M2DP_NUM_Q) * M2DP_NUM_Q + pq % M2DP_NUM_Q;
}

// queue entry of a guard-band evaluation: point i, plane pq, histogram slot, and whether the twin plane of the
// other variant of the pair (slot 1) has to be evaluated as well
__device__ __forceinline__ unsigned long long queue_entry(int i, int pq, int slot, bool twin) {
  return ((unsigned long long)(unsigned)i << 8) | (unsigned)(pq | (slot << 6) | (twin ? 0x80 : 0));
}
// stash != nullptr (first pair of a scan): an entry of a p = 0 plane of variant v also stands for the same plane of
// variant v ^ 1 (second pair), whose p = 0 rows are kept in the stash.
// An entry has up to three jobs -- role 0: its own variant, 1: the twin plane of the pair's other variant, 2: the
// stash -- each one fp64 evaluation (a few thousand clocks of dependent latency); role < 0 runs all of them in turn.
template <bool EXACT>
__device__ __forceinline__ void run_entry(M2Smem &S, const ScanRef &R, unsigned long long w, int var_slot0,
                                          int var_slot1, float iscale, unsigned *stash = nullptr, int role = -1) {
  const int i = (int)(w >> 8), pq = (int)(w & 0x3fu), slot = (int)((w >> 6) & 1u);
  const int var = slot ? var_slot1 : var_slot0;
  if (role < 0 || role == 0) add_exact<EXACT>(S, R, i, pq, var, slot, iscale);
  if ((role < 0 || role == 1) && (w & 0x80u)) add_exact<EXACT>(S, R, i, twin_plane(pq), var_slot1, 1, iscale);
  if ((role < 0 || role == 2) && EXACT && stash && pq < M2DP_NUM_Q) {
    const int idx = m2dp_bin_exact(R, i, pq, var ^ 1);
    if (idx >= 0) {
      atomicAdd(&stash[slot * P0_BINS + idx], 1u);
      atomicAdd(reinterpret_cast<int *>(stash) + 2 * P0_BINS + slot * P0_BINS + idx, (int)(R.gi[i] * iscale));
    }
  }
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// fp32 proposal of the (ring, sector) bin of projected coordinates (xp, yp): sr = ri * 16 + si, `inside` = within the
// 8 rings (M2DP.cpp:66), and `ok` = the proposal is certainly what the fp64 expression gives.
//   error budget: xp, yp good to delta = 2e-5 m (|p| < 128 m, checked by the caller through r2_lim).  With
//   mx = max(|xp|, |yp|), mn = min: a = mn / mx is good to 2 delta / mx + 3e-7, and the three decisions (sign pattern,
//   |yp| > |xp|, a > tan 22.5) are right whenever a is further than that from {0, tan 22.5, 1}; accepted with a 3x
//   margin.  Ring: rho * R_f good to 4e-6 + 1e-6 rings, accepted 3e-5 away from an edge.
// Sector of the fp64 expression: si = floor(8 + phi / 22.5 deg), phi = atan2(yp, xp).  With k = floor(alpha / 22.5)
// of the first-quadrant angle alpha = atan(|yp| / |xp|):  phi = alpha -> 8 + k;  180 - alpha -> 15 - k;
// -alpha -> 7 - k;  alpha - 180 -> k   (never on an edge: those evaluations are not safe).
__device__ __forceinline__ int propose_bin(float xp, float yp, float R_f, float r2_lim, bool &ok, bool &inside) {
  const float ax = fabsf(xp), ay = fabsf(yp);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float r2 = __fmaf_rn(xp, xp, yp * yp);
  const float rf = sqrt_approx(r2) * R_f;
  // floor(rf) for 0 <= rf < 2^22 without a conversion instruction: (rf - 0.5) + 1.5 * 2^23 rounds to the nearest
  // integer, which is floor(rf) unless rf is within rounding of an integer -- and those are inside the ring guard
  const float tr = (rf - 0.5f) + 12582912.0f;
  const float fr = tr - 12582912.0f;
  const int ri = (int)(__float_as_uint(tr) & 0x3fffffu);
  const float er = fabsf((rf - fr) - 0.5f);
  const float inv = rcp_approx(mx);
  const float a = mn * inv;
  const float d = fminf(fminf(a, fabsf(a - TAN22)), 1.0f - a);
  const float eg = __fmaf_rn(1.2e-4f, inv, 1e-6f);
  const bool hi = a > TAN22, swap = ay > ax;
  int k = hi ? 1 : 0;
  k = swap ? 3 - k : k;
  const unsigned sx = __float_as_uint(xp) >> 31, sy = __float_as_uint(yp) >> 31;   // 1 = negative
  const int m7 = (sx != sy) ? 7 : 0;           // upper half plane: xp < 0 mirrors k; lower half: xp > 0 mirrors
  const int sector = (int)((sy ^ 1u) << 3) + (k ^ m7);
  // (a small radius needs no test of its own: eg grows like 1 / mx and d never exceeds 0.293)
  ok = d > eg && er < 0.5f - M2_GUARD_R && r2 < r2_lim;
  inside = rf < (float)M2DP_NUM_R;
  return ri * M2DP_NUM_S + sector;
}

// Y = X X for a symmetric 64 x 64 fp64 matrix (swizzled, gi()) on the fp64 tensor cores (mma.sync.m8n8k4.f64: 256
// multiply-adds per warp instruction instead of 32), by the warps of one SVD group.  The B fragment X[k][j] is read
// as X[j][k] (symmetry), so both operands are 8 x 4 row reads.  Fragments (PTX ISA, m8n8k4 .f64):
// A[lane / 4][lane % 4], B[lane % 4][lane / 4], C[lane / 4][2 (lane % 4) + {0, 1}].
// Y is symmetric too: only the 36 tiles (I, J <= I) of the 8 x 8 tile grid are computed, the others are mirrored.  A
// warp takes a run of up to three tiles of one tile row, which share the A fragment: 15 runs, 51 fragment loads per 36
// DMMAs instead of 72 (a squaring pair costs as many shared-memory wavefronts as tensor clocks: 8.1k -> 6.7k clocks).
// (2 x 2 tile blocks need fewer loads still, 44, but leave 14 warps with a 64-DMMA chain: 7.2k.)
__constant__ unsigned char c_sq_runs[15][3] = {   // (I, first J, tiles)
    {7, 0, 3}, {7, 3, 3}, {6, 0, 3}, {5, 0, 3}, {5, 3, 3}, {4, 0, 3}, {2, 0, 3}, {7, 6, 2},
    {6, 3, 2}, {6, 5, 2}, {4, 3, 2}, {3, 0, 2}, {3, 2, 2}, {1, 0, 2}, {0, 0, 1}};

template <int N>
__device__ __forceinline__ void sq_run(const double *X, double *Y, int I, int J0, int r, int c) {
  const int i0 = I * 8;
  // rows i0 + r and 8 J + r have (row & 7) == r: the swizzle of column k0 + c is ((k0 >> 2) ^ r) << 2 | c
  const double *pa = X + (i0 + r) * M2DP_PQ + c, *pb = X + (J0 * 8 + r) * M2DP_PQ + c;
  double acc[N][2];
#pragma unroll
  for (int t = 0; t < N; t++) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < M2DP_PQ; k0 += 4) {
    const int o = ((k0 >> 2) ^ r) << 2;
    const double a = pa[o];
#pragma unroll
    for (int t = 0; t < N; t++) {
      const double b = pb[t * 8 * M2DP_PQ + o];
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(acc[t][0]), "+d"(acc[t][1])
                   : "d"(a), "d"(b));
    }
  }
#pragma unroll
  for (int t = 0; t < N; t++) {
    const int J = J0 + t, j0 = J * 8;
    *reinterpret_cast<double2 *>(Y + gi(i0 + r, j0 + 2 * c)) = make_double2(acc[t][0], acc[t][1]);
    if (I != J) {
      Y[gi(j0 + 2 * c, i0 + r)] = acc[t][0];
      Y[gi(j0 + 2 * c + 1, i0 + r)] = acc[t][1];
    }
  }
}

__device__ __forceinline__ void sym_square64(const double *X, double *Y, int gwarp, int gwarps) {
  const int lane = threadIdx.x & 31;
  const int r = lane >> 2, c = lane & 3;
  for (int u = gwarp; u < 15; u += gwarps) {
    const int I = c_sq_runs[u][0], J0 = c_sq_runs[u][1], n = c_sq_runs[u][2];   // (warp-uniform)
    if (n == 3)
      sq_run<3>(X, Y, I, J0, r, c);
    else if (n == 2)
      sq_run<2>(X, Y, I, J0, r, c);
    else
      sq_run<1>(X, Y, I, J0, r, c);
  }
}

// Gram matrix G = A A^T of a 64 x 128 count matrix, exact in 64-bit integers: the 528 2 x 2 blocks (bi, bj <= bi) of the
// lower block triangle.  A job is one block (NSPLIT = 1) or one block restricted to the k-steps of one lane
// (leftover jobs, spread over a warp and reduced by shuffles).  Lane l walks k in 16-byte steps starting at step l, so
// the 8 lanes of a quarter-warp always hit 8 different 16-byte bank groups whatever rows they read (rows are 512 B
// apart); integer sums do not care about the order.
// ACC = unsigned when the scan has fewer than 65536 points (every entry of G is at most n * max count <= n^2 < 2^32:
// full-rate 32-bit multiply-adds), unsigned long long otherwise.
template <typename ACC>
__device__ __forceinline__ void gram_block(const unsigned *A, double *G, int blk, int lane, bool whole) {
  int bi = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
  while ((bi + 1) * (bi + 2) / 2 <= blk) bi++;
  while (bi * (bi + 1) / 2 > blk) bi--;
  const int bj = blk - bi * (bi + 1) / 2;
  const uint4 *r0 = reinterpret_cast<const uint4 *>(A + (2 * bi) * M2DP_SR);
  const uint4 *r1 = reinterpret_cast<const uint4 *>(A + (2 * bi + 1) * M2DP_SR);
  const uint4 *c0 = reinterpret_cast<const uint4 *>(A + (2 * bj) * M2DP_SR);
  const uint4 *c1 = reinterpret_cast<const uint4 *>(A + (2 * bj + 1) * M2DP_SR);
  ACC g00 = 0, g01 = 0, g10 = 0, g11 = 0;
  const int steps = whole ? M2DP_SR / 4 : 1;   // whole: every k-step; otherwise the lane's own step, summed over the warp
#pragma unroll 4
  for (int t = 0; t < steps; t++) {
    const int k4 = (t + lane) & (M2DP_SR / 4 - 1);
    const uint4 a0 = r0[k4], a1 = r1[k4], b0 = c0[k4], b1 = c1[k4];
    g00 += (ACC)a0.x * b0.x + (ACC)a0.y * b0.y + (ACC)a0.z * b0.z + (ACC)a0.w * b0.w;
    g01 += (ACC)a0.x * b1.x + (ACC)a0.y * b1.y + (ACC)a0.z * b1.z + (ACC)a0.w * b1.w;
    g10 += (ACC)a1.x * b0.x + (ACC)a1.y * b0.y + (ACC)a1.z * b0.z + (ACC)a1.w * b0.w;
    g11 += (ACC)a1.x * b1.x + (ACC)a1.y * b1.y + (ACC)a1.z * b1.z + (ACC)a1.w * b1.w;
  }
  if (!whole) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      g00 += __shfl_xor_sync(0xffffffffu, g00, o);
      g01 += __shfl_xor_sync(0xffffffffu, g01, o);
      g10 += __shfl_xor_sync(0xffffffffu, g10, o);
      g11 += __shfl_xor_sync(0xffffffffu, g11, o);
    }
    if (lane != 0) return;
  }
  const int i0 = 2 * bi, j0 = 2 * bj;
  G[gi(i0, j0)] = (double)g00;
  G[gi(j0, i0)] = (double)g00;
  G[gi(i0, j0 + 1)] = (double)g01;
  G[gi(j0 + 1, i0)] = (double)g01;
  G[gi(i0 + 1, j0)] = (double)g10;
  G[gi(j0, i0 + 1)] = (double)g10;
  G[gi(i0 + 1, j0 + 1)] = (double)g11;
  G[gi(j0 + 1, i0 + 1)] = (double)g11;
}

// the Gram matrices of nmat (1 or 2) count matrices by all threads of the CTA: one block per thread, the 32 blocks
// that are left over when nmat = 2 (1056 jobs, 1024 threads) one per warp
constexpr int GRAM_BLOCKS = 528;
template <typename ACC>
__device__ __forceinline__ void gram_counts_cta(const unsigned *A0, double *G0, const unsigned *A1, double *G1, int nmat) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int njobs = nmat * GRAM_BLOCKS;
  if (tid < njobs)
    gram_block<ACC>(tid < GRAM_BLOCKS ? A0 : A1, tid < GRAM_BLOCKS ? G0 : G1, tid % GRAM_BLOCKS, lane, true);
  const int left = M2_THREADS + warp;   // (warp-uniform)
  if (left < njobs) gram_block<ACC>(A1, G1, left - GRAM_BLOCKS, lane, false);
}

// The same for scans of fewer than 65536 points (32-bit sums are exact), with 4 x 4 blocks: the 136 blocks (qi, qj <= qi)
// of the lower block triangle, each split along k between the two lanes of a pair -- 272 jobs per matrix.  A 4 x 4 block
// needs 8 row loads per 16 products where the 2 x 2 block needs 4 per 4: the shared-memory wavefronts, which bound
// this phase, are halved.  Lane skew: the lanes of a quarter-warp read k-steps that differ mod 8 (conflict free).
__device__ __forceinline__ void gram_counts_cta_small(const unsigned *A0, double *G0, const unsigned *A1, double *G1,
                                                      int nmat) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid >= nmat * 272) return;   // (whole warps: 272 = 8.5 warps, 544 = 17 warps; partial warp: see the shuffles below)
  const int khalf = tid & 1, job = tid >> 1, m = job >= 136 ? 1 : 0, blk = job - 136 * m;
  const unsigned *A = m ? A1 : A0;
  double *G = m ? G1 : G0;
  int qi = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
  while ((qi + 1) * (qi + 2) / 2 <= blk) qi++;
  while (qi * (qi + 1) / 2 > blk) qi--;
  const int qj = blk - qi * (qi + 1) / 2;
  const uint4 *ra = reinterpret_cast<const uint4 *>(A + (4 * qi) * M2DP_SR);
  const uint4 *rb = reinterpret_cast<const uint4 *>(A + (4 * qj) * M2DP_SR);
  unsigned g[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) g[i][j] = 0u;
  const int skew = (lane >> 1) + 4 * khalf;
#pragma unroll 2
  for (int t = 0; t < M2DP_SR / 8; t++) {   // the 16 k-steps (of 4 columns) of this half
    const int k4 = khalf * (M2DP_SR / 8) + ((t + skew) & (M2DP_SR / 8 - 1));
    uint4 a[4];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = ra[i * (M2DP_SR / 4) + k4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint4 b = rb[j * (M2DP_SR / 4) + k4];
#pragma unroll
      for (int i = 0; i < 4; i++) g[i][j] += a[i].x * b.x + a[i].y * b.y + a[i].z * b.z + a[i].w * b.w;
    }
  }
  // the two halves of k; then lane 0 of the pair stores the block, lane 1 its transpose
  const unsigned pair_mask = 3u << (lane & ~1);
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) g[i][j] += __shfl_xor_sync(pair_mask, g[i][j], 1);
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gi_ = 4 * qi + i, gj_ = 4 * qj + j;
      if (khalf == 0)
        G[gi(gi_, gj_)] = (double)g[i][j];
      else
        G[gi(gj_, gi_)] = (double)g[i][j];
    }
}

// dominant singular pair of a 64 x 128 matrix, by the threads of SVD group g (named barriers).  A: the count matrix
// (u32; its Gram matrix is already in G, gram_counts_cta), or for BINARY the 128-bit row masks.  G / T: the group's
// 64 x 64 fp64 workspaces.  Writes [u (64), v (128)] to out.
//   G = A A^T exactly (64-bit integers), scaled by a power of two to trace ~ 1;  G^16 by four squarings;
//   power iteration with G^16 from the all-ones vector (Perron pair: entrywise non-negative);
//   sigma = |A^T u|, v = A^T u / sigma.
template <bool BINARY>
__device__ void dominant_pair(const unsigned *A, double *G, double *T, M2Smem &S, int g, double *out) {
  const int gsize = g ? SVD_G1 : SVD_G0, tid = threadIdx.x - (g ? SVD_G0 : 0), warp = tid >> 5, lane = tid & 31;
  const int gwarps = gsize >> 5;
  double *uvec = S.uvec[g], *yv = S.yv[g], *red = S.red[g];
  if (BINARY) {
    // 0 / 1 matrix: G[i][j] = popcount(row_i & row_j)
    for (int e = tid; e < M2DP_PQ * M2DP_PQ; e += gsize) {
      const int i = e >> 6, j = e & 63;
      const uint4 a = reinterpret_cast<const uint4 *>(A)[i], b = reinterpret_cast<const uint4 *>(A)[j];
      G[gi(i, j)] = (double)(__popc(a.x & b.x) + __popc(a.y & b.y) + __popc(a.z & b.z) + __popc(a.w & b.w));
    }
    group_sync(g);
  }
  if (g == 0) M2_PROF(6);
  // ---- scale to trace in [1, 2) (exact, power of two), so that G^16 neither overflows nor underflows
  if (warp == 0) {
    double tr = G[gi(lane, lane)] + G[gi(lane + 32, lane + 32)];
    tr = warp_sum(tr);
    tr = __shfl_sync(0xffffffffu, tr, 0);
    if (lane == 0) S.sig[g] = tr > 0.0 ? scalbn(1.0, -ilogb(tr)) : 0.0;
  }
  group_sync(g);
  {
    const double sc = S.sig[g];
    for (int e = tid; e < WS; e += gsize) G[e] *= sc;
  }
  group_sync(g);
  sym_square64(G, T, warp, gwarps);   // G^2
  group_sync(g);
  sym_square64(T, G, warp, gwarps);   // G^4
  group_sync(g);
  sym_square64(G, T, warp, gwarps);   // G^8
  group_sync(g);
  sym_square64(T, G, warp, gwarps);   // G^16
  if (tid < M2DP_PQ) S.ubuf[g][0][tid] = 0.125;  // all-ones / |.| (Perron start)
  if (tid < SVD_WARPS) S.nrm[g][0][tid] = 1.0 / SVD_WARPS;
  // CTA-wide (both groups call dominant_pair the same number of times): the two groups stay in the same phase, the
  // DMMA bursts of one do not sit in front of the dependent fp64 chains of the other's power iteration
  __syncthreads();
  if (g == 0) M2_PROF(7);
  // ---- power iteration with G^16: SVD_WARPS warps, 2 threads per row, ONE named barrier per step.  The iterate is kept
  // unnormalised (w_k, parity k & 1) together with the partial sums of |w_k|^2; the normalisation of step k is applied
  // by the readers in step k + 1 (u_k = w_k / |w_k|).  The barrier also ORs "my row of u moved by more than 2e-15 in
  // this step" (bar.red): all quiet = converged, u_k is the result (w_{k+1} has been computed for nothing).
  // A step is one long chain of dependent instructions (~2.4k clocks): tried and not faster, 8 threads per row on all
  // 16 warps with per-warp partial sums (3.9k: the SM-wide shared-memory and fp64 pipes pay for the redundant sums)
  // and the same with G^16 held in registers and the norm from each row group's own copy of w (2.8k).
  if (warp < SVD_WARPS) {
    const int t = warp * 32 + lane;          // 0..127
    const int row = t >> 1, half = t & 1;
    const int bar_n = SVD_WARPS * 32, bar_id = 10 + g;
    double u_prev = 0.125;                   // u_{k-1}[row]
    int iter = 0;
    for (; iter < 4000; iter++) {
      const int p = iter & 1;
      const double *w = S.ubuf[g][p];
      double nn = 0.0;
#pragma unroll
      for (int k = 0; k < SVD_WARPS; k++) nn += S.nrm[g][p][k];
      if (nn == 0.0) {   // zero matrix (uniform over the group)
        if (half == 0) uvec[row] = 0.0;
        break;
      }
      // 1 / sqrt(nn): fp32 seed + three Newton steps (rel. error 1e-7 -> 1e-14 -> 1e-28 -> rounding), instead of the
      // fp64 square root and division routines
      double inv = (double)rsqrtf((float)nn);
      {
        const double hn = 0.5 * nn;
        inv = inv * fma(-hn * inv, inv, 1.5);
        inv = inv * fma(-hn * inv, inv, 1.5);
        inv = inv * fma(-hn * inv, inv, 1.5);
      }
      const double u_cur = w[row] * inv;     // u_k[row]
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
      for (int k = 0; k < 32; k += 2) {
        const int c = half * 32 + k;
        acc0 = fma(G[gi(c, row)], w[c], acc0);   // G symmetric: column access is conflict-free
        acc1 = fma(G[gi(c + 1, row)], w[c + 1], acc1);
      }
      double acc = acc0 + acc1;
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc *= inv;                            // w_{k+1} = G^16 u_k
      const unsigned moved = iter == 0 || fabs(u_cur - u_prev) > 2e-15 ? 1u : 0u;
      u_prev = u_cur;
      const double sq = warp_sum(half == 0 ? acc * acc : 0.0);
      if (half == 0) S.ubuf[g][p ^ 1][row] = acc;
      if (lane == 0) S.nrm[g][p ^ 1][warp] = sq;
      unsigned any_moved;
      asm volatile(
          "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(any_moved)
          : "r"(moved), "r"(bar_id), "r"(bar_n)
          : "memory");
      if (!any_moved) {                      // |u_k - u_{k-1}| <= 2e-15 in every row
        if (half == 0) uvec[row] = u_cur;
        iter++;
        break;
      }
    }
    if (S.prof_on && t == 0) atomicAdd(&S.prof[11 + g], (unsigned long long)iter);   // steps per group
  }
  __syncthreads();   // (see above)
  if (g == 0) M2_PROF(8);
  // ---- y = A^T u (128), sigma = |y|: NP threads per column (64 / NP rows each), partial sums through the T area
  {
    constexpr int NP = 4;
    static_assert(NP * M2DP_SR <= SVD_G0 && NP * M2DP_SR <= SVD_G1, "threads of a group");
    if (tid < NP * M2DP_SR) {
      const int col = tid & (M2DP_SR - 1), part = tid >> 7;
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < M2DP_PQ / NP; r++) {
        const int rr = part * (M2DP_PQ / NP) + r;
        const double a = BINARY ? (double)((A[rr * 4 + (col >> 5)] >> (col & 31)) & 1u) : (double)A[rr * M2DP_SR + col];
        acc = fma(a, uvec[rr], acc);
      }
      T[part * M2DP_SR + col] = acc;
    }
    group_sync(g);
    if (tid < M2DP_SR) {
      double acc = 0.0;
#pragma unroll
      for (int part = 0; part < NP; part++) acc += T[part * M2DP_SR + tid];
      yv[tid] = acc;
      const double sq = warp_sum(acc * acc);
      if (lane == 0) red[warp] = sq;
    }
  }
  group_sync(g);
  // ---- outputs: u, v = y / sigma  (zero matrix: u = e_0, v = 0, the oracle's convention)
  if (tid < M2DP_SIG) {
    const double sigma = sqrt((red[0] + red[1]) + (red[2] + red[3]));
    double v;
    if (tid < M2DP_PQ)
      v = sigma == 0.0 ? (tid == 0 ? 1.0 : 0.0) : uvec[tid];
    else
      v = sigma == 0.0 ? 0.0 : yv[tid - M2DP_PQ] / sigma;
    out[tid] = v;
  }
  if (g == 0) M2_PROF(9);
}

// sector map of the mirrored twin plane: (xp, yp) -> (-xp, yp), i.e. phi -> 180 deg - phi
__device__ __forceinline__ int mirror_sr(int sr) { return (sr & ~15) | ((7 - (sr & 15)) & 15); }
// sector map between the p = 0 planes of variants v and v ^ 1: (xp, yp) -> (xp, -yp), i.e. phi -> -phi
__device__ __forceinline__ int mirror15_sr(int sr) { return sr ^ 15; }

// One block of 1024 points (FULL: all lanes hold a point) of a binning pass.  EXACT: integer intensity sums.
// Evaluations inside the guard band are pushed to the queue (replayed by the caller in fp64, replay_queue);
// twin_var >= 0: the pass is variant a of a pair, and planes p >= 1 also stand for their twins of variant twin_var.
template <int PQ0, int PQ1, bool EXACT, bool FULL>
__device__ __forceinline__ void bin_block(M2Smem &S, const ScanRef &R, int var, int slot, int twin_var, float R_f,
                                          float iscale, unsigned long long degen_mask, int degen_sr_pos,
                                          int degen_sr_neg, int i0, float coord_lim, bool degen_twin) {
  const int lane = threadIdx.x & 31;
  unsigned *hcnt = S.cnt[slot];
  int *hisum = S.isum[slot];
  double *hsum = reinterpret_cast<double *>(S.isum);   // inexact mode (single slot)
  unsigned long long *queue = reinterpret_cast<unsigned long long *>(S.T);
  double ddx, ddy, ddz;
  variant_signs(var, ddx, ddy, ddz);
  const float fdx = (float)ddx, fdy = (float)ddy, fdz = (float)ddz;
  const int n = R.n;
  {
    const int i = i0 + threadIdx.x;
    const bool act = FULL || i < n;
    float px = 0.0f, py = 0.0f, pz = 0.0f, it = 0.0f;
    double ax = 0.0, ay = 0.0, az = 0.0;
    if (act) {
      it = R.gi[i];
      aligned_point64(R, i, ax, ay, az);
      px = fdx * (float)ax;
      py = fdy * (float)ay;
      pz = fdz * (float)az;
    }
    const int iv = (int)(it * iscale);
    // the fp32 error bound assumes coordinates below coord_lim = 128 m (the staging crops at 45 m); 64 m when the
    // p = 0 rows also stand for the other pair's (see the host check in launch_m2dp_generate)
    const float r2_lim = fmaxf(fabsf(px), fmaxf(fabsf(py), fabsf(pz))) < coord_lim ? 1e10f : -1.0f;
    // planes whose projection vectors are exactly zero (p=2, q=0): xp = yp = -0 if all three
    // coordinates are negative, +0 otherwise, and every point lands in one of two bins -> warp-aggregated
    // degen_twin (exact mode, first variant of a pair): the same for variant twin_var into slot 1, under its own signs
    for (unsigned long long dm = degen_mask; dm; dm &= dm - 1) {
      const int pq = __ffsll((long long)dm) - 1;
      // signs of the flipped fp64 coordinates (a flip of +-0 flips the sign bit too; identity: untouched)
      const bool neg = act && (signbit(ax) != (ddx < 0.0)) && (signbit(ay) != (ddy < 0.0)) && (signbit(az) != (ddz < 0.0));
      if (EXACT && degen_twin) {
        double tdx, tdy, tdz;
        variant_signs(twin_var, tdx, tdy, tdz);
        const bool tneg = act && (signbit(ax) != (tdx < 0.0)) && (signbit(ay) != (tdy < 0.0)) && (signbit(az) != (tdz < 0.0));
#pragma unroll
        for (int cls = 0; cls < 2; cls++) {
          const bool mine = act && (tneg == (cls == 1));
          const unsigned ball = __ballot_sync(0xffffffffu, mine);
          const int sr = cls ? degen_sr_neg : degen_sr_pos;
          if (ball == 0u || sr < 0) continue;
          const int sum = __reduce_add_sync(0xffffffffu, mine ? iv : 0);
          if (lane == 0) {
            atomicAdd(&S.cnt[1][pq * M2DP_SR + sr], (unsigned)__popc(ball));
            atomicAdd(&S.isum[1][pq * M2DP_SR + sr], sum);
          }
        }
      }
#pragma unroll
      for (int cls = 0; cls < 2; cls++) {
        const bool mine = act && (neg == (cls == 1));
        const unsigned ball = __ballot_sync(0xffffffffu, mine);
        const int sr = cls ? degen_sr_neg : degen_sr_pos;
        if (ball == 0u || sr < 0) continue;
        if (EXACT) {
          const int sum = __reduce_add_sync(0xffffffffu, mine ? iv : 0);
          if (lane == 0) {
            atomicAdd(&hcnt[pq * M2DP_SR + sr], (unsigned)__popc(ball));
            atomicAdd(&hisum[pq * M2DP_SR + sr], sum);
          }
        } else {
          // the reference adds the intensities point by point (M2DP.cpp:71); fp64 sums of floats in another
          // order differ only when the exact sum needs more than 53 bits
          const double sum = warp_sum(mine ? (double)it : 0.0);
          if (lane == 0) {
            atomicAdd(&hcnt[pq * M2DP_SR + sr], (unsigned)__popc(ball));
            atomicAdd(&hsum[pq * M2DP_SR + sr], sum);
          }
        }
      }
    }
    static_assert(PQ0 % 4 == 0 && PQ1 % 4 == 0, "planes are walked in groups of four");
    // four planes at a time: the four proposals are independent arithmetic (no branch, no vote in between), then one
    // vote for the rare guard-band case, then the updates
#pragma unroll 1
    for (int pq4 = PQ0; pq4 < PQ1; pq4 += 4) {
      const unsigned dm4 = (unsigned)(degen_mask >> pq4) & 0xfu;   // (warp-uniform)
      int idx4[4];
      bool uns = false;
      unsigned unsafe_bits = 0u;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int pq = pq4 + u;
        const float4 ta = c_tab4[2 * pq], tb = c_tab4[2 * pq + 1];   // (x0, x1, x2, y0), (y1, y2, -, -)
        const float xp = __fmaf_rn(ta.z, pz, __fmaf_rn(ta.y, py, ta.x * px));
        const float yp = __fmaf_rn(tb.y, pz, __fmaf_rn(tb.x, py, ta.w * px));
        bool safe, inside;
        const int sr = propose_bin(xp, yp, R_f, r2_lim, safe, inside);
        const bool live = FULL ? !((dm4 >> u) & 1u) : (act && !((dm4 >> u) & 1u));
        idx4[u] = (safe && inside && live) ? pq * M2DP_SR + sr : -1;
        const bool un = live && !safe;
        uns = uns || un;
        unsafe_bits |= un ? (1u << u) : 0u;
      }
      if (__any_sync(0xffffffffu, uns)) {   // rare
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (!((unsafe_bits >> u) & 1u)) continue;
          const int pq = pq4 + u;
          // deferred: replayed in fp64 by all threads in parallel after the pass (run_entry); a full queue is served
          // on the spot
          const bool twin = twin_var >= 0 && pq >= M2DP_NUM_Q;
          const unsigned long long w = queue_entry(i, pq, slot, twin);
          const int slot_q = atomicAdd(&S.qn, 1);
          if (slot_q < QCAP)
            queue[slot_q] = w;
          else if (!twin) {
            run_entry<EXACT>(S, R, w, var, twin_var, iscale);
            if (pq < M2DP_NUM_Q) S.ibc[1] = 1;   // ... and the p = 0 rows of this pass cannot seed the second pair
          } else
            S.ibc[3] = 1;   // shared evaluations lost: the caller redoes the pair one variant at a time
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int idx = idx4[u];
        if (idx >= 0) {   // (branch-free updates that add 0 to a dump bin were tried: 1 % slower)
          atomicAdd(&hcnt[idx], 1u);
          if (EXACT)
            atomicAdd(&hisum[idx], iv);  // exact integer multiple of 2^emin, |sum| < 2^24
          else
            atomicAdd(&hsum[idx], (double)it);
        }
      }
    }
  }
}

// One binning pass over the points of a scan for variant `var` (var < 0: pre-aligned input, no sign flips):
// planes [PQ0, PQ1) -- and the degenerate planes of the whole table -- into histogram slot `slot`.  Full blocks of
// 1024 points take the variant without per-lane bounds handling.
template <int PQ0, int PQ1, bool EXACT>
__device__ __forceinline__ void bin_pass(M2Smem &S, const ScanRef &R, int var, int slot, int twin_var, float R_f,
                                         float iscale, unsigned long long degen_mask, int degen_sr_pos,
                                         int degen_sr_neg, float coord_lim = 128.0f, bool degen_twin = false) {
  int i0 = 0;
  for (; i0 + M2_THREADS <= R.n; i0 += M2_THREADS)
    bin_block<PQ0, PQ1, EXACT, true>(S, R, var, slot, twin_var, R_f, iscale, degen_mask, degen_sr_pos, degen_sr_neg, i0,
                                     coord_lim, degen_twin);
  if (i0 < R.n)
    bin_block<PQ0, PQ1, EXACT, false>(S, R, var, slot, twin_var, R_f, iscale, degen_mask, degen_sr_pos, degen_sr_neg, i0,
                                      coord_lim, degen_twin);
}

// the deferred evaluations of the passes since the queue was reset, one entry per thread
template <bool EXACT>
__device__ __forceinline__ void replay_queue(M2Smem &S, const ScanRef &R, int var_slot0, int var_slot1, float iscale,
                                             unsigned *stash = nullptr) {
  const unsigned long long *queue = reinterpret_cast<const unsigned long long *>(S.T);
  const int qn = S.qn < QCAP ? S.qn : QCAP;
  // one thread per (entry, role): the up to three evaluations of an entry run side by side
  for (int e = threadIdx.x; e < 3 * qn; e += M2_THREADS)
    run_entry<EXACT>(S, R, queue[e / 3], var_slot0, var_slot1, iscale, stash, e % 3);
}

// binarise the slots [slot0, slot0 + nslot) (M2DP.cpp:84-91) into 128-bit row masks, then per slot the two dominant
// pairs -- count matrix on SVD group 0, binarised matrix on group 1, at the same time -- -> output rows of 2 x 192
template <bool EXACT>
__device__ __forceinline__ void finish_variants(M2Smem &S, int nslot, int n, float ave, double unscale, double *row0,
                                                double *row1) {
  const int lane = threadIdx.x & 31;
  const double *hsum = reinterpret_cast<const double *>(S.isum);
  for (int slot = 0; slot < nslot; slot++) {
    const unsigned *hcnt = S.cnt[slot];
    const int *hisum = S.isum[slot];
#pragma unroll
    for (int k = 0; k < HB / M2_THREADS; k++) {
      const int b = threadIdx.x + k * M2_THREADS;   // a warp holds 32 consecutive columns of one row
      const unsigned c = hcnt[b];
      bool v = false;
      if (c) {
        // M2DP.cpp:87-88: (sum / count) > ave.  Exact mode: sum and ave * count are exact in fp64 (integers in units of
        // 2^emin resp. of ulp(ave); count < 2^29), and a quotient that differs from ave at all differs by more than
        // ulp(ave) / count >> 2^-53 ave, so the rounded quotient compares like the exact one: no division.
        if (EXACT && c < (1u << 29))
          v = (double)hisum[b] * unscale > (double)ave * (double)c;
        else
          v = ((EXACT ? (double)hisum[b] * unscale : hsum[b]) / (double)c) > (double)ave;
      }
      const unsigned word = __ballot_sync(0xffffffffu, v);
      if (lane == 0) S.bits[slot][b >> 5] = word;
    }
  }
  __syncthreads();   // the isum area is free from here on: SVD group 1's workspaces
  M2_PROF(5);
  const int g = threadIdx.x >= SVD_G0 ? 1 : 0;
  double *G1 = reinterpret_cast<double *>(&S.isum[0][0]), *T1 = G1 + WS;
  if (n < 65536)
    gram_counts_cta_small(S.cnt[0], S.G, S.cnt[1], G1, nslot);
  else
    gram_counts_cta<unsigned long long>(S.cnt[0], S.G, S.cnt[1], G1, nslot);
  __syncthreads();
  M2_PROF(6);
  if (nslot == 2) {
    // group g: variant of slot g, count matrix then binarised matrix (M2DP.cpp:94-98,107 and 100-108)
    double *row = g ? row1 : row0, *Gg = g ? G1 : S.G, *Tg = g ? T1 : S.T;
    dominant_pair<false>(S.cnt[g], Gg, Tg, S, g, row);
    dominant_pair<true>(S.bits[g], Gg, Tg, S, g, row + M2DP_SIG);
  } else if (g == 0) {
    dominant_pair<false>(S.cnt[0], S.G, S.T, S, 0, row0);
  } else {
    dominant_pair<true>(S.bits[0], G1, T1, S, 1, row0 + M2DP_SIG);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(M2_THREADS, 1)
m2dp_generate_kernel(const double *__restrict__ xyz, const float *__restrict__ inten,
                     const int64_t *__restrict__ off, int nscan, double S_res_inv, double R_res_inv,
                     int variants, int mirror_ok, double *__restrict__ hist, unsigned long long degen_mask,
                     unsigned *__restrict__ stash_all, unsigned long long *__restrict__ prof) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  M2Smem &S = *reinterpret_cast<M2Smem *>(smem_raw);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    S.prof_on = prof != nullptr;
    S.prof_t = clock64();
    for (int k = 0; k < 16; k++) S.prof[k] = 0ull;
  }
  __syncthreads();
  // the two bins of a plane with zero projection vectors: M2DP.cpp:59-63 evaluated at (+0, +0) and (-0, -0)
  int degen_sr_pos, degen_sr_neg;
  {
    const double PI = 3.14159265358979323846;
    const double ap = (atan2(0.0, 0.0) + PI) * S_res_inv, an = (atan2(-0.0, -0.0) + PI) * S_res_inv;
    const int sp = (int)floor(ap), sn = (int)floor(an);   // ring 0
    degen_sr_pos = sp < M2DP_SR ? sp : -1;
    degen_sr_neg = sn < M2DP_SR ? sn : -1;
  }
  const float R_f = (float)R_res_inv;
  // mirror_ok bit 1: the p = 0 planes of variants v and v ^ 1 are mirror images to within the guard band (host check)
  unsigned *stash = (mirror_ok & 2) && stash_all ? stash_all + (size_t)blockIdx.x * STASH_U32 : nullptr;

  for (int scan = blockIdx.x; scan < nscan; scan += gridDim.x) {
    const int64_t p0 = off[scan];
    const int n = (int)(off[scan + 1] - p0);
    const double *g = xyz + 3 * p0;
    const float *gi = inten + p0;

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
    M2_PROF(10);
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
    M2_PROF(0);
    const float ave = (exact ? (float)s11[9] : __int_as_float(S.ibc[2])) / (float)n;  // M2DP.cpp:81
    const float iscale = exact && emin != (1 << 20) ? (float)ldexp(1.0, -emin) : 0.0f;
    const double unscale = exact && emin != (1 << 20) ? ldexp(1.0, emin) : 0.0;

    ScanRef R;
    R.g = g;
    R.gi = gi;
    R.bc = S.bc;
    R.n = n;
    R.identity = variants ? 0 : 1;
    R.S_res_inv = S_res_inv;
    R.R_res_inv = R_res_inv;
    const int nvar = variants ? 4 : 1;
    double *rows = hist + (size_t)scan * nvar * 2 * M2DP_SIG;

    // one variant on its own (pre-aligned input of the class contract, inexact intensity sums, asymmetric table, or a
    // pair whose guard-band queue overflowed)
    auto single_variant = [&](int var) {
      for (int b = threadIdx.x; b < 2 * HB; b += M2_THREADS) {
        (&S.cnt[0][0])[b] = 0u;
        (&S.isum[0][0])[b] = 0;
      }
      if (threadIdx.x == 0) S.qn = 0;
      __syncthreads();
      const int v = variants ? var : -1;
      if (exact) {
        bin_pass<0, M2DP_PQ, true>(S, R, v, 0, -1, R_f, iscale, degen_mask, degen_sr_pos, degen_sr_neg);
        __syncthreads();
        replay_queue<true>(S, R, v, v, iscale);
        __syncthreads();
        finish_variants<true>(S, 1, n, ave, unscale, rows + (size_t)var * 2 * M2DP_SIG, nullptr);
      } else {
        bin_pass<0, M2DP_PQ, false>(S, R, v, 0, -1, R_f, iscale, degen_mask, degen_sr_pos, degen_sr_neg);
        __syncthreads();
        replay_queue<false>(S, R, v, v, iscale);
        __syncthreads();
        finish_variants<false>(S, 1, n, ave, unscale, rows + (size_t)var * 2 * M2DP_SIG, nullptr);
      }
    };
    if (variants && exact && (mirror_ok & 1)) {
      // ---- variant pairs (a, a + 2): rows of planes 16..63 of variant a + 2 are mirrored copies of variant a's.
      // Across the pairs, the p = 0 rows of variants 1 and 3 are mirrored copies of those of variants 0 and 2: they
      // travel through the stash (evaluations inside the guard band are replayed for each variant on its own).
      bool seeded = false;   // (uniform) the stash holds the p = 0 rows of the second pair
      for (int a = 0; a < 2; a++) {
        for (int b = threadIdx.x; b < 2 * HB; b += M2_THREADS) {
          (&S.cnt[0][0])[b] = 0u;
          (&S.isum[0][0])[b] = 0;
        }
        if (threadIdx.x == 0) {
          S.qn = 0;
          S.ibc[3] = 0;
          if (a == 0) S.ibc[1] = 0;
        }
        __syncthreads();
        if (a == 1 && seeded) {
          for (int b = threadIdx.x; b < 2 * P0_BINS; b += M2_THREADS) {   // (L1 may hold lines of an earlier scan)
            const int slot = b / P0_BINS, r = b % P0_BINS;
            S.cnt[slot][r] = __ldcg(stash + b);
            S.isum[slot][r] = (int)__ldcg(stash + 2 * P0_BINS + b);
          }
          // (this pass also adds the degenerate planes of variant a + 2, which has no pass of its own any more)
          bin_pass<M2DP_NUM_Q, M2DP_PQ, true>(S, R, a, 0, a + 2, R_f, iscale, degen_mask, degen_sr_pos, degen_sr_neg, 128.0f,
                                              true);
        } else {
          bin_pass<0, M2DP_PQ, true>(S, R, a, 0, a + 2, R_f, iscale, degen_mask, degen_sr_pos, degen_sr_neg,
                                     a == 0 && stash ? 64.0f : 128.0f);
        }
        __syncthreads();
        M2_PROF(1);
        if (S.ibc[3]) {   // (uniform) more guard-band evaluations than the queue holds
          __syncthreads();
          single_variant(a);
          single_variant(a + 2);
          continue;
        }
        // twin rows: plane (p, q) of variant a  ->  plane (4 - p, q) of variant a + 2, sectors mirrored (the guard-band
        // evaluations are not in the histogram yet: they are replayed for each variant on its own below)
        for (int b = threadIdx.x; b < (M2DP_PQ - M2DP_NUM_Q) * M2DP_SR; b += M2_THREADS) {
          const int pq = M2DP_NUM_Q + b / M2DP_SR, sr = b % M2DP_SR;
          if ((degen_mask >> pq) & 1ull) continue;
          const int pq2 = twin_plane(pq);
          S.cnt[1][pq2 * M2DP_SR + mirror_sr(sr)] = S.cnt[0][pq * M2DP_SR + sr];
          S.isum[1][pq2 * M2DP_SR + mirror_sr(sr)] = S.isum[0][pq * M2DP_SR + sr];
        }
        __syncthreads();
        M2_PROF(2);
        // the planes of variant a + 2 that have no twin in this pair (p = 0) and its degenerate planes (sign rule of its
        // own); seeded: nothing left to do
        if (!(a == 1 && seeded))
          bin_pass<0, M2DP_NUM_Q, true>(S, R, a + 2, 1, -1, R_f, iscale, degen_mask, degen_sr_pos, degen_sr_neg,
                                        a == 0 && stash ? 64.0f : 128.0f);
        __syncthreads();
        M2_PROF(3);
        unsigned *seed = nullptr;
        if (a == 0 && stash && !S.ibc[1]) {   // (uniform) p = 0 rows of variants 0 / 2 -> variants 1 / 3
          for (int b = threadIdx.x; b < 2 * P0_BINS; b += M2_THREADS) {
            const int slot = b / P0_BINS, r = b % P0_BINS;
            stash[slot * P0_BINS + mirror15_sr(r)] = S.cnt[slot][r];
            stash[2 * P0_BINS + slot * P0_BINS + mirror15_sr(r)] = (unsigned)S.isum[slot][r];
          }
          seed = stash;
          seeded = true;
          __syncthreads();
        }
        replay_queue<true>(S, R, a, a + 2, iscale, seed);
        __syncthreads();
        M2_PROF(4);
        finish_variants<true>(S, 2, n, ave, unscale, rows + (size_t)a * 2 * M2DP_SIG, rows + (size_t)(a + 2) * 2 * M2DP_SIG);
      }
    } else {
      for (int var = 0; var < nvar; var++) single_variant(var);
    }
    __syncthreads();
  }
  if (prof && threadIdx.x == 0)
    for (int k = 0; k < 16; k++)
      if (S.prof[k]) atomicAdd(&prof[k], S.prof[k]);
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

size_t m2dp_generate_workspace_bytes(int, bool) { return (size_t)STASH_MAX_CTAS * STASH_U32 * sizeof(unsigned); }

cudaError_t launch_m2dp_generate(const double *xyz, const float *inten, const int64_t *off, int nscan,
                                 double max_rho, bool do_align_and_variants, double *hist, void *ws, size_t ws_bytes,
                                 int num_sms, cudaStream_t st, int64_t *launches) {
  if (nscan <= 0) return cudaSuccess;
  double xp[3 * M2DP_PQ], yp[3 * M2DP_PQ];
  build_tables(xp, yp);
  cudaError_t e = cudaMemcpyToSymbolAsync(c_xproj, xp, sizeof(xp), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(c_yproj, yp, sizeof(yp), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  float4 tab[2 * M2DP_PQ];
  for (int k = 0; k < M2DP_PQ; k++) {
    tab[2 * k] = make_float4((float)xp[3 * k], (float)xp[3 * k + 1], (float)xp[3 * k + 2], (float)yp[3 * k]);
    tab[2 * k + 1] = make_float4((float)yp[3 * k + 1], (float)yp[3 * k + 2], 0.0f, 0.0f);
  }
  e = cudaMemcpyToSymbolAsync(c_tab4, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(m2dp_generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(M2Smem));
  if (e != cudaSuccess) return e;
  const double S_res_inv = M2DP_NUM_S / (2.0 * 3.14159265358979323846);  // M2DP.cpp:32
  const double R_res_inv = M2DP_NUM_R / max_rho;                          // M2DP.cpp:33
  int grid = nscan < num_sms ? nscan : num_sms;
  if (grid > STASH_MAX_CTAS) grid = STASH_MAX_CTAS;
  unsigned long long degen_mask = 0;   // planes with exactly zero projection vectors (M2DP.cpp:21-25 at p=2, q=0)
  for (int k = 0; k < M2DP_PQ; k++) {
    bool zero = true;
    for (int c = 0; c < 3; c++) zero = zero && xp[3 * k + c] == 0.0 && yp[3 * k + c] == 0.0;
    // the sign rule used by the kernel needs +0 entries
    for (int c = 0; c < 3; c++) zero = zero && !std::signbit(xp[3 * k + c]) && !std::signbit(yp[3 * k + c]);
    if (zero) degen_mask |= 1ull << k;
  }
  // variant sharing (see the file header): the rows of planes (p, q), p = 1..3, of variant a + 2 are mirrored copies of
  // the rows of planes (4 - p, q) of variant a iff  S_{a+2} xProj[4-p][q] == -S_a xProj[p][q]  and
  // S_{a+2} yProj[4-p][q] == S_a yProj[p][q]  hold EXACTLY (then the reference's fp64 expressions are mirror images too)
  int mirror_ok = 1;
  for (int a = 0; a < 2 && mirror_ok; a++) {
    const double sa[3] = {-1.0, a ? 1.0 : -1.0, a ? -1.0 : 1.0};         // variant a:     dx = -1, dy, dx * dy
    const double sb[3] = {1.0, a ? 1.0 : -1.0, a ? 1.0 : -1.0};          // variant a + 2: dx = +1, dy, dx * dy
    for (int p = 1; p < M2DP_NUM_P && mirror_ok; p++)
      for (int q = 0; q < M2DP_NUM_Q && mirror_ok; q++) {
        const int k = p * M2DP_NUM_Q + q, k2 = (M2DP_NUM_P - p) * M2DP_NUM_Q + q;
        for (int c = 0; c < 3; c++)
          if (sb[c] * xp[3 * k2 + c] != -(sa[c] * xp[3 * k + c]) || sb[c] * yp[3 * k2 + c] != sa[c] * yp[3 * k + c]) mirror_ok = 0;
        if (((degen_mask >> k) & 1ull) != ((degen_mask >> k2) & 1ull)) mirror_ok = 0;
      }
  }
  // cross-pair sharing: the p = 0 planes of variants v and v ^ 1 (v = 0, 2) see the projected point mirrored,
  // (xp, yp) -> (xp, -yp), up to the residue of cosf(-pi/2f) = -4.4e-8 in the float table: with eps_tab the largest
  // deviation of a table entry from the exact mirror image, the fp64 coordinates of variant v ^ 1 differ from the
  // mirrored ones of variant v by at most 2 * 64 m * eps_tab = 1.1e-5 m for points inside 64 m (the x row has no
  // deviating x entry).  The fp32 proposal of such a point is good to 1e-5 m and accepted only 6e-5 m or more away from
  // every bin edge, so an accepted proposal of variant v is also the mirrored bin of variant v ^ 1; everything else is
  // replayed in fp64 for each variant on its own.
  if (mirror_ok && ws && ws_bytes >= (size_t)grid * STASH_U32 * sizeof(unsigned)) {
    double eps_tab = 0.0;
    for (int v = 0; v < 4; v += 2) {
      const double s0[3] = {v ? 1.0 : -1.0, -1.0, v ? -1.0 : 1.0};   // variant v:     dx, dy = -1, dx * dy
      const double s1[3] = {v ? 1.0 : -1.0, 1.0, v ? 1.0 : -1.0};    // variant v ^ 1: dx, dy = +1, dx * dy
      for (int q = 0; q < M2DP_NUM_Q; q++)
        for (int c = 0; c < 3; c++) {
          eps_tab = std::fmax(eps_tab, std::fabs(s1[c] * xp[3 * q + c] - s0[c] * xp[3 * q + c]));
          eps_tab = std::fmax(eps_tab, std::fabs(s1[c] * yp[3 * q + c] + s0[c] * yp[3 * q + c]));
        }
    }
    bool p0_degen = false;
    for (int q = 0; q < M2DP_NUM_Q; q++) p0_degen = p0_degen || ((degen_mask >> q) & 1ull);
    if (eps_tab * 128.0 <= 2e-5 && !p0_degen) mirror_ok |= 2;
  }
  m2dp_generate_kernel<<<grid, M2_THREADS, sizeof(M2Smem), st>>>(xyz, inten, off, nscan, S_res_inv, R_res_inv,
                                                                 do_align_and_variants ? 1 : 0, mirror_ok, hist, degen_mask,
                                                                 static_cast<unsigned *>(ws), g_debug.prof);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
