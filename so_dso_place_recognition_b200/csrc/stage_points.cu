// Point staging on the GPU: pts_preprocess (pts_preprocess.h:169-232) = the sliding accumulation of world points,
// the w2c transform + range crop (generate_spherical_points, :135-167) and the voxel-grid "highest point" /
// 1-degree polar "closest point" de-duplication (filterPoints :51-94, filterPointsPolar :96-133).
//
// What is sequential in the reference (the walk over poses: reset rule, INIT_FRAME skip, which points have been
// consumed when) is O(poses + points) bookkeeping and stays on the host (stage_plan below).  What is heavy -- every
// accumulated point is transformed and tested at every frame it is alive, and the survivors of every frame are
// de-duplicated -- runs here:
//   stage_lifetime_kernel   one thread per point: walk the frames from the point's entry until it first falls
//                           outside the range (the reference then drops it for good, :165) or the next reset;
//                           -> the half-open interval of scans [s0, s1) that see the point
//   stage_dedupe_kernel     persistent CTAs over scans: candidates = points whose interval contains the scan;
//                           per voxel the minimum of an order-preserving fp64 key (y for the grid filter, |p| for the
//                           polar one) by 64-bit atomicMin in an L2-resident hash table, ties -> lowest point index
//                           (= the reference's "first one wins"); winners are ranked by voxel index with a
//                           shared-memory bitmap
//   stage_emit_kernel       transform the winners once more and write xyz / intensity
// Output order inside a scan: ascending voxel index.  (The reference's order is the iteration order of a libstdc++
// unordered_map; the point SET is identical, and the descriptors only depend on the order through the
// float average-intensity sum.)
// Compiled with -fmad=false: the transform is the reference's expression, separately rounded.
#include <algorithm>
#include <climits>
#include <cmath>
#include <vector>

#include "../../include/sodso_pr.h"
#include "pca.cuh"

namespace sodso {
namespace {

constexpr int INIT_FRAME = 30;                   // pts_preprocess.h:13
constexpr double RES_GRID = 30;                  // pts_preprocess.h:14
constexpr int ST_THREADS = 256;

struct StageGeom {
  double lidar_range;
  int polar;
  // grid filter (pts_preprocess.h:55-65)
  double steps[3];
  int loc_step[3];
  // polar filter (pts_preprocess.h:101-104)
  double azi_res_inv, ele_res_inv;
  int azi_bins;
  int nvox;   // number of distinct voxel indices (bitmap size)
};

// p_l = w2c * [p; 1]  (pts_preprocess.h:140-142).  Summation order: strictly left to right over the 4 columns, and
// (x^2 + y^2) + z^2 for the norm -- the ORACLE's order (oracle/sodso_oracle.cpp, stage_transform).  Which order Eigen
// 3.x uses for a fixed-size 3x4 * 4 product and for Vector3d::norm() (sequential or its unrolled pairwise
// reduction) could not be checked here (no Eigen in the image), so "bit-identical staged point set" is a statement
// about the oracle: against a real Eigen build a point within 1 ulp of the 45 m crop (pts_preprocess.h:144) or of a
// voxel edge could be kept / dropped differently.
__device__ __forceinline__ void to_camera(const double *__restrict__ w, const double *__restrict__ p, double *l) {
#pragma unroll
  for (int r = 0; r < 3; r++)
    l[r] = ((w[4 * r + 0] * p[0] + w[4 * r + 1] * p[1]) + w[4 * r + 2] * p[2]) + w[4 * r + 3] * 1.0;
}
__device__ __forceinline__ double norm3(const double *l) { return sqrt((l[0] * l[0] + l[1] * l[1]) + l[2] * l[2]); }

// voxel index and selection key of a camera-frame point
__device__ __forceinline__ void voxel_of(const StageGeom &G, const double *l, int &loc, long long &key) {
  if (!G.polar) {   // pts_preprocess.h:71-75, keep the smallest y ("highest" point, :78-79)
    const int xi = (int)floor((l[0] + G.lidar_range) * G.steps[0]);
    const int yi = (int)floor((l[1] + G.lidar_range) * G.steps[1]);
    const int zi = (int)floor((l[2] + G.lidar_range) * G.steps[2]);
    loc = xi * G.loc_step[0] + yi * G.loc_step[1] + zi * G.loc_step[2];
    key = f64_key(l[1]);
  } else {          // pts_preprocess.h:108-114, keep the smallest norm (:117-118)
    const double PI = 3.14159265358979323846;
    const double xz = sqrt(l[0] * l[0] + l[2] * l[2]);
    const int azi = (int)floor((atan2(l[2], l[0]) + PI) * G.azi_res_inv);
    const int ele = (int)floor((atan2(l[1], xz) + PI / 2) * G.ele_res_inv);
    loc = azi + ele * G.azi_bins;
    key = f64_key(norm3(l));
  }
}

// scans [s0, s1) that see point p.  entry[p]: frame at which the reference moves the point into nearby_pts
// (-1: never); scan_of_frame[f]: scan index of a processed frame, -1 for skipped ones; seg_end[f]: first frame
// after f that resets the accumulator (n_pose if none).
__global__ void __launch_bounds__(ST_THREADS)
stage_lifetime_kernel(const double *__restrict__ pt_xyz, const int *__restrict__ entry, int64_t n_pts,
                      const double *__restrict__ w2c, const int *__restrict__ scan_of_frame,
                      const int *__restrict__ seg_end, double lidar_range, int *__restrict__ s0_out,
                      int *__restrict__ s1_out, int *__restrict__ cnt_diff) {
  const int64_t p = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x;
  if (p >= n_pts) return;
  int s0 = 0, s1 = 0;
  const int e = entry[p];
  if (e >= 0) {
    const double x[3] = {pt_xyz[3 * p], pt_xyz[3 * p + 1], pt_xyz[3 * p + 2]};
    const int end = seg_end[e];
    bool started = false;
    for (int f = e; f < end; f++) {
      const int s = scan_of_frame[f];
      if (s < 0) continue;
      double l[3];
      to_camera(w2c + 12 * (size_t)f, x, l);
      if (!(norm3(l) < lidar_range)) break;   // dropped from nearby_pts for good (pts_preprocess.h:144,165)
      if (!started) {
        started = true;
        s0 = s;
      }
      s1 = s + 1;
    }
    if (!started) s0 = s1 = 0;
  }
  s0_out[p] = s0;
  s1_out[p] = s1;
  if (s1 > s0) {
    atomicAdd(&cnt_diff[s0], 1);
    atomicAdd(&cnt_diff[s1], -1);
  }
}

struct DedupeArgs {
  const double *pt_xyz;
  const double *w2c;
  const int *s0, *s1;
  const int *frame_of_scan;
  const int64_t *pt_lo, *pt_hi;     // per scan: range of point indices that can be alive
  const int64_t *cand_off;          // per scan: offset of its winner list (prefix of the candidate counts)
  int nscan;
  // per-CTA workspace
  int *tloc;                        // [grid][nslots]  voxel index of the slot, -1 = empty
  long long *tkey;                  // [grid][nslots]  minimum key
  int *tidx;                        // [grid][nslots]  lowest point index among those with the minimum key
  int nslots;                       // power of two
  int *c_pt, *c_slot;               // [grid][maxcand]
  long long *c_key;                 // [grid][maxcand]
  int maxcand;
  int *win_list;                    // winners, ranked by voxel index, at cand_off[s]
  int *n_out;                       // per scan
};

__global__ void __launch_bounds__(ST_THREADS)
stage_dedupe_kernel(const StageGeom G, const DedupeArgs A) {
  extern __shared__ unsigned bitmap[];   // nvox bits, then ST_THREADS chunk prefixes
  const int nwords = (G.nvox + 31) >> 5;
  unsigned *chunk_prefix = bitmap + nwords;
  __shared__ int ncand;
  int *tloc = A.tloc + (size_t)blockIdx.x * A.nslots;
  long long *tkey = A.tkey + (size_t)blockIdx.x * A.nslots;
  int *tidx = A.tidx + (size_t)blockIdx.x * A.nslots;
  int *c_pt = A.c_pt + (size_t)blockIdx.x * A.maxcand, *c_slot = A.c_slot + (size_t)blockIdx.x * A.maxcand;
  long long *c_key = A.c_key + (size_t)blockIdx.x * A.maxcand;
  const int words_per_thread = (nwords + ST_THREADS - 1) / ST_THREADS;
  const unsigned mask = (unsigned)A.nslots - 1u;

  for (int w = threadIdx.x; w < nwords; w += ST_THREADS) bitmap[w] = 0u;
  for (int scan = blockIdx.x; scan < A.nscan; scan += gridDim.x) {
    if (threadIdx.x == 0) ncand = 0;
    __syncthreads();
    const double *w = A.w2c + 12 * (size_t)A.frame_of_scan[scan];
    // ---- candidates of this frame: transform, voxel, key; per-voxel minimum key
    for (int64_t p = A.pt_lo[scan] + threadIdx.x; p < A.pt_hi[scan]; p += ST_THREADS) {
      if (!(A.s0[p] <= scan && scan < A.s1[p])) continue;
      const double x[3] = {A.pt_xyz[3 * p], A.pt_xyz[3 * p + 1], A.pt_xyz[3 * p + 2]};
      double l[3];
      to_camera(w, x, l);
      int loc;
      long long key;
      voxel_of(G, l, loc, key);
      unsigned h = ((unsigned)loc * 2654435761u) & mask;
      for (;;) {
        const int old = atomicCAS(&tloc[h], -1, loc);
        if (old == -1 || old == loc) break;
        h = (h + 1u) & mask;
      }
      atomicMin(&tkey[h], key);
      const int c = atomicAdd(&ncand, 1);
      c_pt[c] = (int)p;
      c_slot[c] = (int)h;
      c_key[c] = key;
    }
    __syncthreads();
    const int nc = ncand;
    // ---- ties on the key: the lowest point index wins (the reference keeps the first one, :78-79 / :117-118)
    for (int c = threadIdx.x; c < nc; c += ST_THREADS)
      if (c_key[c] == __ldcg(&tkey[c_slot[c]])) atomicMin(&tidx[c_slot[c]], c_pt[c]);
    __syncthreads();
    // ---- winners -> bitmap over voxel indices
    for (int c = threadIdx.x; c < nc; c += ST_THREADS) {
      const int sl = c_slot[c];
      if (__ldcg(&tidx[sl]) == c_pt[c]) {
        const int loc = __ldcg(&tloc[sl]);
        atomicOr(&bitmap[loc >> 5], 1u << (loc & 31));
      }
    }
    __syncthreads();
    // ---- rank by voxel index: per-thread chunk popcounts + block scan
    {
      const int w0 = threadIdx.x * words_per_thread, w1 = min(nwords, w0 + words_per_thread);
      unsigned cnt = 0;
      for (int wv = w0; wv < w1; wv++) cnt += __popc(bitmap[wv]);
      chunk_prefix[threadIdx.x] = cnt;
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int t = 0; t < ST_THREADS; t++) {
          const unsigned v = chunk_prefix[t];
          chunk_prefix[t] = run;
          run += v;
        }
        A.n_out[scan] = (int)run;
      }
      __syncthreads();
    }
    for (int c = threadIdx.x; c < nc; c += ST_THREADS) {
      const int sl = c_slot[c];
      if (__ldcg(&tidx[sl]) == c_pt[c]) {
        const int loc = __ldcg(&tloc[sl]);
        const int wv = loc >> 5, t = wv / words_per_thread;
        unsigned r = chunk_prefix[t];
        for (int k = t * words_per_thread; k < wv; k++) r += __popc(bitmap[k]);
        r += __popc(bitmap[wv] & ((1u << (loc & 31)) - 1u));
        A.win_list[A.cand_off[scan] + r] = c_pt[c];
      }
    }
    __syncthreads();
    // ---- reset what was touched
    for (int c = threadIdx.x; c < nc; c += ST_THREADS) {
      const int sl = c_slot[c];
      const int loc = __ldcg(&tloc[sl]);
      if (loc >= 0) bitmap[loc >> 5] = 0u;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nc; c += ST_THREADS) {
      const int sl = c_slot[c];
      tloc[sl] = -1;
      tkey[sl] = LLONG_MAX;
      tidx[sl] = INT_MAX;
    }
    __threadfence_block();
    __syncthreads();
  }
}

__global__ void __launch_bounds__(ST_THREADS)
stage_emit_kernel(const double *__restrict__ pt_xyz, const float *__restrict__ pt_inten,
                  const double *__restrict__ w2c, const int *__restrict__ frame_of_scan,
                  const int64_t *__restrict__ cand_off, const int *__restrict__ win_list,
                  const int64_t *__restrict__ off, int nscan, double *__restrict__ xyz, float *__restrict__ inten) {
  for (int scan = blockIdx.x; scan < nscan; scan += gridDim.x) {
    const double *w = w2c + 12 * (size_t)frame_of_scan[scan];
    const int64_t o = off[scan];
    const int n = (int)(off[scan + 1] - o);
    for (int r = threadIdx.x; r < n; r += ST_THREADS) {
      const int p = win_list[cand_off[scan] + r];
      const double x[3] = {pt_xyz[3 * (size_t)p], pt_xyz[3 * (size_t)p + 1], pt_xyz[3 * (size_t)p + 2]};
      double l[3];
      to_camera(w, x, l);
      xyz[3 * (o + r) + 0] = l[0];
      xyz[3 * (o + r) + 1] = l[1];
      xyz[3 * (o + r) + 2] = l[2];
      inten[o + r] = pt_inten[p];
    }
  }
}

__global__ void fill_i32(int *p, size_t n, int v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_i64(long long *p, size_t n, long long v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

// ---- host side --------------------------------------------------------------------------------------------------
// The sequential part of pts_preprocess.h:187-216 (no arithmetic on points).
void stage_plan(const int *pose_id, const double *w2c, int n_pose, const int *pt_id, int64_t n_pts, StagePlan &P) {
  P.scan_of_frame.assign((size_t)n_pose, -1);
  P.seg_end.assign((size_t)n_pose, n_pose);
  P.entry.assign((size_t)n_pts, -1);
  P.frame_of_scan.clear();
  P.ids.clear();
  P.pt_lo.clear();
  P.pt_hi.clear();
  int64_t pts_idx = 0, seg_first_pt = 0;
  int frame_from_reset = 0;
  std::vector<int> resets;
  for (int f = 0; f < n_pose; f++) {
    const double *w = w2c + 12 * (size_t)f;
    const double tn = std::sqrt((w[3] * w[3] + w[7] * w[7]) + w[11] * w[11]);
    if (tn < 1.0) {                                                    // :189-193
      frame_from_reset = 0;
      seg_first_pt = pts_idx;
      resets.push_back(f);
    }
    while (pts_idx < n_pts && pt_id[pts_idx] <= pose_id[f]) P.entry[(size_t)pts_idx++] = f;   // :196-200
    if (frame_from_reset < INIT_FRAME) {                               // :203-206
      frame_from_reset++;
      continue;
    }
    P.scan_of_frame[(size_t)f] = (int)P.frame_of_scan.size();
    P.frame_of_scan.push_back(f);
    P.ids.push_back(pose_id[f]);
    P.pt_lo.push_back(seg_first_pt);
    P.pt_hi.push_back(pts_idx);
  }
  // first reset frame strictly after f
  size_t ri = 0;
  for (int f = 0; f < n_pose; f++) {
    while (ri < resets.size() && resets[ri] <= f) ri++;
    P.seg_end[(size_t)f] = ri < resets.size() ? resets[ri] : n_pose;
  }
}

static StageGeom make_geom(double lidar_range, bool polar) {
  StageGeom G{};
  G.lidar_range = lidar_range;
  G.polar = polar ? 1 : 0;
  const double resolution[3] = {RES_GRID, 2 * RES_GRID, RES_GRID};     // pts_preprocess.h:156-157
  int voxel_size[3];
  for (int k = 0; k < 3; k++) {
    const double res = lidar_range / resolution[k];                    // :55-57
    G.steps[k] = 1.0 / res;                                            // :58-59
    voxel_size[k] = static_cast<int>(std::floor(2 * lidar_range * G.steps[k]) + 1);   // :60-63
  }
  G.loc_step[0] = 1;                                                   // :64
  G.loc_step[1] = voxel_size[0];
  G.loc_step[2] = voxel_size[0] * voxel_size[1];
  const double RES_POLAR = 1.0 / 180.0 * M_PI;                         // :15
  G.azi_res_inv = 1.0 / RES_POLAR;                                     // :101-102
  G.ele_res_inv = 1.0 / RES_POLAR;
  G.azi_bins = static_cast<int>(std::floor(2 * M_PI * G.azi_res_inv) + 1);   // :103
  // ||p_l|| < lidar_range bounds every voxel coordinate by voxel_size (one extra layer for the rounding at the border)
  G.nvox = polar ? G.azi_bins * (static_cast<int>(std::floor(M_PI * G.ele_res_inv)) + 2)
                 : (voxel_size[0] + 1) * (voxel_size[1] + 1) * (voxel_size[2] + 1);
  return G;
}

cudaError_t launch_stage_lifetime(const double *pt_xyz, const int *entry, int64_t n_pts, const double *w2c,
                                  const int *scan_of_frame, const int *seg_end, double lidar_range, int *s0, int *s1,
                                  int *cnt_diff, cudaStream_t st, int64_t *launches) {
  if (n_pts <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n_pts + ST_THREADS - 1) / ST_THREADS);
  stage_lifetime_kernel<<<grid, ST_THREADS, 0, st>>>(pt_xyz, entry, n_pts, w2c, scan_of_frame, seg_end, lidar_range, s0,
                                                     s1, cnt_diff);
  if (launches) ++*launches;
  return cudaGetLastError();
}

size_t stage_dedupe_workspace_bytes(int grid, int maxcand, int *nslots_out) {
  int nslots = 1024;
  while (nslots < 2 * maxcand) nslots <<= 1;
  if (nslots_out) *nslots_out = nslots;
  return (size_t)grid * ((size_t)nslots * 16 + (size_t)std::max(maxcand, 1) * 16) + 256;
}

cudaError_t launch_stage_dedupe(const double *pt_xyz, const double *w2c, const int *s0, const int *s1,
                                const int *frame_of_scan, const int64_t *pt_lo, const int64_t *pt_hi,
                                const int64_t *cand_off, int nscan, int maxcand, double lidar_range, bool polar,
                                void *workspace, int grid, int *win_list, int *n_out, cudaStream_t st,
                                int64_t *launches) {
  if (nscan <= 0) return cudaSuccess;
  const StageGeom G = make_geom(lidar_range, polar);
  int nslots = 0;
  stage_dedupe_workspace_bytes(grid, maxcand, &nslots);
  const int mc = std::max(maxcand, 1);
  unsigned char *wsp = reinterpret_cast<unsigned char *>(workspace);
  DedupeArgs A{};
  A.pt_xyz = pt_xyz;
  A.w2c = w2c;
  A.s0 = s0;
  A.s1 = s1;
  A.frame_of_scan = frame_of_scan;
  A.pt_lo = pt_lo;
  A.pt_hi = pt_hi;
  A.cand_off = cand_off;
  A.nscan = nscan;
  A.nslots = nslots;
  A.maxcand = mc;
  A.tkey = reinterpret_cast<long long *>(wsp);
  wsp += (size_t)grid * nslots * 8;
  A.c_key = reinterpret_cast<long long *>(wsp);
  wsp += (size_t)grid * mc * 8;
  A.tloc = reinterpret_cast<int *>(wsp);
  wsp += (size_t)grid * nslots * 4;
  A.tidx = reinterpret_cast<int *>(wsp);
  wsp += (size_t)grid * nslots * 4;
  A.c_pt = reinterpret_cast<int *>(wsp);
  wsp += (size_t)grid * mc * 4;
  A.c_slot = reinterpret_cast<int *>(wsp);
  A.win_list = win_list;
  A.n_out = n_out;
  fill_i64<<<256, 256, 0, st>>>(A.tkey, (size_t)grid * nslots, LLONG_MAX);
  fill_i32<<<256, 256, 0, st>>>(A.tloc, (size_t)grid * nslots, -1);
  fill_i32<<<256, 256, 0, st>>>(A.tidx, (size_t)grid * nslots, INT_MAX);
  const size_t smem = ((size_t)((G.nvox + 31) >> 5) + ST_THREADS) * sizeof(unsigned);
  cudaError_t e = cudaFuncSetAttribute(stage_dedupe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  stage_dedupe_kernel<<<grid, ST_THREADS, smem, st>>>(G, A);
  if (launches) *launches += 4;
  return cudaGetLastError();
}

cudaError_t launch_stage_emit(const double *pt_xyz, const float *pt_inten, const double *w2c, const int *frame_of_scan,
                              const int64_t *cand_off, const int *win_list, const int64_t *off, int nscan, double *xyz,
                              float *inten, int grid, cudaStream_t st, int64_t *launches) {
  if (nscan <= 0) return cudaSuccess;
  stage_emit_kernel<<<grid, ST_THREADS, 0, st>>>(pt_xyz, pt_inten, w2c, frame_of_scan, cand_off, win_list, off, nscan, xyz,
                                                 inten);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
