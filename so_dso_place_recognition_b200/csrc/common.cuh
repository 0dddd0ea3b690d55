// Shared declarations of libsodso_pr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace sodso {

constexpr int SC_NUM_S = 60;   // SC.h:7
constexpr int SC_NUM_R = 20;   // SC.h:8
constexpr int SC_SIZE = SC_NUM_S * SC_NUM_R;
constexpr int M2DP_NUM_S = 16; // M2DP.h:7 (theta)
constexpr int M2DP_NUM_R = 8;  // M2DP.h:8 (rho)
constexpr int M2DP_NUM_P = 4;  // M2DP.h:9
constexpr int M2DP_NUM_Q = 16; // M2DP.h:10
constexpr int M2DP_PQ = M2DP_NUM_P * M2DP_NUM_Q;  // 64
constexpr int M2DP_SR = M2DP_NUM_S * M2DP_NUM_R;  // 128
constexpr int M2DP_SIG = M2DP_PQ + M2DP_SR;       // 192

constexpr double STAT_SHIFT = 0.25;  // centre of the [0, 0.5] distance range, see sodso_db_partial_stats
constexpr int STATS_W = 6;           // per query row: [sum, sum of squares, count] of the non-NaN entries, per channel

#ifdef __CUDACC__
// atan2(num, den) / 2pi + 1/2 in [0, 1] ("turns"), fp32, |error| < 2e-7 turns for finite inputs that are
// not both zero (odd polynomial of degree 13 on the octant-reduced ratio: 5e-7 rad, + a 2-ulp divide).
// Used only to PROPOSE a polar bin; the generation kernels accept it only away from bin edges.
__device__ __forceinline__ float fast_turns(float num, float den) {
  const float an = fabsf(num), ad = fabsf(den);
  const float mx = fmaxf(an, ad), mn = fminf(an, ad);
  const float a = __fdividef(mn, mx);
  const float s = a * a;
  float p = 0.00782548f;
  p = __fmaf_rn(p, s, -0.03689863f);
  p = __fmaf_rn(p, s, 0.08374156f);
  p = __fmaf_rn(p, s, -0.13480406f);
  p = __fmaf_rn(p, s, 0.19879872f);
  p = __fmaf_rn(p, s, -0.33326375f);
  p = __fmaf_rn(p, s, 0.99999933f);
  float t = p * a * 0.15915494309f;  // atan(a) / 2pi in [0, 1/8]
  t = an > ad ? 0.25f - t : t;
  t = den < 0.0f ? 0.5f - t : t;
  t = num < 0.0f ? -t : t;
  return t + 0.5f;
}
#endif

void set_error(const std::string &msg);

// kernel debug switches of tools/ (sodso_debug_set_kernel_flags, include/sodso_pr_debug.h); never read from the environment
struct DebugFlags {
  int tc_flags = 0;    // sc_match_tc_kernel: 1 skip epilogue loads, 2 skip MMAs, 4 force generic mode
  int gen_flags = 2;   // sc_generate_kernel variant bits
  int gen_ctas = 0;    // CTAs per SM of sc_generate_kernel (0 = default)
  unsigned long long *prof = nullptr;   // device buffer of 16 phase clock sums (sodso_debug_phase_profile)
};
extern DebugFlags g_debug;

struct KernelTimer;  // capi.cu

#define SODSO_CUDA_CHECK(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::sodso::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +        \
                         __FILE__ + ":" + std::to_string(__LINE__) + ")");                  \
      return SODSO_E_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

// ---- launchers implemented in the kernel translation units ----------------------------
// Each returns a cudaError_t from the launch (cudaGetLastError) and counts into *launches.

// sc_generate.cu
cudaError_t launch_sc_generate(const double *xyz, const float *inten, const int64_t *off, int nscan,
                               double max_rho, double *hist, int num_sms, cudaStream_t st,
                               int64_t *launches);
cudaError_t launch_align_pca(const double *xyz, const int64_t *off, int nscan, double *out_xyz,
                             double *evec, int num_sms, cudaStream_t st, int64_t *launches);

cudaError_t launch_fast_turns_probe(const float *num, const float *den, int64_t n, float *out, cudaStream_t st);

// m2dp_generate.cu
cudaError_t launch_m2dp_generate(const double *xyz, const float *inten, const int64_t *off, int nscan,
                                 double max_rho, bool do_align_and_variants, double *hist,
                                 void *workspace, size_t workspace_bytes, int num_sms,
                                 cudaStream_t st, int64_t *launches);
size_t m2dp_generate_workspace_bytes(int nscan, bool variants);

// sc_match_simt.cu : fp32 CUDA-core cross-check path
// prep: hist (rows x 2400 fp64) -> per channel, row-normalised fp32, K-major [ch][1200][ld]
cudaError_t launch_sc_prep_simt(const double *hist, int rows, float *out_t, int ld, cudaStream_t st,
                                int64_t *launches);
cudaError_t launch_sc_match_simt(const float *q_t, int m, int ldq, const float *h_t, int n, int ldh,
                                 float *d_p, float *d_i, int ldd, cudaStream_t st, int64_t *launches);

// sc_match_tc.cu : tcgen05 path
struct ScTcDb;     // resident DB operand (device)
struct ScTcQuery;  // query operand (device)
size_t sc_tc_db_bytes(int n);
size_t sc_tc_query_bytes(int m);
cudaError_t launch_sc_tc_prep_db(const double *hist, int n, void *db_buf, cudaStream_t st,
                                 int64_t *launches);
cudaError_t launch_sc_tc_prep_query(const double *hist, int m, void *q_buf, cudaStream_t st,
                                    int64_t *launches);
// streamed / blocked variants: operands are filled row range by row range and matched block by block
cudaError_t launch_sc_tc_clear_flags(void *buf, cudaStream_t st);
cudaError_t launch_sc_tc_prep_db_rows(const double *hist, int n, int row0, int row1, void *db_buf,
                                      cudaStream_t st, int64_t *launches, int n_layout = 0);
cudaError_t launch_sc_tc_prep_query_rows(const double *hist, int m, int row0, int row1, void *q_buf,
                                         cudaStream_t st, int64_t *launches);
int sc_tc_db_rows_padded(int n);
cudaError_t sc_tc_db_relayout(const void *old_buf, int old_cap, void *new_buf, int new_cap, cudaStream_t st);
int sc_tc_query_rows_padded(int m);
cudaError_t launch_sc_match_tc_block(const void *q_buf, int m, int q0, int q1, const void *db_buf, int n,
                                     int r0, int r1, float *d_p, float *d_i, int ldd, int num_sms,
                                     cudaStream_t st, int64_t *launches);
long long sc_tc_self_items(int n, int q0, int q1);
cudaError_t launch_sc_match_tc_self(const void *q_buf, const void *db_buf, int n, int q0, int q1, float *d_p, float *d_i,
                                    int ldd, int num_sms, cudaStream_t st, int64_t *launches);
cudaError_t launch_sc_match_tc_blocks(const void *q_buf, int m, const void *db_buf, int n, int qa0, int qa1,
                                      int ra0, int ra1, int qb0, int qb1, int rb0, int rb1, float *d_p,
                                      float *d_i, int ldd, int num_sms, cudaStream_t st, int64_t *launches,
                                      int n_layout = 0);
// d_p / d_i: fp32 m x ldd.  Returns cudaErrorNotSupported if tensor maps cannot be encoded.
cudaError_t launch_sc_match_tc(const void *q_buf, int m, const void *db_buf, int n, float *d_p,
                               float *d_i, int ldd, int num_sms, cudaStream_t st, int64_t *launches, int n_layout = 0);

// stage_points.cu : pts_preprocess on the GPU
struct StagePlan {                  // host bookkeeping of the sequential pose walk (pts_preprocess.h:187-216)
  std::vector<int> scan_of_frame;   // per pose: scan index, -1 for skipped frames
  std::vector<int> seg_end;         // per pose: first later pose that resets the accumulator (n_pose if none)
  std::vector<int> entry;           // per point: pose at which it enters nearby_pts, -1 = never
  std::vector<int> frame_of_scan;   // per scan
  std::vector<int> ids;             // per scan: incoming id (incoming_id_file.txt)
  std::vector<int64_t> pt_lo, pt_hi;  // per scan: range of point indices that can be alive
};
void stage_plan(const int *pose_id, const double *w2c, int n_pose, const int *pt_id, int64_t n_pts, StagePlan &P);
cudaError_t launch_stage_lifetime(const double *pt_xyz, const int *entry, int64_t n_pts, const double *w2c,
                                  const int *scan_of_frame, const int *seg_end, double lidar_range, int *s0,
                                  int *s1, int *cnt_diff, cudaStream_t st, int64_t *launches);
size_t stage_dedupe_workspace_bytes(int grid, int maxcand, int *nslots_out);
cudaError_t launch_stage_dedupe(const double *pt_xyz, const double *w2c, const int *s0, const int *s1,
                                const int *frame_of_scan, const int64_t *pt_lo, const int64_t *pt_hi,
                                const int64_t *cand_off, int nscan, int maxcand, double lidar_range, bool polar,
                                void *workspace, int grid, int *win_list, int *n_out, cudaStream_t st,
                                int64_t *launches);
cudaError_t launch_stage_emit(const double *pt_xyz, const float *pt_inten, const double *w2c,
                              const int *frame_of_scan, const int64_t *cand_off, const int *win_list,
                              const int64_t *off, int nscan, double *xyz, float *inten, int grid,
                              cudaStream_t st, int64_t *launches);

// delight.cu : DELIGHT descriptor
cudaError_t launch_delight_generate(const double *xyz, const float *inten, const int64_t *off, int nscan,
                                    double *hist, int num_sms, cudaStream_t st, int64_t *launches);
size_t delight_match_workspace_bytes(int m, int n);
cudaError_t launch_delight_match(const double *hist1, int m, const double *hist2, int n, double *dist,
                                 void *workspace, cudaStream_t st, int64_t *launches);
cudaError_t launch_top1_single(const double *d, int m, int n, int mask_width, int32_t *idx, double *score,
                               cudaStream_t st, int64_t *launches);

// m2dp_match.cu
cudaError_t launch_m2dp_match(const double *hist1, int m, const double *hist2, int n, float *d_p,
                              float *d_i, int ldd, void *workspace, cudaStream_t st,
                              int64_t *launches);
size_t m2dp_match_workspace_bytes(int m, int n);
// m2dp_match_tc.cu : tcgen05 path (the product path; the fp32 kernel above is the on-GPU cross-check)
size_t m2dp_match_tc_workspace_bytes(int m, int n);
cudaError_t launch_m2dp_match_tc(const double *hist1, int m, const double *hist2, int n, float *d_p, float *d_i,
                                 int ldd, void *workspace, int num_sms, cudaStream_t st, int64_t *launches);

// fuse_topk.cu
// Peer-memory exchange of a sharded query batch (csrc/sharded.cu): every rank owns a window in HBM that all ranks of
// the box can write over NVLink (cudaIpc / peer access).  A window has two halves (batch parity), each with one slot
// per SOURCE rank; a slot has a FIXED layout (so that stale bytes of an earlier batch shape can never be taken for a
// flag): [sflag: PX_MAX_ROWS u32][lflag: PX_MAX_ROWS u32][stats: PX_MAX_ROWS x 6 f64][lists: 4 x m x k x 8 B].
// Producers write their row's data into the slot `rank` of EVERY rank's window, fence, then write the batch epoch
// into the row's flag; consumers poll their own window.
constexpr int PX_MAX_ROWS = 8192;
constexpr int PX_MAX_RANKS = 16;
constexpr size_t PX_OFF_SFLAG = 0, PX_OFF_LFLAG = (size_t)PX_MAX_ROWS * 4, PX_OFF_STATS = (size_t)PX_MAX_ROWS * 8,
                 PX_OFF_LISTS = PX_OFF_STATS + (size_t)PX_MAX_ROWS * STATS_W * 8;
struct PeerExchange {
  unsigned char *win[PX_MAX_RANKS];   // win[r] = rank r's window as seen from this GPU; nullptr in win[0] = no exchange
  size_t slot_bytes;                  // bytes per (parity, source) slot
  int nranks, rank;
  unsigned epoch;                     // batch number, >= 1
  int *err;                           // device word: set to 1 when a wait timed out
};
cudaError_t launch_row_stats(const float *d_p, const float *d_i, int m, int n, int ldd, double *stats,
                             cudaStream_t st, int64_t *launches, const PeerExchange *px = nullptr);
cudaError_t launch_fuse_topk(const float *d_p, const float *d_i, int m, int n, int ldd,
                             const double *global_stats, int64_t n_global, int64_t q_row0,
                             int64_t db_row0, int mask_width, double p_weight, int k, int64_t *idx,
                             double *score, double *dp_at, double *di_at, cudaStream_t st,
                             int64_t *launches, const PeerExchange *px = nullptr);
// merge of the lists the peers have written into this rank's window (PeerExchange), m x k outputs
cudaError_t launch_topk_merge_px(const PeerExchange &px, int m, int k, int64_t *out_idx, double *out_score,
                                 double *out_d_p, double *out_d_i, cudaStream_t st, int64_t *launches);
// fp64 variants for sodso_fuse_top1 (inputs are caller-supplied fp64 matrices; exact two-pass
// statistics like run_test.m:40)
cudaError_t launch_fuse_top1_f64(const double *d_p, const double *d_i, int m, int n, int mask_width,
                                 double p_weight, int32_t *idx, double *score, cudaStream_t st,
                                 int64_t *launches);
cudaError_t launch_topk_merge(const int64_t *idx, const double *score, const double *d_p, const double *d_i,
                              int nshards, int m, int k, int64_t *out_idx, double *out_score, double *out_d_p,
                              double *out_d_i, cudaStream_t st, int64_t *launches, size_t shard_stride = 0);
cudaError_t launch_gt_loops(const double *gt1, int m, const double *gt2, int n, int mask_width,
                            int32_t *nearest, double *dist2, cudaStream_t st, int64_t *launches);
cudaError_t launch_f32_to_f64(const float *src, int rows, int cols, int ld, double *dst,
                              cudaStream_t st, int64_t *launches);

}  // namespace sodso
