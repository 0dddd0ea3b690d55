/*
 * sodso_pr.h — C ABI of libsodso_pr.so: the B200 (sm_100a) implementation of the
 * descriptor generate + match hot path of IRVLab/so_dso_place_recognition.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference repository root).  The reference has no FFI of its own: its seams are two
 * C++ classes + one free function (generation) and two MATLAB functions + one script
 * section (matching / fusion); see INTEGRATION.md for the bindings a maintainer adds.
 *
 * Conventions
 *  - plain pointers and sizes only; all matrices are row-major.
 *  - every data pointer may be a HOST pointer or a DEVICE pointer (same device as the
 *    context); the library detects which (cudaPointerGetAttributes) and stages host
 *    buffers through its own pinned/HBM workspaces.  Nothing is ever freed for the caller.
 *  - return value: 0 = ok, negative = error (SODSO_E_*); sodso_last_error() gives the text.
 *  - there is NO CPU fallback: without a usable CUDA device every compute call fails.
 *  - calls on one context are stream-ordered on the context's stream and return after that
 *    stream has been synchronised (inputs may be reused, outputs are complete, asynchronous
 *    failures are reported by the call that caused them) -- except the *_sharded calls with
 *    DEVICE output pointers, which only enqueue (see there).  A context is not thread-safe,
 *    use one per thread / per GPU.
 *  - test hooks and the fp32 cross-check kernels are declared in sodso_pr_debug.h, not here.
 */
#ifndef SODSO_PR_H
#define SODSO_PR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SODSO_OK 0
#define SODSO_E_ARG (-1)     /* bad argument */
#define SODSO_E_CUDA (-2)    /* CUDA runtime / driver error */
#define SODSO_E_NODEV (-3)   /* no usable sm_100 device */
#define SODSO_E_STATE (-4)   /* call sequence error (e.g. topk before match) */
#define SODSO_E_NCCL (-5)    /* NCCL error, or libnccl.so.2 could not be bound */

#define SODSO_SC_SIZE 1200   /* SC.h:7-8   numS*numR = 60*20 */
#define SODSO_M2DP_SIZE 192  /* M2DP.cpp:36 numS*numR + numP*numQ = 128 + 64 */

#define SODSO_TYPE_SC 0      /* run_test.m:29 'sc'   */
#define SODSO_TYPE_M2DP 1    /* run_test.m:27 'm2dp' */

typedef struct sodso_ctx sodso_ctx;
typedef struct sodso_db sodso_db;
typedef struct sodso_staged sodso_staged;

/* ---- context ---------------------------------------------------------------------- */
int sodso_ctx_create(int device, sodso_ctx **out);
void sodso_ctx_destroy(sodso_ctx *ctx);
/* The context's stream as a cudaStream_t (void* to keep this header CUDA-free). */
void *sodso_ctx_stream(sodso_ctx *ctx);
/* Use an external stream (e.g. torch's current stream); NULL restores the own stream. */
int sodso_ctx_set_stream(sodso_ctx *ctx, void *cuda_stream);
/* Wait for everything enqueued on the context's stream (the *_sharded calls with DEVICE output pointers return
 * without synchronising). */
int sodso_ctx_sync(sodso_ctx *ctx);
/* sodso_sc_scans_to_loops / sodso_db_stream_match / sodso_db_scans_query_sharded stream HOST point buffers in 512-scan
 * chunks (copy of chunk k+1 overlapped with binning + matching of chunk k) when a call has at least min_scans scans
 * (default 2048, minimum 512); below that one plain copy precedes the compute.  Use page-locked (cudaHostAlloc /
 * cudaHostRegister) buffers: from pageable memory every chunk copy blocks the calling thread and runs at a fraction
 * of the PCIe rate. */
int sodso_ctx_set_stream_threshold(sodso_ctx *ctx, int min_scans);
const char *sodso_last_error(void);
const char *sodso_version(void);
/* Number of kernels of this library launched on the context so far. */
int64_t sodso_ctx_launch_count(sodso_ctx *ctx);
/* Device time (ms, CUDA events on the context's stream) of the dominant kernel of the last
 * generate / match call, and its name. */
double sodso_ctx_last_kernel_ms(sodso_ctx *ctx);
const char *sodso_ctx_last_kernel_name(sodso_ctx *ctx);

/* ---- signature sizes:  SC::getSignatureSize (SC.cpp:10), M2DP::getSignatureSize (M2DP.cpp:36) */
int sodso_sc_signature_size(void);
int sodso_m2dp_signature_size(void);

/* ---- point staging --------------------------------------------------------------- */
/* pts_preprocess(poses_file, pts_file, id_file, lidarRange, out_vec, polar_filter) (pts_preprocess.h:169-232) on
 * the records the two files parse to (PosesPts.h:5-40): pose_id[n_pose], w2c[n_pose x 12] (row-major 3x4),
 * pt_id[n_pts], pt_xyz[n_pts x 3] (world frame), pt_inten[n_pts].  The sequential pose walk (reset rule, the 30
 * skipped frames, which points have entered) is bookkeeping on the host; the per-frame transform + range crop of
 * every accumulated point and the voxel-grid (polar_filter = 0, SC) / 1-degree polar (polar_filter = 1, M2DP)
 * de-duplication run on the GPU.  The staged scans stay in HBM inside the handle; inside a scan the points are
 * ordered by voxel index (the reference's order is the iteration order of a libstdc++ unordered_map; the point set
 * is identical).  sodso_staged_xyz/inten/scan_off are DEVICE pointers that can be handed to sodso_sc_generate /
 * sodso_m2dp_generate / sodso_sc_scans_to_loops; sodso_staged_copy copies out (ids = incoming_id_file.txt). */
int sodso_stage_points(sodso_ctx *ctx, const int32_t *pose_id, const double *w2c, int n_pose,
                       const int32_t *pt_id, const double *pt_xyz, const float *pt_inten, int64_t n_pts,
                       double lidar_range, int polar_filter, sodso_staged **out);
void sodso_staged_destroy(sodso_staged *s);
int sodso_staged_num_scans(sodso_staged *s);
int64_t sodso_staged_num_points(sodso_staged *s);
const double *sodso_staged_xyz(sodso_staged *s);
const float *sodso_staged_inten(sodso_staged *s);
const int64_t *sodso_staged_scan_off(sodso_staged *s);
int sodso_staged_copy(sodso_staged *s, int32_t *ids, int64_t *scan_off, double *xyz, float *inten);

/* ---- generation ------------------------------------------------------------------- */
/* pts_align.h:7-46  align_points_PCA for a batch of scans.
 * xyz: sum(n) x 3 doubles (AoS), scan_off: nscan+1 offsets (in points).  out_xyz same shape.
 * evec (optional): nscan x 9, row-major 3x3, column k = k-th eigenvector (ascending eigenvalue). */
int sodso_align_pca(sodso_ctx *ctx, const double *xyz, const int64_t *scan_off, int nscan,
                    double *out_xyz, double *evec);

/* SC::getSignature (SC.cpp:12-76) over a batch, output in the history_sc layout of
 * test_sc.cpp:52-54: hist row s = [structure(1200), intensity(1200)], nscan x 2400 doubles.
 * inten: sum(n) floats.  max_rho: SC::SC argument (lidarRange, test_sc.cpp:28,35). */
int sodso_sc_generate(sodso_ctx *ctx, const double *xyz, const float *inten,
                      const int64_t *scan_off, int nscan, double max_rho, double *hist);

/* test_m2dp.cpp:41-67 over a batch: PCA once per scan (pts_align.h), the 4 sign variants
 * (dx outer, dy inner), M2DP::getSignature (M2DP.cpp:38-109) for each; output in the
 * history_m2dp layout: rows 4s..4s+3 = [count(192), intensity(192)], 4*nscan x 384 doubles. */
int sodso_m2dp_generate(sodso_ctx *ctx, const double *xyz, const float *inten,
                        const int64_t *scan_off, int nscan, double max_rho, double *hist);

/* M2DP::getSignature (M2DP.cpp:38-109) on ALREADY aligned + sign-flipped points (the class
 * contract, M2DP.h:18-20): sig row s = [count(192), intensity(192)], nscan x 384. */
int sodso_m2dp_signature(sodso_ctx *ctx, const double *xyz_aligned, const float *inten,
                         const int64_t *scan_off, int nscan, double max_rho, double *sig);

/* ---- matching --------------------------------------------------------------------- */
/* [d_p, d_i] = processSC(hist1, hist2)   (processSC.m:1-45).
 * hist1: m x 2400, hist2: n x 2400 doubles; d_p, d_i: m x n (either may be NULL). */
int sodso_sc_match(sodso_ctx *ctx, const double *hist1, int m, const double *hist2, int n,
                   double *d_p, double *d_i);
int sodso_sc_match_f32(sodso_ctx *ctx, const double *hist1, int m, const double *hist2, int n,
                       float *d_p, float *d_i);

/* [d_p, d_i] = processM2DP(hist1, hist2) (processM2DP.m:1-22).
 * hist1: 4m x 384, hist2: 4n x 384 doubles; d_p, d_i: m x n. */
int sodso_m2dp_match(sodso_ctx *ctx, const double *hist1, int m, const double *hist2, int n,
                     double *d_p, double *d_i);
int sodso_m2dp_match_f32(sodso_ctx *ctx, const double *hist1, int m, const double *hist2, int n,
                         float *d_p, float *d_i);

/* run_test.m:38-57 on given distance matrices: fused = p_weight*zscore_row(d_p) +
 * zscore_row(d_i) (std with N-1 over the UNMASKED row), |i-j| < mask_width -> Inf,
 * first-index argmin.  idx: m int32 (0-BASED; MATLAB's diff_idx is 1-based), score: m doubles. */
int sodso_fuse_top1(sodso_ctx *ctx, const double *d_p, const double *d_i, int m, int n,
                    int mask_width, double p_weight, int32_t *idx, double *score);

/* run_test.m:25-57 without materialising the matrices on the host: match (type = SODSO_TYPE_SC
 * -> processSC, SODSO_TYPE_M2DP -> processM2DP), fuse, mask, argmin.  hist1 has m (SC) or 4m
 * (M2DP) rows.  Optional outputs (may be NULL): d_p_at / d_i_at = the two channel distances
 * of the chosen candidate. */
int sodso_loop_top1(sodso_ctx *ctx, int type, const double *hist1, int m, const double *hist2,
                    int n, int mask_width, double p_weight, int32_t *idx, double *score,
                    double *d_p_at, double *d_i_at);

/* test_sc.cpp:36-57 followed by run_test('sc', hist, hist, ...) (run_test.m:25-57; the KITTI self-match of
 * test_kitti.m:28) in one call: scans in, loop candidates out.  hist (optional, nscan x 2400) receives the
 * signatures.  When the point buffers are host memory the scans are streamed: 512-scan chunks are copied on
 * a second stream while earlier chunks are binned and matched (new queries x all DB rows so far, old queries
 * x new DB rows), so the transfer hides behind the tensor-core work.  Results are identical to
 * sodso_sc_generate + sodso_loop_top1. */
int sodso_sc_scans_to_loops(sodso_ctx *ctx, const double *xyz, const float *inten,
                            const int64_t *scan_off, int nscan, double max_rho, int mask_width,
                            double p_weight, double *hist, int32_t *idx, double *score,
                            double *d_p_at, double *d_i_at);

/* ---- DELIGHT -------------------------------------------------------- */
/* DELIGHT::getSignatureSize (DELIGHT.cpp:4) */
int sodso_delight_signature_size(void);
/* DELIGHT::getSignature (DELIGHT.cpp:6-24) over a batch, output in the history_delight layout of
 * test_delight.cpp:42-56: rows 16s..16s+15 = the 16 intensity histograms (8 octants x inside / outside 10 m) of scan s,
 * 16*nscan x 256 doubles.  A point whose int(intensity) falls outside [0, 255] (out of the matrix in the reference) is
 * dropped. */
int sodso_delight_generate(sodso_ctx *ctx, const double *xyz, const float *inten, const int64_t *scan_off,
                           int nscan, double *hist);
/* dist = processDELIGHT(hist1, hist2) (processDELIGHT.m:1-38): hist1 16m x 256, hist2 16n x 256, dist m x n. */
int sodso_delight_match(sodso_ctx *ctx, const double *hist1, int m, const double *hist2, int n, double *dist);
/* run_test.m:47-57 for ONE distance matrix (the delight / gist / bow branch of run_test.m:26-37: no fusion): mask
 * |i - j| < mask_width, first minimum per row (NaN skipped).  idx 0-based. */
int sodso_top1_single(sodso_ctx *ctx, const double *dist, int m, int n, int mask_width, int32_t *idx,
                      double *score);

/* ---- evaluation ------------------------------------------------------ */
/* Ground-truth loop set of run_test.m:3-21.  gt1: m x 3, gt2: n x 3 positions.  nearest[i] = 0-based index of the
 * closest gt2 position with |i - j| >= mask_width (first one on ties, -1 if none); is_loop[i] = 1 if it is closer than
 * loop_diff (optional); *n_loops = number of loops (rows of lp_gt). */
int sodso_gt_loops(sodso_ctx *ctx, const double *gt1, int m, const double *gt2, int n, double loop_diff,
                   int mask_width, int32_t *nearest, int32_t *is_loop, int *n_loops);
/* Precision / recall of run_test.m:56-85 from the per-query decision (diff_v, 0-BASED diff_idx) of
 * sodso_loop_top1 / sodso_fuse_top1: queries ranked by ascending diff_v (stable, NaN last), cumulative true / false
 * positives by ground-truth distance, AUC = trapz(recall, precision), top_recall = recall at the last rank with
 * precision 1, top_count = that rank (the first top_count entries of rank_out are lp_detected(:,1) - 1).
 * n_gt_loops: from sodso_gt_loops (MATLAB's length(lp_gt) quirk for exactly one loop is reproduced).
 * Host pointers only; rank_out / precision_out / recall_out (m each) are optional. */
int sodso_pr_curve(const double *diff_v, const int32_t *diff_idx, const double *gt1, int m, const double *gt2, int n,
                   double loop_diff, int n_gt_loops, double *auc, double *top_recall, int *top_count,
                   int32_t *rank_out, double *precision_out, double *recall_out);

/* ---- resident, row-sharded signature database ------------------------ */
/* A shard holds n_local consecutive DB signatures whose first row has global index
 * global_row0; the signatures stay resident in HBM in MMA operand format.  n_local may be 0 (a shard that is
 * filled by sodso_db_append). */
int sodso_db_create(sodso_ctx *ctx, int type, const double *hist2, int n_local,
                    int64_t global_row0, sodso_db **out);
void sodso_db_destroy(sodso_db *db);
/* Incremental growth (streaming queries against a database that keeps growing, the on-line form of
 * test_sc.cpp:52-54 appending one history row per scan): n_new further signatures behind the shard's last row (global
 * indices global_row0 + n_local ...).  Only the new rows' operands are written; when the capacity runs out it doubles
 * and the operand buffers are re-laid out on the device.  sodso_db_reserve sets the capacity up front. */
int sodso_db_append(sodso_db *db, const double *hist_new, int n_new);
int sodso_db_reserve(sodso_db *db, int capacity);
/* Replace the shard's signatures by n_local new ones (same size, same global_row0): the operand buffers
 * are rewritten in place, nothing is reallocated. */
int sodso_db_reload(sodso_db *db, const double *hist2);
/* sodso_db_reload + sodso_db_match in one streamed pass for a Scan Context shard whose scans are still POINTS:
 * xyz / inten / scan_off describe the shard's n_local scans (test_sc.cpp:36-57 input).  Host buffers are copied in
 * 512-scan chunks on a second stream; each chunk is binned, written into the operand buffers in place and matched
 * against the m query signatures (hist1) as soon as it has landed.  Same state afterwards as reload + match. */
int sodso_db_stream_match(sodso_db *db, const double *xyz, const float *inten, const int64_t *scan_off,
                          double max_rho, const double *hist1, int m);
int sodso_db_size(sodso_db *db);
/* Distances of m queries against the shard (kept on the device inside the handle). */
int sodso_db_match(sodso_db *db, const double *hist1, int m);
/* Per-query partial row statistics over this shard, m x 6 doubles:
 * [sum(d_p-c), sum((d_p-c)^2), count(d_p), sum(d_i-c), sum((d_i-c)^2), count(d_i)] with c = 0.25, over the non-NaN
 * entries (MATLAB's normalize omits NaN: the NaN column of a zero-norm DB signature, processSC.m:15-20, does not
 * poison the row).  Summed over shards (allreduce) they give the row mean / std of run_test.m:40. */
int sodso_db_partial_stats(sodso_db *db, double *stats);
/* Fuse with the GLOBAL stats, mask |q_global - j_global| < mask_width, emit the k best
 * (ascending score, lowest global index first on ties) of this shard:
 * idx (m x k, int64 global; -1 when fewer than k valid), score, d_p, d_i (m x k doubles).
 * q_global_row0: global index of query 0 (run_test.m:47-53 compares row and column numbers). */
int sodso_db_topk(sodso_db *db, const double *global_stats, int64_t n_global,
                  int64_t q_global_row0, int mask_width, double p_weight, int k, int64_t *idx,
                  double *score, double *d_p, double *d_i);
/* Merge R gathered per-shard top-k lists (R x m x k, as produced by an allgather of
 * sodso_db_topk outputs) into the global top-k per query (m x k).  Host-side helper. */
int sodso_topk_merge(const int64_t *idx, const double *score, const double *d_p,
                     const double *d_i, int nshards, int m, int k, int64_t *out_idx,
                     double *out_score, double *out_d_p, double *out_d_i);
/* The same merge on the GPU for gathered lists that live in HBM (what an NCCL all-gather leaves there): device
 * pointers only, at most 16 shards. */
int sodso_topk_merge_device(sodso_ctx *ctx, const int64_t *idx, const double *score, const double *d_p,
                            const double *d_i, int nshards, int m, int k, int64_t *out_idx,
                            double *out_score, double *out_d_p, double *out_d_i);
/* Copy the last sodso_db_match result (fp32, m x n_local each; either may be NULL) out. */
int sodso_db_get_distances(sodso_db *db, float *d_p, float *d_i);

/* ---- multi-GPU: one rank per GPU, NCCL inside the library --------------------------------------
 * The reference is single-process; this is the seam a sharded host (C++ or Python, one thread or process per GPU) binds.
 * Rank 0 calls sodso_comm_unique_id and hands the SODSO_COMM_ID_BYTES bytes to every rank by any means (MPI, a file,
 * torch.distributed); every rank then calls sodso_comm_init on its own context (ncclCommInitRank: collective, blocks
 * until all ranks have joined).  libnccl.so.2 is bound at run time; a process that already has NCCL loaded (PyTorch)
 * shares that copy.  All collectives run on the context's stream. */
#define SODSO_COMM_ID_BYTES 128
int sodso_comm_unique_id(void *id_out);
int sodso_comm_init(sodso_ctx *ctx, const void *unique_id, int nranks, int rank);
int sodso_comm_finalize(sodso_ctx *ctx);
int sodso_comm_nranks(sodso_ctx *ctx);
int sodso_comm_rank(sodso_ctx *ctx);
int sodso_comm_nccl_version(void);   /* e.g. 22809; 0 if NCCL could not be bound */
/* 1 if the per-batch exchange of the sharded queries runs over peer memory: sodso_comm_init also gives every rank a
 * window in HBM that all ranks of the box can write over NVLink (cudaIpc between processes, peer access between threads
 * of one process).  The partial row statistics and the per-shard candidate lists are then written by the producing
 * kernels straight into every rank's window and the consuming kernels poll per-row flags -- no collective launch
 * between the kernels of a batch; the statistics are summed in rank order, so every rank computes identical bits.
 * 0: the windows could not be mapped on some rank (no peer access, IPC not permitted) and the exchange uses
 * ncclAllReduce + ncclAllGather.  The peer-memory exchange is a latency optimisation for streaming batches (per row
 * a system-scope fence and a few NVLink stores; two NCCL collectives cost ~45 us each whatever the batch size):
 * batches of more than 512 queries take the NCCL path, as does the bulk exchange of query signatures. */
int sodso_comm_exchange(sodso_ctx *ctx);

/* run_test.m:25-57 for m queries against a database whose rows are sharded over the ranks of the context's
 * communicator.  COLLECTIVE: every rank calls it with the same queries and parameters and receives the same result,
 * the k best candidates per query over the WHOLE database (ascending fused score, lowest global index first on ties --
 * MATLAB's first minimum, run_test.m:57): idx m x k int64 global 0-based (-1 when fewer than k valid), score / d_p /
 * d_i m x k doubles (d_p / d_i optional).  On the stream, without any host synchronisation in between:
 * match -> partial row statistics -> [exchange 1: sum over the shards] -> fuse + mask + per-shard top-k ->
 * [exchange 2: all shards' lists] -> merge.  The exchanges are peer-memory writes from the producing kernels
 * (sodso_comm_exchange) or ncclAllReduce / ncclAllGather.
 * DEVICE output pointers: the call returns as soon as the work is enqueued (sodso_ctx_sync, or stream order, before
 * the results are read; several batches can be in flight).  HOST output pointers: copied out, one synchronisation.
 * Without a communicator (or nranks == 1) the same path runs on the one shard, without collectives.
 * q_global_row0: global index of query 0 for the temporal mask |q - j| < mask_width (run_test.m:47-53). */
int sodso_db_query_sharded(sodso_db *db, const double *hist1, int m, int64_t q_global_row0, int mask_width,
                           double p_weight, int k, int64_t *idx, double *score, double *d_p, double *d_i);
/* The statistics / top-k / exchange / merge part alone, for the last match of the handle (sodso_db_match,
 * sodso_db_stream_match). */
int sodso_db_finish_sharded(sodso_db *db, int64_t q_global_row0, int mask_width, double p_weight, int k, int64_t *idx,
                            double *score, double *d_p, double *d_i);
/* One step of the sharded pipeline from POINTS (test_sc.cpp:36-57 + run_test.m:25-57), COLLECTIVE:
 *  - queries: this rank bins scans [q_first, q_first + m_slice) of the m_total queries (q_xyz / q_inten / q_off describe
 *    the slice only).  m_slice == m_total: every rank bins all queries, nothing is exchanged.  Otherwise the slices must
 *    be the contiguous block partition of m_total over the ranks (rank r: base = m_total / R rows, the first
 *    m_total % R ranks one more) and the signatures (19 KB per scan instead of 115 KB of points) are exchanged by NCCL;
 *  - shard: db_xyz / db_inten / db_off non-NULL: the shard's n_local scans are binned and its operand is rewritten in
 *    place, streamed chunk-wise from HOST buffers as in sodso_db_stream_match; NULL: the resident operand is used;
 *    If the query pointers ARE the shard pointers (q_xyz == db_xyz, q_inten == db_inten, q_off == db_off, m_slice ==
 *    m_total == n_local: on this rank the queries are the shard's own scans, the KITTI self-match of test_kitti.m:28)
 *    the scans are copied and binned once and matched block-wise while they stream in.  Unlike
 *    sodso_sc_scans_to_loops the sharded entry points compute every pair (the mirrored half of a self-match triangle
 *    does not exist for the other ranks' shards);
 *  - then as sodso_db_query_sharded.  q_hist (optional, m_total x 2400): the query signatures. */
int sodso_db_scans_query_sharded(sodso_db *db, const double *db_xyz, const float *db_inten, const int64_t *db_off,
                                 const double *q_xyz, const float *q_inten, const int64_t *q_off, int m_total,
                                 int q_first, int m_slice, double max_rho, int64_t q_global_row0, int mask_width,
                                 double p_weight, int k, double *q_hist, int64_t *idx, double *score, double *d_p,
                                 double *d_i);

#ifdef __cplusplus
}
#endif
#endif /* SODSO_PR_H */
