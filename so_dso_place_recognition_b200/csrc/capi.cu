// extern "C" surface of libsodso_pr.so (include/sodso_pr.h): context, host<->HBM staging,
// call sequencing.  All compute is in the kernel translation units; there is no CPU path.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "capi_internal.cuh"

namespace sodso {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
const char *last_error_cstr() { return g_err.c_str(); }
DebugFlags g_debug;

}  // namespace sodso

using namespace sodso;

struct sodso_staged {
  sodso_ctx *ctx = nullptr;
  std::vector<int> ids;
  std::vector<int64_t> off;   // nscan + 1
  Buf d_off, d_xyz, d_inten;
};

namespace sodso {

struct TimedRegion {
  sodso_ctx *c;
  TimedRegion(sodso_ctx *ctx, const char *name) : c(ctx) {
    c->kname = name;
    c->ev_valid = cudaEventRecord(c->ev0, c->stream) == cudaSuccess;
  }
  ~TimedRegion() {
    if (c->ev_valid) c->ev_valid = cudaEventRecord(c->ev1, c->stream) == cudaSuccess;
  }
};

int check_offsets_host(const int64_t *off, int nscan, int64_t *total) {
  // off may be host or device; fetch the last element
  int64_t last = 0, first = 0;
  if (is_device_ptr(off)) {
    SODSO_CUDA_CHECK(cudaMemcpy(&last, off + nscan, sizeof(int64_t), cudaMemcpyDeviceToHost));
    SODSO_CUDA_CHECK(cudaMemcpy(&first, off, sizeof(int64_t), cudaMemcpyDeviceToHost));
  } else {
    last = off[nscan];
    first = off[0];
    for (int s = 0; s < nscan; s++)
      if (off[s + 1] < off[s] || off[s + 1] - off[s] > INT32_MAX) {
        set_error("scan_off must be non-decreasing with < 2^31 points per scan");
        return SODSO_E_ARG;
      }
  }
  if (first != 0 || last < 0) {
    set_error("scan_off[0] must be 0");
    return SODSO_E_ARG;
  }
  *total = last;
  return SODSO_OK;
}

// SC operands for `rows` signatures in the format of the selected algorithm
int sc_prepare(sodso_ctx *c, int algo, const double *hist_dev, int rows, Buf &op, bool is_db) {
  if (algo == SODSO_ALGO_SIMT) {
    int ld = (rows + 31) & ~31;
    SODSO_CUDA_CHECK(op.reserve((size_t)2 * SC_SIZE * ld * sizeof(float)));
    SODSO_CUDA_CHECK(launch_sc_prep_simt(hist_dev, rows, op.as<float>(), ld, c->stream, &c->launches));
  } else {
    size_t bytes = is_db ? sc_tc_db_bytes(rows) : sc_tc_query_bytes(rows);
    SODSO_CUDA_CHECK(op.reserve(bytes));
    if (is_db)
      SODSO_CUDA_CHECK(launch_sc_tc_prep_db(hist_dev, rows, op.p, c->stream, &c->launches));
    else
      SODSO_CUDA_CHECK(launch_sc_tc_prep_query(hist_dev, rows, op.p, c->stream, &c->launches));
  }
  return SODSO_OK;
}

// A self-match (hist1 and hist2 are the same n rows) has a symmetric distance matrix: the tcgen05 matcher then computes
// the lower block triangle only and stores every value at its transposed position as well.  The test hook
// sodso_debug_set_sc_symmetry (include/sodso_pr_debug.h) switches this off (every pair computed, as for distinct operands).

int sc_match_core(sodso_ctx *c, int algo, const Buf &q_op, int m, const Buf &db_op, int n, float *dp,
                  float *di, int ldd, bool self = false, int n_layout = 0) {
  TimedRegion tr(c, algo == SODSO_ALGO_SIMT ? "sc_match_simt_kernel" : "sc_match_tc_kernel");
  if (algo == SODSO_ALGO_SIMT) {
    int ldq = (m + 31) & ~31, ldh = (n + 31) & ~31;
    SODSO_CUDA_CHECK(launch_sc_match_simt(q_op.as<float>(), m, ldq, db_op.as<float>(), n, ldh, dp, di,
                                          ldd, c->stream, &c->launches));
  } else if (self && m == n && c->sc_symmetry) {
    SODSO_CUDA_CHECK(launch_sc_match_tc_self(q_op.p, db_op.p, n, 0, n, dp, di, ldd, c->num_sms, c->stream, &c->launches));
  } else {
    SODSO_CUDA_CHECK(launch_sc_match_tc(q_op.p, m, db_op.p, n, dp, di, ldd, c->num_sms, c->stream,
                                        &c->launches, n_layout));
  }
  return SODSO_OK;
}

// distances (fp32, device, ld = n) of hist1 vs hist2 for either descriptor type
int match_to_device(sodso_ctx *c, int type, const double *hist1, int m, const double *hist2, int n,
                    float *dp, float *di) {
  const size_t w = type == SODSO_TYPE_SC ? 2 * SC_SIZE : 2 * M2DP_SIG;
  const size_t r1 = type == SODSO_TYPE_SC ? m : 4 * (size_t)m, r2 = type == SODSO_TYPE_SC ? n : 4 * (size_t)n;
  const double *h1d, *h2d;
  int rc;
  if ((rc = stage_in(c, hist1, r1 * w, c->h1, &h1d))) return rc;
  if ((rc = stage_in(c, hist2, r2 * w, c->h2, &h2d))) return rc;
  if (type == SODSO_TYPE_SC) {
    if ((rc = sc_prepare(c, c->algo, h2d, n, c->db_op, true))) return rc;
    if ((rc = sc_prepare(c, c->algo, h1d, m, c->q_op, false))) return rc;
    return sc_match_core(c, c->algo, c->q_op, m, c->db_op, n, dp, di, n, hist1 == hist2 && m == n);
  }
  if (c->algo == SODSO_ALGO_SIMT) {
    SODSO_CUDA_CHECK(c->m2dp_ws.reserve(m2dp_match_workspace_bytes(m, n)));
    TimedRegion tr(c, "m2dp_match_kernel");
    SODSO_CUDA_CHECK(launch_m2dp_match(h1d, m, h2d, n, dp, di, n, c->m2dp_ws.p, c->stream, &c->launches));
  } else {
    SODSO_CUDA_CHECK(c->m2dp_ws.reserve(m2dp_match_tc_workspace_bytes(m, n)));
    TimedRegion tr(c, "m2dp_match_tc_kernel");
    SODSO_CUDA_CHECK(launch_m2dp_match_tc(h1d, m, h2d, n, dp, di, n, c->m2dp_ws.p, c->num_sms, c->stream, &c->launches));
  }
  return SODSO_OK;
}

template <class T>
int match_api(sodso_ctx *c, int type, const double *hist1, int m, const double *hist2, int n, T *d_p,
              T *d_i) {
  CTX_CHECK(c);
  if (m < 0 || n < 0 || (m > 0 && !hist1) || (n > 0 && !hist2)) {
    set_error("bad match arguments");
    return SODSO_E_ARG;
  }
  if (m == 0 || n == 0) return SODSO_OK;
  const size_t cnt = (size_t)m * n;
  int rc;
  float *dp32, *di32;
  constexpr bool is_f32 = sizeof(T) == 4;
  const bool direct_p = is_f32 && is_device_ptr(d_p), direct_i = is_f32 && is_device_ptr(d_i);
  if (direct_p)
    dp32 = reinterpret_cast<float *>(d_p);
  else {
    SODSO_CUDA_CHECK(c->dp32.reserve(cnt * 4));
    dp32 = c->dp32.as<float>();
  }
  if (direct_i)
    di32 = reinterpret_cast<float *>(d_i);
  else {
    SODSO_CUDA_CHECK(c->di32.reserve(cnt * 4));
    di32 = c->di32.as<float>();
  }
  if ((rc = match_to_device(c, type, hist1, m, hist2, n, dp32, di32))) return rc;
  if constexpr (is_f32) {
    if (d_p && !direct_p && (rc = finish_out(c, reinterpret_cast<float *>(d_p), cnt, dp32))) return rc;
    if (d_i && !direct_i && (rc = finish_out(c, reinterpret_cast<float *>(d_i), cnt, di32))) return rc;
  } else {
    double *o;
    if (d_p) {
      if ((rc = stage_out(c, reinterpret_cast<double *>(d_p), cnt, c->dp64, &o))) return rc;
      SODSO_CUDA_CHECK(launch_f32_to_f64(dp32, m, n, n, o, c->stream, &c->launches));
      if ((rc = finish_out(c, reinterpret_cast<double *>(d_p), cnt, o))) return rc;
    }
    if (d_i) {
      if ((rc = stage_out(c, reinterpret_cast<double *>(d_i), cnt, c->di64, &o))) return rc;
      SODSO_CUDA_CHECK(launch_f32_to_f64(di32, m, n, n, o, c->stream, &c->launches));
      if ((rc = finish_out(c, reinterpret_cast<double *>(d_i), cnt, o))) return rc;
    }
  }
  return sync_ctx(c);
}

int generate_common(sodso_ctx *c, const double *xyz, const float *inten, const int64_t *off, int nscan,
                    const double **xd, const float **id, const int64_t **od) {
  if (nscan < 0 || !off || (nscan > 0 && !xyz)) {
    set_error("bad generate arguments");
    return SODSO_E_ARG;
  }
  int64_t total = 0;
  int rc;
  if ((rc = check_offsets_host(off, nscan, &total))) return rc;
  if ((rc = stage_in(c, xyz, (size_t)total * 3, c->in_xyz, xd))) return rc;
  if (inten && (rc = stage_in(c, inten, (size_t)total, c->in_inten, id))) return rc;
  if ((rc = stage_in(c, off, (size_t)nscan + 1, c->in_off, od))) return rc;
  return SODSO_OK;
}

}  // namespace sodso

extern "C" {

const char *sodso_last_error(void) { return g_err.c_str(); }
const char *sodso_version(void) { return "sodso_pr 0.1 (sm_100a)"; }
int sodso_sc_signature_size(void) { return SC_SIZE; }
int sodso_m2dp_signature_size(void) { return M2DP_SIG; }

int sodso_ctx_create(int device, sodso_ctx **out) {
  if (!out) {
    set_error("out is null");
    return SODSO_E_ARG;
  }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) +
              " — libsodso_pr has no CPU fallback");
    return SODSO_E_NODEV;
  }
  if (device < 0 || device >= ndev) {
    set_error("device index out of range");
    return SODSO_E_ARG;
  }
  cudaDeviceProp prop;
  SODSO_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error(std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major * 10 + prop.minor) +
              "; libsodso_pr is built for sm_100a only");
    return SODSO_E_NODEV;
  }
  SODSO_CUDA_CHECK(cudaSetDevice(device));
  sodso_ctx *c = new sodso_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  SODSO_CUDA_CHECK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  SODSO_CUDA_CHECK(cudaEventCreate(&c->ev0));
  SODSO_CUDA_CHECK(cudaEventCreate(&c->ev1));
  *out = c;
  return SODSO_OK;
}

void sodso_ctx_destroy(sodso_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->xchg_stream) cudaStreamSynchronize(c->xchg_stream);
  comm_release(c);
  for (Buf *b : {&c->in_xyz, &c->in_inten, &c->in_off, &c->out_hist, &c->out_xyz, &c->out_evec, &c->h1,
                 &c->h2, &c->q_op, &c->db_op, &c->dp32, &c->di32, &c->dp64, &c->di64, &c->stats,
                 &c->idx64, &c->idx32, &c->score, &c->dpat, &c->diat, &c->gen_ws, &c->m2dp_ws})
    b->release();
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->own_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->xchg_stream) cudaStreamDestroy(c->xchg_stream);
  for (int f = 0; f < 2; f++) {
    if (c->ev_stats[f]) cudaEventDestroy(c->ev_stats[f]);
    if (c->ev_done[f]) cudaEventDestroy(c->ev_done[f]);
  }
  delete c;
}

void *sodso_ctx_stream(sodso_ctx *c) { return c ? (void *)c->stream : nullptr; }

int sodso_ctx_set_stream(sodso_ctx *c, void *s) {
  if (!c) return SODSO_E_ARG;
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return SODSO_OK;
}

int sodso_ctx_sync(sodso_ctx *c) {
  CTX_CHECK(c);
  int rc = sync_ctx(c);
  return rc ? rc : comm_check(c);
}

int sodso_ctx_set_stream_threshold(sodso_ctx *c, int min_scans) {
  if (!c || min_scans < 0) {
    set_error("bad stream threshold");
    return SODSO_E_ARG;
  }
  c->stream_min_scans = std::max(min_scans, 512);
  return SODSO_OK;
}

// ---- include/sodso_pr_debug.h: test hooks, not part of the reference surface ----
int sodso_debug_set_match_algo(sodso_ctx *c, int algo) {
  if (!c || (algo != SODSO_ALGO_TC && algo != SODSO_ALGO_SIMT)) {
    set_error("bad algo");
    return SODSO_E_ARG;
  }
  c->algo = algo;
  return SODSO_OK;
}

int sodso_debug_set_sc_symmetry(sodso_ctx *c, int on) {
  if (!c) return SODSO_E_ARG;
  c->sc_symmetry = on != 0;
  return SODSO_OK;
}

int sodso_debug_set_kernel_flags(int tc_flags, int gen_flags, int gen_ctas) {
  g_debug.tc_flags = tc_flags;
  g_debug.gen_flags = gen_flags < 0 ? 2 : gen_flags;
  g_debug.gen_ctas = gen_ctas;
  return SODSO_OK;
}

int sodso_debug_phase_profile(int enable, unsigned long long *out16) {
  if (enable && !g_debug.prof) {
    SODSO_CUDA_CHECK(cudaMalloc(&g_debug.prof, 16 * sizeof(unsigned long long)));
    SODSO_CUDA_CHECK(cudaMemset(g_debug.prof, 0, 16 * sizeof(unsigned long long)));
  }
  if (out16) {
    if (!g_debug.prof) return SODSO_E_ARG;
    SODSO_CUDA_CHECK(cudaDeviceSynchronize());
    SODSO_CUDA_CHECK(cudaMemcpy(out16, g_debug.prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    SODSO_CUDA_CHECK(cudaMemset(g_debug.prof, 0, 16 * sizeof(unsigned long long)));
  }
  if (!enable && g_debug.prof) {
    cudaFree(g_debug.prof);
    g_debug.prof = nullptr;
  }
  return SODSO_OK;
}

int64_t sodso_ctx_launch_count(sodso_ctx *c) { return c ? c->launches : 0; }

double sodso_ctx_last_kernel_ms(sodso_ctx *c) {
  if (!c || !c->ev_valid) return -1.0;
  cudaSetDevice(c->device);
  if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}

const char *sodso_ctx_last_kernel_name(sodso_ctx *c) { return c ? c->kname.c_str() : ""; }

int sodso_align_pca(sodso_ctx *c, const double *xyz, const int64_t *off, int nscan, double *out_xyz,
                    double *evec) {
  CTX_CHECK(c);
  const double *xd;
  const float *id = nullptr;
  const int64_t *od;
  int rc;
  if ((rc = generate_common(c, xyz, nullptr, off, nscan, &xd, &id, &od))) return rc;
  if (nscan == 0) return SODSO_OK;
  int64_t total = 0;
  if ((rc = check_offsets_host(off, nscan, &total))) return rc;
  double *o, *ev;
  if ((rc = stage_out(c, out_xyz, (size_t)total * 3, c->out_xyz, &o))) return rc;
  if ((rc = stage_out(c, evec, (size_t)nscan * 9, c->out_evec, &ev))) return rc;
  if (!o) {
    set_error("out_xyz is null");
    return SODSO_E_ARG;
  }
  {
    TimedRegion tr(c, "align_pca_kernel");
    SODSO_CUDA_CHECK(launch_align_pca(xd, od, nscan, o, ev, c->num_sms, c->stream, &c->launches));
  }
  if ((rc = finish_out(c, out_xyz, (size_t)total * 3, o))) return rc;
  if ((rc = finish_out(c, evec, (size_t)nscan * 9, ev))) return rc;
  return sync_ctx(c);
}

int sodso_sc_generate(sodso_ctx *c, const double *xyz, const float *inten, const int64_t *off,
                      int nscan, double max_rho, double *hist) {
  CTX_CHECK(c);
  const double *xd;
  const float *id = nullptr;
  const int64_t *od;
  int rc;
  if (nscan > 0 && (!inten || !hist)) {
    set_error("inten / hist is null");
    return SODSO_E_ARG;
  }
  if ((rc = generate_common(c, xyz, inten, off, nscan, &xd, &id, &od))) return rc;
  if (nscan == 0) return SODSO_OK;
  double *h;
  const size_t cnt = (size_t)nscan * 2 * SC_SIZE;
  if ((rc = stage_out(c, hist, cnt, c->out_hist, &h))) return rc;
  {
    TimedRegion tr(c, "sc_generate_kernel");
    SODSO_CUDA_CHECK(launch_sc_generate(xd, id, od, nscan, max_rho, h, c->num_sms, c->stream, &c->launches));
  }
  if ((rc = finish_out(c, hist, cnt, h))) return rc;
  return sync_ctx(c);
}

static int m2dp_generate_impl(sodso_ctx *c, const double *xyz, const float *inten, const int64_t *off,
                              int nscan, double max_rho, double *hist, bool variants) {
  CTX_CHECK(c);
  const double *xd;
  const float *id = nullptr;
  const int64_t *od;
  int rc;
  if (nscan > 0 && (!inten || !hist)) {
    set_error("inten / hist is null");
    return SODSO_E_ARG;
  }
  if ((rc = generate_common(c, xyz, inten, off, nscan, &xd, &id, &od))) return rc;
  if (nscan == 0) return SODSO_OK;
  double *h;
  const size_t cnt = (size_t)nscan * (variants ? 4 : 1) * 2 * M2DP_SIG;
  if ((rc = stage_out(c, hist, cnt, c->out_hist, &h))) return rc;
  const size_t wsb = m2dp_generate_workspace_bytes(nscan, variants);
  SODSO_CUDA_CHECK(c->gen_ws.reserve(wsb));
  {
    TimedRegion tr(c, "m2dp_hist_kernel");
    SODSO_CUDA_CHECK(launch_m2dp_generate(xd, id, od, nscan, max_rho, variants, h, c->gen_ws.p, wsb,
                                          c->num_sms, c->stream, &c->launches));
  }
  if ((rc = finish_out(c, hist, cnt, h))) return rc;
  return sync_ctx(c);
}

int sodso_m2dp_generate(sodso_ctx *c, const double *xyz, const float *inten, const int64_t *off,
                        int nscan, double max_rho, double *hist) {
  return m2dp_generate_impl(c, xyz, inten, off, nscan, max_rho, hist, true);
}

int sodso_m2dp_signature(sodso_ctx *c, const double *xyz, const float *inten, const int64_t *off,
                         int nscan, double max_rho, double *sig) {
  return m2dp_generate_impl(c, xyz, inten, off, nscan, max_rho, sig, false);
}

int sodso_sc_match(sodso_ctx *c, const double *h1, int m, const double *h2, int n, double *d_p,
                   double *d_i) {
  return match_api<double>(c, SODSO_TYPE_SC, h1, m, h2, n, d_p, d_i);
}
int sodso_sc_match_f32(sodso_ctx *c, const double *h1, int m, const double *h2, int n, float *d_p,
                       float *d_i) {
  return match_api<float>(c, SODSO_TYPE_SC, h1, m, h2, n, d_p, d_i);
}
int sodso_m2dp_match(sodso_ctx *c, const double *h1, int m, const double *h2, int n, double *d_p,
                     double *d_i) {
  return match_api<double>(c, SODSO_TYPE_M2DP, h1, m, h2, n, d_p, d_i);
}
int sodso_m2dp_match_f32(sodso_ctx *c, const double *h1, int m, const double *h2, int n, float *d_p,
                         float *d_i) {
  return match_api<float>(c, SODSO_TYPE_M2DP, h1, m, h2, n, d_p, d_i);
}

int sodso_fuse_top1(sodso_ctx *c, const double *d_p, const double *d_i, int m, int n, int mask_width,
                    double p_weight, int32_t *idx, double *score) {
  CTX_CHECK(c);
  if (m < 0 || n <= 0 || !d_p || !d_i || !idx || !score) {
    set_error("bad fuse_top1 arguments");
    return SODSO_E_ARG;
  }
  if (m == 0) return SODSO_OK;
  const double *pd, *qd;
  int rc;
  const size_t cnt = (size_t)m * n;
  if ((rc = stage_in(c, d_p, cnt, c->dp64, &pd))) return rc;
  if ((rc = stage_in(c, d_i, cnt, c->di64, &qd))) return rc;
  int32_t *id;
  double *sd;
  if ((rc = stage_out(c, idx, (size_t)m, c->idx32, &id))) return rc;
  if ((rc = stage_out(c, score, (size_t)m, c->score, &sd))) return rc;
  {
    TimedRegion tr(c, "fuse_top1_f64_kernel");
    SODSO_CUDA_CHECK(launch_fuse_top1_f64(pd, qd, m, n, mask_width, p_weight, id, sd, c->stream, &c->launches));
  }
  if ((rc = finish_out(c, idx, (size_t)m, id))) return rc;
  if ((rc = finish_out(c, score, (size_t)m, sd))) return rc;
  return sync_ctx(c);
}

static int fuse_top1_tail(sodso_ctx *c, int m, int n, int mask_width, double p_weight, int32_t *idx,
                          double *score, double *d_p_at, double *d_i_at);

int sodso_loop_top1(sodso_ctx *c, int type, const double *hist1, int m, const double *hist2, int n,
                    int mask_width, double p_weight, int32_t *idx, double *score, double *d_p_at,
                    double *d_i_at) {
  CTX_CHECK(c);
  if ((type != SODSO_TYPE_SC && type != SODSO_TYPE_M2DP) || m < 0 || n <= 0 || !hist1 || !hist2 || !idx ||
      !score) {
    set_error("bad loop_top1 arguments");
    return SODSO_E_ARG;
  }
  if (m == 0) return SODSO_OK;
  const size_t cnt = (size_t)m * n;
  int rc;
  SODSO_CUDA_CHECK(c->dp32.reserve(cnt * 4));
  SODSO_CUDA_CHECK(c->di32.reserve(cnt * 4));
  if ((rc = match_to_device(c, type, hist1, m, hist2, n, c->dp32.as<float>(), c->di32.as<float>()))) return rc;
  return fuse_top1_tail(c, m, n, mask_width, p_weight, idx, score, d_p_at, d_i_at);
}

// shared tail of sodso_loop_top1 / sodso_sc_scans_to_loops: row statistics, fusion, top-1, outputs
static int fuse_top1_tail(sodso_ctx *c, int m, int n, int mask_width, double p_weight, int32_t *idx,
                          double *score, double *d_p_at, double *d_i_at) {
  int rc;
  SODSO_CUDA_CHECK(c->stats.reserve((size_t)m * STATS_W * sizeof(double)));
  SODSO_CUDA_CHECK(launch_row_stats(c->dp32.as<float>(), c->di32.as<float>(), m, n, n, c->stats.as<double>(),
                                    c->stream, &c->launches));
  SODSO_CUDA_CHECK(c->idx64.reserve((size_t)m * sizeof(int64_t)));
  double *sd, *pa, *ia;
  if ((rc = stage_out(c, score, (size_t)m, c->score, &sd))) return rc;
  if ((rc = stage_out(c, d_p_at, (size_t)m, c->dpat, &pa))) return rc;
  if ((rc = stage_out(c, d_i_at, (size_t)m, c->diat, &ia))) return rc;
  SODSO_CUDA_CHECK(launch_fuse_topk(c->dp32.as<float>(), c->di32.as<float>(), m, n, n, c->stats.as<double>(), n,
                                    0, 0, mask_width, p_weight, 1, c->idx64.as<int64_t>(), sd, pa, ia,
                                    c->stream, &c->launches));
  // idx: int64 (-1 = none) -> int32 0-based (MATLAB returns the first index for an all-NaN row)
  std::vector<int64_t> tmp((size_t)m);
  SODSO_CUDA_CHECK(cudaMemcpyAsync(tmp.data(), c->idx64.p, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                   c->stream));
  if ((rc = finish_out(c, score, (size_t)m, sd))) return rc;
  if ((rc = finish_out(c, d_p_at, (size_t)m, pa))) return rc;
  if ((rc = finish_out(c, d_i_at, (size_t)m, ia))) return rc;
  if ((rc = sync_ctx(c))) return rc;
  std::vector<int32_t> t32((size_t)m);
  for (int i = 0; i < m; i++) t32[i] = tmp[i] < 0 ? 0 : (int32_t)tmp[i];
  if (is_device_ptr(idx))
    SODSO_CUDA_CHECK(cudaMemcpy(idx, t32.data(), (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice));
  else
    std::memcpy(idx, t32.data(), (size_t)m * sizeof(int32_t));
  return SODSO_OK;
}

int sodso_sc_scans_to_loops(sodso_ctx *c, const double *xyz, const float *inten, const int64_t *off, int nscan,
                            double max_rho, int mask_width, double p_weight, double *hist, int32_t *idx,
                            double *score, double *d_p_at, double *d_i_at) {
  CTX_CHECK(c);
  if (nscan <= 0 || !xyz || !inten || !off || !idx || !score) {
    set_error("bad scans_to_loops arguments");
    return SODSO_E_ARG;
  }
  if (c->algo != SODSO_ALGO_TC) {
    set_error("sodso_sc_scans_to_loops needs the tensor-core match (SODSO_ALGO_TC)");
    return SODSO_E_STATE;
  }
  int rc;
  int64_t total = 0;
  if ((rc = check_offsets_host(off, nscan, &total))) return rc;
  const bool host_pts = !is_device_ptr(xyz), host_int = !is_device_ptr(inten), host_off = !is_device_ptr(off);
  // scans are streamed in chunks (a multiple of the 256-row DB tile) when the points live in host memory
  // chunk boundaries (multiples of the 256-row DB tile): two 256-scan chunks first so that matching starts early --
  // the work that can be done grows with the square of what has arrived -- then 512-scan chunks
  const int CH = 512;
  const bool streamed = host_pts && host_int && host_off && nscan >= c->stream_min_scans;
  std::vector<int> bounds{0};
  if (streamed) {
    for (int b = 256; b < nscan; b += (b < 512 || nscan - b <= 1024) ? 256 : CH) bounds.push_back(b);
  }
  bounds.push_back(nscan);
  const int nchunk = (int)bounds.size() - 1;

  const double *xd = xyz;
  const float *id = inten;
  const int64_t *od = off;
  if (host_pts) {
    SODSO_CUDA_CHECK(c->in_xyz.reserve((size_t)total * 3 * sizeof(double)));
    xd = c->in_xyz.as<double>();
  }
  if (host_int) {
    SODSO_CUDA_CHECK(c->in_inten.reserve((size_t)total * sizeof(float)));
    id = c->in_inten.as<float>();
  }
  if ((rc = stage_in(c, off, (size_t)nscan + 1, c->in_off, &od))) return rc;
  double *hd;
  const size_t hcnt = (size_t)nscan * 2 * SC_SIZE;
  if (hist && is_device_ptr(hist)) {
    hd = hist;
  } else {  // host output, or no signature output requested: they are still needed on the device
    SODSO_CUDA_CHECK(c->out_hist.reserve(hcnt * sizeof(double) + 16));
    hd = c->out_hist.as<double>();
  }
  const size_t cnt = (size_t)nscan * nscan;
  SODSO_CUDA_CHECK(c->dp32.reserve(cnt * 4));
  SODSO_CUDA_CHECK(c->di32.reserve(cnt * 4));
  SODSO_CUDA_CHECK(c->db_op.reserve(sc_tc_db_bytes(nscan)));
  SODSO_CUDA_CHECK(c->q_op.reserve(sc_tc_query_bytes(nscan)));
  SODSO_CUDA_CHECK(launch_sc_tc_clear_flags(c->db_op.p, c->stream));
  SODSO_CUDA_CHECK(launch_sc_tc_clear_flags(c->q_op.p, c->stream));
  float *dp = c->dp32.as<float>(), *di = c->di32.as<float>();

  // events of the chunk copies: destroyed on every exit path
  struct Events {
    std::vector<cudaEvent_t> v;
    ~Events() {
      for (cudaEvent_t e : v)
        if (e) cudaEventDestroy(e);
    }
  } evs;
  auto enqueue_copy = [&](int k) -> cudaError_t {   // chunk k of the host buffers -> HBM, on the copy stream
    const int s0 = bounds[k], s1 = bounds[k + 1];
    const int64_t p0 = off[s0], p1 = off[s1];
    cudaError_t e = cudaSuccess;
    if (p1 > p0) {
      e = cudaMemcpyAsync(c->in_xyz.as<double>() + 3 * p0, xyz + 3 * p0, (size_t)(p1 - p0) * 3 * sizeof(double),
                          cudaMemcpyHostToDevice, c->copy_stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(c->in_inten.as<float>() + p0, inten + p0, (size_t)(p1 - p0) * sizeof(float),
                            cudaMemcpyHostToDevice, c->copy_stream);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&evs.v[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(evs.v[k], c->copy_stream);
    return e;
  };
  if (streamed) {
    if (!c->copy_stream) SODSO_CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    evs.v.assign(nchunk, nullptr);
    // the staging buffers may still be read by earlier work on the compute stream
    cudaEvent_t e0;
    SODSO_CUDA_CHECK(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(e0, c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->copy_stream, e0, 0);
    cudaEventDestroy(e0);
    SODSO_CUDA_CHECK(e);
    // copies are enqueued one chunk ahead of the compute below: with pageable host memory cudaMemcpyAsync returns only
    // once the chunk has been staged, and the previous chunk's kernels are already in the queue by then
    SODSO_CUDA_CHECK(enqueue_copy(0));
  } else {
    if (host_pts)
      SODSO_CUDA_CHECK(cudaMemcpyAsync(c->in_xyz.p, xyz, (size_t)total * 3 * sizeof(double), cudaMemcpyHostToDevice,
                                       c->stream));
    if (host_int)
      SODSO_CUDA_CHECK(cudaMemcpyAsync(c->in_inten.p, inten, (size_t)total * sizeof(float), cudaMemcpyHostToDevice,
                                       c->stream));
  }

  const int n_pad = sc_tc_db_rows_padded(nscan), m_pad = sc_tc_query_rows_padded(nscan);
  rc = SODSO_OK;
  c->kname = "sc_match_tc_kernel";
  c->ev_valid = false;
  for (int k = 0; k < nchunk && rc == SODSO_OK; k++) {
    const int s0 = bounds[k], s1 = bounds[k + 1];
    const bool last = k == nchunk - 1;
    cudaError_t e = cudaSuccess;
    if (streamed && !last) e = enqueue_copy(k + 1);
    if (streamed && e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, evs.v[k], 0);
    // test_sc.cpp:40-57 for the scans of this chunk
    if (e == cudaSuccess)
      e = launch_sc_generate(xd, id, od + s0, s1 - s0, max_rho, hd + (size_t)s0 * 2 * SC_SIZE, c->num_sms, c->stream,
                             &c->launches);
    // their rows of both match operands (the last chunk also writes the zero padding)
    if (e == cudaSuccess)
      e = launch_sc_tc_prep_db_rows(hd, nscan, s0, last ? n_pad : s1, c->db_op.p, c->stream, &c->launches);
    if (e == cudaSuccess)
      e = launch_sc_tc_prep_query_rows(hd, nscan, s0, last ? m_pad : s1, c->q_op.p, c->stream, &c->launches);
    // processSC.m:22-33 for every (query, DB) pair that has become available: new queries x all DB rows so far,
    // old queries x new DB rows
    if (!streamed && e == cudaSuccess) c->ev_valid = cudaEventRecord(c->ev0, c->stream) == cudaSuccess;
    if (e == cudaSuccess) {
      if (c->sc_symmetry)   // self-match: the new queries against the DB rows up to their own block, transposes stored
        e = launch_sc_match_tc_self(c->q_op.p, c->db_op.p, nscan, s0, s1, dp, di, nscan, c->num_sms, c->stream,
                                    &c->launches);
      else                         // one launch for the L-shaped region
        e = launch_sc_match_tc_blocks(c->q_op.p, nscan, c->db_op.p, nscan, s0, s1, 0, s1, 0, s0, s0, s1, dp, di, nscan,
                                      c->num_sms, c->stream, &c->launches);
    }
    if (!streamed && c->ev_valid) c->ev_valid = cudaEventRecord(c->ev1, c->stream) == cudaSuccess;
    if (e != cudaSuccess) {
      set_error(std::string("scans_to_loops: ") + cudaGetErrorString(e));
      rc = SODSO_E_CUDA;
    }
  }
  if (rc) {
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    return rc;
  }
  if (hist && hist != hd && (rc = finish_out(c, hist, hcnt, hd))) return rc;
  return fuse_top1_tail(c, nscan, nscan, mask_width, p_weight, idx, score, d_p_at, d_i_at);
}

// ---- DELIGHT --------------------------------------------------------------------
int sodso_delight_signature_size(void) { return 256; }

int sodso_delight_generate(sodso_ctx *c, const double *xyz, const float *inten, const int64_t *off, int nscan,
                           double *hist) {
  CTX_CHECK(c);
  const double *xd;
  const float *id = nullptr;
  const int64_t *od;
  int rc;
  if (nscan > 0 && (!inten || !hist)) {
    set_error("inten / hist is null");
    return SODSO_E_ARG;
  }
  if ((rc = generate_common(c, xyz, inten, off, nscan, &xd, &id, &od))) return rc;
  if (nscan == 0) return SODSO_OK;
  double *h;
  const size_t cnt = (size_t)nscan * 16 * 256;
  if ((rc = stage_out(c, hist, cnt, c->out_hist, &h))) return rc;
  {
    TimedRegion tr(c, "delight_generate_kernel");
    SODSO_CUDA_CHECK(launch_delight_generate(xd, id, od, nscan, h, c->num_sms, c->stream, &c->launches));
  }
  if ((rc = finish_out(c, hist, cnt, h))) return rc;
  return sync_ctx(c);
}

int sodso_delight_match(sodso_ctx *c, const double *hist1, int m, const double *hist2, int n, double *dist) {
  CTX_CHECK(c);
  if (m < 0 || n < 0 || (m > 0 && !hist1) || (n > 0 && !hist2) || !dist) {
    set_error("bad delight_match arguments");
    return SODSO_E_ARG;
  }
  if (m == 0 || n == 0) return SODSO_OK;
  const double *h1d, *h2d;
  double *dd;
  int rc;
  if ((rc = stage_in(c, hist1, (size_t)m * 16 * 256, c->h1, &h1d))) return rc;
  if ((rc = stage_in(c, hist2, (size_t)n * 16 * 256, c->h2, &h2d))) return rc;
  if ((rc = stage_out(c, dist, (size_t)m * n, c->dp64, &dd))) return rc;
  SODSO_CUDA_CHECK(c->m2dp_ws.reserve(delight_match_workspace_bytes(m, n)));
  {
    TimedRegion tr(c, "delight_match_kernel");
    SODSO_CUDA_CHECK(launch_delight_match(h1d, m, h2d, n, dd, c->m2dp_ws.p, c->stream, &c->launches));
  }
  if ((rc = finish_out(c, dist, (size_t)m * n, dd))) return rc;
  return sync_ctx(c);
}

int sodso_top1_single(sodso_ctx *c, const double *dist, int m, int n, int mask_width, int32_t *idx, double *score) {
  CTX_CHECK(c);
  if (m < 0 || n <= 0 || !dist || !idx || !score) {
    set_error("bad top1_single arguments");
    return SODSO_E_ARG;
  }
  if (m == 0) return SODSO_OK;
  const double *dd;
  int32_t *id;
  double *sd;
  int rc;
  if ((rc = stage_in(c, dist, (size_t)m * n, c->dp64, &dd))) return rc;
  if ((rc = stage_out(c, idx, (size_t)m, c->idx32, &id))) return rc;
  if ((rc = stage_out(c, score, (size_t)m, c->score, &sd))) return rc;
  SODSO_CUDA_CHECK(launch_top1_single(dd, m, n, mask_width, id, sd, c->stream, &c->launches));
  if ((rc = finish_out(c, idx, (size_t)m, id))) return rc;
  if ((rc = finish_out(c, score, (size_t)m, sd))) return rc;
  return sync_ctx(c);
}

// ---- evaluation (run_test.m:2-22, 58-85) -------------------------------------------------------
int sodso_gt_loops(sodso_ctx *c, const double *gt1, int m, const double *gt2, int n, double loop_diff,
                   int mask_width, int32_t *nearest, int32_t *is_loop, int *n_loops) {
  CTX_CHECK(c);
  if (m < 0 || n < 0 || (m > 0 && !gt1) || (n > 0 && !gt2) || !nearest) {
    set_error("bad gt_loops arguments");
    return SODSO_E_ARG;
  }
  if (n_loops) *n_loops = 0;
  if (m == 0) return SODSO_OK;
  const double *g1, *g2;
  int rc;
  if ((rc = stage_in(c, gt1, (size_t)m * 3, c->h1, &g1))) return rc;
  if ((rc = stage_in(c, gt2, (size_t)n * 3, c->h2, &g2))) return rc;
  SODSO_CUDA_CHECK(c->idx32.reserve((size_t)m * 4));
  SODSO_CUDA_CHECK(c->score.reserve((size_t)m * 8));
  SODSO_CUDA_CHECK(launch_gt_loops(g1, m, g2, n, mask_width, c->idx32.as<int32_t>(), c->score.as<double>(), c->stream,
                                   &c->launches));
  std::vector<int32_t> near((size_t)m);
  std::vector<double> d2((size_t)m);
  SODSO_CUDA_CHECK(cudaMemcpyAsync(near.data(), c->idx32.p, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
  SODSO_CUDA_CHECK(cudaMemcpyAsync(d2.data(), c->score.p, (size_t)m * 8, cudaMemcpyDeviceToHost, c->stream));
  if ((rc = sync_ctx(c))) return rc;
  int cnt = 0;
  std::vector<int32_t> flag((size_t)m);
  for (int i = 0; i < m; i++) {
    flag[(size_t)i] = d2[(size_t)i] < loop_diff * loop_diff ? 1 : 0;   // run_test.m:19
    cnt += flag[(size_t)i];
  }
  auto put = [&](int32_t *dst, const std::vector<int32_t> &src) -> cudaError_t {
    if (!dst) return cudaSuccess;
    if (is_device_ptr(dst)) return cudaMemcpy(dst, src.data(), src.size() * 4, cudaMemcpyHostToDevice);
    std::memcpy(dst, src.data(), src.size() * 4);
    return cudaSuccess;
  };
  SODSO_CUDA_CHECK(put(nearest, near));
  SODSO_CUDA_CHECK(put(is_loop, flag));
  if (n_loops) *n_loops = cnt;
  return SODSO_OK;
}

// run_test.m:56-85 from the per-query decision (host pointers; sequential cumulative counts)
int sodso_pr_curve(const double *diff_v, const int32_t *diff_idx, const double *gt1, int m, const double *gt2, int n,
                   double loop_diff, int n_gt_loops, double *auc, double *top_recall, int *top_count,
                   int32_t *rank_out, double *precision_out, double *recall_out) {
  if (m < 0 || (m > 0 && (!diff_v || !diff_idx || !gt1 || !gt2))) {
    set_error("bad pr_curve arguments");
    return SODSO_E_ARG;
  }
  for (const void *p : {(const void *)diff_v, (const void *)diff_idx, (const void *)gt1, (const void *)gt2})
    if (p && is_device_ptr(p)) {
      set_error("pr_curve takes host pointers");
      return SODSO_E_ARG;
    }
  // [~, diff_rank] = sort(diff_v): ascending, stable, NaN last (run_test.m:58)
  std::vector<int32_t> rank((size_t)m);
  for (int i = 0; i < m; i++) rank[(size_t)i] = i;
  std::stable_sort(rank.begin(), rank.end(), [&](int32_t a, int32_t b) {
    const double x = diff_v[a], y = diff_v[b];
    const bool xn = x != x, yn = y != y;
    if (xn || yn) return !xn && yn;
    return x < y;
  });
  // length(lp_gt) of an L x 2 matrix (run_test.m:22): L, except that one loop reports 2 and none reports 0
  const double total_lp = n_gt_loops == 1 ? 2.0 : (double)n_gt_loops;
  double tp = 0, fp = 0, a = 0.0, tr = 0.0, prev_p = 0.0, prev_r = 0.0;
  int tc = 0;
  for (int i = 0; i < m; i++) {
    const int32_t q = rank[(size_t)i], b = diff_idx[q];
    bool hit = false;
    if (b >= 0 && b < n) {
      const double dx = gt1[3 * (size_t)q] - gt2[3 * (size_t)b], dy = gt1[3 * (size_t)q + 1] - gt2[3 * (size_t)b + 1],
                   dz = gt1[3 * (size_t)q + 2] - gt2[3 * (size_t)b + 2];
      hit = (dx * dx + dy * dy) + dz * dz < loop_diff * loop_diff;   // :68-70
    }
    if (hit) tp += 1; else fp += 1;                                  // :70-74
    const double pr = tp / (tp + fp), rc = tp / total_lp;            // :75-76
    if (precision_out) precision_out[i] = pr;
    if (recall_out) recall_out[i] = rc;
    if (pr == 1.0) {                                                 // :78-81
      tc = i + 1;
      tr = rc;
    }
    if (i > 0) a += (rc - prev_r) * (pr + prev_p) * 0.5;             // trapz(recall, precision), :84
    prev_p = pr;
    prev_r = rc;
  }
  if (rank_out) std::memcpy(rank_out, rank.data(), (size_t)m * 4);
  if (auc) *auc = a;
  if (top_recall) *top_recall = tr;
  if (top_count) *top_count = tc;
  return SODSO_OK;
}

// test hook (see include/sodso_pr.h)
int64_t sodso_debug_sc_self_items(int64_t n, int64_t q0, int64_t q1) {
  if (n > INT32_MAX || q0 > INT32_MAX || q1 > INT32_MAX) return -1;
  return (int64_t)sc_tc_self_items((int)n, (int)q0, (int)q1);
}

int sodso_debug_fast_turns(sodso_ctx *c, const float *num, const float *den, int64_t n, float *out) {
  CTX_CHECK(c);
  if (n < 0 || (n > 0 && (!num || !den || !out))) {
    set_error("bad debug_fast_turns arguments");
    return SODSO_E_ARG;
  }
  if (n == 0) return SODSO_OK;
  const float *dn, *dd;
  float *dout;
  int rc;
  if ((rc = stage_in(c, num, (size_t)n, c->h1, &dn))) return rc;
  if ((rc = stage_in(c, den, (size_t)n, c->h2, &dd))) return rc;
  if ((rc = stage_out(c, out, (size_t)n, c->dp32, &dout))) return rc;
  SODSO_CUDA_CHECK(launch_fast_turns_probe(dn, dd, n, dout, c->stream));
  if ((rc = finish_out(c, out, (size_t)n, dout))) return rc;
  return sync_ctx(c);
}

// ---- point staging (pts_preprocess.h) ----------------------------------------------------------
int sodso_stage_points(sodso_ctx *c, const int32_t *pose_id, const double *w2c, int n_pose, const int32_t *pt_id,
                       const double *pt_xyz, const float *pt_inten, int64_t n_pts, double lidar_range,
                       int polar_filter, sodso_staged **out) {
  CTX_CHECK(c);
  if (!out || n_pose < 0 || n_pts < 0 || (n_pose > 0 && (!pose_id || !w2c)) ||
      (n_pts > 0 && (!pt_id || !pt_xyz || !pt_inten)) || !(lidar_range > 0)) {
    set_error("bad stage_points arguments");
    return SODSO_E_ARG;
  }
  *out = nullptr;
  // the pose walk is sequential bookkeeping on ids and translations: host copies of the small arrays
  std::vector<int32_t> h_pose_id, h_pt_id;
  std::vector<double> h_w2c;
  auto to_host = [&](const void *src, size_t bytes, void *dst) -> cudaError_t {
    if (bytes == 0) return cudaSuccess;
    if (is_device_ptr(src)) return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
    std::memcpy(dst, src, bytes);
    return cudaSuccess;
  };
  h_pose_id.resize((size_t)n_pose);
  h_w2c.resize((size_t)n_pose * 12);
  h_pt_id.resize((size_t)n_pts);
  SODSO_CUDA_CHECK(to_host(pose_id, (size_t)n_pose * 4, h_pose_id.data()));
  SODSO_CUDA_CHECK(to_host(w2c, (size_t)n_pose * 96, h_w2c.data()));
  SODSO_CUDA_CHECK(to_host(pt_id, (size_t)n_pts * 4, h_pt_id.data()));
  StagePlan P;
  stage_plan(h_pose_id.data(), h_w2c.data(), n_pose, h_pt_id.data(), n_pts, P);
  const int nscan = (int)P.frame_of_scan.size();

  sodso_staged *S = new sodso_staged();
  S->ctx = c;
  S->ids = P.ids;
  S->off.assign((size_t)nscan + 1, 0);
  auto fail = [&](int rc) {
    sodso_staged_destroy(S);
    return rc;
  };
  if (nscan == 0 || n_pts == 0) {
    *out = S;
    return SODSO_OK;
  }
  int rc;
  const double *d_pts;
  const float *d_int;
  if ((rc = stage_in(c, pt_xyz, (size_t)n_pts * 3, c->in_xyz, &d_pts))) return fail(rc);
  if ((rc = stage_in(c, pt_inten, (size_t)n_pts, c->in_inten, &d_int))) return fail(rc);
  Buf b_w2c, b_entry, b_sof, b_segend, b_s0, b_s1, b_diff, b_fos, b_lo, b_hi, b_coff, b_win, b_nout, b_ws;
  auto release_all = [&]() {
    for (Buf *b : {&b_w2c, &b_entry, &b_sof, &b_segend, &b_s0, &b_s1, &b_diff, &b_fos, &b_lo, &b_hi, &b_coff, &b_win,
                   &b_nout, &b_ws})
      b->release();
  };
  auto up = [&](Buf &b, const void *src, size_t bytes) -> cudaError_t {
    cudaError_t e = b.reserve(bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    return bytes ? cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
  };
#define ST_CHECK(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      set_error(std::string("stage_points: ") + #expr + ": " + cudaGetErrorString(_e));         \
      cudaStreamSynchronize(c->stream);                                                         \
      release_all();                                                                            \
      return fail(SODSO_E_CUDA);                                                                \
    }                                                                                           \
  } while (0)
  ST_CHECK(up(b_w2c, h_w2c.data(), h_w2c.size() * 8));
  ST_CHECK(up(b_entry, P.entry.data(), P.entry.size() * 4));
  ST_CHECK(up(b_sof, P.scan_of_frame.data(), P.scan_of_frame.size() * 4));
  ST_CHECK(up(b_segend, P.seg_end.data(), P.seg_end.size() * 4));
  ST_CHECK(b_s0.reserve((size_t)n_pts * 4));
  ST_CHECK(b_s1.reserve((size_t)n_pts * 4));
  ST_CHECK(b_diff.reserve(((size_t)nscan + 1) * 4));
  ST_CHECK(cudaMemsetAsync(b_diff.p, 0, ((size_t)nscan + 1) * 4, c->stream));
  {
    TimedRegion tr(c, "stage_dedupe_kernel");
    ST_CHECK(launch_stage_lifetime(d_pts, b_entry.as<int>(), n_pts, b_w2c.as<double>(), b_sof.as<int>(),
                                   b_segend.as<int>(), lidar_range, b_s0.as<int>(), b_s1.as<int>(), b_diff.as<int>(),
                                   c->stream, &c->launches));
    std::vector<int> diff((size_t)nscan + 1);
    ST_CHECK(cudaMemcpyAsync(diff.data(), b_diff.p, diff.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    ST_CHECK(cudaStreamSynchronize(c->stream));
    std::vector<int64_t> cand_off((size_t)nscan + 1, 0);
    int run = 0, maxcand = 0;
    for (int s = 0; s < nscan; s++) {
      run += diff[(size_t)s];
      maxcand = std::max(maxcand, run);
      cand_off[(size_t)s + 1] = cand_off[(size_t)s] + run;
    }
    const int grid = std::min(nscan, 2 * c->num_sms);
    ST_CHECK(up(b_fos, P.frame_of_scan.data(), P.frame_of_scan.size() * 4));
    ST_CHECK(up(b_lo, P.pt_lo.data(), P.pt_lo.size() * 8));
    ST_CHECK(up(b_hi, P.pt_hi.data(), P.pt_hi.size() * 8));
    ST_CHECK(up(b_coff, cand_off.data(), cand_off.size() * 8));
    ST_CHECK(b_win.reserve((size_t)std::max<int64_t>(cand_off[(size_t)nscan], 1) * 4));
    ST_CHECK(b_nout.reserve((size_t)nscan * 4));
    ST_CHECK(b_ws.reserve(stage_dedupe_workspace_bytes(grid, maxcand, nullptr)));
    ST_CHECK(launch_stage_dedupe(d_pts, b_w2c.as<double>(), b_s0.as<int>(), b_s1.as<int>(), b_fos.as<int>(),
                                 b_lo.as<int64_t>(), b_hi.as<int64_t>(), b_coff.as<int64_t>(), nscan, maxcand,
                                 lidar_range, polar_filter != 0, b_ws.p, grid, b_win.as<int>(), b_nout.as<int>(),
                                 c->stream, &c->launches));
    std::vector<int> n_out((size_t)nscan);
    ST_CHECK(cudaMemcpyAsync(n_out.data(), b_nout.p, n_out.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    ST_CHECK(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < nscan; s++) S->off[(size_t)s + 1] = S->off[(size_t)s] + n_out[(size_t)s];
    const int64_t total = S->off[(size_t)nscan];
    ST_CHECK(S->d_off.reserve(((size_t)nscan + 1) * 8));
    ST_CHECK(cudaMemcpyAsync(S->d_off.p, S->off.data(), ((size_t)nscan + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    ST_CHECK(S->d_xyz.reserve((size_t)std::max<int64_t>(total, 1) * 24));
    ST_CHECK(S->d_inten.reserve((size_t)std::max<int64_t>(total, 1) * 4));
    ST_CHECK(launch_stage_emit(d_pts, d_int, b_w2c.as<double>(), b_fos.as<int>(), b_coff.as<int64_t>(), b_win.as<int>(),
                               S->d_off.as<int64_t>(), nscan, S->d_xyz.as<double>(), S->d_inten.as<float>(), grid,
                               c->stream, &c->launches));
  }
  ST_CHECK(cudaStreamSynchronize(c->stream));
#undef ST_CHECK
  release_all();
  *out = S;
  return SODSO_OK;
}

void sodso_staged_destroy(sodso_staged *S) {
  if (!S) return;
  if (S->ctx) cudaSetDevice(S->ctx->device);
  for (Buf *b : {&S->d_off, &S->d_xyz, &S->d_inten}) b->release();
  delete S;
}
int sodso_staged_num_scans(sodso_staged *S) { return S ? (int)S->ids.size() : 0; }
int64_t sodso_staged_num_points(sodso_staged *S) { return S && !S->off.empty() ? S->off.back() : 0; }
const double *sodso_staged_xyz(sodso_staged *S) { return S ? S->d_xyz.as<double>() : nullptr; }
const float *sodso_staged_inten(sodso_staged *S) { return S ? S->d_inten.as<float>() : nullptr; }
const int64_t *sodso_staged_scan_off(sodso_staged *S) { return S ? S->d_off.as<int64_t>() : nullptr; }

int sodso_staged_copy(sodso_staged *S, int32_t *ids, int64_t *scan_off, double *xyz, float *inten) {
  if (!S) {
    set_error("null staged handle");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = S->ctx;
  CTX_CHECK(c);
  const size_t ns = S->ids.size();
  const int64_t total = S->off.empty() ? 0 : S->off.back();
  auto put = [&](void *dst, const void *host_src, const void *dev_src, size_t bytes) -> cudaError_t {
    if (!dst || bytes == 0) return cudaSuccess;
    if (is_device_ptr(dst))
      return dev_src ? cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToDevice, c->stream)
                     : cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, c->stream);
    if (host_src) {
      std::memcpy(dst, host_src, bytes);
      return cudaSuccess;
    }
    return cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, c->stream);
  };
  SODSO_CUDA_CHECK(put(ids, S->ids.data(), nullptr, ns * 4));
  SODSO_CUDA_CHECK(put(scan_off, S->off.data(), nullptr, S->off.size() * 8));
  SODSO_CUDA_CHECK(put(xyz, nullptr, S->d_xyz.p, (size_t)total * 24));
  SODSO_CUDA_CHECK(put(inten, nullptr, S->d_inten.p, (size_t)total * 4));
  return sync_ctx(c);
}

// ---- resident row-sharded database ---------------------------------------------------------
// (re)builds the operand of rows [row0, row0 + rows) of a shard laid out for db->cap rows from `rows` signatures on
// the device; row0 == 0 also clears the binary-channel flags
static int db_write_rows(sodso_db *db, const double *hist_dev, int row0, int rows, bool pad_tail) {
  sodso_ctx *c = db->ctx;
  int jrc = join_xchg(c);   // (a pipelined query may still be reading the operand / results on the exchange stream)
  if (jrc) return jrc;
  const size_t w = db->type == SODSO_TYPE_SC ? 2 * SC_SIZE : 2 * M2DP_SIG;
  if (db->type == SODSO_TYPE_M2DP) {
    SODSO_CUDA_CHECK(cudaMemcpyAsync(db->op.as<double>() + (size_t)4 * row0 * w, hist_dev, (size_t)4 * rows * w * sizeof(double),
                                     cudaMemcpyDeviceToDevice, c->stream));
    return SODSO_OK;
  }
  if (db->op_algo == SODSO_ALGO_SIMT) {   // cross-check kernels: whole-operand rebuild only
    if (row0 != 0) {
      set_error("the fp32 cross-check database cannot grow");
      return SODSO_E_STATE;
    }
    return sc_prepare(c, db->op_algo, hist_dev, rows, db->op, true);
  }
  if (row0 == 0) SODSO_CUDA_CHECK(launch_sc_tc_clear_flags(db->op.p, c->stream));
  const int n_valid = row0 + rows;
  const int row1 = pad_tail ? sc_tc_db_rows_padded(db->cap) : n_valid;
  // the kernel indexes hist by absolute row: hand it the virtual base of row 0 (only rows >= row0 are read)
  SODSO_CUDA_CHECK(launch_sc_tc_prep_db_rows(hist_dev - (size_t)row0 * w, n_valid, row0, row1, db->op.p, c->stream,
                                             &c->launches, db->cap));
  return SODSO_OK;
}

static size_t db_op_bytes(const sodso_db *db, int cap) {
  if (db->type == SODSO_TYPE_M2DP) return (size_t)4 * cap * 2 * M2DP_SIG * sizeof(double);
  if (db->op_algo == SODSO_ALGO_SIMT) return (size_t)2 * SC_SIZE * ((cap + 31) & ~31) * sizeof(float);
  return sc_tc_db_bytes(cap);
}

int sodso_db_create(sodso_ctx *c, int type, const double *hist2, int n_local, int64_t global_row0,
                    sodso_db **out) {
  CTX_CHECK(c);
  if (!out || (type != SODSO_TYPE_SC && type != SODSO_TYPE_M2DP) || n_local < 0 || (n_local > 0 && !hist2)) {
    set_error("bad db_create arguments");
    return SODSO_E_ARG;
  }
  *out = nullptr;
  sodso_db *db = new sodso_db();
  db->ctx = c;
  db->type = type;
  db->n = n_local;
  db->cap = std::max(n_local, 1);
  db->row0 = global_row0;
  db->op_algo = c->algo;
  const size_t w = type == SODSO_TYPE_SC ? 2 * SC_SIZE : 2 * M2DP_SIG;
  const size_t rows = type == SODSO_TYPE_SC ? n_local : 4 * (size_t)n_local;
  const double *hd = nullptr;
  int rc = SODSO_OK;
  cudaError_t e = db->op.reserve(db_op_bytes(db, db->cap));
  if (e == cudaSuccess) e = cudaMemsetAsync(db->op.p, 0, db->op.cap, c->stream);
  if (e != cudaSuccess) {
    set_error(std::string("db_create: ") + cudaGetErrorString(e));
    rc = SODSO_E_CUDA;
  }
  if (!rc && n_local > 0) rc = stage_in(c, hist2, rows * w, c->h2, &hd);
  if (!rc && n_local > 0) rc = db_write_rows(db, hd, 0, n_local, true);
  if (!rc) rc = sync_ctx(c);
  if (rc) {
    sodso_db_destroy(db);
    return rc;
  }
  *out = db;
  return SODSO_OK;
}

// capacity: the operand buffers are laid out for `cap` rows; growing re-lays them out on the device (no host traffic)
int sodso_db_reserve(sodso_db *db, int capacity) {
  if (!db || capacity < 0) {
    set_error("bad db_reserve arguments");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  if (capacity <= db->cap) return SODSO_OK;
  {
    int jrc = join_xchg(c);
    if (jrc) return jrc;
  }
  if (db->type == SODSO_TYPE_SC && db->op_algo == SODSO_ALGO_SIMT) {
    set_error("the fp32 cross-check database cannot grow");
    return SODSO_E_STATE;
  }
  Buf nb;
  SODSO_CUDA_CHECK(nb.reserve(db_op_bytes(db, capacity)));
  cudaError_t e = cudaMemsetAsync(nb.p, 0, nb.cap, c->stream);
  if (e == cudaSuccess) {
    if (db->type == SODSO_TYPE_M2DP)
      e = cudaMemcpyAsync(nb.p, db->op.p, (size_t)4 * db->n * 2 * M2DP_SIG * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
    else
      e = sc_tc_db_relayout(db->op.p, db->cap, nb.p, capacity, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    nb.release();
    set_error(std::string("db_reserve: ") + cudaGetErrorString(e));
    return SODSO_E_CUDA;
  }
  db->op.release();
  db->op = nb;
  db->cap = capacity;
  db->matched = false;
  return SODSO_OK;
}

// n_new further signatures behind the shard's last row (global indices global_row0 + n ...): only their operand rows
// are written; capacity doubles when it runs out
int sodso_db_append(sodso_db *db, const double *hist_new, int n_new) {
  if (!db || n_new < 0 || (n_new > 0 && !hist_new)) {
    set_error("bad db_append arguments");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  if (n_new == 0) return SODSO_OK;
  if ((int64_t)db->n + n_new > INT32_MAX / 4) {
    set_error("db_append: too many rows");
    return SODSO_E_ARG;
  }
  int rc;
  if (db->n + n_new > db->cap && (rc = sodso_db_reserve(db, std::max(db->n + n_new, 2 * db->cap)))) return rc;
  const size_t w = db->type == SODSO_TYPE_SC ? 2 * SC_SIZE : 2 * M2DP_SIG;
  const size_t rows = db->type == SODSO_TYPE_SC ? n_new : 4 * (size_t)n_new;
  const double *hd;
  if ((rc = stage_in(c, hist_new, rows * w, c->h2, &hd))) return rc;
  if ((rc = db_write_rows(db, hd, db->n, n_new, false))) return rc;
  db->n += n_new;
  db->matched = false;
  return sync_ctx(c);
}

int sodso_db_reload(sodso_db *db, const double *hist2) {
  if (!db || !hist2) {
    set_error("bad db_reload arguments");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  db->matched = false;
  const size_t w = db->type == SODSO_TYPE_SC ? 2 * SC_SIZE : 2 * M2DP_SIG;
  const size_t rows = db->type == SODSO_TYPE_SC ? db->n : 4 * (size_t)db->n;
  const double *hd;
  int rc;
  if ((rc = stage_in(c, hist2, rows * w, c->h2, &hd))) return rc;
  if ((rc = db_write_rows(db, hd, 0, db->n, true))) return rc;
  return sync_ctx(c);
}

// reload + match in one streamed pass: the shard's scans arrive as HOST point buffers, are copied in 512-scan chunks
// on the copy stream, binned and written into the operand buffers in place, and every chunk is matched against the m
// queries as soon as it has landed.  Afterwards the handle is in the state sodso_db_reload + sodso_db_match leave it in.
// The query operand (db->q_op, m rows) must already be enqueued on the context's stream; nothing is synchronised.
}  // extern "C"
namespace sodso {
int db_stream_match_async(sodso_db *db, const double *xyz, const float *inten, const int64_t *off, double max_rho, int m,
                          bool self, double *hist_dev) {
  // self: the m == n queries ARE the shard's scans (same buffers): they are binned once, both operands are filled
  // chunk by chunk and every chunk is matched as an L-shaped region (new queries x all DB rows so far, old queries x
  // new DB rows).  Every pair is computed: the sharded entry points never use the self-match triangle, whose mirrored
  // half does not exist for the other ranks' shards.
  sodso_ctx *c = db->ctx;
  const int n = db->n;
  int rc;
  int64_t total = 0;
  if ((rc = check_offsets_host(off, n, &total))) return rc;
  db->matched = false;
  const bool host_pts = !is_device_ptr(xyz) && !is_device_ptr(inten) && !is_device_ptr(off);
  const int CH = 512;
  const bool streamed = host_pts && n >= std::max(c->stream_min_scans, 2 * CH);
  // chunk boundaries (multiples of the 256-row DB tile).  self: two 256-scan chunks first, the work that can be done
  // grows with the square of what has arrived
  std::vector<int> bounds{0};
  if (streamed)   // (self: the last kilo-scan also goes in 256-scan chunks -- what is left to do after the last byte
                  //  has landed is the L-shaped region of the final chunk)
    for (int b = self ? 256 : CH; b < n; b += (self && (b < 512 || n - b <= 1024)) ? 256 : CH) bounds.push_back(b);
  bounds.push_back(n);
  const int nchunk = (int)bounds.size() - 1;
  const double *xd = xyz;
  const float *id = inten;
  const int64_t *od;
  if (host_pts) {
    SODSO_CUDA_CHECK(c->in_xyz.reserve((size_t)total * 3 * sizeof(double)));
    SODSO_CUDA_CHECK(c->in_inten.reserve((size_t)total * sizeof(float)));
    xd = c->in_xyz.as<double>();
    id = c->in_inten.as<float>();
  }
  if ((rc = stage_in(c, off, (size_t)n + 1, c->in_off, &od))) return rc;
  double *hd = hist_dev;
  if (!hd) {
    SODSO_CUDA_CHECK(c->out_hist.reserve((size_t)n * 2 * SC_SIZE * sizeof(double)));
    hd = c->out_hist.as<double>();
  }
  const size_t cnt = (size_t)m * n;
  SODSO_CUDA_CHECK(db->dp.reserve(cnt * 4));
  SODSO_CUDA_CHECK(db->di.reserve(cnt * 4));
  SODSO_CUDA_CHECK(launch_sc_tc_clear_flags(db->op.p, c->stream));
  if (self) {
    SODSO_CUDA_CHECK(db->q_op.reserve(sc_tc_query_bytes(m)));
    SODSO_CUDA_CHECK(launch_sc_tc_clear_flags(db->q_op.p, c->stream));
  }
  // events of the chunk copies: destroyed on every exit path
  struct Events {
    std::vector<cudaEvent_t> v;
    ~Events() {
      for (cudaEvent_t e : v)
        if (e) cudaEventDestroy(e);
    }
  } evs;
  auto enqueue_copy = [&](int k) -> cudaError_t {   // chunk k of the host buffers -> HBM, on the copy stream
    const int s0 = bounds[k], s1 = bounds[k + 1];
    const int64_t p0 = off[s0], p1 = off[s1];
    cudaError_t e = cudaSuccess;
    if (p1 > p0) {
      e = cudaMemcpyAsync(c->in_xyz.as<double>() + 3 * p0, xyz + 3 * p0, (size_t)(p1 - p0) * 3 * sizeof(double),
                          cudaMemcpyHostToDevice, c->copy_stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(c->in_inten.as<float>() + p0, inten + p0, (size_t)(p1 - p0) * sizeof(float),
                            cudaMemcpyHostToDevice, c->copy_stream);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&evs.v[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(evs.v[k], c->copy_stream);
    return e;
  };
  if (streamed) {
    if (!c->copy_stream) SODSO_CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    cudaEvent_t e0;
    SODSO_CUDA_CHECK(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(e0, c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->copy_stream, e0, 0);
    cudaEventDestroy(e0);
    SODSO_CUDA_CHECK(e);
    evs.v.assign(nchunk, nullptr);
    // copies run one chunk ahead of the compute that is enqueued below (with pageable host memory cudaMemcpyAsync
    // returns only once the chunk has been staged: the compute of the previous chunk is already in the queue then)
    SODSO_CUDA_CHECK(enqueue_copy(0));
  } else if (host_pts) {
    SODSO_CUDA_CHECK(cudaMemcpyAsync(c->in_xyz.p, xyz, (size_t)total * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SODSO_CUDA_CHECK(cudaMemcpyAsync(c->in_inten.p, inten, (size_t)total * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  const int n_pad = sc_tc_db_rows_padded(db->cap), m_pad = sc_tc_query_rows_padded(m);
  c->kname = "sc_match_tc_kernel";
  c->ev_valid = false;
  for (int k = 0; k < nchunk; k++) {
    const int s0 = bounds[k], s1 = bounds[k + 1];
    const bool last = k == nchunk - 1;
    cudaError_t e = cudaSuccess;
    if (streamed && !last) e = enqueue_copy(k + 1);
    if (streamed && e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, evs.v[k], 0);
    if (e == cudaSuccess)
      e = launch_sc_generate(xd, id, od + s0, s1 - s0, max_rho, hd + (size_t)s0 * 2 * SC_SIZE, c->num_sms, c->stream,
                             &c->launches);
    if (e == cudaSuccess)
      e = launch_sc_tc_prep_db_rows(hd, n, s0, last ? n_pad : s1, db->op.p, c->stream, &c->launches, db->cap);
    if (self && e == cudaSuccess)
      e = launch_sc_tc_prep_query_rows(hd, m, s0, last ? m_pad : s1, db->q_op.p, c->stream, &c->launches);
    if (!streamed && e == cudaSuccess) c->ev_valid = cudaEventRecord(c->ev0, c->stream) == cudaSuccess;
    if (e == cudaSuccess) {
      if (self)
        e = launch_sc_match_tc_blocks(db->q_op.p, m, db->op.p, n, s0, s1, 0, s1, 0, s0, s0, s1, db->dp.as<float>(),
                                      db->di.as<float>(), n, c->num_sms, c->stream, &c->launches, db->cap);
      else
        e = launch_sc_match_tc_blocks(db->q_op.p, m, db->op.p, n, 0, m, s0, s1, 0, 0, 0, 0, db->dp.as<float>(),
                                      db->di.as<float>(), n, c->num_sms, c->stream, &c->launches, db->cap);
    }
    if (!streamed && c->ev_valid) c->ev_valid = cudaEventRecord(c->ev1, c->stream) == cudaSuccess;
    if (e != cudaSuccess) {
      set_error(std::string("db_stream_match: ") + cudaGetErrorString(e));
      cudaStreamSynchronize(c->stream);
      if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
      return SODSO_E_CUDA;
    }
  }
  db->m = m;
  db->matched = true;
  return SODSO_OK;
}
}  // namespace sodso
extern "C" {

int sodso_db_stream_match(sodso_db *db, const double *xyz, const float *inten, const int64_t *off, double max_rho,
                          const double *hist1, int m) {
  if (!db || !xyz || !inten || !off || !hist1 || m <= 0) {
    set_error("bad db_stream_match arguments");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  if (db->type != SODSO_TYPE_SC || db->op_algo != SODSO_ALGO_TC || db->n <= 0) {
    set_error("db_stream_match: non-empty Scan Context shards with the tensor-core matcher only");
    return SODSO_E_STATE;
  }
  int rc;
  if ((rc = join_xchg(c))) return rc;
  // queries: operand first (it gates every block)
  const double *hq;
  if ((rc = stage_in(c, hist1, (size_t)m * 2 * SC_SIZE, db->q_in, &hq))) return rc;
  if ((rc = sc_prepare(c, db->op_algo, hq, m, db->q_op, false))) return rc;
  if ((rc = db_stream_match_async(db, xyz, inten, off, max_rho, m, false, nullptr))) return rc;
  return sync_ctx(c);
}

void sodso_db_destroy(sodso_db *db) {
  if (!db) return;
  cudaSetDevice(db->ctx->device);
  cudaStreamSynchronize(db->ctx->stream);
  if (db->ctx->xchg_stream) cudaStreamSynchronize(db->ctx->xchg_stream);
  for (Buf *b : {&db->op, &db->q_in, &db->q_op, &db->dp, &db->di, &db->stats, &db->gstats, &db->idx,
                 &db->score, &db->dpat, &db->diat, &db->ws, &db->q_hist, &db->pack, &db->gather, &db->q_xyz, &db->q_inten, &db->q_off, &db->dp2, &db->di2, &db->stats2, &db->pack2, &db->gstats2, &db->gather2})
    b->release();
  delete db;
}

int sodso_db_size(sodso_db *db) { return db ? db->n : 0; }

}  // extern "C"
namespace sodso {
int db_match_prepared_async(sodso_db *db, int m) {
  sodso_ctx *c = db->ctx;
  db->matched = false;
  const size_t cnt = (size_t)m * db->n;
  int rc;
  SODSO_CUDA_CHECK(db->dp.reserve(cnt * 4));
  SODSO_CUDA_CHECK(db->di.reserve(cnt * 4));
  if ((rc = sc_match_core(c, db->op_algo, db->q_op, m, db->op, db->n, db->dp.as<float>(), db->di.as<float>(), db->n,
                          false, db->cap)))
    return rc;
  db->m = m;
  db->matched = true;
  return SODSO_OK;
}

// distances of m queries against the shard, everything enqueued on the context's stream, nothing synchronised
int db_match_async(sodso_db *db, const double *hist1, int m) {
  sodso_ctx *c = db->ctx;
  db->matched = false;
  if (db->n <= 0) {
    set_error("the database is empty");
    return SODSO_E_STATE;
  }
  const size_t w = db->type == SODSO_TYPE_SC ? 2 * SC_SIZE : 2 * M2DP_SIG;
  const size_t rows = db->type == SODSO_TYPE_SC ? m : 4 * (size_t)m;
  const double *hd;
  int rc;
  if ((rc = stage_in(c, hist1, rows * w, db->q_in, &hd))) return rc;
  const size_t cnt = (size_t)m * db->n;
  SODSO_CUDA_CHECK(db->dp.reserve(cnt * 4));
  SODSO_CUDA_CHECK(db->di.reserve(cnt * 4));
  if (db->type == SODSO_TYPE_SC) {
    if ((rc = sc_prepare(c, db->op_algo, hd, m, db->q_op, false))) return rc;
    return db_match_prepared_async(db, m);
  } else {
    if (db->op_algo == SODSO_ALGO_SIMT) {
      SODSO_CUDA_CHECK(db->ws.reserve(m2dp_match_workspace_bytes(m, db->n)));
      TimedRegion tr(c, "m2dp_match_kernel");
      SODSO_CUDA_CHECK(launch_m2dp_match(hd, m, db->op.as<double>(), db->n, db->dp.as<float>(),
                                         db->di.as<float>(), db->n, db->ws.p, c->stream, &c->launches));
    } else {
      SODSO_CUDA_CHECK(db->ws.reserve(m2dp_match_tc_workspace_bytes(m, db->n)));
      TimedRegion tr(c, "m2dp_match_tc_kernel");
      SODSO_CUDA_CHECK(launch_m2dp_match_tc(hd, m, db->op.as<double>(), db->n, db->dp.as<float>(), db->di.as<float>(),
                                            db->n, db->ws.p, c->num_sms, c->stream, &c->launches));
    }
  }
  db->m = m;
  db->matched = true;
  return SODSO_OK;
}
}  // namespace sodso
extern "C" {

int sodso_db_match(sodso_db *db, const double *hist1, int m) {
  if (!db) {
    set_error("null db");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  if (m <= 0 || !hist1) {
    set_error("bad db_match arguments");
    return SODSO_E_ARG;
  }
  int rc;
  if ((rc = join_xchg(c))) return rc;
  if ((rc = db_match_async(db, hist1, m))) return rc;
  // hist1 may be pageable / pinned host memory or a device tensor the caller reuses: the call returns once the library
  // has finished reading it (and asynchronous kernel failures surface here, not in a later unrelated call)
  return sync_ctx(c);
}

int sodso_db_partial_stats(sodso_db *db, double *stats) {
  if (!db || !stats) {
    set_error("bad db_partial_stats arguments");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  if (!db->matched) {
    set_error("db_partial_stats before db_match");
    return SODSO_E_STATE;
  }
  {
    int jrc = join_xchg(c);
    if (jrc) return jrc;
  }
  double *sd;
  int rc;
  const size_t cnt = (size_t)db->m * STATS_W;
  if ((rc = stage_out(c, stats, cnt, db->stats, &sd))) return rc;
  SODSO_CUDA_CHECK(launch_row_stats(db->dp.as<float>(), db->di.as<float>(), db->m, db->n, db->n, sd, c->stream,
                                    &c->launches));
  if ((rc = finish_out(c, stats, cnt, sd))) return rc;
  return sync_ctx(c);
}

int sodso_db_topk(sodso_db *db, const double *global_stats, int64_t n_global, int64_t q_global_row0,
                  int mask_width, double p_weight, int k, int64_t *idx, double *score, double *d_p,
                  double *d_i) {
  if (!db || !global_stats || !idx || !score || k <= 0 || n_global < db->n) {
    set_error("bad db_topk arguments");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  if (!db->matched) {
    set_error("db_topk before db_match");
    return SODSO_E_STATE;
  }
  {
    int jrc = join_xchg(c);
    if (jrc) return jrc;
  }
  const double *gs;
  int rc;
  const size_t cnt = (size_t)db->m * k;
  if ((rc = stage_in(c, global_stats, (size_t)db->m * STATS_W, db->gstats, &gs))) return rc;
  int64_t *id;
  double *sd, *pa, *ia;
  if ((rc = stage_out(c, idx, cnt, db->idx, &id))) return rc;
  if ((rc = stage_out(c, score, cnt, db->score, &sd))) return rc;
  if ((rc = stage_out(c, d_p, cnt, db->dpat, &pa))) return rc;
  if ((rc = stage_out(c, d_i, cnt, db->diat, &ia))) return rc;
  SODSO_CUDA_CHECK(launch_fuse_topk(db->dp.as<float>(), db->di.as<float>(), db->m, db->n, db->n, gs, n_global,
                                    q_global_row0, db->row0, mask_width, p_weight, k, id, sd, pa, ia,
                                    c->stream, &c->launches));
  if ((rc = finish_out(c, idx, cnt, id))) return rc;
  if ((rc = finish_out(c, score, cnt, sd))) return rc;
  if ((rc = finish_out(c, d_p, cnt, pa))) return rc;
  if ((rc = finish_out(c, d_i, cnt, ia))) return rc;
  return sync_ctx(c);
}

int sodso_topk_merge_device(sodso_ctx *c, const int64_t *idx, const double *score, const double *d_p,
                            const double *d_i, int nshards, int m, int k, int64_t *out_idx, double *out_score,
                            double *out_d_p, double *out_d_i) {
  CTX_CHECK(c);
  if (!idx || !score || !out_idx || !out_score || nshards <= 0 || nshards > 16 || m < 0 || k <= 0) {
    set_error("bad topk_merge_device arguments (at most 16 shards)");
    return SODSO_E_ARG;
  }
  for (const void *p : {(const void *)idx, (const void *)score, (const void *)out_idx, (const void *)out_score})
    if (!is_device_ptr(p)) {
      set_error("topk_merge_device takes device pointers");
      return SODSO_E_ARG;
    }
  SODSO_CUDA_CHECK(launch_topk_merge(idx, score, d_p, d_i, nshards, m, k, out_idx, out_score, out_d_p, out_d_i,
                                     c->stream, &c->launches));
  return sync_ctx(c);
}

int sodso_topk_merge(const int64_t *idx, const double *score, const double *d_p, const double *d_i,
                     int nshards, int m, int k, int64_t *out_idx, double *out_score, double *out_d_p,
                     double *out_d_i) {
  if (!idx || !score || !out_idx || !out_score || nshards <= 0 || m < 0 || k <= 0) {
    set_error("bad topk_merge arguments");
    return SODSO_E_ARG;
  }
  std::vector<int> cur((size_t)nshards);
  for (int q = 0; q < m; q++) {
    std::fill(cur.begin(), cur.end(), 0);
    for (int r = 0; r < k; r++) {
      int best = -1;
      size_t bo = 0;
      for (int s = 0; s < nshards; s++) {
        if (cur[s] >= k) continue;
        size_t o = ((size_t)s * m + q) * k + cur[s];
        if (idx[o] < 0) continue;
        if (best < 0 || score[o] < score[bo] || (score[o] == score[bo] && idx[o] < idx[bo])) {
          best = s;
          bo = o;
        }
      }
      size_t oo = (size_t)q * k + r;
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
        cur[best]++;
      }
    }
  }
  return SODSO_OK;
}

int sodso_db_get_distances(sodso_db *db, float *d_p, float *d_i) {
  if (!db) {
    set_error("null db");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  if (!db->matched) {
    set_error("db_get_distances before db_match");
    return SODSO_E_STATE;
  }
  {
    int jrc = join_xchg(c);
    if (jrc) return jrc;
  }
  const size_t bytes = (size_t)db->m * db->n * sizeof(float);
  if (d_p) SODSO_CUDA_CHECK(cudaMemcpyAsync(d_p, db->dp.p, bytes, cudaMemcpyDefault, c->stream));
  if (d_i) SODSO_CUDA_CHECK(cudaMemcpyAsync(d_i, db->di.p, bytes, cudaMemcpyDefault, c->stream));
  return sync_ctx(c);
}

}  // extern "C"
