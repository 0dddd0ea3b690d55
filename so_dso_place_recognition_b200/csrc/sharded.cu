// Row-sharded signature database over the GPUs of one box, behind the C ABI: the NCCL communicator
// of a context and the per-batch exchange of a sharded query.
//
// Every (query, DB row) distance is independent (processSC.m:30-32), so no collective touches the distance matrices.
// run_test.m:40 z-scores every query row over the WHOLE database before the arg-min (run_test.m:57); a query batch
// therefore needs exactly two small exchanges, both enqueued on the context's stream between the kernels they
// connect -- no host synchronisation anywhere before the final m x k copy-out:
//   match -> row_stats_kernel -> ncclAllReduce(m x 6 fp64 partial row sums / counts)
//         -> fuse_topk_kernel (global statistics, global mask indices) -> ncclAllGather(per-shard top-k, 32 B / entry)
//         -> topk_merge_kernel (lowest global index on ties, like MATLAB's first minimum)
// Payloads are KBs: the exchange is latency-bound, NVLink bandwidth does not matter.
//
// NCCL is bound at run time (dlopen): inside a PyTorch process the already loaded libnccl.so.2 is reused (two NCCL
// copies in one process do not mix), a plain C++ host gets the system library.  <nccl.h> supplies types only.
#include <dlfcn.h>
#include <nccl.h>

#include "capi_internal.cuh"

namespace sodso {

namespace {

struct NcclApi {
  void *handle = nullptr;
  std::string err;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  bool ok() const { return handle && err.empty(); }
};

NcclApi load_nccl() {
  NcclApi A;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    A.handle = dlopen(name, RTLD_NOW | RTLD_NOLOAD);   // the copy this process already uses (e.g. PyTorch's)
    if (A.handle) break;
  }
  if (!A.handle)
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      A.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (A.handle) break;
    }
  if (!A.handle) {
    A.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
    return A;
  }
#define SODSO_BIND(sym)                                                       \
  A.sym = reinterpret_cast<decltype(A.sym)>(dlsym(A.handle, "nccl" #sym));    \
  if (!A.sym) A.err += std::string(" missing nccl" #sym);
  SODSO_BIND(GetUniqueId)
  SODSO_BIND(CommInitRank)
  SODSO_BIND(CommDestroy)
  SODSO_BIND(AllReduce)
  SODSO_BIND(AllGather)
  SODSO_BIND(Broadcast)
  SODSO_BIND(GroupStart)
  SODSO_BIND(GroupEnd)
  SODSO_BIND(GetErrorString)
  SODSO_BIND(GetVersion)
#undef SODSO_BIND
  return A;
}

NcclApi &nccl() {
  static NcclApi A = load_nccl();
  return A;
}

#define SODSO_NCCL_CHECK(expr)                                                                        \
  do {                                                                                                \
    ncclResult_t _r = (expr);                                                                         \
    if (_r != ncclSuccess) {                                                                          \
      set_error(std::string(#expr) + ": " + nccl().GetErrorString(_r) + " (" + __FILE__ + ":" +       \
                std::to_string(__LINE__) + ")");                                                      \
      return SODSO_E_NCCL;                                                                            \
    }                                                                                                 \
  } while (0)

int need_nccl() {
  if (nccl().ok()) return SODSO_OK;
  set_error("NCCL is not available: " + nccl().err);
  return SODSO_E_NCCL;
}

}  // namespace

struct CommState {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};

void comm_release(sodso_ctx *c) {
  if (!c || !c->comm) return;
  if (c->comm->comm && nccl().ok()) nccl().CommDestroy(c->comm->comm);
  delete c->comm;
  c->comm = nullptr;
}

// contiguous block partition of `total` rows over `world` ranks (the rule of so_dso_place_recognition_b200/sharded.py)
static void block_partition(int total, int world, int rank, int *first, int *count) {
  const int base = total / world, rem = total % world;
  *count = base + (rank < rem ? 1 : 0);
  *first = rank * base + std::min(rank, rem);
}

// statistics -> (all-reduce) -> per-shard top-k -> (all-gather, merge) for the last match of `db`; outputs are DEVICE
// pointers (idx, score required); everything is enqueued on the context's stream
static int finish_sharded_async(sodso_db *db, int64_t q_row0, int mask_width, double p_weight, int k, int64_t *idx,
                                double *score, double *d_p, double *d_i) {
  sodso_ctx *c = db->ctx;
  const int m = db->m, R = c->comm ? c->comm->nranks : 1;
  const size_t mk = (size_t)m * k;
  SODSO_CUDA_CHECK(db->stats.reserve((size_t)m * STATS_W * sizeof(double)));
  SODSO_CUDA_CHECK(launch_row_stats(db->dp.as<float>(), db->di.as<float>(), m, db->n, db->n, db->stats.as<double>(),
                                    c->stream, &c->launches));
  const double *gs = db->stats.as<double>();
  if (R > 1) {
    SODSO_CUDA_CHECK(db->gstats.reserve((size_t)m * STATS_W * sizeof(double)));
    SODSO_NCCL_CHECK(nccl().AllReduce(db->stats.p, db->gstats.p, (size_t)m * STATS_W, ncclDouble, ncclSum, c->comm->comm,
                                      c->stream));
    gs = db->gstats.as<double>();
  }
  if (R == 1) {
    SODSO_CUDA_CHECK(launch_fuse_topk(db->dp.as<float>(), db->di.as<float>(), m, db->n, db->n, gs, db->n, q_row0, db->row0,
                                      mask_width, p_weight, k, idx, score, d_p, d_i, c->stream, &c->launches));
    return SODSO_OK;
  }
  // packed lists [4][m][k] of 8-byte entries (global index, fused score, d_p, d_i): one all-gather moves everything
  SODSO_CUDA_CHECK(db->pack.reserve(4 * mk * 8));
  SODSO_CUDA_CHECK(db->gather.reserve((size_t)R * 4 * mk * 8));
  int64_t *pi = db->pack.as<int64_t>();
  double *ps = db->pack.as<double>() + mk, *pp = ps + mk, *pd = pp + mk;
  SODSO_CUDA_CHECK(launch_fuse_topk(db->dp.as<float>(), db->di.as<float>(), m, db->n, db->n, gs, db->n, q_row0, db->row0,
                                    mask_width, p_weight, k, pi, ps, pp, pd, c->stream, &c->launches));
  SODSO_NCCL_CHECK(nccl().AllGather(db->pack.p, db->gather.p, 4 * mk * 8, ncclChar, c->comm->comm, c->stream));
  const int64_t *gi = db->gather.as<int64_t>();
  const double *gsc = db->gather.as<double>() + mk, *gp = gsc + mk, *gd = gp + mk;
  SODSO_CUDA_CHECK(launch_topk_merge(gi, gsc, gp, gd, R, m, k, idx, score, d_p, d_i, c->stream, &c->launches, 4 * mk));
  return SODSO_OK;
}

// host-or-device outputs around finish_sharded_async; synchronises only when something is copied to the host
static int finish_sharded(sodso_db *db, int64_t q_row0, int mask_width, double p_weight, int k, int64_t *idx,
                          double *score, double *d_p, double *d_i) {
  sodso_ctx *c = db->ctx;
  if (c->comm && c->comm->nranks > 16) {
    set_error("sharded query: at most 16 ranks");
    return SODSO_E_ARG;
  }
  const size_t cnt = (size_t)db->m * k;
  int64_t *id;
  double *sd, *pa, *ia;
  int rc;
  if ((rc = stage_out(c, idx, cnt, db->idx, &id))) return rc;
  if ((rc = stage_out(c, score, cnt, db->score, &sd))) return rc;
  if ((rc = stage_out(c, d_p, cnt, db->dpat, &pa))) return rc;
  if ((rc = stage_out(c, d_i, cnt, db->diat, &ia))) return rc;
  if ((rc = finish_sharded_async(db, q_row0, mask_width, p_weight, k, id, sd, pa, ia))) return rc;
  const bool any_host = id != idx || sd != score || (d_p && pa != d_p) || (d_i && ia != d_i);
  if ((rc = finish_out(c, idx, cnt, id))) return rc;
  if ((rc = finish_out(c, score, cnt, sd))) return rc;
  if ((rc = finish_out(c, d_p, cnt, pa))) return rc;
  if ((rc = finish_out(c, d_i, cnt, ia))) return rc;
  return any_host ? sync_ctx(c) : SODSO_OK;
}

}  // namespace sodso

using namespace sodso;

extern "C" {

int sodso_comm_unique_id(void *id_out) {
  if (!id_out) {
    set_error("id_out is null");
    return SODSO_E_ARG;
  }
  int rc;
  if ((rc = need_nccl())) return rc;
  static_assert(sizeof(ncclUniqueId) == SODSO_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  SODSO_NCCL_CHECK(nccl().GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return SODSO_OK;
}

int sodso_comm_init(sodso_ctx *c, const void *unique_id, int nranks, int rank) {
  CTX_CHECK(c);
  if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !unique_id)) {
    set_error("bad comm_init arguments");
    return SODSO_E_ARG;
  }
  comm_release(c);
  CommState *S = new CommState();
  S->nranks = nranks;
  S->rank = rank;
  if (nranks > 1) {
    int rc;
    if ((rc = need_nccl())) {
      delete S;
      return rc;
    }
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = nccl().CommInitRank(&S->comm, nranks, id, rank);
    if (r != ncclSuccess) {
      set_error(std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
      delete S;
      return SODSO_E_NCCL;
    }
  }
  c->comm = S;
  return SODSO_OK;
}

int sodso_comm_finalize(sodso_ctx *c) {
  CTX_CHECK(c);
  SODSO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  comm_release(c);
  return SODSO_OK;
}

int sodso_comm_nranks(sodso_ctx *c) { return c && c->comm ? c->comm->nranks : 1; }
int sodso_comm_rank(sodso_ctx *c) { return c && c->comm ? c->comm->rank : 0; }

int sodso_comm_nccl_version(void) {
  int v = 0;
  if (!nccl().ok() || nccl().GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

int sodso_db_finish_sharded(sodso_db *db, int64_t q_global_row0, int mask_width, double p_weight, int k, int64_t *idx,
                            double *score, double *d_p, double *d_i) {
  if (!db || !idx || !score || k <= 0) {
    set_error("bad db_finish_sharded arguments");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  if (!db->matched) {
    set_error("db_finish_sharded before a match");
    return SODSO_E_STATE;
  }
  return finish_sharded(db, q_global_row0, mask_width, p_weight, k, idx, score, d_p, d_i);
}

int sodso_db_query_sharded(sodso_db *db, const double *hist1, int m, int64_t q_global_row0, int mask_width,
                           double p_weight, int k, int64_t *idx, double *score, double *d_p, double *d_i) {
  if (!db || !hist1 || m <= 0 || !idx || !score || k <= 0) {
    set_error("bad db_query_sharded arguments");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  int rc;
  if ((rc = db_match_async(db, hist1, m))) return rc;
  return finish_sharded(db, q_global_row0, mask_width, p_weight, k, idx, score, d_p, d_i);
}

int sodso_db_scans_query_sharded(sodso_db *db, const double *db_xyz, const float *db_inten, const int64_t *db_off,
                                 const double *q_xyz, const float *q_inten, const int64_t *q_off, int m_total,
                                 int q_first, int m_slice, double max_rho, int64_t q_global_row0, int mask_width,
                                 double p_weight, int k, double *q_hist, int64_t *idx, double *score, double *d_p,
                                 double *d_i) {
  if (!db || m_total <= 0 || q_first < 0 || m_slice < 0 || q_first + m_slice > m_total ||
      (m_slice > 0 && (!q_xyz || !q_inten || !q_off)) || !idx || !score || k <= 0 ||
      ((db_xyz != nullptr) != (db_inten != nullptr)) || ((db_xyz != nullptr) != (db_off != nullptr))) {
    set_error("bad db_scans_query_sharded arguments");
    return SODSO_E_ARG;
  }
  sodso_ctx *c = db->ctx;
  CTX_CHECK(c);
  if (db->type != SODSO_TYPE_SC || db->op_algo != SODSO_ALGO_TC || db->n <= 0) {
    set_error("db_scans_query_sharded: non-empty Scan Context shards with the tensor-core matcher only");
    return SODSO_E_STATE;
  }
  const int R = c->comm ? c->comm->nranks : 1, rank = c->comm ? c->comm->rank : 0;
  const bool gather = m_slice != m_total;
  if (gather) {
    int first, count;
    block_partition(m_total, R, rank, &first, &count);
    if (R == 1 || first != q_first || count != m_slice) {
      set_error("db_scans_query_sharded: a query slice must be this rank's block of the contiguous partition of m_total");
      return SODSO_E_ARG;
    }
  }
  int rc;
  // the queries are the shard's own scans (same buffers, a self-match on this rank): they are binned once, while the
  // shard streams in.  Every pair is still computed -- see db_stream_match_async.
  const bool self = db_xyz && q_xyz == db_xyz && q_inten == db_inten && q_off == db_off && m_slice == m_total &&
                    m_total == db->n;
  // ---- query signatures: this rank's slice is binned here (test_sc.cpp:40-57); slices travel as signatures
  // (19 KB per scan instead of 115 KB of points) over NVLink
  double *qh;
  const size_t qcnt = (size_t)m_total * 2 * SC_SIZE;
  if (q_hist && is_device_ptr(q_hist))
    qh = q_hist;
  else {
    SODSO_CUDA_CHECK(db->q_hist.reserve(qcnt * sizeof(double)));
    qh = db->q_hist.as<double>();
  }
  if (self) {
    if ((rc = db_stream_match_async(db, db_xyz, db_inten, db_off, max_rho, m_total, true, qh))) return rc;
  } else {
    if (m_slice > 0) {
      int64_t total = 0;
      if ((rc = check_offsets_host(q_off, m_slice, &total))) return rc;
      const double *xd;
      const float *id;
      const int64_t *od;
      if ((rc = stage_in(c, q_xyz, (size_t)total * 3, db->q_xyz, &xd))) return rc;
      if ((rc = stage_in(c, q_inten, (size_t)total, db->q_inten, &id))) return rc;
      if ((rc = stage_in(c, q_off, (size_t)m_slice + 1, db->q_off, &od))) return rc;
      SODSO_CUDA_CHECK(launch_sc_generate(xd, id, od, m_slice, max_rho, qh + (size_t)q_first * 2 * SC_SIZE, c->num_sms,
                                          c->stream, &c->launches));
    }
    if (gather) {
      SODSO_NCCL_CHECK(nccl().GroupStart());
      for (int r = 0; r < R; r++) {
        int first, count;
        block_partition(m_total, R, r, &first, &count);
        if (count == 0) continue;
        double *p = qh + (size_t)first * 2 * SC_SIZE;
        SODSO_NCCL_CHECK(nccl().Broadcast(p, p, (size_t)count * 2 * SC_SIZE, ncclDouble, r, c->comm->comm, c->stream));
      }
      SODSO_NCCL_CHECK(nccl().GroupEnd());
    }
    if ((rc = sc_prepare(c, db->op_algo, qh, m_total, db->q_op, false))) return rc;
    // ---- the shard: new scans (binned, operand rewritten in place, matched chunk by chunk) or the resident operand
    if (db_xyz) {
      if ((rc = db_stream_match_async(db, db_xyz, db_inten, db_off, max_rho, m_total, false, nullptr))) return rc;
    } else {
      if ((rc = db_match_prepared_async(db, m_total))) return rc;
    }
  }
  const bool host_qh = q_hist && q_hist != qh;
  if (host_qh && (rc = finish_out(c, q_hist, qcnt, qh))) return rc;
  if ((rc = finish_sharded(db, q_global_row0, mask_width, p_weight, k, idx, score, d_p, d_i))) return rc;
  return host_qh ? sync_ctx(c) : SODSO_OK;
}

}  // extern "C"
