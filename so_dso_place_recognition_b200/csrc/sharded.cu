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
#include <unistd.h>

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
  // peer-memory exchange (PeerExchange, common.cuh): this rank's window and every rank's window as mapped here
  void *win = nullptr;
  void *peer[PX_MAX_RANKS] = {};
  bool peer_ipc[PX_MAX_RANKS] = {};
  size_t slot_bytes = 0;
  int *h_err = nullptr;   // mapped pinned word: a consumer kernel whose wait timed out sets it
  int *d_err = nullptr;
  bool p2p = false;
};

constexpr size_t PX_WINDOW_BYTES = (size_t)64 << 20;
constexpr int PX_FAST_ROWS = 512;   // batches up to this many queries use the peer-memory exchange

void comm_release(sodso_ctx *c) {
  if (!c || !c->comm) return;
  CommState *S = c->comm;
  for (int r = 0; r < S->nranks && r < PX_MAX_RANKS; r++)
    if (S->peer_ipc[r] && S->peer[r]) cudaIpcCloseMemHandle(S->peer[r]);
  if (S->win) cudaFree(S->win);
  if (S->h_err) cudaFreeHost(S->h_err);
  if (S->comm && nccl().ok()) nccl().CommDestroy(S->comm);
  delete S;
  c->comm = nullptr;
}

// The window every rank exposes to the others: allocated here, its IPC handle (or, for ranks that are threads of this
// process, its address) exchanged through the communicator, mapped on every rank.  Any failure (no peer access between
// two GPUs, IPC not permitted in the container) leaves p2p off on ALL ranks and the exchange on NCCL.
struct PxInfo {
  cudaIpcMemHandle_t handle;
  unsigned long long ptr;
  int pid, dev;
};
static int px_setup(sodso_ctx *c) {
  CommState *S = c->comm;
  const int R = S->nranks;
  if (R > PX_MAX_RANKS) return SODSO_OK;
  int ok = 1;
  S->slot_bytes = (PX_WINDOW_BYTES / (2 * (size_t)R)) & ~(size_t)255;
  if (S->slot_bytes <= PX_OFF_LISTS + 4096) ok = 0;
  if (cudaMalloc(&S->win, PX_WINDOW_BYTES) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void **>(&S->h_err), sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(reinterpret_cast<void **>(&S->d_err), S->h_err, 0) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  PxInfo mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    cudaMemset(S->win, 0, PX_WINDOW_BYTES);
    *S->h_err = 0;
    if (cudaIpcGetMemHandle(&mine.handle, S->win) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
    }
  }
  mine.ptr = (unsigned long long)(uintptr_t)S->win;
  mine.pid = (int)getpid();
  mine.dev = c->device;
  // exchange: [info | ok] of every rank
  Buf send, recv;
  const size_t rec = sizeof(PxInfo) + 8;
  SODSO_CUDA_CHECK(send.reserve(rec));
  SODSO_CUDA_CHECK(recv.reserve(rec * R));
  std::vector<unsigned char> h(rec * R, 0);
  memcpy(h.data(), &mine, sizeof(mine));
  memcpy(h.data() + sizeof(PxInfo), &ok, sizeof(int));
  SODSO_CUDA_CHECK(cudaMemcpyAsync(send.p, h.data(), rec, cudaMemcpyHostToDevice, c->stream));
  SODSO_NCCL_CHECK(nccl().AllGather(send.p, recv.p, rec, ncclChar, S->comm, c->stream));
  SODSO_CUDA_CHECK(cudaMemcpyAsync(h.data(), recv.p, rec * R, cudaMemcpyDeviceToHost, c->stream));
  SODSO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < R; r++) {
    int okr;
    memcpy(&okr, h.data() + r * rec + sizeof(PxInfo), sizeof(int));
    ok = ok && okr;
  }
  if (ok)
    for (int r = 0; r < R && ok; r++) {
      PxInfo pi;
      memcpy(&pi, h.data() + r * rec, sizeof(pi));
      if (r == S->rank) {
        S->peer[r] = S->win;
      } else if (pi.pid == mine.pid) {   // a thread of this process: its address is valid here once peer access is on
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, c->device, pi.dev) != cudaSuccess || !can) ok = 0;
        if (ok) {
          cudaError_t e = cudaDeviceEnablePeerAccess(pi.dev, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
          cudaGetLastError();
        }
        S->peer[r] = (void *)(uintptr_t)pi.ptr;
      } else {
        if (cudaIpcOpenMemHandle(&S->peer[r], pi.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          S->peer[r] = nullptr;
          ok = 0;
        } else {
          S->peer_ipc[r] = true;
        }
      }
    }
  // second round: everybody must have mapped everybody
  {
    int *d = send.as<int>();
    SODSO_CUDA_CHECK(cudaMemcpyAsync(d, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    SODSO_NCCL_CHECK(nccl().AllReduce(d, d, 1, ncclInt, ncclMin, S->comm, c->stream));
    SODSO_CUDA_CHECK(cudaMemcpyAsync(&ok, d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SODSO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  S->p2p = ok != 0;
  send.release();
  recv.release();
  return SODSO_OK;
}

static bool px_enabled = true;   // sodso_debug_set_peer_exchange (tests compare the two transports)

// after a synchronisation: did a consumer kernel of the peer-memory exchange give up waiting for a rank?
int comm_check(sodso_ctx *c) {
  if (c && c->comm && c->comm->h_err && *c->comm->h_err) {
    *c->comm->h_err = 0;
    set_error("sharded query: a rank did not deliver its part of the exchange (timeout)");
    return SODSO_E_NCCL;
  }
  return SODSO_OK;
}

// contiguous block partition of `total` rows over `world` ranks (the rule of so_dso_place_recognition_b200/sharded.py)
static void block_partition(int total, int world, int rank, int *first, int *count) {
  const int base = total / world, rem = total % world;
  *count = base + (rank < rem ? 1 : 0);
  *first = rank * base + std::min(rank, rem);
}

// statistics -> (all-reduce) -> per-shard top-k -> (all-gather, merge) for the last match of `db`; outputs are DEVICE
// pointers (idx, score required); everything is enqueued on the context's stream
// pipelined: everything after the partial statistics (the exchanges, fusion, merge) goes to the context's exchange
// stream, so that the caller's next batch can be matched on the main stream meanwhile (pipeline_begin has chosen the
// buffer set and made the main stream wait for the batch that used it last).
static int finish_sharded_async(sodso_db *db, int64_t q_row0, int mask_width, double p_weight, int k, int64_t *idx,
                                double *score, double *d_p, double *d_i, bool pipelined) {
  sodso_ctx *c = db->ctx;
  const int m = db->m, R = c->comm ? c->comm->nranks : 1;
  const size_t mk = (size_t)m * k;
  const unsigned batch = ++c->batch_no;
  const int f = (int)(batch & 1u);
  cudaStream_t st2 = pipelined ? c->xchg_stream : c->stream;
  // the statistics kernel (a producer of the exchange) stays on the main stream; `after_stats` moves to the other one
  auto after_stats = [&]() -> int {
    if (!pipelined) return SODSO_OK;
    SODSO_CUDA_CHECK(cudaEventRecord(c->ev_stats[f], c->stream));
    SODSO_CUDA_CHECK(cudaStreamWaitEvent(c->xchg_stream, c->ev_stats[f], 0));
    return SODSO_OK;
  };
  auto done = [&]() -> int {
    if (!pipelined) return SODSO_OK;
    SODSO_CUDA_CHECK(cudaEventRecord(c->ev_done[f], c->xchg_stream));
    c->ev_done_rec[f] = true;
    return SODSO_OK;
  };
  int rc;
  SODSO_CUDA_CHECK(db->stats.reserve((size_t)m * STATS_W * sizeof(double)));
  // ---- peer-memory exchange: the partial statistics and the candidate lists are written by row_stats_kernel /
  // fuse_topk_kernel straight into every rank's window over NVLink, the consumers poll per-row flags: no collective
  // launch between the kernels (see PeerExchange in common.cuh).  Falls back to NCCL when the windows could not be
  // mapped or the batch does not fit a slot.
  // It is a latency optimisation for streaming batches: per row it costs a system-scope fence and a few NVLink
  // stores, the two NCCL collectives cost ~45 us each whatever the batch size.  Measured break-even ~800 rows
  // (5 000-row batch: +0.46 ms; 128-row batch: -0.06 ms at 2 ranks), so larger batches take the collectives.
  if (R > 1 && c->comm->p2p && m <= PX_FAST_ROWS && PX_OFF_LISTS + 4 * mk * 8 <= c->comm->slot_bytes) {
    CommState *S = c->comm;
    PeerExchange px;
    memset(&px, 0, sizeof(px));
    for (int r = 0; r < R; r++) px.win[r] = reinterpret_cast<unsigned char *>(S->peer[r]);
    px.slot_bytes = S->slot_bytes;
    px.nranks = R;
    px.rank = S->rank;
    px.epoch = batch;
    px.err = S->d_err;
    SODSO_CUDA_CHECK(db->pack.reserve(4 * mk * 8));
    int64_t *pi = db->pack.as<int64_t>();
    double *ps = db->pack.as<double>() + mk, *pp = ps + mk, *pd = pp + mk;
    SODSO_CUDA_CHECK(launch_row_stats(db->dp.as<float>(), db->di.as<float>(), m, db->n, db->n, db->stats.as<double>(),
                                      c->stream, &c->launches, &px));
    if ((rc = after_stats())) return rc;
    SODSO_CUDA_CHECK(launch_fuse_topk(db->dp.as<float>(), db->di.as<float>(), m, db->n, db->n, db->stats.as<double>(), db->n,
                                      q_row0, db->row0, mask_width, p_weight, k, pi, ps, pp, pd, st2, &c->launches, &px));
    SODSO_CUDA_CHECK(launch_topk_merge_px(px, m, k, idx, score, d_p, d_i, st2, &c->launches));
    return done();
  }
  SODSO_CUDA_CHECK(launch_row_stats(db->dp.as<float>(), db->di.as<float>(), m, db->n, db->n, db->stats.as<double>(),
                                    c->stream, &c->launches));
  if ((rc = after_stats())) return rc;
  const double *gs = db->stats.as<double>();
  if (R > 1) {
    SODSO_CUDA_CHECK(db->gstats.reserve((size_t)m * STATS_W * sizeof(double)));
    SODSO_NCCL_CHECK(nccl().AllReduce(db->stats.p, db->gstats.p, (size_t)m * STATS_W, ncclDouble, ncclSum, c->comm->comm, st2));
    gs = db->gstats.as<double>();
  }
  if (R == 1) {
    SODSO_CUDA_CHECK(launch_fuse_topk(db->dp.as<float>(), db->di.as<float>(), m, db->n, db->n, gs, db->n, q_row0, db->row0,
                                      mask_width, p_weight, k, idx, score, d_p, d_i, st2, &c->launches));
    return done();
  }
  // packed lists [4][m][k] of 8-byte entries (global index, fused score, d_p, d_i): one all-gather moves everything
  SODSO_CUDA_CHECK(db->pack.reserve(4 * mk * 8));
  SODSO_CUDA_CHECK(db->gather.reserve((size_t)R * 4 * mk * 8));
  int64_t *pi = db->pack.as<int64_t>();
  double *ps = db->pack.as<double>() + mk, *pp = ps + mk, *pd = pp + mk;
  SODSO_CUDA_CHECK(launch_fuse_topk(db->dp.as<float>(), db->di.as<float>(), m, db->n, db->n, gs, db->n, q_row0, db->row0,
                                    mask_width, p_weight, k, pi, ps, pp, pd, st2, &c->launches));
  SODSO_NCCL_CHECK(nccl().AllGather(db->pack.p, db->gather.p, 4 * mk * 8, ncclChar, c->comm->comm, st2));
  const int64_t *gi = db->gather.as<int64_t>();
  const double *gsc = db->gather.as<double>() + mk, *gp = gsc + mk, *gd = gp + mk;
  SODSO_CUDA_CHECK(launch_topk_merge(gi, gsc, gp, gd, R, m, k, idx, score, d_p, d_i, st2, &c->launches, 4 * mk));
  return done();
}

// Pipelined mode of a sharded query: all outputs are device memory (nothing has to be copied out, so the call only
// enqueues) and the context runs on its own stream.  Chooses the buffer set of this batch -- the one batch b - 2 used --
// and makes the main stream wait until that batch has left the exchange stream; that wait is also the flow control of
// the peer-memory exchange (a rank publishes batch b only after its merge of batch b - 2, i.e. after every rank has
// consumed what batch b - 2 left in the windows).
static bool can_pipeline(sodso_ctx *c, const int64_t *idx, const double *score, const double *d_p, const double *d_i) {
  return c->stream == c->own_stream && is_device_ptr(idx) && is_device_ptr(score) && (!d_p || is_device_ptr(d_p)) &&
         (!d_i || is_device_ptr(d_i));
}
static int pipeline_begin(sodso_db *db) {
  sodso_ctx *c = db->ctx;
  if (!c->xchg_stream) {
    SODSO_CUDA_CHECK(cudaStreamCreateWithFlags(&c->xchg_stream, cudaStreamNonBlocking));
    for (int f = 0; f < 2; f++) {
      SODSO_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_stats[f], cudaEventDisableTiming));
      SODSO_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_done[f], cudaEventDisableTiming));
    }
  }
  const int f = (int)((c->batch_no + 1) & 1u);
  if (c->ev_done_rec[f]) SODSO_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_done[f], 0));
  std::swap(db->dp, db->dp2);
  std::swap(db->di, db->di2);
  std::swap(db->stats, db->stats2);
  std::swap(db->gstats, db->gstats2);
  std::swap(db->pack, db->pack2);
  std::swap(db->gather, db->gather2);
  return SODSO_OK;
}

// host-or-device outputs around finish_sharded_async; synchronises only when something is copied to the host
static int finish_sharded(sodso_db *db, int64_t q_row0, int mask_width, double p_weight, int k, int64_t *idx,
                          double *score, double *d_p, double *d_i, bool pipelined = false) {
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
  if ((rc = finish_sharded_async(db, q_row0, mask_width, p_weight, k, id, sd, pa, ia, pipelined))) return rc;
  const bool any_host = id != idx || sd != score || (d_p && pa != d_p) || (d_i && ia != d_i);
  if ((rc = finish_out(c, idx, cnt, id))) return rc;
  if ((rc = finish_out(c, score, cnt, sd))) return rc;
  if ((rc = finish_out(c, d_p, cnt, pa))) return rc;
  if ((rc = finish_out(c, d_i, cnt, ia))) return rc;
  if (!any_host) return SODSO_OK;
  if ((rc = sync_ctx(c))) return rc;
  return comm_check(c);
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
  if (nranks > 1 && px_enabled) {
    int rc = px_setup(c);
    if (rc) return rc;
  }
  return SODSO_OK;
}

int sodso_comm_exchange(sodso_ctx *c) { return c && c->comm && c->comm->p2p ? 1 : 0; }

int sodso_debug_set_peer_exchange(int on) {
  px_enabled = on != 0;
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
  int rc;
  if ((rc = join_xchg(c))) return rc;
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
  const bool pipelined = can_pipeline(c, idx, score, d_p, d_i);
  if ((rc = pipelined ? pipeline_begin(db) : join_xchg(c))) return rc;
  if ((rc = db_match_async(db, hist1, m))) return rc;
  return finish_sharded(db, q_global_row0, mask_width, p_weight, k, idx, score, d_p, d_i, pipelined);
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
  if ((rc = join_xchg(c))) return rc;
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
