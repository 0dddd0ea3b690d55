// Private declarations shared by the extern "C" translation units (capi.cu, sharded.cu): the opaque handle
// structs, device workspaces and the host<->HBM staging helpers.  Not installed; include/sodso_pr.h is the ABI.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sodso_pr.h"
#include "../../include/sodso_pr_debug.h"
#include "common.cuh"

namespace sodso {

struct Buf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

inline bool is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

struct CommState;   // sharded.cu: NCCL communicator of a context (row-sharded database)

}  // namespace sodso

struct sodso_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host -> HBM chunk copies of the streamed path
  int algo = SODSO_ALGO_TC;
  bool sc_symmetry = true;             // self-match symmetry of the tcgen05 matcher (sodso_debug_set_sc_symmetry)
  int stream_min_scans = 2048;         // sodso_ctx_set_stream_threshold
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  std::string kname;
  sodso::CommState *comm = nullptr;
  // pipelined sharded queries (device outputs): the exchange / fusion / merge of batch b runs on a second stream while
  // the match of batch b + 1 runs on the main one
  cudaStream_t xchg_stream = nullptr;
  cudaEvent_t ev_stats[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  bool ev_done_rec[2] = {false, false};
  unsigned batch_no = 0;               // sharded query batches on this context (epoch of the peer-memory exchange)
  // workspaces
  sodso::Buf in_xyz, in_inten, in_off, out_hist, out_xyz, out_evec;
  sodso::Buf h1, h2, q_op, db_op, dp32, di32, dp64, di64;
  sodso::Buf stats, idx64, idx32, score, dpat, diat, gen_ws, m2dp_ws;
};

struct sodso_db {
  sodso_ctx *ctx = nullptr;
  int type = 0;
  int n = 0;           // valid rows
  int cap = 0;         // rows the operand buffer is laid out for (>= n; sodso_db_reserve / sodso_db_append)
  int64_t row0 = 0;
  sodso::Buf op;       // SC: MMA operand (TC) or normalised K-major fp32 (SIMT); M2DP: raw fp64 rows
  int op_algo = 0;
  sodso::Buf q_in, q_op, dp, di, stats, gstats, idx, score, dpat, diat, ws;
  sodso::Buf q_hist, pack, gather;   // sharded query: generated query signatures, local / gathered top-k lists
  sodso::Buf q_xyz, q_inten, q_off;  // sharded query from scans: staged query points
  sodso::Buf dp2, di2, stats2, pack2, gstats2, gather2;   // second buffer set of the pipelined sharded queries
  int m = 0;           // rows of the last match
  bool matched = false;
};

namespace sodso {

#define CTX_CHECK(ctx)                                   \
  if (!(ctx)) {                                          \
    set_error("null context");                           \
    return SODSO_E_ARG;                                  \
  }                                                      \
  SODSO_CUDA_CHECK(cudaSetDevice((ctx)->device))

// host-or-device input -> device pointer (copies through `ws` if host)
template <class T>
int stage_in(sodso_ctx *c, const T *src, size_t count, Buf &ws, const T **out) {
  if (count == 0 || is_device_ptr(src)) {
    *out = src;
    return SODSO_OK;
  }
  SODSO_CUDA_CHECK(ws.reserve(count * sizeof(T)));
  SODSO_CUDA_CHECK(cudaMemcpyAsync(ws.p, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  *out = ws.as<T>();
  return SODSO_OK;
}

// host-or-device output -> device pointer to write to
template <class T>
int stage_out(sodso_ctx *c, T *dst, size_t count, Buf &ws, T **out) {
  if (!dst) {
    *out = nullptr;
    return SODSO_OK;
  }
  if (is_device_ptr(dst)) {
    *out = dst;
    return SODSO_OK;
  }
  SODSO_CUDA_CHECK(ws.reserve(count * sizeof(T) + 16));
  *out = ws.as<T>();
  return SODSO_OK;
}

template <class T>
int finish_out(sodso_ctx *c, T *dst, size_t count, const T *dev) {
  if (!dst || dst == dev) return SODSO_OK;
  SODSO_CUDA_CHECK(cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
  return SODSO_OK;
}

inline int sync_ctx(sodso_ctx *c) {
  SODSO_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (c->xchg_stream) SODSO_CUDA_CHECK(cudaStreamSynchronize(c->xchg_stream));
  return SODSO_OK;
}

// the main stream waits for everything a pipelined sharded query has put on the exchange stream
inline int join_xchg(sodso_ctx *c) {
  for (int f = 0; f < 2; f++)
    if (c->ev_done_rec[f]) SODSO_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_done[f], 0));
  return SODSO_OK;
}

// capi.cu
int check_offsets_host(const int64_t *off, int nscan, int64_t *total);
int sc_prepare(sodso_ctx *c, int algo, const double *hist_dev, int rows, Buf &op, bool is_db);
int db_match_async(sodso_db *db, const double *hist1, int m);   // sodso_db_match without the final synchronisation
int db_match_prepared_async(sodso_db *db, int m);               // Scan Context: db->q_op already holds the m queries
int db_stream_match_async(sodso_db *db, const double *xyz, const float *inten, const int64_t *off, double max_rho, int m,
                          bool self, double *hist_dev);
// sharded.cu
void comm_release(sodso_ctx *c);
int comm_check(sodso_ctx *c);

}  // namespace sodso
