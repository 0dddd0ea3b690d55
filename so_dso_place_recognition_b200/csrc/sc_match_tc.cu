// Scan Context all-pairs x all-shifts match (processSC.m:12-45) on the 5th-generation tensor cores.
//
// For one channel, d(i, j) = min over the 120 variants v of (1 - <variant_v(q_i), h_j>) / 2
// (processSC.m:24-31).  The 120 variants are the 60 circular sector shifts of the query image x
// and the 60 circular sector shifts of its sector-reversed image y (reverse shift k of x ==
// forward shift (61-k) mod 60 of y).  So per (query, DB row) we need the 2 x 60 correlations
//      corr_b[s] = sum_{c, r} b[(c + s) mod 60][r] * h[c][r],        b in {x, y}
// i.e. a dense contraction  D[j, (b, s)] = sum_k A[j, k] * B[(b, s), k]  with
//      A = DB signatures           (M side: 256 DB rows per CTA pair, TMA-fed, SWIZZLE_128B)
//      B = Hankel matrix of shifts (N side, never materialised)
//
// The Hankel operand.  K is ordered so that one 16-byte "unit" holds slots of ONE sector, and a
// base vector is stored doubled ([b, b], unit u = sector u mod 60).  In the canonical K-major
// no-swizzle UMMA layout ((8,n),2):((16 B, SBO), LBO) rows inside an 8-row core matrix are 16 B
// apart, so a descriptor with SBO = 128 B reads row r at unit (base + r): overlapping windows of one
// small buffer ARE the shifted copies.  Two queries are interleaved unit-wise (Z[2u + b] = q_b[u]);
// with LBO = 32 B, row r = 2 s + b then is shift s of query b, and one N = 256 MMA of a CTA pair
// (cta_group::2, M = 256, N = 240; CTA 0 supplies the 120 x-rows, CTA 1 the 120 y-rows of B)
// produces exactly the 120 variants of two queries against 256 DB rows.  A query costs 16 KB (fp16) or
// 4 KB (fp8) of shared memory per CTA instead of a 120 x 1200 expanded tile per K-block.
//
// Two arithmetic modes, chosen per channel on the device:
//  * generic (any real values): rows are normalised in fp64 (processSC.m:15-20), scaled by 64 and
//    split v = hi + lo into two fp16; the product is evaluated as hi*lo + lo*hi + hi*hi (lo*lo ~
//    2^-22 dropped) with fp32 accumulation in TMEM.  The three terms are interleaved per sector
//    into 64 fp16 slots (20 + 20 + 20 + 4 pad), cross terms first so that the large hi*hi partial
//    sums come last.  |d - d_ref| ~ 1e-6 (measured), bar 1e-5.
//  * binary (every value 0 or 1 on both sides -- the intensity channel, SC.cpp:67-72): the raw bits
//    go in as e2m1 under kind::mxf4 (32 slots per 16-byte unit; block scales all 1.0, parked in the
//    TMEM columns the N = 240 accumulators leave free), TMEM accumulates exact integer overlap
//    counts, and the epilogue applies 1/(|q| |h|).  7.5x fewer MMAs, exact up to the final fp32
//    rounding.  (An e4m3 kind::f8f6f4 variant of the same path is kept behind debug flag 8.)
//
// Roles per CTA (256 threads): warp 0 TMA producer (DB tiles), warp 1 MMA issuer (leader CTA; the
// warp runs its loop uniformly, one elected lane issues; K loop specialised per operand format),
// warp 2 TMEM allocator, warp 3 query-operand loader, warps 4-7 epilogue (tcgen05.ld -> max over
// the shift columns -> (1 - x)/2 -> coalesced fp32 stores).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "../../include/sodso_pr.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace sodso {
namespace {
using namespace tc;

constexpr int UNITS_PER_CHUNK = SC_NUM_S;            // 60 units (16 B) of K per chunk
constexpr int F16_CHUNKS = 8;                        // 64 fp16 slots per sector
constexpr int F8_CHUNKS = 2;                         // 32 fp8 slots per sector
constexpr int K_F16 = F16_CHUNKS * UNITS_PER_CHUNK * 8;    // 3840 elements
constexpr int K_F8 = F8_CHUNKS * UNITS_PER_CHUNK * 16;     // 1920 elements (bytes)
constexpr int F4_ROW_BYTES = 1024;                   // e2m1: 60 units x 32 slots (20 rings + 12 pad) + 4 zero units
constexpr int F4_KB = F4_ROW_BYTES / 128;            // 8 K-blocks
constexpr int KB_UNITS = 8;                          // units per K-block (128 B TMA box row)
constexpr int Q_UNITS = 256;                         // units per (query pair, base, chunk): 2 x doubled vector
constexpr int CHUNK_BYTES = Q_UNITS * 16;            // 4 KB
constexpr int QG = 4;                                // queries per tile = 2 interleaved pairs
constexpr int TILE_M = 256, CTA_M = 128;             // DB rows per CTA pair / per CTA
constexpr int N_MMA = 256;                           // TMEM column stride between the two query pairs
constexpr int N_INST = 240;                          // MMA N: 2 bases x 60 shifts x 2 interleaved queries
constexpr int A_STAGE_BYTES = CTA_M * 128;           // 16 KB
constexpr int NSTAGE = 8;
constexpr int B_BYTES = 2 * F16_CHUNKS * CHUNK_BYTES;   // 64 KB (two pairs, generic mode)
constexpr float VAL_SCALE = 64.0f;                   // operand scale (generic mode)
constexpr float ACC_SCALE = 1.0f / (VAL_SCALE * VAL_SCALE);
constexpr int TC_THREADS = 256;
constexpr int HEADER_BYTES = 256;

struct __align__(8) TcBarriers {
  uint64_t full[NSTAGE], empty[NSTAGE];
  uint64_t b_full, b_peer, b_empty, tmem_full, tmem_empty;
  uint32_t tmem_ptr;
  uint32_t pad;
};
constexpr int SMEM_BYTES = 1024 /*align*/ + NSTAGE * A_STAGE_BYTES + B_BYTES + (int)sizeof(TcBarriers);

inline int pad_to(int v, int a) { return (v + a - 1) / a * a; }

// Operand buffers in HBM.  header: int nonbinary[2] (per channel: some value is not 0/1).
struct DbLayout {
  size_t off_f16, off_f8, off_f4, off_norm, total;
  int n_pad;
  explicit DbLayout(int n) {
    n_pad = pad_to(n, TILE_M);
    off_f16 = HEADER_BYTES;                                      // [ch][n_pad][3840] fp16
    off_f8 = off_f16 + (size_t)2 * n_pad * K_F16 * 2;            // [ch][n_pad][1920] e4m3
    off_f4 = off_f8 + (size_t)2 * n_pad * K_F8;                  // [ch][n_pad][1024 B] e2m1, two per byte
    off_norm = off_f4 + (size_t)2 * n_pad * F4_ROW_BYTES;        // [ch][n_pad] float 1/|h|
    total = off_norm + (size_t)2 * n_pad * 4;
  }
};
struct QLayout {
  size_t off_f16, off_f8, off_f4, off_norm, total;
  int m_pad;
  explicit QLayout(int m) {
    m_pad = pad_to(m, QG);
    off_f16 = HEADER_BYTES;                                      // [ch][base][m_pad/2][8][256][16 B]
    off_f8 = off_f16 + (size_t)2 * 2 * (m_pad / 2) * F16_CHUNKS * CHUNK_BYTES;
    off_f4 = off_f8 + (size_t)2 * 2 * (m_pad / 2) * F8_CHUNKS * CHUNK_BYTES;     // [ch][base][m_pad/2][256][16 B]
    off_norm = off_f4 + (size_t)2 * 2 * (m_pad / 2) * CHUNK_BYTES;               // [ch][m_pad] float
    total = off_norm + (size_t)2 * m_pad * 4;
  }
};

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_fp16(double v, __half &hi, __half &lo) {
  hi = __float2half_rn((float)v);
  lo = __float2half_rn((float)(v - (double)__half2float(hi)));
}

// per-row preparation shared by both operands: norms, fp16 split, binary test
struct RowPrep {
  __half hi[2][SC_SIZE], lo[2][SC_SIZE];
  unsigned char bits[2][SC_SIZE];   // e4m3 encoding of the raw value when it is 0 or 1
  double red[64];
  int nonbin[2];
  float inv_norm[2];
};

__device__ inline void prep_row(const double *h, bool valid, RowPrep &S) {
  if (threadIdx.x < 2) S.nonbin[threadIdx.x] = 0;
  double ss[2] = {0.0, 0.0};
  int nb[2] = {0, 0};
  if (valid)
    for (int ch = 0; ch < 2; ch++)
      for (int k = threadIdx.x; k < SC_SIZE; k += blockDim.x) {
        double v = h[ch * SC_SIZE + k];
        ss[ch] += v * v;
        if (v != 0.0 && v != 1.0) nb[ch] = 1;
      }
  for (int ch = 0; ch < 2; ch++) {
    double s = ss[ch];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) S.red[ch * 32 + (threadIdx.x >> 5)] = s;
  }
  __syncthreads();
  for (int ch = 0; ch < 2; ch++)
    if (nb[ch]) atomicOr(&S.nonbin[ch], 1);
  double nrm[2];
  for (int ch = 0; ch < 2; ch++) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += S.red[ch * 32 + w];
    nrm[ch] = sqrt(s);  // processSC.m:16,19
  }
  for (int ch = 0; ch < 2; ch++)
    for (int k = threadIdx.x; k < SC_SIZE; k += blockDim.x) {
      if (valid) {
        double v = h[ch * SC_SIZE + k];
        split_fp16(v / nrm[ch] * (double)VAL_SCALE, S.hi[ch][k], S.lo[ch][k]);
        S.bits[ch][k] = v == 1.0 ? 0x38 : 0x00;  // e4m3 1.0 / 0.0
      } else {
        S.hi[ch][k] = S.lo[ch][k] = __float2half(0.0f);
        S.bits[ch][k] = 0;
      }
    }
  if (threadIdx.x < 2) S.inv_norm[threadIdx.x] = valid ? (float)(1.0 / nrm[threadIdx.x]) : 0.0f;
  __syncthreads();
}

// fp16 slot s of a sector: which part of the split value an operand supplies
//   DB   : slots 0..19 hi, 20..39 lo, 40..59 hi, 60..63 zero
//   query: slots 0..19 lo, 20..39 hi, 40..59 hi, 60..63 zero          (hi*lo + lo*hi + hi*hi)
__device__ __forceinline__ __half slot_value(const RowPrep &S, int ch, int sector, int s, bool is_db) {
  if (s >= 60) return __float2half(0.0f);
  const int r = s < 20 ? s : (s < 40 ? s - 20 : s - 40);
  const bool use_lo = is_db ? (s >= 20 && s < 40) : (s < 20);
  return use_lo ? S.lo[ch][sector * SC_NUM_R + r] : S.hi[ch][sector * SC_NUM_R + r];
}
// fp8 slot s (0..31) of a sector: rings 0..19, then zero padding
__device__ __forceinline__ unsigned char slot_bits(const RowPrep &S, int ch, int sector, int s) {
  return s < SC_NUM_R ? S.bits[ch][sector * SC_NUM_R + s] : (unsigned char)0;
}

// e2m1 byte t (0..15) of a sector's unit: slots 2t (low nibble) and 2t + 1 (high nibble), 1.0 = 0b0010
__device__ __forceinline__ unsigned char slot_nibbles(const RowPrep &S, int ch, int sector, int t) {
  const unsigned lo = slot_bits(S, ch, sector, 2 * t) ? 0x2u : 0u, hi = slot_bits(S, ch, sector, 2 * t + 1) ? 0x20u : 0u;
  return (unsigned char)(lo | hi);
}

__global__ void __launch_bounds__(256)
sc_tc_prep_db_kernel(const double *__restrict__ hist, int n, int n_pad, unsigned char *__restrict__ buf,
                     size_t off_f16, size_t off_f8, size_t off_f4, size_t off_norm, int row0, int want_f8) {
  __shared__ RowPrep S;
  const int row = row0 + blockIdx.x;
  prep_row(hist + (size_t)row * 2 * SC_SIZE, row < n, S);
  if (threadIdx.x < 2 && S.nonbin[threadIdx.x]) atomicOr(reinterpret_cast<int *>(buf) + threadIdx.x, 1);
  for (int ch = 0; ch < 2; ch++) {
    // byte k = unit * 16 + t ; unit = sector (units 60..63: zero padding of the 1024-byte row)
    unsigned char *o4 = buf + off_f4 + ((size_t)ch * n_pad + row) * F4_ROW_BYTES;
    for (int k = threadIdx.x; k < F4_ROW_BYTES; k += blockDim.x)
      o4[k] = (k >> 4) < SC_NUM_S ? slot_nibbles(S, ch, k >> 4, k & 15) : (unsigned char)0;
    // k = chunk*480 + sector*8 + t ; slot = chunk*8 + t
    __half *o = reinterpret_cast<__half *>(buf + off_f16) + ((size_t)ch * n_pad + row) * K_F16;
    for (int k = threadIdx.x; k < K_F16; k += blockDim.x) {
      const int j = k / (UNITS_PER_CHUNK * 8), rem = k - j * (UNITS_PER_CHUNK * 8);
      o[k] = slot_value(S, ch, rem >> 3, j * 8 + (rem & 7), true);
    }
    // k = chunk*960 + sector*16 + t ; slot = chunk*16 + t
    unsigned char *o8 = buf + off_f8 + ((size_t)ch * n_pad + row) * K_F8;
    for (int k = threadIdx.x; want_f8 && k < K_F8; k += blockDim.x) {
      const int j = k / (UNITS_PER_CHUNK * 16), rem = k - j * (UNITS_PER_CHUNK * 16);
      o8[k] = slot_bits(S, ch, rem >> 4, j * 16 + (rem & 15));
    }
    if (threadIdx.x == 0) reinterpret_cast<float *>(buf + off_norm)[(size_t)ch * n_pad + row] = S.inv_norm[ch];
  }
}

// Query operand: per (channel, base, query pair): [chunk][unit 2u+b][16 B], unit u = sector u % 60 of
// the base vector of query b of the pair (x: the query image, y: its sector reversal y[c] = x[(60-c)%60]).
// One CTA prepares a query PAIR, so that the interleaved units are written as one contiguous, fully coalesced
// stream of 16-byte stores.
__global__ void __launch_bounds__(256)
sc_tc_prep_query_kernel(const double *__restrict__ hist, int m, int m_pad, unsigned char *__restrict__ buf,
                        size_t off_f16, size_t off_f8, size_t off_f4, size_t off_norm, int pair0, int want_f8) {
  __shared__ RowPrep S[2];
  const int pair = pair0 + blockIdx.x, npairs = m_pad >> 1;
  for (int b = 0; b < 2; b++) {
    const int row = 2 * pair + b;
    prep_row(hist + (size_t)row * 2 * SC_SIZE, row < m, S[b]);
    if (threadIdx.x < 2 && S[b].nonbin[threadIdx.x]) atomicOr(reinterpret_cast<int *>(buf) + threadIdx.x, 1);
    if (threadIdx.x < 2)
      reinterpret_cast<float *>(buf + off_norm)[(size_t)threadIdx.x * m_pad + row] = S[b].inv_norm[threadIdx.x];
  }
  for (int ch = 0; ch < 2; ch++)
    for (int base = 0; base < 2; base++) {
      const size_t pb = ((size_t)ch * 2 + base) * npairs + pair;
      // fp16: 8 chunks x 256 units of 8 halves
      uint4 *o = reinterpret_cast<uint4 *>(buf + off_f16 + pb * F16_CHUNKS * CHUNK_BYTES);
      for (int e = threadIdx.x; e < F16_CHUNKS * Q_UNITS; e += blockDim.x) {
        const int b = e & 1, u = (e >> 1) & 127, j = e >> 8;
        const int cs = u % SC_NUM_S, c = base == 0 ? cs : (SC_NUM_S - cs) % SC_NUM_S;
        __align__(16) __half v[8];
#pragma unroll
        for (int t = 0; t < 8; t++) v[t] = slot_value(S[b], ch, c, j * 8 + t, false);
        o[e] = *reinterpret_cast<const uint4 *>(v);
      }
      if (want_f8) {
        uint4 *o8 = reinterpret_cast<uint4 *>(buf + off_f8 + pb * F8_CHUNKS * CHUNK_BYTES);
        for (int e = threadIdx.x; e < F8_CHUNKS * Q_UNITS; e += blockDim.x) {
          const int b = e & 1, u = (e >> 1) & 127, j = e >> 8;
          const int cs = u % SC_NUM_S, c = base == 0 ? cs : (SC_NUM_S - cs) % SC_NUM_S;
          __align__(16) unsigned char v[16];
#pragma unroll
          for (int t = 0; t < 16; t++) v[t] = slot_bits(S[b], ch, c, j * 16 + t);
          o8[e] = *reinterpret_cast<const uint4 *>(v);
        }
      }
      uint4 *o4 = reinterpret_cast<uint4 *>(buf + off_f4 + pb * CHUNK_BYTES);
      for (int e = threadIdx.x; e < Q_UNITS; e += blockDim.x) {
        const int b = e & 1, u = e >> 1;
        const int cs = u % SC_NUM_S, c = base == 0 ? cs : (SC_NUM_S - cs) % SC_NUM_S;
        __align__(16) unsigned char v[16];
#pragma unroll
        for (int t = 0; t < 16; t++) v[t] = slot_nibbles(S[b], ch, c, t);
        o4[e] = *reinterpret_cast<const uint4 *>(v);
      }
    }
}

struct TcParams {
  const unsigned char *q_buf;   // QLayout
  const unsigned char *db_buf;  // DbLayout
  size_t q_off_f16, q_off_f8, q_off_f4, q_off_norm, db_off_norm;
  float *d_out[2];              // per channel, m x ldd
  int m, n, m_pad, n_pad, ldd;
  // Work of one launch: up to two rectangles of (query group, DB tile) items (the streamed path matches an L-shaped
  // region per chunk: new queries x all DB rows so far + old queries x new DB rows).  Per rectangle: units = 2 channels x
  // query groups of 4, tiles of 256 DB rows, first query group / first tile.  w0 = number of items of rectangle 0.
  int n_units[2], n_tiles[2], qg0[2], tile0[2];
  long long w0, w_total;
  int flags;                    // debug: 1 skip epilogue loads, 2 skip MMAs, 4 force generic mode, 8 binary channel in e4m3
};

// One work item's K loop on the issuing thread, specialised per operand format so that the loop body is
// straight-line: descriptors are advanced by integer adds on their 16-byte address field.
//   MODE 0: fp16 3-term split (kind::f16), 1: e4m3 (kind::f8f6f4), 2: e2m1 (kind::mxf4, block scales = 1)
template <int MODE>
__device__ __forceinline__ void issue_k_loop(TcBarriers *bars, uint32_t sA, uint32_t sB, uint32_t tmem_base,
                                             uint32_t tmem_sf, int &stage, uint32_t &phase, bool skip) {
  constexpr int NCHUNK = MODE == 0 ? F16_CHUNKS : (MODE == 1 ? F8_CHUNKS : 1);
  constexpr int NUM_KB = MODE == 2 ? F4_KB : NCHUNK * UNITS_PER_CHUNK / KB_UNITS;
  constexpr uint32_t PAIR_UNITS = NCHUNK * CHUNK_BYTES / 16;     // second query pair, in 16-byte descriptor units
  constexpr uint32_t idesc = MODE == 2 ? make_idesc_mxf4(TILE_M, N_INST) : make_idesc(TILE_M, N_INST);
  // A: SWIZZLE_128B K-major, 8-row groups 1024 B apart; a K-step advances the start by 32 B (2 units)
  const uint64_t adesc0 = make_desc(sA, 16, 1024, 2);
  // B: Hankel view, no swizzle: row r = 2 s + b, k-group g -> unit 2 (c + s + g) + b of chunk j
  const uint64_t bdesc0 = make_desc(sB, 32, 128, 0);
  uint32_t boff = 0;    // (chunk j) * 256 + 2 * (sector c): the K position of the next MMA in the Hankel buffer
  uint32_t c2 = 0;      // 2 * c
  uint32_t acc = 0;
#pragma unroll 1
  for (int kb = 0; kb < NUM_KB; kb++) {
    mbar_wait(smem_u32(&bars->full[stage]), phase, 7);
    tc_fence_after();
    const uint64_t ad = adesc0 + (uint64_t)(stage * (A_STAGE_BYTES >> 4));
    if (!skip) {
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const uint64_t a = ad + (uint64_t)(kk * 2), b = bdesc0 + (uint64_t)boff;
        if (elect_one()) {
          if (MODE == 0) {
            umma_f16_2sm(tmem_base, a, b, idesc, acc);
            umma_f16_2sm(tmem_base + N_MMA, a, b + PAIR_UNITS, idesc, acc);
          } else if (MODE == 1) {
            umma_f8_2sm(tmem_base, a, b, idesc, acc);
            umma_f8_2sm(tmem_base + N_MMA, a, b + PAIR_UNITS, idesc, acc);
          } else {
            umma_mxf4_2sm(tmem_base, a, b, idesc, acc, tmem_sf, tmem_sf + 8);
            umma_mxf4_2sm(tmem_base + N_MMA, a, b + PAIR_UNITS, idesc, acc, tmem_sf, tmem_sf + 8);
          }
        }
        acc = 1;
        boff += 4;
        if (MODE != 2) {   // e2m1: one chunk; units 60..63 of the DB row are zero, what the view reads there is moot
          c2 += 4;
          if (c2 >= 2 * UNITS_PER_CHUNK) {
            c2 -= 2 * UNITS_PER_CHUNK;
            boff += CHUNK_BYTES / 16 - 2 * UNITS_PER_CHUNK;
          }
        }
      }
    }
    if (elect_one()) umma_commit_2sm(smem_u32(&bars->empty[stage]));
    __syncwarp();
    if (++stage == NSTAGE) {
      stage = 0;
      phase ^= 1;
    }
  }
}

// work item -> (unit id unique within the launch, channel, query group, DB tile)
__device__ __forceinline__ void tc_decode(const TcParams &P, long long it, int &unit_id, int &ch, int &qg, int &tile) {
  const int r = it >= P.w0;
  const long long l = it - (r ? P.w0 : 0);
  const int nt = P.n_tiles[r];
  const int unit = (int)(l / nt);
  tile = P.tile0[r] + (int)(l - (long long)unit * nt);
  ch = unit & 1;
  qg = P.qg0[r] + (unit >> 1);
  unit_id = unit | (r << 28);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
sc_match_tc_kernel(const __grid_constant__ CUtensorMap map_f16_0, const __grid_constant__ CUtensorMap map_f16_1,
                   const __grid_constant__ CUtensorMap map_f8_0, const __grid_constant__ CUtensorMap map_f8_1,
                   const __grid_constant__ CUtensorMap map_f4_0, const __grid_constant__ CUtensorMap map_f4_1,
                   const TcParams P) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_raw = smem_u32(smem_dyn);
  const uint32_t base = (base_raw + 1023u) & ~1023u;
  unsigned char *smem = smem_dyn + (base - base_raw);
  const uint32_t sA = base;                               // NSTAGE x 16 KB (1024-aligned)
  const uint32_t sB = base + NSTAGE * A_STAGE_BYTES;      // 64 KB
  TcBarriers *bars = reinterpret_cast<TcBarriers *>(smem + NSTAGE * A_STAGE_BYTES + B_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  // binary mode per channel: both operands hold only 0/1 in that channel
  const int *qf = reinterpret_cast<const int *>(P.q_buf), *df = reinterpret_cast<const int *>(P.db_buf);
  bool binary[2];
  for (int ch = 0; ch < 2; ch++) binary[ch] = !(P.flags & 4) && qf[ch] == 0 && df[ch] == 0;

  // work items of this CTA pair: contiguous range in unit-major order
  const long long W = P.w_total;
  const long long it_begin = W * pair_id / npairs, it_end = W * (pair_id + 1) / npairs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; s++) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->b_full), 1);
    mbar_init(smem_u32(&bars->b_peer), 1);
    mbar_init(smem_u32(&bars->b_empty), 1);
    mbar_init(smem_u32(&bars->tmem_full), 1);
    mbar_init(smem_u32(&bars->tmem_empty), 8);  // 4 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_ptr)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_ptr;
  // binary channels run as e2m1 under kind::mxf4 (2x the e4m3 rate); its block scales (ue8m0, all 1.0 = 0x7f)
  // live in the 16 TMEM columns that the N = 240 accumulators leave free in each 256-column slot
  const bool use_f4 = !(P.flags & 8);
  const uint32_t tmem_sf = tmem_base + (uint32_t)N_INST;
  if (warp >= 4) {
    tmem_st16_const(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)N_INST, 0x7f7f7f7fu);
    tmem_st16_const(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(N_MMA + N_INST), 0x7f7f7f7fu);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();

  if (warp == 0) {
    // ===== TMA producer: this CTA's 128 DB rows of every K-block =====
    if (lane == 0) {
      const uint32_t leader_full0 = map_to_cta(smem_u32(&bars->full[0]), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (long long it = it_begin; it < it_end; ++it) {
        int unit, ch, qg, tile;
        tc_decode(P, it, unit, ch, qg, tile);
        const bool bin = binary[ch];
        const CUtensorMap *map = bin ? (use_f4 ? (ch == 0 ? &map_f4_0 : &map_f4_1) : (ch == 0 ? &map_f8_0 : &map_f8_1))
                                     : (ch == 0 ? &map_f16_0 : &map_f16_1);
        const int num_kb = bin ? (use_f4 ? F4_KB : F8_CHUNKS * UNITS_PER_CHUNK / KB_UNITS) : F16_CHUNKS * UNITS_PER_CHUNK / KB_UNITS;
        const int kb_elems = bin ? 128 : 64;
        const int row0 = tile * TILE_M + (int)rank * CTA_M;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1, 1);
          if (leader) mbar_expect_tx(smem_u32(&bars->full[stage]), 2 * A_STAGE_BYTES);
          tma_load_2d_2sm(sA + stage * A_STAGE_BYTES, map, kb * kb_elems, row0, leader_full0 + stage * 8);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      // drain: every stage released (all multicast commits have landed in this CTA) before exit
      if (it_end > it_begin)
        for (int i = 0; i < NSTAGE; i++) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1, 9);
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
    }
  } else if (warp == 3) {
    // ===== query operand loader: 2 interleaved query pairs of this CTA's base (x: rank 0, y: rank 1) =====
    if (lane == 0) {
      uint32_t phase = 0;
      int prev_unit = -1;
      const int qpairs = P.m_pad >> 1;
      for (long long it = it_begin; it < it_end; ++it) {
        int unit, ch, qg, tile;
        tc_decode(P, it, unit, ch, qg, tile);
        if (unit == prev_unit) continue;
        prev_unit = unit;
        const bool bin = binary[ch];
        const uint32_t pair_bytes = (bin ? (use_f4 ? 1 : F8_CHUNKS) : F16_CHUNKS) * CHUNK_BYTES;
        mbar_wait(smem_u32(&bars->b_empty), phase ^ 1, 2);
        mbar_expect_tx(smem_u32(&bars->b_full), 2 * pair_bytes);
        const unsigned char *src = P.q_buf + (bin ? (use_f4 ? P.q_off_f4 : P.q_off_f8) : P.q_off_f16) +
                                   (((size_t)ch * 2 + rank) * qpairs + (size_t)qg * 2) * pair_bytes;
        for (int p = 0; p < 2; p++)
          bulk_load_1d(sB + p * pair_bytes, src + (size_t)p * pair_bytes, pair_bytes, smem_u32(&bars->b_full));
        if (!leader) {
          mbar_wait(smem_u32(&bars->b_full), phase, 3);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive_remote(map_to_cta(smem_u32(&bars->b_peer), 0));
        }
        phase ^= 1;
      }
      if (prev_unit >= 0) mbar_wait(smem_u32(&bars->b_empty), phase ^ 1, 10);  // drain the last commit
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA; the warp runs the loop uniformly, one elected lane issues) =====
    if (leader) {
      int stage = 0;
      uint32_t phase = 0, b_phase = 0, t_phase = 0;
      int prev_unit = -1;
      const bool skip = (P.flags & 2) != 0;
      for (long long it = it_begin; it < it_end; ++it) {
        int unit, ch, qg, tile;
        tc_decode(P, it, unit, ch, qg, tile);
        if (unit != prev_unit) {
          prev_unit = unit;
          mbar_wait(smem_u32(&bars->b_full), b_phase, 4);
          mbar_wait(smem_u32(&bars->b_peer), b_phase, 5);
          b_phase ^= 1;
        }
        mbar_wait(smem_u32(&bars->tmem_empty), t_phase ^ 1, 6);
        tc_fence_after();
        if (!binary[ch])
          issue_k_loop<0>(bars, sA, sB, tmem_base, tmem_sf, stage, phase, skip);
        else if (use_f4)
          issue_k_loop<2>(bars, sA, sB, tmem_base, tmem_sf, stage, phase, skip);
        else
          issue_k_loop<1>(bars, sA, sB, tmem_base, tmem_sf, stage, phase, skip);
        bool last_of_unit = it + 1 == it_end;
        if (!last_of_unit) {
          int u2, c2, g2, t2;
          tc_decode(P, it + 1, u2, c2, g2, t2);
          last_of_unit = u2 != unit;
        }
        if (elect_one()) {
          umma_commit_2sm(smem_u32(&bars->tmem_full));
          if (last_of_unit) umma_commit_2sm(smem_u32(&bars->b_empty));
        }
        __syncwarp();
        t_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> max over the shift columns -> d = (1 - dot)/2 -> global =====
    const int ew = warp & 3;  // TMEM lane quarter accessible to this warp
    const uint32_t leader_tmem_empty = map_to_cta(smem_u32(&bars->tmem_empty), 0);
    const float *q_norm = reinterpret_cast<const float *>(P.q_buf + P.q_off_norm);
    const float *db_norm = reinterpret_cast<const float *>(P.db_buf + P.db_off_norm);
    uint32_t t_phase = 0;
    for (long long it = it_begin; it < it_end; ++it) {
      int unit, ch, qg, tile;
      tc_decode(P, it, unit, ch, qg, tile);
      const bool bin = binary[ch];
      mbar_wait(smem_u32(&bars->tmem_full), t_phase, 8);
      tc_fence_after();
      const int row = tile * TILE_M + (int)rank * CTA_M + ew * 32 + lane;
      float *out = P.d_out[ch];
      const float rd = bin ? db_norm[(size_t)ch * P.n_pad + row] : 0.0f;
      if (!(P.flags & 1)) {
#pragma unroll 1
        for (int p = 0; p < 2; p++) {
          // column 120 h + 2 s + b : base h, shift s, query b of the pair
          float best0 = __int_as_float(0x7fc00000), best1 = best0;  // NaN: min ignores NaN (processSC.m:31)
          const uint32_t tcol = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(p * N_MMA);
#pragma unroll 1
          for (int cb = 0; cb + 32 <= N_INST; cb += 32) {
            uint32_t r[32];
            tmem_ld32(tcol + (uint32_t)cb, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              best0 = fmaxf(best0, __uint_as_float(r[i]));
              best1 = fmaxf(best1, __uint_as_float(r[i + 1]));
            }
          }
          {
            static_assert(N_INST % 32 == 16, "tail of 16 columns");
            uint32_t r[16];
            tmem_ld16(tcol + (uint32_t)(N_INST - 16), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              best0 = fmaxf(best0, __uint_as_float(r[i]));
              best1 = fmaxf(best1, __uint_as_float(r[i + 1]));
            }
          }
          const int qi = qg * QG + p * 2;
          if (row < P.n && out) {
            if (bin) {
              if (qi < P.m) out[(size_t)qi * P.ldd + row] = (1.0f - best0 * q_norm[(size_t)ch * P.m_pad + qi] * rd) * 0.5f;
              if (qi + 1 < P.m)
                out[(size_t)(qi + 1) * P.ldd + row] = (1.0f - best1 * q_norm[(size_t)ch * P.m_pad + qi + 1] * rd) * 0.5f;
            } else {
              if (qi < P.m) out[(size_t)qi * P.ldd + row] = (1.0f - best0 * ACC_SCALE) * 0.5f;
              if (qi + 1 < P.m) out[(size_t)(qi + 1) * P.ldd + row] = (1.0f - best1 * ACC_SCALE) * 0.5f;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(leader_tmem_empty);
      t_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
}  // namespace

// debug flags (SODSO_TC_FLAGS): 1 skip epilogue loads, 2 skip MMAs, 4 force generic mode, 8 binary channel in e4m3
static int tc_flags() {
  const char *e = getenv("SODSO_TC_FLAGS");
  return e ? atoi(e) : 0;
}

size_t sc_tc_db_bytes(int n) { return DbLayout(n).total; }
size_t sc_tc_query_bytes(int m) { return QLayout(m).total; }

cudaError_t launch_sc_tc_prep_db(const double *hist, int n, void *db_buf, cudaStream_t st, int64_t *launches) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e = launch_sc_tc_clear_flags(db_buf, st);
  if (e != cudaSuccess) return e;
  return launch_sc_tc_prep_db_rows(hist, n, 0, DbLayout(n).n_pad, db_buf, st, launches);
}

cudaError_t launch_sc_tc_clear_flags(void *buf, cudaStream_t st) { return cudaMemsetAsync(buf, 0, HEADER_BYTES, st); }

// rows [row0, row1) of an n-row operand (row1 may run into the zero padding up to n_pad)
cudaError_t launch_sc_tc_prep_db_rows(const double *hist, int n, int row0, int row1, void *db_buf, cudaStream_t st,
                                      int64_t *launches) {
  if (row1 <= row0) return cudaSuccess;
  DbLayout L(n);
  sc_tc_prep_db_kernel<<<row1 - row0, 256, 0, st>>>(hist, n, L.n_pad, reinterpret_cast<unsigned char *>(db_buf),
                                                    L.off_f16, L.off_f8, L.off_f4, L.off_norm, row0, tc_flags() & 8);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_sc_tc_prep_query_rows(const double *hist, int m, int row0, int row1, void *q_buf, cudaStream_t st,
                                         int64_t *launches) {
  if (row1 <= row0) return cudaSuccess;
  QLayout L(m);
  if (row0 & 1) return cudaErrorInvalidValue;   // one CTA per query pair
  sc_tc_prep_query_kernel<<<(row1 - row0 + 1) / 2, 256, 0, st>>>(hist, m, L.m_pad, reinterpret_cast<unsigned char *>(q_buf),
                                                                 L.off_f16, L.off_f8, L.off_f4, L.off_norm, row0 / 2,
                                                                 tc_flags() & 8);
  if (launches) ++*launches;
  return cudaGetLastError();
}

int sc_tc_db_rows_padded(int n) { return DbLayout(n).n_pad; }
int sc_tc_query_rows_padded(int m) { return QLayout(m).m_pad; }

cudaError_t launch_sc_tc_prep_query(const double *hist, int m, void *q_buf, cudaStream_t st, int64_t *launches) {
  if (m <= 0) return cudaSuccess;
  cudaError_t e = launch_sc_tc_clear_flags(q_buf, st);
  if (e != cudaSuccess) return e;
  return launch_sc_tc_prep_query_rows(hist, m, 0, QLayout(m).m_pad, q_buf, st, launches);
}

cudaError_t launch_sc_match_tc(const void *q_buf, int m, const void *db_buf, int n, float *d_p, float *d_i, int ldd,
                               int num_sms, cudaStream_t st, int64_t *launches) {
  return launch_sc_match_tc_block(q_buf, m, 0, m, db_buf, n, 0, n, d_p, d_i, ldd, num_sms, st, launches);
}

// queries [q0, q1) x DB rows [r0, r1) of the m x n problem; q0 must be a multiple of 4 and r0 of 256
cudaError_t launch_sc_match_tc_block(const void *q_buf, int m, int q0, int q1, const void *db_buf, int n, int r0, int r1,
                                     float *d_p, float *d_i, int ldd, int num_sms, cudaStream_t st,
                                     int64_t *launches) {
  return launch_sc_match_tc_blocks(q_buf, m, db_buf, n, q0, q1, r0, r1, 0, 0, 0, 0, d_p, d_i, ldd, num_sms, st, launches);
}

// two rectangles [qa0, qa1) x [ra0, ra1) and [qb0, qb1) x [rb0, rb1) in one launch (either may be empty)
cudaError_t launch_sc_match_tc_blocks(const void *q_buf, int m, const void *db_buf, int n, int qa0, int qa1, int ra0, int ra1,
                                      int qb0, int qb1, int rb0, int rb1, float *d_p, float *d_i, int ldd, int num_sms,
                                      cudaStream_t st, int64_t *launches) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  const int q0s[2] = {qa0, qb0}, q1s[2] = {qa1, qb1}, r0s[2] = {ra0, rb0}, r1s[2] = {ra1, rb1};
  long long wr[2];
  for (int r = 0; r < 2; r++) {
    const bool empty = q1s[r] <= q0s[r] || r1s[r] <= r0s[r];
    if (!empty && ((q0s[r] % QG) || (r0s[r] % TILE_M))) return cudaErrorInvalidValue;
    wr[r] = empty ? 0 : 2LL * ((q1s[r] - q0s[r] + QG - 1) / QG) * ((r1s[r] - r0s[r] + TILE_M - 1) / TILE_M);
  }
  if (wr[0] + wr[1] == 0) return cudaSuccess;
  PFN_encodeTiled enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  DbLayout DL(n);
  QLayout QL(m);
  const unsigned char *dbb = reinterpret_cast<const unsigned char *>(db_buf);
  CUtensorMap maps[6];
  for (int fmt = 0; fmt < 3; fmt++)
    for (int ch = 0; ch < 2; ch++) {
      const size_t kbytes = fmt == 0 ? (size_t)K_F16 * 2 : (fmt == 1 ? (size_t)K_F8 : (size_t)F4_ROW_BYTES);
      cuuint64_t dims[2] = {(cuuint64_t)(fmt == 0 ? K_F16 : (fmt == 1 ? K_F8 : F4_ROW_BYTES)), (cuuint64_t)DL.n_pad};
      cuuint64_t strides[1] = {(cuuint64_t)kbytes};
      cuuint32_t box[2] = {(cuuint32_t)(fmt == 0 ? 64 : 128), (cuuint32_t)CTA_M};
      cuuint32_t estr[2] = {1, 1};
      void *gaddr = (void *)(dbb + (fmt == 0 ? DL.off_f16 : (fmt == 1 ? DL.off_f8 : DL.off_f4)) + (size_t)ch * DL.n_pad * kbytes);
      CUresult r = enc(&maps[fmt * 2 + ch], fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                       gaddr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
  TcParams P;
  P.q_buf = reinterpret_cast<const unsigned char *>(q_buf);
  P.db_buf = dbb;
  P.q_off_f16 = QL.off_f16;
  P.q_off_f8 = QL.off_f8;
  P.q_off_f4 = QL.off_f4;
  P.q_off_norm = QL.off_norm;
  P.db_off_norm = DL.off_norm;
  P.d_out[0] = d_p;
  P.d_out[1] = d_i;
  P.m = m;
  P.n = n;
  P.m_pad = QL.m_pad;
  P.n_pad = DL.n_pad;
  P.ldd = ldd;
  for (int r = 0; r < 2; r++) {
    P.qg0[r] = q0s[r] / QG;
    P.tile0[r] = r0s[r] / TILE_M;
    P.n_units[r] = wr[r] ? 2 * ((q1s[r] - q0s[r] + QG - 1) / QG) : 0;
    P.n_tiles[r] = wr[r] ? (r1s[r] - r0s[r] + TILE_M - 1) / TILE_M : 1;
  }
  P.w0 = wr[0];
  P.w_total = wr[0] + wr[1];
  P.flags = tc_flags();
  const long long W = P.w_total;
  int npairs = num_sms / 2;
  if (npairs > W) npairs = (int)W;
  if (npairs < 1) npairs = 1;
  cudaError_t e = cudaFuncSetAttribute(sc_match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  sc_match_tc_kernel<<<2 * npairs, TC_THREADS, SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], P);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
