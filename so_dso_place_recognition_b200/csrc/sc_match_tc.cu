// Scan Context all-pairs x all-shifts match (processSC.m:12-45) on the 5th-generation tensor cores.
//
// For one channel, d(i, j) = min over the 120 variants v of (1 - <variant_v(q_i), h_j>) / 2
// (processSC.m:24-31).  The 120 variants are the 60 circular sector shifts of the query image x
// and the 60 circular sector shifts of its sector-reversed image y (reverse shift k of x ==
// forward shift (61-k) mod 60 of y).  So per (query, DB row) we need the 2 x 60 correlations
//      corr_b[s] = sum_{c < 60, r} b[(c + s) mod 60][r] * h[c][r],        b in {x, y}.
//
// Even / odd halving.  With  e[u] = b[u] + b[(u+30) mod 60],  o[u] = b[u] - b[(u+30) mod 60]  (u = 0..59; e is
// 30-periodic, o 30-antiperiodic) and  he[c] = h[c] + h[c+30],  ho[c] = h[c] - h[c+30]  (c = 0..29),
//      E[s] = sum_{c < 30, r} e[c + s][r] he[c][r] = corr[s] + corr[s + 30]
//      O[s] = sum_{c < 30, r} o[c + s][r] ho[c][r] = corr[s] - corr[s + 30]          s = 0..29
// so  max_s corr[s] = max_{s < 30} (E[s] + |O[s]|) / 2 : two contractions of 30 shifts x K = 600 instead of one of
// 60 shifts x K = 1200 -- half the multiply-adds for the same 120 outputs per (query, DB row).
// Each is a dense contraction  D[j, (b, s)] = sum_k A[j, k] * B[(b, s), k]  with
//      A = transformed DB signatures (M side: 256 DB rows per CTA pair, TMA-fed, SWIZZLE_128B)
//      B = Hankel matrix of shifts   (N side, never materialised)
//
// The Hankel operand.  K is ordered so that one 16-byte "unit" holds slots of ONE sector; the 60 units of e (or o)
// already contain every window of 30 sectors.  In the canonical K-major no-swizzle UMMA layout
// ((8,n),2):((16 B, SBO), LBO) rows inside an 8-row core matrix are 16 B apart, so a descriptor with SBO = 128 B reads
// row r at unit (base + r): overlapping windows of one small buffer ARE the shifted copies.  Four queries are
// interleaved unit-wise (Z[4u + b] = q_b[u]); with LBO = 64 B, row r = 4 s + b then is shift s of query b, and one
// N = 240 MMA of a CTA pair (cta_group::2, M = 256; CTA 0 supplies the 120 x-rows, CTA 1 the 120 y-rows of B)
// produces E (TMEM columns 0..239) or O (columns 256..495) of four queries against 256 DB rows.
//
// Two arithmetic modes, chosen per channel on the device:
//  * generic (any real values): rows are normalised in fp64 (processSC.m:15-20), scaled by 64, transformed to e / o in
//    fp64 and split v = hi + lo into two fp16; the product is evaluated as hi*lo + lo*hi + hi*hi (lo*lo ~ 2^-22
//    dropped) with fp32 accumulation in TMEM.  The three terms are interleaved per sector into 64 fp16 slots
//    (20 + 20 + 20 + 4 pad), cross terms first so that the large hi*hi partial sums come last.
//    |d - d_ref| ~ 1e-6 (measured), bar 1e-5.
//  * binary (every value 0 or 1 on both sides -- the intensity channel, SC.cpp:67-72): e in {0,1,2} and o in {-1,0,1}
//    go in as e2m1 under kind::mxf4 (32 slots per 16-byte unit; block scales all 1.0, parked in the TMEM columns the
//    N = 240 accumulators leave free), TMEM accumulates exact integers, and the epilogue applies 1/(|q| |h|).
//
// Self-match.  When queries and DB are the same n signatures, d is symmetric (the variant set is closed under swapping
// the two images), so only the (query group of 4, DB tile of 256) work items with tile_start <= group_end are run and
// the epilogue also stores a value at its transposed position whenever the item owning that position is not run
// (launch_sc_match_tc_self; TcParams::tri).  Work items are split over the CTA pairs by estimated cost.
//
// Roles per CTA (256 threads): warp 0 TMA producer (DB tiles), warp 1 MMA issuer (leader CTA; the
// warp runs its loop uniformly, one elected lane issues; K loop specialised per operand format),
// warp 2 TMEM allocator, warp 3 query-operand loader, warps 4-7 epilogue (tcgen05.ld of E and O -> max of E + |O|
// over the shift columns -> (1 - x/2)/2 -> coalesced fp32 stores).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "../../include/sodso_pr.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace sodso {
namespace {
using namespace tc;

constexpr int HALF_S = SC_NUM_S / 2;                 // 30 sectors per contraction, 30 shifts
constexpr int F16_CHUNKS = 8;                        // 64 fp16 slots per sector = 8 units of 8
constexpr int NCOMP = 2;                             // E, O
constexpr int K_F16 = NCOMP * F16_CHUNKS * HALF_S * 8;     // 3840 elements per DB row: [comp][chunk][sector][8]
constexpr int F4_ROW_BYTES = 1024;                   // e2m1: [E: 30 units][O: 30 units][4 zero units], 32 slots per unit
constexpr int F4_KB = F4_ROW_BYTES / 128;            // 8 K-blocks
constexpr int KB_UNITS = 8;                          // units per K-block (128 B TMA box row)
constexpr int F16_KB = K_F16 / (KB_UNITS * 8);       // 60 K-blocks
constexpr int QG = 4;                                // queries per tile, interleaved unit-wise
constexpr int Q_UNITS = QG * SC_NUM_S;               // 240 units per (query group, base, comp, chunk)
constexpr int BLOCK_BYTES = Q_UNITS * 16;            // 3840 B
constexpr int Q_F16_BYTES = NCOMP * F16_CHUNKS * BLOCK_BYTES;   // 61 440 B per (query group, base, channel)
constexpr int Q_F4_BYTES = NCOMP * BLOCK_BYTES;                 // 7 680 B
constexpr int TILE_M = 256, CTA_M = 128;             // DB rows per CTA pair / per CTA
constexpr int N_MMA = 256;                           // TMEM column stride between the E and the O accumulators
constexpr int N_INST = 240;                          // MMA N: 2 bases x 30 shifts x 4 interleaved queries
constexpr int A_STAGE_BYTES = CTA_M * 128;           // 16 KB
constexpr int NSTAGE = 8;
constexpr int B_BYTES = Q_F16_BYTES;
constexpr float VAL_SCALE = 64.0f;                   // operand scale (generic mode)
constexpr float ACC_SCALE = 0.5f / (VAL_SCALE * VAL_SCALE);   // (E + |O|) / 2, unscaled
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
  size_t off_f16, off_f4, off_norm, total;
  int n_pad;
  explicit DbLayout(int n) {
    n_pad = pad_to(n, TILE_M);
    off_f16 = HEADER_BYTES;                                      // [ch][n_pad][3840] fp16
    off_f4 = off_f16 + (size_t)2 * n_pad * K_F16 * 2;            // [ch][n_pad][1024 B] e2m1, two per byte
    off_norm = off_f4 + (size_t)2 * n_pad * F4_ROW_BYTES;        // [ch][n_pad] float 1/|h|
    total = off_norm + (size_t)2 * n_pad * 4;
  }
};
struct QLayout {
  size_t off_f16, off_f4, off_norm, total;
  int m_pad;
  explicit QLayout(int m) {
    m_pad = pad_to(m, QG);
    off_f16 = HEADER_BYTES;                                      // [ch][base][m_pad/4][comp][chunk][240][16 B]
    off_f4 = off_f16 + (size_t)2 * 2 * (m_pad / QG) * Q_F16_BYTES;   // [ch][base][m_pad/4][comp][240][16 B]
    off_norm = off_f4 + (size_t)2 * 2 * (m_pad / QG) * Q_F4_BYTES;   // [ch][m_pad] float
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

// One (row, channel) as the operands want it: per contraction (E, O) and sequence position u = 0..59 (sector indices
// mod 60; the DB operand uses u < 30, the query windows all 60) the 64 fp16 slots of that sector and its 16 e2m1 bytes.
//   fp16 slots   DB   : 0..19 hi, 20..39 lo, 40..59 hi, 60..63 zero
//                query: 0..19 lo, 20..39 hi, 40..59 hi, 60..63 zero          (hi*lo + lo*hi + hi*hi)
//   e2m1 byte t: rings 2t (low nibble) and 2t + 1 (high nibble) of e in {0,1,2} / o in {-1,0,1}; bytes 10..15 zero
// Sectors are 144 B apart and images 96 B (mod 128) so that the 16-byte reads of the copy-out loops (consecutive lanes:
// consecutive queries of a group, then consecutive sectors) are free of bank conflicts.
constexpr int IMG_SECTOR_HALVES = 72;
struct __align__(16) RowImg {
  __half f16[NCOMP][SC_NUM_S][IMG_SECTOR_HALVES];
  unsigned char f4[NCOMP][SC_NUM_S][16];
  double red[8];
  float inv_norm;
  int nonbin;
  int pad[6];
};
static_assert(sizeof(RowImg) % 128 == 96, "bank staggering of consecutive images: b * 96 mod 128 = 0, 96, 64, 32");

// builds the images of channel ch of NB signature rows at once (nsect = 30 for the DB operand, 60 for queries): the
// loads of all rows are in flight together and there is one reduction / two barriers for the NB rows, not per row
template <int NB>
__device__ inline void build_imgs(const double *h0, int row0, int nrows, int ch, bool is_db, int nsect, RowImg *S) {
  double ss[NB];
  int nb[NB];
#pragma unroll
  for (int b = 0; b < NB; b++) {
    ss[b] = 0.0;
    nb[b] = 0;
  }
  for (int k = threadIdx.x; k < SC_SIZE; k += blockDim.x) {
    double v[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) v[b] = row0 + b < nrows ? h0[(size_t)b * 2 * SC_SIZE + ch * SC_SIZE + k] : 0.0;
#pragma unroll
    for (int b = 0; b < NB; b++) {
      ss[b] += v[b] * v[b];
      if (v[b] != 0.0 && v[b] != 1.0) nb[b] = 1;
    }
  }
#pragma unroll
  for (int b = 0; b < NB; b++) {
    for (int o = 16; o > 0; o >>= 1) ss[b] += __shfl_down_sync(0xffffffffu, ss[b], o);
    if (threadIdx.x == 0) S[b].nonbin = 0;
    if ((threadIdx.x & 31) == 0) S[b].red[threadIdx.x >> 5] = ss[b];
  }
  __syncthreads();
  double scale[NB], nrm[NB];
#pragma unroll
  for (int b = 0; b < NB; b++) {
    if (nb[b]) atomicOr(&S[b].nonbin, 1);
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += S[b].red[w];
    nrm[b] = sqrt(s);  // processSC.m:16,19
    // h / |h| as h * (1 / |h|): one fp64 division per thread instead of four per work item; the operand is rounded to
    // 22 bits (fp16 hi + lo) right after, so the last-bit difference from the quotient is invisible
    scale[b] = (double)VAL_SCALE / nrm[b];
  }
  const __half2 zero2 = __floats2half2_rn(0.0f, 0.0f);
  const int per_row = nsect * (SC_NUM_R / 2);
  // one work item = rings 2t, 2t + 1 of sequence position u of row b
  for (int it0 = threadIdx.x; it0 < per_row; it0 += blockDim.x) {
    const int u = it0 / (SC_NUM_R / 2), t = it0 - u * (SC_NUM_R / 2);
    const int u2 = u < HALF_S ? u + HALF_S : u - HALF_S;      // sector (u + 30) mod 60
#pragma unroll
    for (int b = 0; b < NB; b++) {
      const bool valid = row0 + b < nrows;
      const double *hc = h0 + (size_t)b * 2 * SC_SIZE + ch * SC_SIZE;
      __half2 hi[NCOMP], lo[NCOMP];
      unsigned nib[NCOMP] = {0u, 0u};
      if (valid) {
        __half h_hi[NCOMP][2], h_lo[NCOMP][2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const double r1 = hc[u * SC_NUM_R + 2 * t + q], r2 = hc[u2 * SC_NUM_R + 2 * t + q];
          const double v1 = r1 * scale[b], v2 = r2 * scale[b];
          split_fp16(v1 + v2, h_hi[0][q], h_lo[0][q]);
          split_fp16(v1 - v2, h_hi[1][q], h_lo[1][q]);
          const int b1 = r1 == 1.0, b2 = r2 == 1.0;
          nib[0] |= (b1 + b2 == 2 ? 0x4u : (b1 + b2 == 1 ? 0x2u : 0x0u)) << (4 * q);   // e2m1 2.0 / 1.0 / 0
          nib[1] |= (b1 == b2 ? 0x0u : (b1 ? 0x2u : 0xAu)) << (4 * q);                 // 0 / +1.0 / -1.0
        }
#pragma unroll
        for (int comp = 0; comp < NCOMP; comp++) {
          hi[comp] = __halves2half2(h_hi[comp][0], h_hi[comp][1]);
          lo[comp] = __halves2half2(h_lo[comp][0], h_lo[comp][1]);
        }
      } else {
#pragma unroll
        for (int comp = 0; comp < NCOMP; comp++) hi[comp] = lo[comp] = zero2;
      }
#pragma unroll
      for (int comp = 0; comp < NCOMP; comp++) {
        __half2 *dst = reinterpret_cast<__half2 *>(&S[b].f16[comp][u][0]);
        dst[t] = is_db ? hi[comp] : lo[comp];
        dst[10 + t] = is_db ? lo[comp] : hi[comp];
        dst[20 + t] = hi[comp];
        if (t < 2) dst[30 + t] = zero2;                          // slots 60..63
        S[b].f4[comp][u][t] = (unsigned char)nib[comp];
        if (t < 6) S[b].f4[comp][u][10 + t] = 0;                 // rings 20..31: padding
      }
    }
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int b = 0; b < NB; b++) S[b].inv_norm = row0 + b < nrows ? (float)(1.0 / nrm[b]) : 0.0f;
  }
  __syncthreads();
}

// one CTA per (DB row, channel)
__global__ void __launch_bounds__(128)
sc_tc_prep_db_kernel(const double *__restrict__ hist, int n, int n_pad, unsigned char *__restrict__ buf,
                     size_t off_f16, size_t off_f4, size_t off_norm, int row0) {
  __shared__ RowImg S;
  const int row = row0 + blockIdx.x, ch = blockIdx.y;
  build_imgs<1>(hist + (size_t)row * 2 * SC_SIZE, row, n, ch, true, HALF_S, &S);
  if (threadIdx.x == 0) {
    if (S.nonbin) atomicOr(reinterpret_cast<int *>(buf) + ch, 1);
    reinterpret_cast<float *>(buf + off_norm)[(size_t)ch * n_pad + row] = S.inv_norm;
  }
  // e2m1 row: units 0..29 E sectors, 30..59 O sectors, 60..63 zero padding of the 1024-byte row
  uint4 *o4 = reinterpret_cast<uint4 *>(buf + off_f4 + ((size_t)ch * n_pad + row) * F4_ROW_BYTES);
  for (int unit = threadIdx.x; unit < F4_ROW_BYTES / 16; unit += blockDim.x) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (unit < 2 * HALF_S) v = *reinterpret_cast<const uint4 *>(&S.f4[unit >= HALF_S][unit >= HALF_S ? unit - HALF_S : unit][0]);
    o4[unit] = v;
  }
  // fp16 row: unit = (comp*8 + chunk)*30 + sector holds slots chunk*8 .. chunk*8 + 7 of that sector
  uint4 *o = reinterpret_cast<uint4 *>(buf + off_f16 + ((size_t)ch * n_pad + row) * K_F16 * 2);
  for (int unit = threadIdx.x; unit < K_F16 / 8; unit += blockDim.x) {
    const int blk = unit / HALF_S, c = unit - blk * HALF_S;
    o[unit] = *reinterpret_cast<const uint4 *>(&S.f16[blk >> 3][c][(blk & 7) * 8]);
  }
}

// Query operand: per (channel, base, query group of 4): [comp][chunk][unit 4u+b][16 B], unit u = sector u (0..59) of
// the e / o sequence of the base vector of query b of the group (x: the query image, y: its sector reversal
// y[c] = x[(60-c)%60]; e and o of y are the reversals of e and o of x).  One CTA prepares one channel of a query GROUP,
// so that the interleaved units are written as one contiguous, fully coalesced stream of 16-byte stores.
__global__ void __launch_bounds__(256)
sc_tc_prep_query_kernel(const double *__restrict__ hist, int m, int m_pad, unsigned char *__restrict__ buf,
                        size_t off_f16, size_t off_f4, size_t off_norm, int group0) {
  extern __shared__ __align__(16) unsigned char prep_smem[];
  RowImg *S = reinterpret_cast<RowImg *>(prep_smem);
  const int group = group0 + blockIdx.x, ngroups = m_pad / QG, ch = blockIdx.y;
  build_imgs<QG>(hist + (size_t)QG * group * 2 * SC_SIZE, QG * group, m, ch, false, SC_NUM_S, S);
  if (threadIdx.x < QG) {
    const int b = threadIdx.x, row = QG * group + b;
    if (S[b].nonbin) atomicOr(reinterpret_cast<int *>(buf) + ch, 1);
    reinterpret_cast<float *>(buf + off_norm)[(size_t)ch * m_pad + row] = S[b].inv_norm;
  }
  for (int base = 0; base < 2; base++) {
    const size_t gb = ((size_t)ch * 2 + base) * ngroups + group;
    uint4 *o = reinterpret_cast<uint4 *>(buf + off_f16 + gb * Q_F16_BYTES);
    for (int e = threadIdx.x; e < NCOMP * F16_CHUNKS * Q_UNITS; e += blockDim.x) {
      const int blk = e / Q_UNITS, w = e - blk * Q_UNITS;
      const int b = w & 3, u = w >> 2;
      const int c = base == 0 ? u : (u == 0 ? 0 : SC_NUM_S - u);
      o[e] = *reinterpret_cast<const uint4 *>(&S[b].f16[blk >> 3][c][(blk & 7) * 8]);
    }
    uint4 *o4 = reinterpret_cast<uint4 *>(buf + off_f4 + gb * Q_F4_BYTES);
    for (int e = threadIdx.x; e < NCOMP * Q_UNITS; e += blockDim.x) {
      const int comp = e / Q_UNITS, w = e - comp * Q_UNITS;
      const int b = w & 3, u = w >> 2;
      const int c = base == 0 ? u : (u == 0 ? 0 : SC_NUM_S - u);
      o4[e] = *reinterpret_cast<const uint4 *>(&S[b].f4[comp][c][0]);
    }
  }
}

struct TcParams {
  const unsigned char *q_buf;   // QLayout
  const unsigned char *db_buf;  // DbLayout
  size_t q_off_f16, q_off_f4, q_off_norm, db_off_norm;
  float *d_out[2];              // per channel, m x ldd
  int m, n, m_pad, n_pad, ldd;
  // Work of one launch: up to two rectangles of (query group, DB tile) items (the streamed path matches an L-shaped
  // region per chunk: new queries x all DB rows so far + old queries x new DB rows).  Per rectangle: units = 2 channels x
  // query groups of 4, tiles of 256 DB rows, first query group / first tile.  w0 = number of items of rectangle 0.
  int n_units[2], n_tiles[2], qg0[2], tile0[2];
  long long w0, w_total;
  int flags;                    // debug: 1 skip epilogue loads, 2 skip MMAs, 4 force generic mode
  // Self-match (queries == DB rows, m == n): d is symmetric, so only the (query group, DB tile) items with
  // tile_start <= group_end are computed (rectangle 0 = queries [q0, q1), q0 a multiple of 256, tiles 0 .. diagonal) and
  // every value whose transposed position belongs to an item that is not computed is stored there as well.
  int tri, tri_nt0;             // tri_nt0: tiles of the first query block (q0 / 256 + 1)
};
constexpr int TRI_UNITS = 2 * TILE_M / QG;   // 128 units (2 channels x 64 query groups) share a tile count

// items of a triangular region of `units` units (2 channels x query groups) whose first query block has nt0 tiles
__host__ __device__ __forceinline__ long long tri_prefix(int lb, int nt0);
__host__ __device__ __forceinline__ long long tri_total_items(int units, int nt0) {
  const int lb_full = units / TRI_UNITS;
  return tri_prefix(lb_full, nt0) + (long long)(units - lb_full * TRI_UNITS) * (nt0 + lb_full);
}

// items before query block lb of a triangular region
__host__ __device__ __forceinline__ long long tri_prefix(int lb, int nt0) {
  return (long long)TRI_UNITS * ((long long)lb * nt0 + (long long)lb * (lb - 1) / 2);
}

// One work item's K loop on the issuing thread, specialised per operand format so that the loop body is
// straight-line: descriptors are advanced by integer adds on their 16-byte address field.
//   MODE 0: fp16 3-term split (kind::f16), 2: e2m1 (kind::mxf4, block scales = 1)
// K runs over [E: chunks x 30 sectors][O: chunks x 30 sectors]; an MMA covers 2 sectors (2 units of the DB row, 8 units of
// the 4-way interleaved Hankel buffer); after the 15 MMAs of a 30-sector run the Hankel offset skips the 120 units the
// windows of that block extend over, and after the last E run the destination switches to the O accumulators.
template <int MODE>
__device__ __forceinline__ void issue_k_loop(TcBarriers *bars, uint32_t sA, uint32_t sB, uint32_t tmem_base,
                                             uint32_t tmem_sf, int &stage, uint32_t &phase, bool skip) {
  constexpr int NUM_KB = MODE == 2 ? F4_KB : F16_KB;
  constexpr int RUNS_E = MODE == 2 ? 1 : F16_CHUNKS;            // 30-sector runs of the E contraction
  constexpr uint32_t idesc = MODE == 2 ? make_idesc_mxf4(TILE_M, N_INST) : make_idesc(TILE_M, N_INST);
  // A: SWIZZLE_128B K-major, 8-row groups 1024 B apart; a K-step advances the start by 32 B (2 units)
  const uint64_t adesc0 = make_desc(sA, 16, 1024, 2);
  // B: Hankel view, no swizzle: row r = 4 s + b, k-group g -> unit 4 (c + s + g) + b of block (comp, chunk)
  const uint64_t bdesc0 = make_desc(sB, 64, 128, 0);
  uint32_t boff = 0;    // block * 240 + 4 * (sector c): the K position of the next MMA in the Hankel buffer
  uint32_t c4 = 0;      // 4 * c
  int run = 0;
  uint32_t acc = 0, dst = tmem_base;
#pragma unroll 1
  for (int kb = 0; kb < NUM_KB; kb++) {
    mbar_wait(smem_u32(&bars->full[stage]), phase, 7);
    tc_fence_after();
    const uint64_t ad = adesc0 + (uint64_t)(stage * (A_STAGE_BYTES >> 4));
    if (!skip) {
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        if (MODE == 2 && run >= 2 * RUNS_E) break;   // e2m1: units 60..63 of the DB row are zero padding
        const uint64_t a = ad + (uint64_t)(kk * 2), b = bdesc0 + (uint64_t)boff;
        if (elect_one()) {
          if (MODE == 0)
            umma_f16_2sm(dst, a, b, idesc, acc);
          else
            umma_mxf4_2sm(dst, a, b, idesc, acc, tmem_sf, tmem_sf + 8);
        }
        acc = 1;
        boff += 2 * QG;
        c4 += 2 * QG;
        if (c4 >= (uint32_t)(QG * HALF_S)) {
          c4 = 0;
          boff += Q_UNITS - QG * HALF_S;
          if (++run == RUNS_E) {
            dst = tmem_base + N_MMA;
            acc = 0;
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

// Work items of a CTA pair in launch order: (unit id unique within the launch, channel, query group, DB tile).
// One 64-bit division at the start (and at the rectangle switch), then counters.
struct ItemIter {
  int r, unit, tile_l, nt;
  __device__ __forceinline__ void seek(const TcParams &P, long long it) {
    if (P.tri) {
      r = 0;
      int lb = 0;
      while (tri_prefix(lb + 1, P.tri_nt0) <= it) ++lb;
      const long long rem = it - tri_prefix(lb, P.tri_nt0);
      nt = P.tri_nt0 + lb;
      const int u = (int)(rem / nt);
      unit = lb * TRI_UNITS + u;
      tile_l = (int)(rem - (long long)u * nt);
      return;
    }
    r = it >= P.w0;
    const long long l = it - (r ? P.w0 : 0);
    nt = P.n_tiles[r];
    unit = (int)(l / nt);
    tile_l = (int)(l - (long long)unit * nt);
  }
  __device__ __forceinline__ void get(const TcParams &P, int &unit_id, int &ch, int &qg, int &tile) const {
    tile = P.tile0[r] + tile_l;
    ch = unit & 1;
    qg = P.qg0[r] + (unit >> 1);
    unit_id = unit | (r << 28);
  }
  // advance to item it_next = current + 1
  __device__ __forceinline__ void next(const TcParams &P, long long it_next) {
    if (!P.tri && r == 0 && it_next == P.w0) {
      seek(P, it_next);
    } else if (++tile_l == nt) {
      tile_l = 0;
      ++unit;
      if (P.tri && (unit & (TRI_UNITS - 1)) == 0) ++nt;
    }
  }
  // is the current item the last one of its unit?  (more: item it + 1 exists)
  __device__ __forceinline__ bool last_of_unit(const TcParams &P, long long it, bool more) const {
    return !more || tile_l + 1 == nt || (!P.tri && r == 0 && it + 1 == P.w0);
  }
};

// Split of the item sequence over the CTA pairs by COST, not by count: an item of a generic channel (240 MMAs) takes
// about 5.5 x as long as one of a binary channel (30 MMAs + the same hand-off), and the sequence alternates runs of up
// to 20 items of either kind, so equal counts leave the slowest pair up to 0.3 ms behind.  Items of one query group g:
// [channel 0: nt(g) tiles][channel 1: nt(g) tiles]; cost of a group = nt(g) (w0 + w1).  Returns the index of the item
// at fraction num / den of the total cost.
__device__ __forceinline__ long long item_at_cost(const TcParams &P, int w0, int w1, long long num, long long den) {
  const long long tg_total = P.w_total / 2;                       // (group, tile) pairs of the launch
  if (num >= den) return P.w_total;
  const long long x = tg_total * (w0 + w1) * num / den;          // cost position (< 2^63 for any realistic launch)
  // group containing cost x: tile-pair prefix T(g) <= x / (w0 + w1)
  const long long tq = x / (w0 + w1);
  long long t_before;   // T(g)
  long long item_base;  // index of the first item of group g
  int nt;
  if (P.tri) {
    int lb = 0;                                   // query block (64 groups share a tile count)
    while (tri_prefix(lb + 1, P.tri_nt0) / 2 <= tq) ++lb;
    nt = P.tri_nt0 + lb;
    const long long k = (tq - tri_prefix(lb, P.tri_nt0) / 2) / nt;
    t_before = tri_prefix(lb, P.tri_nt0) / 2 + k * nt;
    item_base = 2 * t_before;
  } else {
    const long long tg0 = P.w0 / 2;
    const int r = tq >= tg0;
    nt = P.n_tiles[r];
    const long long k = (tq - (r ? tg0 : 0)) / nt;
    t_before = (r ? tg0 : 0) + k * nt;
    item_base = 2 * t_before;
  }
  const long long rem = x - t_before * (w0 + w1);                 // cost inside the group, < nt (w0 + w1)
  if (rem < (long long)nt * w0) return item_base + rem / w0;
  return item_base + nt + (rem - (long long)nt * w0) / w1;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
sc_match_tc_kernel(const __grid_constant__ CUtensorMap map_f16_0, const __grid_constant__ CUtensorMap map_f16_1,
                   const __grid_constant__ CUtensorMap map_f4_0, const __grid_constant__ CUtensorMap map_f4_1,
                   const TcParams P) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_raw = smem_u32(smem_dyn);
  const uint32_t base = (base_raw + 1023u) & ~1023u;
  unsigned char *smem = smem_dyn + (base - base_raw);
  const uint32_t sA = base;                               // NSTAGE x 16 KB (1024-aligned)
  const uint32_t sB = base + NSTAGE * A_STAGE_BYTES;      // 60 KB
  TcBarriers *bars = reinterpret_cast<TcBarriers *>(smem + NSTAGE * A_STAGE_BYTES + B_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  // binary mode per channel: both operands hold only 0/1 in that channel
  const int *qf = reinterpret_cast<const int *>(P.q_buf), *df = reinterpret_cast<const int *>(P.db_buf);
  bool binary[2];
  for (int ch = 0; ch < 2; ch++) binary[ch] = !(P.flags & 4) && qf[ch] == 0 && df[ch] == 0;

  // work items of this CTA pair: contiguous range in unit-major order, equal shares of the estimated cost
  constexpr int W_GENERIC = 11, W_BINARY = 2;   // measured: 30.8 k vs 5.6 k clocks per item incl. the hand-off
  const int wc0 = binary[0] ? W_BINARY : W_GENERIC, wc1 = binary[1] ? W_BINARY : W_GENERIC;
  const long long it_begin = item_at_cost(P, wc0, wc1, pair_id, npairs), it_end = item_at_cost(P, wc0, wc1, pair_id + 1, npairs);

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
  // binary channels run as e2m1 under kind::mxf4; its block scales (ue8m0, all 1.0 = 0x7f) live in the 16 TMEM
  // columns that the N = 240 accumulators leave free in each 256-column slot
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
      ItemIter iter;
      iter.seek(P, it_begin);
      for (long long it = it_begin; it < it_end; ++it, iter.next(P, it)) {
        int unit, ch, qg, tile;
        iter.get(P, unit, ch, qg, tile);
        const bool bin = binary[ch];
        const CUtensorMap *map = bin ? (ch == 0 ? &map_f4_0 : &map_f4_1) : (ch == 0 ? &map_f16_0 : &map_f16_1);
        const int num_kb = bin ? F4_KB : F16_KB;
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
    // ===== query operand loader: the 4 interleaved queries of this CTA's base (x: rank 0, y: rank 1) =====
    if (lane == 0) {
      uint32_t phase = 0;
      int prev_unit = -1;
      const int qgroups = P.m_pad / QG;
      ItemIter iter;
      iter.seek(P, it_begin);
      for (long long it = it_begin; it < it_end; ++it, iter.next(P, it)) {
        int unit, ch, qg, tile;
        iter.get(P, unit, ch, qg, tile);
        if (unit == prev_unit) continue;
        prev_unit = unit;
        const bool bin = binary[ch];
        const uint32_t half_bytes = (bin ? Q_F4_BYTES : Q_F16_BYTES) / 2;
        mbar_wait(smem_u32(&bars->b_empty), phase ^ 1, 2);
        mbar_expect_tx(smem_u32(&bars->b_full), 2 * half_bytes);
        const unsigned char *src = P.q_buf + (bin ? P.q_off_f4 : P.q_off_f16) +
                                   (((size_t)ch * 2 + rank) * qgroups + (size_t)qg) * (2 * half_bytes);
        for (int p = 0; p < 2; p++)
          bulk_load_1d(sB + p * half_bytes, src + (size_t)p * half_bytes, half_bytes, smem_u32(&bars->b_full));
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
      ItemIter iter;
      iter.seek(P, it_begin);
      for (long long it = it_begin; it < it_end; ++it, iter.next(P, it)) {
        int unit, ch, qg, tile;
        iter.get(P, unit, ch, qg, tile);
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
        else
          issue_k_loop<2>(bars, sA, sB, tmem_base, tmem_sf, stage, phase, skip);
        const bool last_of_unit = iter.last_of_unit(P, it, it + 1 < it_end);
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
    ItemIter iter;
    iter.seek(P, it_begin);
    for (long long it = it_begin; it < it_end; ++it, iter.next(P, it)) {
      int unit, ch, qg, tile;
      iter.get(P, unit, ch, qg, tile);
      const bool bin = binary[ch];
      mbar_wait(smem_u32(&bars->tmem_full), t_phase, 8);
      tc_fence_after();
      const int row = tile * TILE_M + (int)rank * CTA_M + ew * 32 + lane;
      float *out = P.d_out[ch];
      const float rd = bin ? db_norm[(size_t)ch * P.n_pad + row] : 0.0f;
      if (!(P.flags & 1)) {
        // column 120 h + 4 s + b of E (slot 0) and of O (slot 1): base h, shift s, query b of the group;
        // max over (h, s) of E + |O| = 2 max over the 120 variants.  The loads of block i + 1 are in flight while
        // block i is reduced.
        float best[QG];
#pragma unroll
        for (int b = 0; b < QG; b++) best[b] = __int_as_float(0x7fc00000);  // NaN: min ignores NaN (processSC.m:31)
        const uint32_t tcol = tmem_base + ((uint32_t)(ew * 32) << 16);
        static_assert(N_INST == 7 * 32 + 16, "7 blocks of 32 columns + a tail of 16");
        uint32_t e[2][32], o[2][32];
        tmem_ld32(tcol, e[0]);
        tmem_ld32(tcol + (uint32_t)N_MMA, o[0]);
        tmem_ld_wait();
#pragma unroll
        for (int blk = 0; blk < 7; blk++) {
          const int cur = blk & 1, nxt = cur ^ 1;
          if (blk + 1 < 7) {
            tmem_ld32(tcol + (uint32_t)((blk + 1) * 32), e[nxt]);
            tmem_ld32(tcol + (uint32_t)(N_MMA + (blk + 1) * 32), o[nxt]);
          } else {
            tmem_ld16(tcol + (uint32_t)(N_INST - 16), reinterpret_cast<uint32_t(&)[16]>(e[nxt]));
            tmem_ld16(tcol + (uint32_t)(N_MMA + N_INST - 16), reinterpret_cast<uint32_t(&)[16]>(o[nxt]));
          }
#pragma unroll
          for (int i = 0; i < 32; i++)
            best[i & 3] = fmaxf(best[i & 3], __uint_as_float(e[cur][i]) + fabsf(__uint_as_float(o[cur][i])));
          tmem_ld_wait();
        }
        // every accumulator of this item is in registers: hand TMEM back before the last reduction and the stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(leader_tmem_empty);
#pragma unroll
        for (int i = 0; i < 16; i++)
          best[i & 3] = fmaxf(best[i & 3], __uint_as_float(e[1][i]) + fabsf(__uint_as_float(o[1][i])));
        if (row < P.n && out) {
          float d[QG];
#pragma unroll
          for (int b = 0; b < QG; b++) {
            const int qi = qg * QG + b;
            d[b] = bin ? (1.0f - 0.5f * best[b] * (qi < P.m ? q_norm[(size_t)ch * P.m_pad + qi] : 0.0f) * rd) * 0.5f
                       : (1.0f - best[b] * ACC_SCALE) * 0.5f;
            if (qi < P.m) out[(size_t)qi * P.ldd + row] = d[b];
          }
          // self-match: d(row, qi) = d(qi, row) when the item (query group of `row`, DB tile of qi) is not computed,
          // i.e. when the tile of the queries starts after the group of `row` ends
          const int q0 = qg * QG;
          if (P.tri && (q0 / TILE_M) * TILE_M > (row | (QG - 1))) {
            float *mo = out + (size_t)row * P.ldd + q0;
            if (q0 + QG <= P.m && (reinterpret_cast<uintptr_t>(mo) & 15) == 0) {
              *reinterpret_cast<float4 *>(mo) = make_float4(d[0], d[1], d[2], d[3]);
            } else {
#pragma unroll
              for (int b = 0; b < QG; b++)
                if (q0 + b < P.m) mo[b] = d[b];
            }
          }
        }
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(leader_tmem_empty);
      }
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

// debug flags (sodso_debug_set_kernel_flags): 1 skip epilogue loads, 2 skip MMAs, 4 force generic mode
static int tc_flags() { return g_debug.tc_flags; }

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
// n_layout > 0: the operand buffer is laid out for n_layout >= n rows (a database with spare capacity, sodso_db_append)
cudaError_t launch_sc_tc_prep_db_rows(const double *hist, int n, int row0, int row1, void *db_buf, cudaStream_t st,
                                      int64_t *launches, int n_layout) {
  if (row1 <= row0) return cudaSuccess;
  DbLayout L(n_layout > 0 ? n_layout : n);
  sc_tc_prep_db_kernel<<<dim3(row1 - row0, 2), 128, 0, st>>>(hist, n, L.n_pad, reinterpret_cast<unsigned char *>(db_buf),
                                                             L.off_f16, L.off_f4, L.off_norm, row0);
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_sc_tc_prep_query_rows(const double *hist, int m, int row0, int row1, void *q_buf, cudaStream_t st,
                                         int64_t *launches) {
  if (row1 <= row0) return cudaSuccess;
  QLayout L(m);
  if (row0 % QG) return cudaErrorInvalidValue;   // one CTA per query group
  const int smem = QG * (int)sizeof(RowImg);
  cudaError_t e = cudaFuncSetAttribute(sc_tc_prep_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  sc_tc_prep_query_kernel<<<dim3((row1 - row0 + QG - 1) / QG, 2), 256, smem, st>>>(
      hist, m, L.m_pad, reinterpret_cast<unsigned char *>(q_buf), L.off_f16, L.off_f4, L.off_norm, row0 / QG);
  if (launches) ++*launches;
  return cudaGetLastError();
}

int sc_tc_db_rows_padded(int n) { return DbLayout(n).n_pad; }

// moves the rows of a DB operand laid out for old_cap rows into a (zeroed) buffer laid out for new_cap >= old_cap rows
cudaError_t sc_tc_db_relayout(const void *old_buf, int old_cap, void *new_buf, int new_cap, cudaStream_t st) {
  const DbLayout A(old_cap), B(new_cap);
  const unsigned char *a = reinterpret_cast<const unsigned char *>(old_buf);
  unsigned char *b = reinterpret_cast<unsigned char *>(new_buf);
  cudaError_t e = cudaMemcpyAsync(b, a, HEADER_BYTES, cudaMemcpyDeviceToDevice, st);
  const size_t row_bytes[3] = {(size_t)K_F16 * 2, (size_t)F4_ROW_BYTES, sizeof(float)};
  const size_t offa[3] = {A.off_f16, A.off_f4, A.off_norm}, offb[3] = {B.off_f16, B.off_f4, B.off_norm};
  for (int part = 0; part < 3 && e == cudaSuccess; part++)
    for (int ch = 0; ch < 2 && e == cudaSuccess; ch++)
      e = cudaMemcpyAsync(b + offb[part] + (size_t)ch * B.n_pad * row_bytes[part],
                          a + offa[part] + (size_t)ch * A.n_pad * row_bytes[part], (size_t)A.n_pad * row_bytes[part],
                          cudaMemcpyDeviceToDevice, st);
  return e;
}
int sc_tc_query_rows_padded(int m) { return QLayout(m).m_pad; }

cudaError_t launch_sc_tc_prep_query(const double *hist, int m, void *q_buf, cudaStream_t st, int64_t *launches) {
  if (m <= 0) return cudaSuccess;
  cudaError_t e = launch_sc_tc_clear_flags(q_buf, st);
  if (e != cudaSuccess) return e;
  return launch_sc_tc_prep_query_rows(hist, m, 0, QLayout(m).m_pad, q_buf, st, launches);
}

cudaError_t launch_sc_match_tc(const void *q_buf, int m, const void *db_buf, int n, float *d_p, float *d_i, int ldd,
                               int num_sms, cudaStream_t st, int64_t *launches, int n_layout) {
  return launch_sc_match_tc_blocks(q_buf, m, db_buf, n, 0, m, 0, n, 0, 0, 0, 0, d_p, d_i, ldd, num_sms, st, launches, n_layout);
}

// queries [q0, q1) x DB rows [r0, r1) of the m x n problem; q0 must be a multiple of 4 and r0 of 256
cudaError_t launch_sc_match_tc_block(const void *q_buf, int m, int q0, int q1, const void *db_buf, int n, int r0, int r1,
                                     float *d_p, float *d_i, int ldd, int num_sms, cudaStream_t st,
                                     int64_t *launches) {
  return launch_sc_match_tc_blocks(q_buf, m, db_buf, n, q0, q1, r0, r1, 0, 0, 0, 0, d_p, d_i, ldd, num_sms, st, launches);
}

// number of (query group, DB tile, channel) work items launch_sc_match_tc_self runs for queries [q0, q1); -1: bad arguments
long long sc_tc_self_items(int n, int q0, int q1) {
  if (n <= 0 || q0 < 0 || q1 > n || q1 <= q0 || q0 % TILE_M) return -1;
  return tri_total_items(2 * ((q1 - q0 + QG - 1) / QG), q0 / TILE_M + 1);
}

// self-match of an n x n problem, queries [q0, q1) (q0 a multiple of 256) against the DB rows up to their own block:
// the part of the lower block triangle that belongs to these queries; the transposed values are stored as well
cudaError_t launch_sc_match_tc_self(const void *q_buf, const void *db_buf, int n, int q0, int q1, float *d_p, float *d_i,
                                    int ldd, int num_sms, cudaStream_t st, int64_t *launches) {
  if (q0 % TILE_M) return cudaErrorInvalidValue;
  return launch_sc_match_tc_blocks(q_buf, n, db_buf, n, q0, q1, 0, n, -1, 0, 0, 0, d_p, d_i, ldd, num_sms, st, launches);
}

// two rectangles [qa0, qa1) x [ra0, ra1) and [qb0, qb1) x [rb0, rb1) in one launch (either may be empty);
// qb0 < 0: rectangle A is the triangular region of a self-match (launch_sc_match_tc_self)
cudaError_t launch_sc_match_tc_blocks(const void *q_buf, int m, const void *db_buf, int n, int qa0, int qa1, int ra0, int ra1,
                                      int qb0, int qb1, int rb0, int rb1, float *d_p, float *d_i, int ldd, int num_sms,
                                      cudaStream_t st, int64_t *launches, int n_layout) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  const bool tri = qb0 < 0;
  if (tri) qb0 = 0;
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
  DbLayout DL(n_layout > 0 ? n_layout : n);   // rows >= n are never stored (P.n), whatever the padding holds
  QLayout QL(m);
  const unsigned char *dbb = reinterpret_cast<const unsigned char *>(db_buf);
  CUtensorMap maps[4];
  for (int fmt = 0; fmt < 2; fmt++)
    for (int ch = 0; ch < 2; ch++) {
      const size_t kbytes = fmt == 0 ? (size_t)K_F16 * 2 : (size_t)F4_ROW_BYTES;
      cuuint64_t dims[2] = {(cuuint64_t)(fmt == 0 ? K_F16 : F4_ROW_BYTES), (cuuint64_t)DL.n_pad};
      cuuint64_t strides[1] = {(cuuint64_t)kbytes};
      cuuint32_t box[2] = {(cuuint32_t)(fmt == 0 ? 64 : 128), (cuuint32_t)CTA_M};
      cuuint32_t estr[2] = {1, 1};
      void *gaddr = (void *)(dbb + (fmt == 0 ? DL.off_f16 : DL.off_f4) + (size_t)ch * DL.n_pad * kbytes);
      CUresult r = enc(&maps[fmt * 2 + ch], fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                       gaddr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
  TcParams P;
  P.q_buf = reinterpret_cast<const unsigned char *>(q_buf);
  P.db_buf = dbb;
  P.q_off_f16 = QL.off_f16;
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
  P.tri = tri ? 1 : 0;
  P.tri_nt0 = qa0 / TILE_M + 1;
  if (tri) {
    P.w0 = P.w_total = tri_total_items(P.n_units[0], P.tri_nt0);
    P.tile0[0] = 0;
  }
  P.flags = tc_flags();
  const long long W = P.w_total;
  int npairs = num_sms / 2;
  if (npairs > W) npairs = (int)W;
  if (npairs < 1) npairs = 1;
  cudaError_t e = cudaFuncSetAttribute(sc_match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  sc_match_tc_kernel<<<2 * npairs, TC_THREADS, SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], P);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
