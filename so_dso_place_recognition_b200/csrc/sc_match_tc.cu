// Scan Context all-pairs x all-shifts match (processSC.m:12-45) on the 5th-generation tensor cores.
//
// For one channel, d(i, j) = min over the 120 variants v of (1 - <variant_v(q_i), h_j>) / 2
// (processSC.m:24-31).  The 120 variants are the 60 circular sector shifts of the query image x
// and the 60 circular sector shifts of its sector-reversed image y (reverse shift k of x ==
// forward shift (61-k) mod 60 of y).  So per (query, DB row) we need the 2 x 60 correlations
//      corr_b[s] = sum_{c, r} b[(c + s) mod 60][r] * h[c][r],        b in {x, y}
// i.e. a dense contraction  D[j, (b, s)] = sum_k A[j, k] * B[(b, s), k]  with
//      A = DB signatures           (M side: 256 DB rows per CTA pair, TMA-fed, SWIZZLE_128B)
//      B = Hankel matrix of shifts (N side: 64 shift rows per base vector, never materialised)
//
// The Hankel operand.  B[(b,s)][c, r] = b[(c+s) mod 60][r] is a *view* of the doubled vector
// [b, b]: in the canonical K-major no-swizzle UMMA layout ((8,n),2):((16 B, SBO), LBO) rows
// inside an 8-row core matrix are 16 B apart; choosing SBO = 128 B and LBO = 16 B makes
// address(row s, k-group g) = base + 16 (s + g): row s reads the 16-byte unit s + g.  With a
// K ordering in which one 16-byte unit = 8 slots of ONE sector, the shift by one sector is a
// shift by one unit, so all 64 shift rows of a K-step are overlapping windows of one 2 KB
// buffer.  The x rows live in CTA 0 of the pair and the y rows in CTA 1 (cta_group::2,
// M = 256, N = 128: each CTA contributes N/2 = 64 rows of B), so one query costs 16 KB of
// shared memory per CTA and channel instead of a 120 x 1200 expanded tile per K-block.
//
// Precision (target |d - d_ref| <= 1e-5 against the fp64 reference).  Values are row-normalised
// in fp64 (processSC.m:15-20), scaled by 64 and split v = hi + lo into two fp16; the product is
// evaluated as hi*lo + lo*hi + hi*hi (lo*lo ~ 2^-22 dropped) with fp32 accumulation in TMEM.
// The three terms are interleaved per sector into 64 fp16 "slots" (20 + 20 + 20 + 4 pad), the
// cross terms first so that the large hi*hi partial sums come last (accumulation error).
//
// Roles per CTA (256 threads): warp 0 TMA producer (DB tiles), warp 1 MMA issuer (leader CTA),
// warp 2 TMEM allocator, warp 3 query-operand loader, warps 4-7 epilogue (tcgen05.ld -> max over
// the 128 shift columns -> (1 - x)/2 -> coalesced fp32 stores).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "../../include/sodso_pr.h"
#include "common.cuh"

namespace sodso {
namespace {

constexpr int KS_CHUNKS = 8;                         // 16-byte units per sector
constexpr int KS_SLOTS = KS_CHUNKS * 8;              // 64 fp16 slots per sector
constexpr int K_TOTAL = KS_SLOTS * SC_NUM_S;         // 3840
constexpr int K_BLOCK = 64;                          // fp16 elements per TMA box row (128 B)
constexpr int NUM_KB = K_TOTAL / K_BLOCK;            // 60
constexpr int CHUNK_K = 8 * SC_NUM_S;                // 480 K elements per chunk
constexpr int Q_UNITS = 128;                         // 16-byte units per (query, base, chunk): doubled vector
constexpr int Q_BASE_BYTES = KS_CHUNKS * Q_UNITS * 16;  // 16 KB per (query, base, channel)
constexpr int QG = 4;                                // queries per tile (4 x 128 TMEM columns)
constexpr int TILE_M = 256, CTA_M = 128;             // DB rows per CTA pair / per CTA
constexpr int N_PER_Q = 128;                         // accumulator columns per query: 64 x-shifts + 64 y-shifts
constexpr int A_STAGE_BYTES = CTA_M * K_BLOCK * 2;   // 16 KB
constexpr int NSTAGE = 7;
constexpr int B_BYTES = QG * Q_BASE_BYTES;           // 64 KB
constexpr float VAL_SCALE = 64.0f;                   // operand scale
constexpr float ACC_SCALE = 1.0f / (VAL_SCALE * VAL_SCALE);
constexpr int TC_THREADS = 256;

struct __align__(8) TcBarriers {
  uint64_t full[NSTAGE], empty[NSTAGE];
  uint64_t b_full, b_peer, b_empty, tmem_full, tmem_empty;
  uint32_t tmem_ptr;
  uint32_t pad;
};
constexpr int SMEM_BYTES = 1024 /*align*/ + NSTAGE * A_STAGE_BYTES + B_BYTES + (int)sizeof(TcBarriers);

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_fp16(double v, __half &hi, __half &lo) {
  hi = __float2half_rn((float)v);
  lo = __float2half_rn((float)(v - (double)__half2float(hi)));
}

__device__ __forceinline__ void row_norms(const double *h, double nrm[2], double *red) {
  double ss[2] = {0.0, 0.0};
  for (int ch = 0; ch < 2; ch++)
    for (int k = threadIdx.x; k < SC_SIZE; k += blockDim.x) {
      double v = h[ch * SC_SIZE + k];
      ss[ch] += v * v;
    }
  for (int ch = 0; ch < 2; ch++) {
    double s = ss[ch];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[ch * 32 + (threadIdx.x >> 5)] = s;
  }
  __syncthreads();
  for (int ch = 0; ch < 2; ch++) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[ch * 32 + w];
    nrm[ch] = sqrt(s);  // processSC.m:16,19
  }
}

// A operand: [ch][n_pad][3840] fp16.  k = chunk*480 + sector*8 + t, slot = chunk*8 + t:
//   slots  0..19: hi[r]   (pairs with the query's lo)
//   slots 20..39: lo[r]   (pairs with the query's hi)
//   slots 40..59: hi[r]   (pairs with the query's hi)
//   slots 60..63: 0
__global__ void __launch_bounds__(256)
sc_tc_prep_db_kernel(const double *__restrict__ hist, int n, int n_pad, __half *__restrict__ out) {
  __shared__ double red[64];
  __shared__ __half s_hi[2][SC_SIZE], s_lo[2][SC_SIZE];
  const int row = blockIdx.x;
  if (row < n) {
    const double *h = hist + (size_t)row * 2 * SC_SIZE;
    double nrm[2];
    row_norms(h, nrm, red);
    for (int ch = 0; ch < 2; ch++)
      for (int k = threadIdx.x; k < SC_SIZE; k += blockDim.x)
        split_fp16(h[ch * SC_SIZE + k] / nrm[ch] * (double)VAL_SCALE, s_hi[ch][k], s_lo[ch][k]);
  } else {
    for (int ch = 0; ch < 2; ch++)
      for (int k = threadIdx.x; k < SC_SIZE; k += blockDim.x) s_hi[ch][k] = s_lo[ch][k] = __float2half(0.0f);
  }
  __syncthreads();
  for (int ch = 0; ch < 2; ch++) {
    __half *o = out + ((size_t)ch * n_pad + row) * K_TOTAL;
    for (int k = threadIdx.x; k < K_TOTAL; k += blockDim.x) {
      const int j = k / CHUNK_K, rem = k - j * CHUNK_K, c = rem >> 3, t = rem & 7, s = j * 8 + t;
      __half v = __float2half(0.0f);
      if (s < 20) v = s_hi[ch][c * SC_NUM_R + s];
      else if (s < 40) v = s_lo[ch][c * SC_NUM_R + s - 20];
      else if (s < 60) v = s_hi[ch][c * SC_NUM_R + s - 40];
      o[k] = v;
    }
  }
}

// B operand: [ch][base][m_pad][chunk 8][unit 128][8] fp16; unit u holds sector u % 60 of the base
// vector (x: the query image, y: its sector reversal y[c] = x[(60 - c) % 60]).
//   slots  0..19: lo[r],  slots 20..39: hi[r],  slots 40..59: hi[r],  slots 60..63: 0
__global__ void __launch_bounds__(256)
sc_tc_prep_query_kernel(const double *__restrict__ hist, int m, int m_pad, __half *__restrict__ out) {
  __shared__ double red[64];
  __shared__ __half s_hi[2][SC_SIZE], s_lo[2][SC_SIZE];
  const int row = blockIdx.x;
  if (row < m) {
    const double *h = hist + (size_t)row * 2 * SC_SIZE;
    double nrm[2];
    row_norms(h, nrm, red);
    for (int ch = 0; ch < 2; ch++)
      for (int k = threadIdx.x; k < SC_SIZE; k += blockDim.x)
        split_fp16(h[ch * SC_SIZE + k] / nrm[ch] * (double)VAL_SCALE, s_hi[ch][k], s_lo[ch][k]);
  } else {
    for (int ch = 0; ch < 2; ch++)
      for (int k = threadIdx.x; k < SC_SIZE; k += blockDim.x) s_hi[ch][k] = s_lo[ch][k] = __float2half(0.0f);
  }
  __syncthreads();
  constexpr int PER_BASE = Q_BASE_BYTES / 2;  // halves
  for (int ch = 0; ch < 2; ch++)
    for (int b = 0; b < 2; b++) {
      __half *o = out + (((size_t)ch * 2 + b) * m_pad + row) * PER_BASE;
      for (int e = threadIdx.x; e < PER_BASE; e += blockDim.x) {
        const int t = e & 7, u = (e >> 3) & (Q_UNITS - 1), j = e >> 10, s = j * 8 + t;
        const int cs = u % SC_NUM_S;
        const int c = b == 0 ? cs : (SC_NUM_S - cs) % SC_NUM_S;
        __half v = __float2half(0.0f);
        if (s < 20) v = s_lo[ch][c * SC_NUM_R + s];
        else if (s < 40) v = s_hi[ch][c * SC_NUM_R + s - 20];
        else if (s < 60) v = s_hi[ch][c * SC_NUM_R + s - 40];
        o[e] = v;
      }
    }
}

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (surfacing as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s
      printf("sc_match_tc: barrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, (int)blockIdx.x,
             (int)threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, int c0, int c1,
                                                uint32_t cluster_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(cluster_bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// completion of all prior MMAs of this thread -> arrive on the barrier at the same offset in both CTAs
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptors (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

// instruction descriptor (InstrDescriptor): c_format F32 (1) [4,6), a/b format F16 (0), K-major both,
// n_dim = N>>3 [17,23), m_dim = M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcParams {
  const __half *q_op;   // [ch][base][m_pad][16 KB]
  float *d_out[2];      // per channel, m x ldd
  int m, n, m_pad, n_pad, ldd;
  int n_units, n_tiles;  // units = 2 channels x (m_pad / 4) query groups; tiles = n_pad / 256
  int dbg_lbo, dbg_sbo;  // Hankel descriptor strides in bytes (16 / 128)
};

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
sc_match_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
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
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  // work items of this CTA pair: contiguous range in unit-major order
  const long long W = (long long)P.n_units * P.n_tiles;
  const long long it_begin = W * pair / npairs, it_end = W * (pair + 1) / npairs;

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

  if (warp == 0) {
    // ===== TMA producer: this CTA's 128 DB rows of every K-block =====
    if (lane == 0) {
      const uint32_t leader_full0 = map_to_cta(smem_u32(&bars->full[0]), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (long long it = it_begin; it < it_end; ++it) {
        const int unit = (int)(it / P.n_tiles), tile = (int)(it - (long long)unit * P.n_tiles);
        const int ch = unit & 1;
        const CUtensorMap *map = ch == 0 ? &map_a0 : &map_a1;
        const int row0 = tile * TILE_M + (int)rank * CTA_M;
        for (int kb = 0; kb < NUM_KB; kb++) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1, 1);
          if (leader) mbar_expect_tx(smem_u32(&bars->full[stage]), 2 * A_STAGE_BYTES);
          tma_load_2d_2sm(sA + stage * A_STAGE_BYTES, map, kb * K_BLOCK, row0, leader_full0 + stage * 8);
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
    // ===== query operand loader: 4 queries x 16 KB of this CTA's base vector (x: rank 0, y: rank 1) =====
    if (lane == 0) {
      uint32_t phase = 0;
      int prev_unit = -1;
      for (long long it = it_begin; it < it_end; ++it) {
        const int unit = (int)(it / P.n_tiles);
        if (unit == prev_unit) continue;
        prev_unit = unit;
        const int ch = unit & 1, qg = unit >> 1;
        mbar_wait(smem_u32(&bars->b_empty), phase ^ 1, 2);
        mbar_expect_tx(smem_u32(&bars->b_full), B_BYTES);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(P.q_op) +
                                   (((size_t)ch * 2 + rank) * P.m_pad + (size_t)qg * QG) * Q_BASE_BYTES;
        for (int q = 0; q < QG; q++)
          bulk_load_1d(sB + q * Q_BASE_BYTES, src + (size_t)q * Q_BASE_BYTES, Q_BASE_BYTES, smem_u32(&bars->b_full));
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
    // ===== MMA issuer (leader CTA, one thread) =====
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(TILE_M, N_PER_Q);
      int stage = 0;
      uint32_t phase = 0, b_phase = 0, t_phase = 0;
      int prev_unit = -1;
      for (long long it = it_begin; it < it_end; ++it) {
        const int unit = (int)(it / P.n_tiles);
        if (unit != prev_unit) {
          prev_unit = unit;
          mbar_wait(smem_u32(&bars->b_full), b_phase, 4);
          mbar_wait(smem_u32(&bars->b_peer), b_phase, 5);
          b_phase ^= 1;
        }
        mbar_wait(smem_u32(&bars->tmem_empty), t_phase ^ 1, 6);
        tc_fence_after();
        for (int kb = 0; kb < NUM_KB; kb++) {
          mbar_wait(smem_u32(&bars->full[stage]), phase, 7);
          tc_fence_after();
          const uint32_t a_base = sA + stage * A_STAGE_BYTES;
#pragma unroll
          for (int kk = 0; kk < K_BLOCK / 16; kk++) {
            const int k = kb * K_BLOCK + kk * 16;
            const int j = k / CHUNK_K, c = (k - j * CHUNK_K) >> 3;
            // A: SWIZZLE_128B K-major, 8-row groups 1024 B apart; K-step advances the start by 32 B
            const uint64_t adesc = make_desc(a_base + kk * 32, 16, 1024, 2);
#pragma unroll
            for (int q = 0; q < QG; q++) {
              // B: Hankel view, no swizzle: row s, k-group g -> unit (c + s + g) of chunk j
              const uint64_t bdesc =
                  make_desc(sB + q * Q_BASE_BYTES + j * (Q_UNITS * 16) + c * 16, P.dbg_lbo, P.dbg_sbo, 0);
              umma_f16_2sm(tmem_base + q * N_PER_Q, adesc, bdesc, idesc, (kb | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2sm(smem_u32(&bars->empty[stage]));
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(smem_u32(&bars->tmem_full));
        const bool last_of_unit = (it + 1 == it_end) || ((int)((it + 1) / P.n_tiles) != unit);
        if (last_of_unit) umma_commit_2sm(smem_u32(&bars->b_empty));
        t_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> max over 128 shift columns -> d = (1 - dot)/2 -> global =====
    const int ew = warp & 3;  // TMEM lane quarter accessible to this warp
    const uint32_t leader_tmem_empty = map_to_cta(smem_u32(&bars->tmem_empty), 0);
    uint32_t t_phase = 0;
    for (long long it = it_begin; it < it_end; ++it) {
      const int unit = (int)(it / P.n_tiles), tile = (int)(it - (long long)unit * P.n_tiles);
      const int ch = unit & 1, qg = unit >> 1;
      mbar_wait(smem_u32(&bars->tmem_full), t_phase, 8);
      tc_fence_after();
      const int row = tile * TILE_M + (int)rank * CTA_M + ew * 32 + lane;
      float *out = P.d_out[ch];
#pragma unroll 1
      for (int q = 0; q < QG; q++) {
        float best = __int_as_float(0x7fc00000);  // NaN: min over variants ignores NaN (processSC.m:31)
#pragma unroll 1
        for (int cb = 0; cb < N_PER_Q; cb += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(q * N_PER_Q + cb), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) best = fmaxf(best, __uint_as_float(r[i]));
        }
        const int qi = qg * QG + q;
        if (qi < P.m && row < P.n && out) out[(size_t)qi * P.ldd + row] = (1.0f - best * ACC_SCALE) * 0.5f;
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
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

inline int pad_to(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

size_t sc_tc_db_bytes(int n) { return (size_t)2 * pad_to(n, TILE_M) * K_TOTAL * sizeof(__half); }
size_t sc_tc_query_bytes(int m) { return (size_t)2 * 2 * pad_to(m, QG) * Q_BASE_BYTES; }

cudaError_t launch_sc_tc_prep_db(const double *hist, int n, void *db_buf, cudaStream_t st, int64_t *launches) {
  if (n <= 0) return cudaSuccess;
  const int n_pad = pad_to(n, TILE_M);
  sc_tc_prep_db_kernel<<<n_pad, 256, 0, st>>>(hist, n, n_pad, reinterpret_cast<__half *>(db_buf));
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_sc_tc_prep_query(const double *hist, int m, void *q_buf, cudaStream_t st, int64_t *launches) {
  if (m <= 0) return cudaSuccess;
  const int m_pad = pad_to(m, QG);
  sc_tc_prep_query_kernel<<<m_pad, 256, 0, st>>>(hist, m, m_pad, reinterpret_cast<__half *>(q_buf));
  if (launches) ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_sc_match_tc(const void *q_buf, int m, const void *db_buf, int n, float *d_p, float *d_i, int ldd,
                               int num_sms, cudaStream_t st, int64_t *launches) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  PFN_encodeTiled enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  const int n_pad = pad_to(n, TILE_M), m_pad = pad_to(m, QG);
  CUtensorMap maps[2];
  for (int ch = 0; ch < 2; ch++) {
    cuuint64_t dims[2] = {(cuuint64_t)K_TOTAL, (cuuint64_t)n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)K_TOTAL * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)K_BLOCK, (cuuint32_t)CTA_M};
    cuuint32_t estr[2] = {1, 1};
    void *gaddr = (void *)(reinterpret_cast<const __half *>(db_buf) + (size_t)ch * n_pad * K_TOTAL);
    CUresult r = enc(&maps[ch], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, gaddr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  }
  TcParams P;
  P.q_op = reinterpret_cast<const __half *>(q_buf);
  P.d_out[0] = d_p;
  P.d_out[1] = d_i;
  P.m = m;
  P.n = n;
  P.m_pad = m_pad;
  P.n_pad = n_pad;
  P.ldd = ldd;
  P.n_units = 2 * (m_pad / QG);
  P.n_tiles = n_pad / TILE_M;
  P.dbg_lbo = 16;
  P.dbg_sbo = 128;
  if (const char *e = getenv("SODSO_TC_LBO")) P.dbg_lbo = atoi(e);
  if (const char *e = getenv("SODSO_TC_SBO")) P.dbg_sbo = atoi(e);
  const long long W = (long long)P.n_units * P.n_tiles;
  int npairs = num_sms / 2;
  if (npairs > W) npairs = (int)W;
  if (npairs < 1) npairs = 1;
  cudaError_t e = cudaFuncSetAttribute(sc_match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return e;
  sc_match_tc_kernel<<<2 * npairs, TC_THREADS, SMEM_BYTES, st>>>(maps[0], maps[1], P);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
