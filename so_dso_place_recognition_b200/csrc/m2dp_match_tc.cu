// M2DP all-pairs match (processM2DP.m:12-22) on the 5th-generation tensor cores.
//
//   diff_full = (1 - hist1 * hist2') / 2 on the 4m x 4n sign-variant rows (NO normalisation, processM2DP.m:15),
//   diff(i, j) = min of the 4 x 4 block (processM2DP.m:17-21)  =  (1 - max of the 16 dot products) / 2.
//
// A plain K-major GEMM D[r, c] = sum_k A[r, k] B[c, k] per channel: A = DB variant rows (M side, 256 per CTA pair),
// B = query variant rows (N = 256), K = 192 evaluated as the 3-term fp16 split hi*lo + lo*hi + hi*hi (K' = 576, cross
// terms first; values scaled by 64, the dropped lo*lo is 2^-22) with fp32 accumulation in TMEM -- the same precision
// scheme as the Scan Context matcher.  tcgen05.mma.cta_group::2, both operands by TMA (SWIZZLE_128B, 5-stage mbarrier
// ring), a 256 x 256 tile uses half of TMEM, so the two halves are double-buffered: the epilogue of tile t (tcgen05.ld,
// max over 4 adjacent columns in registers and over 4 adjacent lanes by shuffles, (1 - x)/2, store) overlaps the MMAs
// of tile t + 1.
// Roles per CTA (256 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA, warp-uniform loop, one elected
// lane), warp 2 TMEM allocator, warps 4-7 epilogue.
#include <cstdlib>

#include "../../include/sodso_pr.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace sodso {
namespace {
using namespace tc;

constexpr int MK = M2DP_SIG;            // 192
constexpr int MK3 = 3 * MK;             // 576 fp16 per operand row
constexpr int ROW_BYTES = MK3 * 2;      // 1152
constexpr int NUM_KB = MK3 / 64;        // 9 K-blocks of 64 halves (128 B)
constexpr int TILE = 256, CTA_ROWS = 128;
constexpr int STAGE_BYTES = 2 * CTA_ROWS * 128;   // A half-tile + B half-tile: 32 KB per CTA
constexpr int M2_NSTAGE = 5;
constexpr float M2_SCALE = 64.0f, M2_ACC_SCALE = 1.0f / (M2_SCALE * M2_SCALE);
constexpr int M2TC_THREADS = 256;

struct __align__(8) M2Bars {
  uint64_t full[M2_NSTAGE], empty[M2_NSTAGE];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_ptr, pad;
};
constexpr int M2_SMEM_BYTES = 1024 + M2_NSTAGE * STAGE_BYTES + (int)sizeof(M2Bars);

inline int pad_to(int v, int a) { return (v + a - 1) / a * a; }

// hist rows (rows x 384 fp64: [count 192 | intensity 192]) -> per channel fp16 split rows of 576:
//   DB side   [hi | lo | hi],   query side [lo | hi | hi]      (hi*lo + lo*hi + hi*hi)
__global__ void __launch_bounds__(192)
m2dp_tc_prep_kernel(const double *__restrict__ hist, int rows, int rows_pad, __half *__restrict__ out, int is_db) {
  const int r = blockIdx.x, k = threadIdx.x;
  for (int ch = 0; ch < 2; ch++) {
    __half hi = __float2half(0.0f), lo = hi;
    if (r < rows) {
      const double v = hist[(size_t)r * 2 * MK + ch * MK + k] * (double)M2_SCALE;
      hi = __float2half_rn((float)v);
      lo = __float2half_rn((float)(v - (double)__half2float(hi)));
    }
    __half *o = out + ((size_t)ch * rows_pad + r) * MK3;
    o[k] = is_db ? hi : lo;
    o[MK + k] = is_db ? lo : hi;
    o[2 * MK + k] = hi;
  }
}

struct M2TcParams {
  float *d_out[2];
  int m, n, ldd;            // queries, DB scans
  int tiles_q, tiles_d;     // 256-row tiles of variant rows
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(M2TC_THREADS, 1)
m2dp_match_tc_kernel(const __grid_constant__ CUtensorMap map_db0, const __grid_constant__ CUtensorMap map_db1,
                     const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                     const M2TcParams P) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_raw = smem_u32(smem_dyn);
  const uint32_t base = (base_raw + 1023u) & ~1023u;
  unsigned char *smem = smem_dyn + (base - base_raw);
  const uint32_t sS = base;   // stages: [A 16 KB | B 16 KB]
  M2Bars *bars = reinterpret_cast<M2Bars *>(smem + M2_NSTAGE * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  // work items: (channel, query tile, DB tile), contiguous range per CTA pair
  const long long W = 2LL * P.tiles_q * P.tiles_d;
  const long long it_begin = W * pair_id / npairs, it_end = W * (pair_id + 1) / npairs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < M2_NSTAGE; s++) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(smem_u32(&bars->tmem_full[s]), 1);
      mbar_init(smem_u32(&bars->tmem_empty[s]), 8);   // 4 epilogue warps x 2 CTAs
    }
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
    // ===== TMA producer: this CTA's 128 DB rows and 128 query rows of every K-block =====
    if (lane == 0) {
      const uint32_t leader_full0 = map_to_cta(smem_u32(&bars->full[0]), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (long long it = it_begin; it < it_end; ++it) {
        const int ch = (int)(it / ((long long)P.tiles_q * P.tiles_d));
        const int rem = (int)(it - (long long)ch * P.tiles_q * P.tiles_d);
        const int tq = rem / P.tiles_d, td = rem - tq * P.tiles_d;
        const CUtensorMap *ma = ch == 0 ? &map_db0 : &map_db1, *mb = ch == 0 ? &map_q0 : &map_q1;
        const int ra = td * TILE + (int)rank * CTA_ROWS, rb = tq * TILE + (int)rank * CTA_ROWS;
        for (int kb = 0; kb < NUM_KB; kb++) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1, 1);
          if (leader) mbar_expect_tx(smem_u32(&bars->full[stage]), 2 * STAGE_BYTES);
          tma_load_2d_2sm(sS + stage * STAGE_BYTES, ma, kb * 64, ra, leader_full0 + stage * 8);
          tma_load_2d_2sm(sS + stage * STAGE_BYTES + CTA_ROWS * 128, mb, kb * 64, rb, leader_full0 + stage * 8);
          if (++stage == M2_NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (it_end > it_begin)
        for (int i = 0; i < M2_NSTAGE; i++) {   // drain: all multicast commits have landed before exit
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1, 9);
          if (++stage == M2_NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc(TILE, TILE);
      const uint64_t adesc0 = make_desc(sS, 16, 1024, 2), bdesc0 = make_desc(sS + CTA_ROWS * 128, 16, 1024, 2);
      int stage = 0;
      uint32_t phase = 0, t_phase[2] = {0, 0};
      int slot = 0;
      for (long long it = it_begin; it < it_end; ++it) {
        mbar_wait(smem_u32(&bars->tmem_empty[slot]), t_phase[slot] ^ 1, 6);
        tc_fence_after();
        uint32_t acc = 0;
#pragma unroll 1
        for (int kb = 0; kb < NUM_KB; kb++) {
          mbar_wait(smem_u32(&bars->full[stage]), phase, 7);
          tc_fence_after();
          const uint64_t so = (uint64_t)(stage * (STAGE_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              umma_f16_2sm(tmem_base + slot * TILE, adesc0 + so + (uint64_t)(kk * 2), bdesc0 + so + (uint64_t)(kk * 2), idesc,
                           acc);
              acc = 1;
            }
            umma_commit_2sm(smem_u32(&bars->empty[stage]));
          }
          acc = 1;
          __syncwarp();
          if (++stage == M2_NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit_2sm(smem_u32(&bars->tmem_full[slot]));
        __syncwarp();
        t_phase[slot] ^= 1;
        slot ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> max over the 4 x 4 variant block -> (1 - dot)/2 -> global =====
    const int ew = warp & 3;
    const uint32_t leader_tmem_empty0 = map_to_cta(smem_u32(&bars->tmem_empty[0]), 0);
    uint32_t t_phase[2] = {0, 0};
    int slot = 0;
    for (long long it = it_begin; it < it_end; ++it) {
      const int ch = (int)(it / ((long long)P.tiles_q * P.tiles_d));
      const int rem = (int)(it - (long long)ch * P.tiles_q * P.tiles_d);
      const int tq = rem / P.tiles_d, td = rem - tq * P.tiles_d;
      mbar_wait(smem_u32(&bars->tmem_full[slot]), t_phase[slot], 8);
      tc_fence_after();
      // lane = DB variant row inside this CTA's 128; DB scan = row / 4
      const int db_scan = (td * TILE + (int)rank * CTA_ROWS + ew * 32 + lane) >> 2;
      float *out = P.d_out[ch];
      const uint32_t tcol = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(slot * TILE);
#pragma unroll 1
      for (int cb = 0; cb < TILE; cb += 32) {
        uint32_t r[32];
        tmem_ld32(tcol + (uint32_t)cb, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; g++) {
          float best = fmaxf(fmaxf(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1])),
                             fmaxf(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])));
          best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 1));
          best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 2));
          const int qi = (tq * TILE + cb) / 4 + g;
          if ((lane & 3) == 0 && qi < P.m && db_scan < P.n && out)
            out[(size_t)qi * P.ldd + db_scan] = (1.0f - best * M2_ACC_SCALE) * 0.5f;   // processM2DP.m:15,19
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(leader_tmem_empty0 + slot * 8);
      t_phase[slot] ^= 1;
      slot ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

size_t m2dp_match_tc_workspace_bytes(int m, int n) {
  return ((size_t)2 * pad_to(4 * m, TILE) + (size_t)2 * pad_to(4 * n, TILE)) * ROW_BYTES + 1024;
}

cudaError_t launch_m2dp_match_tc(const double *hist1, int m, const double *hist2, int n, float *d_p, float *d_i, int ldd,
                                 void *workspace, int num_sms, cudaStream_t st, int64_t *launches) {
  if (m <= 0 || n <= 0) return cudaSuccess;
  PFN_encodeTiled enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  const int mq = pad_to(4 * m, TILE), nd = pad_to(4 * n, TILE);
  // 1024-byte aligned operand buffers inside the workspace
  unsigned char *w = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  __half *op_db = reinterpret_cast<__half *>(w);
  __half *op_q = reinterpret_cast<__half *>(w + (size_t)2 * nd * ROW_BYTES);
  m2dp_tc_prep_kernel<<<nd, MK, 0, st>>>(hist2, 4 * n, nd, op_db, 1);
  m2dp_tc_prep_kernel<<<mq, MK, 0, st>>>(hist1, 4 * m, mq, op_q, 0);
  if (launches) *launches += 2;
  CUtensorMap maps[4];
  for (int side = 0; side < 2; side++)
    for (int ch = 0; ch < 2; ch++) {
      const int rows = side == 0 ? nd : mq;
      cuuint64_t dims[2] = {(cuuint64_t)MK3, (cuuint64_t)rows};
      cuuint64_t strides[1] = {(cuuint64_t)ROW_BYTES};
      cuuint32_t box[2] = {64, (cuuint32_t)CTA_ROWS};
      cuuint32_t estr[2] = {1, 1};
      void *gaddr = (void *)((side == 0 ? op_db : op_q) + (size_t)ch * rows * MK3);
      CUresult r = enc(&maps[side * 2 + ch], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, gaddr, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
  M2TcParams P;
  P.d_out[0] = d_p;
  P.d_out[1] = d_i;
  P.m = m;
  P.n = n;
  P.ldd = ldd;
  P.tiles_q = mq / TILE;
  P.tiles_d = nd / TILE;
  const long long W = 2LL * P.tiles_q * P.tiles_d;
  int npairs = num_sms / 2;
  if (npairs > W) npairs = (int)W;
  if (npairs < 1) npairs = 1;
  cudaError_t e = cudaFuncSetAttribute(m2dp_match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  m2dp_match_tc_kernel<<<2 * npairs, M2TC_THREADS, M2_SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], P);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace sodso
