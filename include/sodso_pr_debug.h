/*
 * sodso_pr_debug.h -- test hooks of libsodso_pr.so.  NOT part of the reference-facing surface (include/sodso_pr.h):
 * nothing here replaces a reference interface.  The product path is the tcgen05 matcher with the self-match symmetry
 * on; these switches exist so that tests can cross-check it on the GPU and tools/ can profile kernel variants.  The
 * library never reads behaviour from the environment.
 */
#ifndef SODSO_PR_DEBUG_H
#define SODSO_PR_DEBUG_H

#include "sodso_pr.h"

#ifdef __cplusplus
extern "C" {
#endif

/* matcher used by a context: TC = the tcgen05 kernels (default, the product path); SIMT = plain fp32 CUDA-core
 * kernels (csrc/sc_match_simt.cu, csrc/m2dp_match.cu) kept as an independent on-GPU cross-check */
#define SODSO_ALGO_TC 0
#define SODSO_ALGO_SIMT 1
int sodso_debug_set_match_algo(sodso_ctx *ctx, int algo);

/* on = 0: a self-match (hist1 == hist2) computes every pair like distinct operands instead of the lower block triangle
 * + mirrored stores (csrc/sc_match_tc.cu, launch_sc_match_tc_self) */
int sodso_debug_set_sc_symmetry(sodso_ctx *ctx, int on);

/* kernel variant switches used by tools/ (process-wide): tc_flags 1 = skip epilogue loads, 2 = skip MMAs, 4 = force
 * the generic fp16 mode for binary channels; gen_flags = sc_generate_kernel variant bits (-1 = default);
 * gen_ctas = CTAs per SM of sc_generate_kernel (0 = default) */
int sodso_debug_set_kernel_flags(int tc_flags, int gen_flags, int gen_ctas);

/* per-phase clock sums of m2dp_generate_kernel (thread 0 of every CTA, summed over CTAs and scans), used by
 * tools/profile_m2dp.py: enable != 0 allocates the 16-counter device buffer (the kernel then records), out16 != NULL
 * synchronises, copies the counters out and clears them, enable = 0 frees the buffer (process-wide, current device) */
int sodso_debug_phase_profile(int enable, unsigned long long *out16);

/* on = 0: communicators created afterwards do not set up the peer-memory windows, so the sharded exchange runs on
 * ncclAllReduce / ncclAllGather (tests compare the two transports) */
int sodso_debug_set_peer_exchange(int on);

/* the fp32 angle proposal atan2(num, den)/2pi + 1/2 that the generation kernels use to PROPOSE a polar bin (accepted
 * only outside an error-derived guard band around bin edges, otherwise SC.cpp:37 / M2DP.cpp:59 in fp64 decides);
 * exposed so that tests can check the error bound the guard band rests on */
int sodso_debug_fast_turns(sodso_ctx *ctx, const float *num, const float *den, int64_t n, float *out);

/* host only, no GPU work: the number of (query group of 4, DB tile of 256, channel) work items the tcgen05 matcher runs
 * for queries [q0, q1) of an n x n SELF-match (q0 a multiple of 256) -- the lower block triangle
 * tile_start <= group_end of processSC.m:22-33's all-pairs loop; -1 for bad arguments */
int64_t sodso_debug_sc_self_items(int64_t n, int64_t q0, int64_t q1);

#ifdef __cplusplus
}
#endif
#endif /* SODSO_PR_DEBUG_H */
