/* Hardware probes (measurement scaffolding, not part of the product ABI): compiled into libsfd2_b200.so only when
 * the library is built with SFD2_WITH_PROBES=1 (python -m sfd2_b200.build --force).  See tools/umma_probe.py and
 * tools/mma_rate_probe.py. */
#pragma once
#include "../../include/sfd2_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Which UMMA descriptor settings let a 3x3 tap read a shifted view of one TMA halo tile. */
SFD2_API int sfd2_debug_umma_probe(int pitch, int ky, int kx, int use_base_offset, int pattern, float* out_host);
/* Cycles per CTA for `iters` back-to-back SMEM-operand tcgen05.mma of shape M128 x n x K (kind 0: f16, K = 16;
 * kind 1: f8f6f4 / E4M3, K = 32) on `grid` CTAs. */
SFD2_API int sfd2_debug_mma_rate(int n, int kind, int iters, int grid, unsigned long long* cycles_host);
#ifdef __cplusplus
}
#endif
