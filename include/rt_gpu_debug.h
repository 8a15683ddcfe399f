/* rt_gpu_debug.h — tooling entry points of librt_b200.so (tools/render_once.py, tools/timeline.py).
 *
 * Not part of the drop-in boundary (include/rt_gpu.h): nothing here replaces a reference interface.  They read
 * the per-round bookkeeping of the wavefront after a render call made with rt_gpu_time_kernels(ctx, 1), so that
 * the round structure can be inspected without a profiler.  Every function synchronises the context. */
#ifndef RT_GPU_DEBUG_H
#define RT_GPU_DEBUG_H

#include "rt_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Entries of each round of the last batch on pipe 0 (counts[k], k < max_rounds) and the walk-kernel time of
 * each timed bracket (ms[k]); counts[max_rounds - 1] is overwritten with the longest single walk (node steps).
 * Returns the number of timed brackets, or a negative rt_status. */
int rt_gpu_debug_rounds(rt_gpu_ctx* ctx, uint32_t* counts, float* ms, int32_t max_rounds);

/* Begin / end of every timed walk bracket of the last call in ms since the call began (launch order: chunk by
 * chunk, round by round).  Returns the number of brackets written (<= cap), or a negative rt_status. */
int rt_gpu_debug_timeline(rt_gpu_ctx* ctx, float* begin_ms, float* end_ms, int32_t cap);

/* Long-walk queue sizes per round of the last batch on pipe 0. */
int rt_gpu_debug_long(rt_gpu_ctx* ctx, uint32_t* lcounts, int32_t max_rounds);

#ifdef __cplusplus
}
#endif
#endif
