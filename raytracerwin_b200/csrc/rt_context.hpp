// rt_context.hpp — the opaque context behind include/rt_gpu.h and the error helpers every translation unit of
// the library uses (rt_gpu.cu: lifetime / upload / render / readback; rt_exchange.cu: multi-GPU tile exchange;
// rt_hooks.cu: test and tooling hooks; rt_bvh_build.cu: device BVH builder).
#pragma once
#include "rt_wave_types.hpp"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <string>
#include <vector>

struct rt_gpu_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    // per-launch timing of the path kernel inside the last render_tile (one pair per pass chunk)
    std::vector<cudaEvent_t> kev;
    int kev_used = 0;
    std::string err;
    int num_sms = 0;

    bool has_scene = false;
    bool needs_table = false;                   // scene has Diffuse materials (RandomHemisphereDirection)
    DevScene scene;
    std::vector<void*> scene_allocs;
    std::vector<cudaArray_t> arrays;
    std::vector<cudaTextureObject_t> texobjs;
    std::vector<DevTexture> host_textures;      // flat list of every texture (test hook)
    std::vector<char> host_shape_is_mesh;
    bool all_bounded = false;
    size_t scene_bytes = 0;
    size_t texel_upload_bytes = 0;              // host-to-device bytes the last upload spent on texels

    int width = 0, height = 0;
    float4* accum = nullptr;
    uint32_t* display = nullptr;
    int2* prim_ids = nullptr;
    float* prim_dist = nullptr;
    float4* preview = nullptr;                  // linear colour of the last preview pass (allocated by the first RT_MODE_PREVIEW call)
    unsigned long long* counters = nullptr;     // rt_counters
    // Frame slots (rt_gpu_set_frame_slot): the members above — stream, timing events, frame size and the five
    // frame buffers — are those of the CURRENT slot; the other slots' sets wait here.  Calls that address
    // different slots are not ordered against each other on the device, so a driver can enqueue frame k+1
    // while the thin last bounce rounds, the exchange and the read-back of frame k are still running.
    struct FrameSlot
    {
        cudaStream_t stream = nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        bool timed = false;
        int width = 0, height = 0;
        float4* accum = nullptr; uint32_t* display = nullptr; int2* prim_ids = nullptr; float* prim_dist = nullptr; float4* preview = nullptr;
    };
    FrameSlot parked[RT_FRAME_SLOTS];
    int slot = 0;
    unsigned slot_mask = 1u; int slots_touched = 1;   // which slots a driver has addressed so far
    bool slots_used = false;                    // rt_gpu_set_frame_slot has been called: calls rotate through the pipes
    int pipe_cursor = 0;
    // wavefront state: path pool, round queues, round counters
    // Batches of a call are dealt round-robin to RT_PIPES pipes, each with its own stream, pool and
    // queues, so the thin late rounds of one batch (few long walks: latency bound) overlap the
    // dense early rounds of the next.
    struct Pipe
    {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        PathPool pool;
        unsigned* queue[2] = { nullptr, nullptr };
        unsigned* round_counters = nullptr;     // counts[RT_MAX_ROUNDS + 1] then heads[RT_MAX_ROUNDS]
        unsigned* longq = nullptr;              // parked long walks of the current round
        unsigned* slowq = nullptr;              // walks of incoherent packets, handed to the lane-per-walk kernel
        unsigned* retry[2] = { nullptr, nullptr };   // items turned away by a full pool (ping-pong)
        unsigned* retry_counts = nullptr;       // one per retry pass
        unsigned* seen_counts = nullptr;        // PINNED HOST copy of the round sizes of the last batch on this pipe (first generate pass)
        unsigned long long seen_signature = 0;  // which call shape they belong to
        volatile unsigned* seen_retry = nullptr;   // PINNED HOST copy of the retry-list sizes of that batch (items each generate pass turned away)
        float4* samples = nullptr;              // radiance samples of the chunk this pipe is rendering
        size_t samples_cap = 0;                 // float4s
        size_t retry_cap = 0;
        std::vector<void*> allocs;
    };
    Pipe pipes[RT_PIPES];
    cudaEvent_t fork = nullptr;
    size_t pool_cap = 0, pool_levels = 0;       // per pipe
    bool pool_whitted = false;
    size_t max_pool_paths = RT_POOL_MAX_PATHS;  // per call, over all pipes
    int tune_pipes = RT_PIPES;
    int walk_blocks_per_sm = 0;
    struct TileTable { int width, height, tile_size, tile_count, rank; long long* offsets; };
    std::vector<TileTable> tile_tables;
    float4* gather_staging = nullptr;
    size_t gather_staging_cap = 0;
    unsigned long long launches = 0;            // kernels launched by this context
    unsigned tune_window = RT_WORK_WINDOW;
    int tune_min_lanes = RT_MIN_LANES;
    int tune_leaf_wait = RT_LEAF_WAIT;
    int tune_finish_round = RT_FINISH_ROUND;
    bool time_walks = false;                    // record an event pair around every walk launch (rt_gpu_time_kernels)
    bool time_classes = false;                  // rt_gpu_time_kernels(ctx, 2): one event before every launch, by kernel class
    std::vector<cudaEvent_t> cev;
    std::vector<int> cev_cls;
    int cev_used = 0;
    unsigned tune_long_limit = RT_LONG_LIMIT;
    unsigned tune_small_round = RT_SMALL_ROUND;
    unsigned tune_thin_count = RT_THIN_COUNT;
    int tune_long_group = RT_LONG_GROUP;
    int tune_packet_rounds = -1;            // -1: by mode
    unsigned tune_packet_probe = RT_PACKET_PROBE;
    unsigned tune_packet_min_lanes = RT_PACKET_MIN_LANES;
    unsigned tune_thin_limit = RT_THIN_LIMIT;
    int tune_thin_from_round = 0;               // RT_THIN_FROM_ROUND: 0 = by the sizes last seen (frames in flight only), k > 0 = from round k, -1 = never
    unsigned tune_thin_grid_count = RT_THIN_GRID_COUNT;   // a round that had fewer entries than this gets quarter grids
    bool tune_octo = false;                     // RT_OCTO=1: bounce rounds of the culled traversal walk the 8-wide tree (measured slower: opt-in)
    int octo_blocks_per_sm = 0;
    bool tune_top_stage = false;                // RT_TOP_STAGE: the walk kernel serves the top of mesh 0's tree from shared memory
    int tune_few_chunks = 0;                    // RT_FEW_CHUNKS: pass chunks of a call of few camera rays (0: 2, or 1 with frame slots)
    size_t tune_sample_budget = 0;              // RT_SAMPLE_BUDGET_MB at create (0: by call size)
    bool tune_time_long = false;                // RT_TIME_LONG at create: timed brackets span walk + long walk
};

extern thread_local std::string g_create_error;

static inline int fail(rt_gpu_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define RT_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

// No C++ exception crosses the C ABI (include/rt_gpu.h: every entry returns 0 or a negative rt_status): the
// bodies of the entry points that allocate host memory (std::vector / std::string) run inside this guard.
template <typename F>
static inline int rt_guard(rt_gpu_ctx* ctx, F&& body) noexcept
{
    try { return body(); }
    catch (const std::bad_alloc&) { try { return fail(ctx, RT_ERR_NOMEM, "out of host memory"); } catch (...) { return RT_ERR_NOMEM; } }
    catch (const std::exception& e) { try { return fail(ctx, RT_ERR_INVALID, e.what()); } catch (...) { return RT_ERR_INVALID; } }
    catch (...) { return RT_ERR_INVALID; }
}
