// rt_gpu.cu — the B200 (sm_100a) implementation behind include/rt_gpu.h.
//
// Replaces ThreadWorker_Render + everything beneath it (RayTracerProgram.cpp:131-188,
// RayTracerScene.cpp:31-175, KdTree.cpp:128-232, MeshShape.cpp:280-331, SurfaceMaterials.cpp,
// RRay.cpp, Texture.cpp:23-57) and the ThreadTaskQueue dispatch (ThreadTaskQueue.h) that feeds it.
//
// Execution model (not the reference's): a WAVEFRONT.  The work list of a render call is every
// (sample, pixel) pair, enumerated as 8x4-pixel blocks so that a warp gets coherent camera rays.
//   generate  one thread per item: camera ray, shape list up to the first mesh whose bounds the ray
//             enters.  Rays that end there having hit nothing (most: they see the sky) are retired
//             on the spot at full warp width; the rest become PATHS — a record in a pool, an entry
//             in the round-0 queue, compacted per warp with __ballot_sync/__popc/__shfl_sync.
//   walk      persistent warps pop path ids from the round's queue and walk the mesh (the
//             reference's KdNode recursion, flattened); a lane that finishes pops the next id, so
//             live walks stay packed in full warps although one ray visits 3 nodes and its
//             neighbour 2000.  Round 0 (camera rays, coherent) is walked by the packet kernel
//             instead — a warp's 32 rays share one node fetch per step — and walks that would hold
//             a round (very long ones, thin rounds) by the long-walk kernel, a warp per walk.
//   shade     one thread per queue entry: hit attributes, texture, rest of the shape list, material
//             bounce / alpha / light loop; ends the path (fold + sample) or starts its next segment
//             and pushes it — ballot-compacted again — into the next round's queue.
// walk and shade alternate once per (segment x mesh); nothing returns to the host in between.
// Each finished path writes one float4 radiance sample; a streaming kernel folds the samples of a
// pixel into accuBuffer[] in the reference's order (4 sub-samples -> /4 -> AddPixel per pass,
// RayTracerProgram.cpp:155-185), which keeps the accumulation bit-identical for any GPU count and
// any scheduling.
//
// The whole file is compiled with -fmad=false; see rt_device.cuh.
#include "rt_device.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

using namespace rtdev;

#define RT_WORK_WINDOW 32u                  // queue entries a warp takes from the round's pop cursor at once
#define RT_POOL_MAX_PATHS (32u << 20)       // path records per pool (~0.5 KB each with a 10-level stack)
#define RT_LEAF_WAIT 12                     // leaves that wait before the walkers are interrupted
#define RT_MIN_LANES 28                     // refill threshold of the mesh walk
#define RT_SAMPLE_BUDGET_BYTES (3ull << 30) // sample buffer cap; longer calls are split in pass chunks
#define RT_SAMPLE_BUDGET_FEW_BYTES (12ull << 30)    // the cap for calls of fewer than RT_FEW_ITEMS camera rays (two chunks)
#define RT_FEW_ITEMS 400000000ull

// ---- work-list geometry ----------------------------------------------------------------------------
struct RenderArgs
{
    int width, height, start, end, mode, max_bounce, antialias;
    uint32_t seed;
    int pass_begin;             // first pass of this chunk
    int spp;                    // samples per pass: 4 (antialias) or 1
    int num_samples;            // pass_count_chunk * spp
    // pixel blocks
    int tiled;                  // 0: one region (rows row0..), 1: round-robin tiles
    int row0, rows;             // untiled region
    int tile_size, tile_count, tile_rank, tiles_x;
    int blocks_x;               // 8-wide blocks per region/tile row
    int blocks_per_tile;
    unsigned num_blocks;
    unsigned num_items;         // num_samples * num_blocks * 32 (the host keeps it below 2^32)
    float4* samples;            // [num_samples][width*height]
    float4* accum;
    uint32_t* display;
    int2* prim_ids;
    float* prim_dist;
    unsigned long long* counters;
    int exact;                  // traverse == RT_TRAVERSE_EXACT: node_tests/tri_tests are the visits
    int all_bounded;            // every shape has culling bounds (no plane): rays that miss them all see the sky
};

__device__ __forceinline__ bool owns_pixel(const RenderArgs& a, int x, int y)
{
    if (!a.tiled) return true;
    int tile = (y / a.tile_size) * a.tiles_x + x / a.tile_size;
    return tile % a.tile_count == a.tile_rank;
}

// block index + lane -> pixel (or -1 when the lane falls outside the region / image / task range)
__device__ __forceinline__ int block_pixel(const RenderArgs& a, unsigned block, int lane, int& x, int& y)
{
    x = 0; y = 0;
    int ox, oy, w, h, b;
    if (a.tiled)
    {
        unsigned k = block / (unsigned)a.blocks_per_tile;
        b = (int)(block - k * (unsigned)a.blocks_per_tile);
        int tile = a.tile_rank + (int)k * a.tile_count;
        ox = (tile % a.tiles_x) * a.tile_size; oy = (tile / a.tiles_x) * a.tile_size;
        w = a.tile_size; h = a.tile_size;
    }
    else { b = (int)block; ox = 0; oy = a.row0; w = a.width; h = a.rows; }
    int bx = b % a.blocks_x, by = b / a.blocks_x;
    int lx = bx * 8 + (lane & 7), ly = by * 4 + (lane >> 3);
    if (lx >= w || ly >= h) return -1;
    x = ox + lx; y = oy + ly;
    if (x >= a.width || y >= a.height) return -1;
    int pixel = y * a.width + x;
    if (pixel < a.start || pixel > a.end) return -1;
    return pixel;
}

__device__ __forceinline__ void flush_counters(const Counters& c, unsigned long long* g, int exact)
{
    unsigned long long v[6] = { c.rays, c.camera_rays, c.shadow_rays, c.node_visits, c.tri_visits, c.mesh_hits };
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
        unsigned long long x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(RT_FULL_MASK, x, o);
        v[k] = x;
    }
    if ((threadIdx.x & 31) == 0)
    {
        // rt_counters: rays, camera_rays, shadow_rays, node_tests, tri_tests, node_visits, tri_visits, mesh_hits
        if (v[0]) atomicAdd(g + 0, v[0]);
        if (v[1]) atomicAdd(g + 1, v[1]);
        if (v[2]) atomicAdd(g + 2, v[2]);
        if (v[3]) { atomicAdd(g + 5, v[3]); if (exact) atomicAdd(g + 3, v[3]); }
        if (v[4]) { atomicAdd(g + 6, v[4]); if (exact) atomicAdd(g + 4, v[4]); }
        if (v[5]) atomicAdd(g + 7, v[5]);
    }
}

// ---- path pool ---------------------------------------------------------------------------------------
// Every camera ray that cannot be retired on the spot becomes a PATH with a record in this pool
// (structure of arrays, one 16-byte word per field group so a warp reads/writes whole lines).
// Queues hold path ids; a path keeps its id for its whole life.
struct PathPool
{
    float4* ro;         // TestRay origin, .w = current Distance (shrinks as hits are accepted)
    float4* rd;         // direction, .w = Distance of the segment as it was shot
    int4* cur;          // x: shape cursor si, y: best leaf slot, z: state | any << 8 | sky_on_miss << 9, w: hit shape
    float4* bp;         // position of the last accepted triangle
    float4* h0;         // RayHitResult: HitPosition, Distance
    float4* h1;         //               HitNormal, SampledAlpha
    float4* h2;         //               SampledColor, .w = triangle id of the hit (int bits)
    int4* pa;           // pixel, sample slot, rng key, rng draw counter
    int4* pb;           // depth_left, stack height, pass-through mask, light cursor
    float4* w0;         // Whitted only: primary hit position / normal / surface colour / running sum
    float4* w1;
    float4* w2;
    float4* w3;
    float4* st0;        // unwinding stack, [level * cap + path]: att.xyz col.x | col.yz emi.xy | emi.z
    float4* st1;
    float* st2;
    unsigned cap;
};

struct PathState
{
    int pixel, slot;
    Rng rng;
    int depth_left, sp;
    unsigned pass_mask;
    int light;
    float seg_dist;
    float3 w_pos, w_nrm, w_surface, w_sum;
};

struct WaveArgs
{
    PathPool pool;
    unsigned* queue[2];         // path ids of round r live in queue[r & 1]
    unsigned* counts;           // counts[r]: entries of round r;  counts[RT_MAX_ROUNDS]: paths allocated
    unsigned* heads;            // heads[r]: pop cursor of the walk kernel in round r
    unsigned* longq;            // walks the walk kernel gave up on (too long): finished one-warp-per-walk
    unsigned* lcounts;          // lcounts[r] / lheads[r]: entries and pop cursor of longq in round r
    unsigned* lheads;
    unsigned long_limit;        // node steps after which a lane hands its walk to the long-walk kernel
    unsigned thin_count;        // a round with fewer entries than this is latency-bound (its longest walk decides):
    unsigned thin_limit;        //   its walks are parked after thin_limit steps already
    unsigned* slowq;            // walks the packet kernel hands back to the lane-per-walk kernel (incoherent packets)
    unsigned* scounts;          // scounts[r], sheads[r]: that queue's size and pop cursor in round r
    unsigned* sheads;
    unsigned packet_probe;      // a packet is judged every this many steps:
    unsigned packet_min_lanes;  //   fewer lane-tests than packet_min_lanes (of 32, scaled to the packet's rays) per step -> not coherent
    int packets;                // round 0 is allocated in aligned packets of 32 (one generate warp each)
    unsigned small_round;       // a round with fewer entries than this is walked entirely one-warp-per-walk
    unsigned item_begin, item_count;   // slice of the work list this batch generates
    const unsigned* retry_in;          // retry pass: the items to generate (else null) and how many
    const unsigned* retry_in_count;
    unsigned* retry_out;               // items that found the pool full
    unsigned* retry_out_count;
    int min_lanes, leaf_wait;
    unsigned window;
};

#define RT_MAX_ROUNDS 512
#ifndef RT_PIPES
#define RT_PIPES 4
#endif
#define RT_MAX_RETRIES 64
#define RT_SMALL_ROUND 24000u               // rounds thinner than this are walked one-warp-per-walk only (frontier kernel)
#ifndef RT_SHADE_BLOCKS
#define RT_SHADE_BLOCKS 2
#endif
#ifndef RT_GEN_BLOCKS
#define RT_GEN_BLOCKS 3
#endif
#define RT_LONG_LIMIT 2048u                 // node steps after which a lane parks its walk for the long-walk kernel
#define RT_THIN_COUNT 200000u               // rounds thinner than this park after RT_THIN_LIMIT steps (0 = never):
#define RT_PACKET_PROBE 24u                 // a packet is judged every this many steps ...
#define RT_PACKET_MIN_LANES 10u             // ... and goes on lane by lane if fewer lanes than this tested a node per step
#define RT_THIN_LIMIT 256u                  //   their time is their longest walk, and the frontier kernel shortens exactly that
#ifndef RT_LONG_BLOCKS
#define RT_LONG_BLOCKS 4
#endif
#ifndef RT_LEAF_SLOTS
#define RT_LEAF_SLOTS 2                     // leaves a lane may hold before its walk has to wait for the triangle phase
#endif
#define RT_FINISH_ROUND 0                   // rounds run as walk/shade waves; the rest in one finishing launch (0: never)
#ifndef RT_WALK_BLOCKS
#define RT_WALK_BLOCKS 4                    // resident 256-thread CTAs per SM of the walk kernel (64 registers)
#endif

template <int MODE>
__device__ __forceinline__ void pool_store(const PathPool& p, unsigned id, const Query& q, int state, const PathState& s, bool sky_on_miss = false)
{
    p.ro[id] = make_float4(q.r.o.x, q.r.o.y, q.r.o.z, q.r.dist);
    p.rd[id] = make_float4(q.r.d.x, q.r.d.y, q.r.d.z, s.seg_dist);
    p.cur[id] = make_int4(q.si, q.best, state | (q.any ? 256 : 0) | (sky_on_miss ? 512 : 0), q.hit_shape);
    p.bp[id] = make_float4(q.bpos.x, q.bpos.y, q.bpos.z, 0.0f);
    p.h0[id] = make_float4(q.h.pos.x, q.h.pos.y, q.h.pos.z, q.h.dist);
    p.h1[id] = make_float4(q.h.nrm.x, q.h.nrm.y, q.h.nrm.z, q.h.alpha);
    p.h2[id] = make_float4(q.h.color.x, q.h.color.y, q.h.color.z, __int_as_float(q.tri));
    p.pa[id] = make_int4(s.pixel, s.slot, (int)s.rng.key, (int)s.rng.n);
    p.pb[id] = make_int4(s.depth_left, s.sp, (int)s.pass_mask, s.light);
    if (MODE == RT_MODE_WHITTED)
    {
        p.w0[id] = make_float4(s.w_pos.x, s.w_pos.y, s.w_pos.z, 0.0f);
        p.w1[id] = make_float4(s.w_nrm.x, s.w_nrm.y, s.w_nrm.z, 0.0f);
        p.w2[id] = make_float4(s.w_surface.x, s.w_surface.y, s.w_surface.z, 0.0f);
        p.w3[id] = make_float4(s.w_sum.x, s.w_sum.y, s.w_sum.z, 0.0f);
    }
}

template <int MODE>
__device__ __forceinline__ void pool_load(const PathPool& p, unsigned id, Query& q, int& state, PathState& s)
{
    const float4 ro = p.ro[id], rd = p.rd[id], bp = p.bp[id], h0 = p.h0[id], h1 = p.h1[id], h2 = p.h2[id];
    const int4 cur = p.cur[id], pa = p.pa[id], pb = p.pb[id];
    q.r.o = xyz(ro); q.r.dist = ro.w; q.r.d = xyz(rd); s.seg_dist = rd.w;
    q.pre = ray_pre(q.r);
    q.weird = !(q.pre.ex && q.pre.ey && q.pre.ez && finite3(q.r.o) && finite3(q.r.d));
    q.si = cur.x; q.best = cur.y; state = cur.z & 255; q.any = (cur.z & 256) != 0; q.hit_shape = cur.w;
    q.node = 0;
    q.bpos = xyz(bp); q.tri = __float_as_int(h2.w);
    q.h.pos = xyz(h0); q.h.dist = h0.w; q.h.nrm = xyz(h1); q.h.alpha = h1.w; q.h.color = xyz(h2);
    s.pixel = pa.x; s.slot = pa.y; s.rng.key = (uint32_t)pa.z; s.rng.n = (uint32_t)pa.w;
    s.depth_left = pb.x; s.sp = pb.y; s.pass_mask = (unsigned)pb.z; s.light = pb.w;
    if (MODE == RT_MODE_WHITTED)
    {
        s.w_pos = xyz(p.w0[id]); s.w_nrm = xyz(p.w1[id]); s.w_surface = xyz(p.w2[id]); s.w_sum = xyz(p.w3[id]);
    }
    else { s.w_pos = s.w_nrm = s.w_surface = s.w_sum = V3(0, 0, 0); }
}

// append the calling lanes' path ids to a queue: one atomic per warp (ballot -> leader add -> shuffle)
__device__ __forceinline__ void queue_push(unsigned* queue, unsigned* count, bool push, unsigned id)
{
    const unsigned active = __activemask();
    const unsigned mask = __ballot_sync(active, push);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(mask));
    base = __shfl_sync(active, base, leader);
    if (push) queue[base + (unsigned)__popc(mask & ((1u << lane) - 1u))] = id;
}

// allocate path ids the same way
__device__ __forceinline__ unsigned path_alloc(unsigned* counter, bool want)
{
    const unsigned active = __activemask();
    const unsigned mask = __ballot_sync(active, want);
    if (mask == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(mask));
    base = __shfl_sync(active, base, leader);
    return base + (unsigned)__popc(mask & ((1u << lane) - 1u));
}

// The same, in whole packets: a warp with at least one taker allocates 32 ids, takers first, so that every
// aligned group of 32 queue entries comes from ONE warp of the generate kernel (one 8x4-pixel block) and
// the packet walk finds coherent rays.  `spare` is the id a non-taker has to mark as unused (or ~0u).
__device__ __forceinline__ unsigned path_alloc_packet(unsigned* counter, bool want, unsigned& spare)
{
    spare = 0xffffffffu;
    const unsigned active = __activemask();
    const unsigned mask = __ballot_sync(active, want);
    if (mask == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counter, 32u);
    base = __shfl_sync(active, base, leader);
    const unsigned below = (1u << lane) - 1u;
    if (!want) spare = base + (unsigned)__popc(mask) + (unsigned)__popc(~mask & below);
    return base + (unsigned)__popc(mask & below);
}

// ---- shading of one completed query ---------------------------------------------------------------------
// RayTracerScene::RayTrace's body after FindIntersectionWithScene (RayTracerScene.cpp:44-97), the
// light loop of the Whitted configuration (CalculateLightColor, :127-175), or the id dump.
// Per-path unwinding record: RayTrace combines the radiance of the NEXT segment as
//   final = 0 + (att * L_next) * SampledColor; final += emissive
// on the way back up its recursion.  The path runs the recursion forwards and keeps (att, colour,
// emissive) per level in the pool so the fold runs in exactly the reference's order and rounding;
// pass-through levels (:79-85, final = 0 + L_next) only set a bit.
// Returns true when the path continues with `next` (query not yet begun); otherwise the path has
// ended and its sample has been written.
template <int MODE>
__device__ __forceinline__ bool shade_query(const DevScene& sc, const RenderArgs& a, const PathPool& pool, unsigned id,
                                            const Query& q, PathState& s, Ray& next, bool& next_any)
{
    bool done = false, newseg = false;
    next_any = false;
    float3 L = V3(0, 0, 0);
    const int shape = q.hit_shape;
    Ray in; in.o = q.r.o; in.d = q.r.d; in.dist = s.seg_dist;
    if (MODE == RT_MODE_PRIMARY)
    {
        a.prim_ids[s.pixel] = make_int2(shape, shape >= 0 ? q.tri : -1);
        a.prim_dist[s.pixel] = shape >= 0 ? q.h.dist : 0.0f;
        return false;
    }
    else if (MODE == RT_MODE_WHITTED)
    {
        bool next_light = false;
        if (!q.any)
        {
            if (shape == -1) { L = sky_color(in.d); done = true; }
            else
            {
                s.w_pos = q.h.pos; s.w_nrm = q.h.nrm; s.w_surface = q.h.color; s.w_sum = V3(0, 0, 0);
                s.light = 0; next_light = true;
            }
        }
        else
        {
            // CalculateLightColor: black if occluded, else SurfaceColor * max(0, N.L)
            float3 c = V3(0, 0, 0);
            if (shape == -1) c = mulf3(s.w_surface, max_ref(0.0f, dot3(s.w_nrm, in.d)));
            s.w_sum = add3(s.w_sum, c);
            s.light++; next_light = true;
        }
        if (next_light)
        {
            if (s.light >= sc.num_lights) { L = s.w_sum; done = true; }
            else
            {
                const rt_light* l = sc.lights + s.light;
                float3 ldir = ld3(l->pos_or_dir);
                float dist = 0.0f;
                if (l->type == RT_LIGHT_POINT)
                {
                    ldir = normalized3(sub3(ld3(l->pos_or_dir), s.w_pos));
                    dist = magnitude3(sub3(s.w_pos, ld3(l->pos_or_dir)));
                }
                else if (l->type == RT_LIGHT_DIRECTIONAL) dist = 1000.0f;
                next.o = add3(s.w_pos, mulf3(ldir, sc.bounce_offset)); next.d = ldir; next.dist = dist;
                newseg = true; next_any = true;
            }
        }
    }
    else if (shape == -1) { L = sky_color(in.d); done = true; }
    else
    {
        const int mat = sc.shapes[shape].material;
        if (MODE == RT_MODE_PREVIEW)
        {
            if (mat >= 0)
            {
                Ray unused = in;
                const Bounce b = material_eval(sc, mat, true, in, q.h, unused, s.rng);
                L = add3(L, mul3(b.att, q.h.color));
            }
            done = true;
        }
        else if (mat < 0) done = true;
        else
        {
            const Bounce b = material_eval(sc, mat, false, in, q.h, next, s.rng);
            if (rng_random(s.rng) <= q.h.alpha)
            {
                if (is_non_zero(b.att))
                {
                    const size_t k = (size_t)s.sp * pool.cap + id;
                    pool.st0[k] = make_float4(b.att.x, b.att.y, b.att.z, q.h.color.x);
                    pool.st1[k] = make_float4(q.h.color.y, q.h.color.z, b.emi.x, b.emi.y);
                    pool.st2[k] = b.emi.z;
                    s.sp++;
                    newseg = true;
                }
                else { L = add3(L, b.emi); done = true; }
            }
            else
            {
                // alpha pass-through (RayTracerScene.cpp:79-85): same direction, unattenuated
                next.o = add3(q.h.pos, mulf3(in.d, sc.bounce_offset)); next.d = in.d; next.dist = in.dist - q.h.dist;
                s.pass_mask |= 1u << s.sp;
                s.sp++;
                newseg = true;
            }
            if (newseg)
            {
                s.depth_left--;
                // RayTrace(ray, 0) returns black before any query (RayTracerScene.cpp:39-42)
                if (s.depth_left == 0) { newseg = false; done = true; }
            }
        }
    }
    if (done)
    {
        // fold the levels back in recursion order, emit the sample
        for (int k = s.sp - 1; k >= 0; k--)
        {
            if ((s.pass_mask >> k) & 1u) L = add3(V3(0, 0, 0), L);
            else
            {
                const size_t e = (size_t)k * pool.cap + id;
                const float4 s0 = pool.st0[e], s1 = pool.st1[e];
                const float s2 = pool.st2[e];
                const float3 att = V3(s0.x, s0.y, s0.z), col = V3(s0.w, s1.x, s1.y), emi = V3(s1.z, s1.w, s2);
                const float3 f = add3(V3(0, 0, 0), mul3(mul3(att, L), col));
                L = add3(f, emi);
            }
        }
        a.samples[(size_t)s.slot * ((size_t)a.width * a.height) + s.pixel] = make_float4(L.x, L.y, L.z, 0.0f);
        return false;
    }
    return newseg;
}

// ---- kernel A: generate -----------------------------------------------------------------------------------
// One thread per work item (sample, 8x4 pixel block, lane).  Camera ray (RayTracerProgram.cpp:133-165),
// then the shape list up to the first mesh whose bounds the ray enters.  A ray that ends there having
// hit nothing — most of them: they miss every bound and see the sky — is retired on the spot; the
// rest become paths: pool record + an entry in the round-0 queue, compacted per warp by ballot.
template <bool CULL, int MODE>
__global__ void __launch_bounds__(256, RT_GEN_BLOCKS)
rt_generate_kernel(const DevScene sc, const RenderArgs a, const WaveArgs w)
{
    Counters cnt = { 0, 0, 0, 0, 0, 0 };
    // One thread per PIXEL of the work list (8x4 block, lane), looping over the chunk's samples: the
    // pixel decode, the base direction and the pixel half of the RNG key are computed once.  Grid-stride,
    // whole warps together (the queue pushes want converged lanes).  A retry pass instead takes one
    // turned-away item per thread.
    const unsigned stride = gridDim.x * blockDim.x;
    const bool retry = w.retry_in != nullptr;
    const unsigned nthreads_needed = retry ? (*w.retry_in_count < w.item_count ? *w.retry_in_count : w.item_count)
                                           : a.num_blocks * 32u;
    if (nthreads_needed == 0) return;
    const unsigned rounded = (nthreads_needed + 31u) & ~31u;
    const int sample_count = retry ? 1 : a.num_samples;
    const size_t frame = (size_t)a.width * a.height;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < rounded; t += stride)
    {
        unsigned bl = 0, first_sample = 0;
        int lane_in_block = 0, px = -1, cx = 0, cy = 0;
        if (t < nthreads_needed)
        {
            if (retry)
            {
                const unsigned item = w.retry_in[t];
                const unsigned blk = item >> 5;
                first_sample = blk / a.num_blocks;
                bl = blk - first_sample * a.num_blocks;
                lane_in_block = (int)(item & 31u);
            }
            else { bl = t >> 5; lane_in_block = (int)(t & 31u); }
            px = block_pixel(a, bl, lane_in_block, cx, cy);
        }
        float base_dx = 0.0f, base_dy = 0.0f;
        camera_base(a.width, a.height, cx, cy, base_dx, base_dy);
        const uint32_t pixel_key = rt_rng_key_pixel(a.seed, (uint32_t)px);
        // when every shape has culling bounds, a ray that misses them all needs no query state at all
        const bool all_bounded = a.all_bounded != 0;
        float3 b0min = V3(0, 0, 0), b0max = V3(0, 0, 0);
        if (all_bounded && sc.num_shapes > 0) { b0min = ld3(sc.shapes[0].bounds_min); b0max = ld3(sc.shapes[0].bounds_max); }
        for (int k = 0; k < sample_count; k++)
        {
            const unsigned smp = first_sample + (unsigned)k;
            bool live = false;
            Query q;
            PathState s;
            int state = ST_IDLE;
            const Counters before = cnt;
            if (px >= 0)
            {
                s.pixel = px; s.slot = (int)smp;
                // spp is 4 (antialias) or 1
                const int pass = a.pass_begin + (a.antialias ? (int)(smp >> 2) : (int)smp);
                const int sub = a.antialias ? (int)(smp & 3u) : -1;
                s.rng.key = rt_rng_key_sample(pixel_key, (uint32_t)(a.antialias ? pass * 4 + sub : pass));
                s.rng.n = 0;
                const Ray cam = camera_ray_from_base(sc, a.width, base_dx, base_dy, MODE == RT_MODE_PRIMARY ? -1 : sub, s.rng);
                cnt.camera_rays++;
                s.depth_left = a.max_bounce; s.sp = 0; s.pass_mask = 0; s.light = 0; s.seg_dist = cam.dist;
                s.w_pos = s.w_nrm = s.w_surface = s.w_sum = V3(0, 0, 0);
                if ((MODE == RT_MODE_PATH || MODE == RT_MODE_PREVIEW) && a.max_bounce == 0)
                    a.samples[(size_t)s.slot * frame + s.pixel] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                else
                {
                    bool enters = !all_bounded;
                    if (all_bounded)
                    {
                        // FindIntersectionWithScene's bounds tests only (RayTracerScene.cpp:107-110)
                        const RayPre pre = ray_pre(cam);
                        float tlo, thi;
                        for (int si = 0; si < sc.num_shapes && !enters; si++)
                            enters = si == 0 ? slab_general(cam, pre, b0min, b0max, tlo, thi)
                                             : slab_general(cam, pre, ld3(sc.shapes[si].bounds_min), ld3(sc.shapes[si].bounds_max), tlo, thi);
                    }
                    if (!enters)
                    {
                        cnt.rays++;
                        cnt.node_visits += (unsigned)sc.num_shapes;
                        state = ST_SHADE;
                        q.hit_shape = -1;
                    }
                    else
                    {
                        query_begin(q, cam, false, cnt);
                        state = ST_SHAPES;
                        query_shapes<CULL>(sc, q, state, cnt);
                    }
                    live = true;
                    if (state == ST_SHADE && q.hit_shape == -1)
                    {
                        // nothing hit and no mesh to walk: RayTrace's miss branch (RayTracerScene.cpp:90-94)
                        if (MODE == RT_MODE_PRIMARY)
                        {
                            a.prim_ids[s.pixel] = make_int2(-1, -1);
                            a.prim_dist[s.pixel] = 0.0f;
                        }
                        else
                        {
                            const float3 L = sky_color(cam.d);
                            a.samples[(size_t)s.slot * frame + s.pixel] = make_float4(L.x, L.y, L.z, 0.0f);
                        }
                        live = false;
                    }
                }
            }
            // round 0's queue is the identity: path id == queue position, one atomic per warp
            unsigned spare = 0xffffffffu;
            const unsigned id = w.packets ? path_alloc_packet(w.counts + 0, live, spare) : path_alloc(w.counts + 0, live);
            const bool full = live && id >= w.pool.cap;
            if (spare < w.pool.cap)
            {
                // filler of a packet: an entry every kernel skips
                w.pool.cur[spare] = make_int4(0, -1, ST_IDLE, 0);
                w.queue[0][spare] = spare;
            }
            if (live && !full)
            {
                // a camera ray whose only remaining chance is this last mesh: if the walk finds nothing the
                // walk kernel itself retires it with the sky colour (no trip through the shade kernel)
                const bool sky_on_miss = (MODE == RT_MODE_PATH || MODE == RT_MODE_PREVIEW) && state == ST_TRAVERSE &&
                                         q.hit_shape == -1 && q.si == sc.num_shapes - 1;
                pool_store<MODE>(w.pool, id, q, state, s, sky_on_miss);
                w.queue[0][id] = id;
            }
            // pool full: the item is turned away untouched (its counters too) and generated again by the retry pass
            if (full) cnt = before;
            queue_push(w.retry_out, w.retry_out_count, full, ((smp * a.num_blocks + bl) << 5) | (unsigned)lane_in_block);
        }
    }
    flush_counters(cnt, a.counters, a.exact);
}

// ---- kernel T: walk ------------------------------------------------------------------------------------------
// Persistent warps.  A lane pops a path id from the round's queue, loads the ray, and walks the mesh its
// shape cursor points at — KdNode::TestRayIntersection (KdTree.cpp:128-195) on the pre-order,
// escape-threaded node array, see rt_device.cuh — to the end; then it stores (best leaf, position,
// shrunken Distance) and pops the next id, so a warp's 32 lanes stay on walks of their own length.
// Rounds of "node steps until the walking lanes hold a leaf, then those triangle tests together".
template <bool CULL>
__global__ void __launch_bounds__(256, RT_WALK_BLOCKS)
rt_walk_kernel(const DevScene sc, const RenderArgs a, const WaveArgs w, int round, int resumed)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    // resumed: the walks the packet kernel handed back (they continue at their cursor); else the round's queue
    const unsigned round_count = w.counts[round] < w.pool.cap ? w.counts[round] : w.pool.cap;
    const unsigned count = resumed ? w.scounts[round] : round_count;
    const unsigned* __restrict__ queue = resumed ? w.slowq : w.queue[round & 1];
    unsigned* head = resumed ? w.sheads + round : w.heads + round;
    if (count == 0 || round_count < w.small_round) return;   // empty, or thin: the long-walk kernel takes all of it
    const unsigned long_limit = round_count < w.thin_count ? w.thin_limit : w.long_limit;
    Counters cnt = { 0, 0, 0, 0, 0, 0 };
    unsigned win_pos = 0, win_end = 0;
    bool exhausted = count == 0;

    bool have = false;
    unsigned id = 0;
    Ray r; r.o = V3(0, 0, 0); r.d = V3(0, 0, 1); r.dist = 0.0f;
    RayPre pre = ray_pre(r);
    bool any = false, weird = false, wide = false, sky_on_miss = false;
    float3 pad3 = V3(0, 0, 0);
    float growth = 0.0f;
    const float4* __restrict__ nodes = nullptr;
    const float4* __restrict__ tris = nullptr;
    int n = 0, i = 0, best = -1;
    float3 bpos = V3(0, 0, 0);
    unsigned nodes_seen = 0, tris_seen = 0;
    unsigned walk_start = 0, walk_max = 0;

    for (;;)
    {
        // ---- refill: lanes without a walk pop ids (ballot -> rank -> window item) ----------------------
        for (;;)
        {
            const unsigned idle = __ballot_sync(RT_FULL_MASK, !have);
            if (idle == 0) break;
            if (win_pos >= win_end)
            {
                if (exhausted) break;
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(head, w.window);
                base = __shfl_sync(RT_FULL_MASK, base, 0);
                if (base >= count) { exhausted = true; break; }
                win_pos = base;
                win_end = count - base < w.window ? count : base + w.window;
            }
            const unsigned item = win_pos + (unsigned)__popc(idle & lt_mask);
            if (!have && item < win_end)
            {
                id = queue[item];
                const int4 cur = w.pool.cur[id];
                if ((cur.z & 255) == ST_TRAVERSE)
                {
                    const float4 ro = w.pool.ro[id], rd = w.pool.rd[id];
                    r.o = xyz(ro); r.dist = ro.w; r.d = xyz(rd);
                    pre = ray_pre(r);
                    weird = !(pre.ex && pre.ey && pre.ez && finite3(r.o) && finite3(r.d));
                    any = (cur.z & 256) != 0;
                    sky_on_miss = (cur.z & 512) != 0;
                    const DevMesh* m = sc.meshes + sc.shapes[cur.x].mesh;
                    nodes = m->nodes; tris = m->tris; n = m->num_nodes;
                    if (CULL)
                    {
                        pre.cull_pad = cull_pad_for(r, pre, m->cull_scale);
                        growth = cull_growth(r, m->cull_scale);
                        // a disabled axis is not constrained (inv = 0 there: its interval is [-pad, pad] around 0)
                        pad3.x = pre.ex ? growth * fabsf(pre.inv.x) + growth : FLT_MAX;
                        pad3.y = pre.ey ? growth * fabsf(pre.inv.y) + growth : FLT_MAX;
                        pad3.z = pre.ez ? growth * fabsf(pre.inv.z) + growth : FLT_MAX;
                        const bool finite = finite3(r.o) && finite3(r.d) && pre.cull_pad < FLT_MAX;
                        wide = finite && (pre.cull_pad > 4096.0f * growth || !(pre.ex && pre.ey && pre.ez));    // |d| < 2.4e-4 on some axis
                    }
                    i = 0; best = -1; bpos = V3(0, 0, 0);
                    if (resumed)
                    {
                        const float4 bp = w.pool.bp[id];
                        i = __float_as_int(bp.w); best = cur.y; bpos = xyz(bp);
                    }
                    walk_start = nodes_seen;
                    have = true;
                }
            }
            const unsigned taken = win_pos + (unsigned)__popc(idle);
            win_pos = taken < win_end ? taken : win_end;
        }
        if (__ballot_sync(RT_FULL_MASK, have) == 0) break;

        // ---- walk until too few lanes are left walking ---------------------------------------------------
        const bool verbatim = __any_sync(RT_FULL_MASK, have && weird);
        const bool widewarp = CULL && __any_sync(RT_FULL_MASK, have && wide);
        const int min_lanes = exhausted ? 1 : w.min_lanes;
        for (;;)
        {
            // Node phase.  The leaves a walk reaches do not depend on the hits it has accepted (the
            // reference's box test is a line test, KdTree.cpp:131; the culling above only drops leaves
            // that would be rejected anyway), so a lane that has found a leaf keeps walking to its NEXT
            // leaf while its neighbours are still looking for their first: up to two leaves are held and
            // then tested in walk order.  Fewer lanes wait, and the triangle phase runs fuller.
            int leaf[RT_LEAF_SLOTS];
#pragma unroll
            for (int k = 0; k < RT_LEAF_SLOTS; k++) leaf[k] = -1;
            for (;;)
            {
                const unsigned stepping = __ballot_sync(RT_FULL_MASK, have && leaf[RT_LEAF_SLOTS - 1] < 0 && i < n);
                if (stepping == 0) break;
                if (w.leaf_wait > 0 && __popc(stepping) < w.leaf_wait &&
                    __ballot_sync(RT_FULL_MASK, leaf[0] >= 0) != 0) break;
                // two node steps per vote: the loop control above costs as much as half a step
#pragma unroll
                for (int u = 0; u < 2; u++)
                {
                    if (have && leaf[RT_LEAF_SLOTS - 1] < 0 && i < n)
                    {
                        const float4 na = __ldg(nodes + 2 * (size_t)i);
                        const float4 nb = __ldg(nodes + 2 * (size_t)i + 1);
                        const int escape = __float_as_int(na.w);
                        const int tri = __float_as_int(nb.w);
                        nodes_seen++;
                        float tlo, thi;
                        bool enter = verbatim ? slab_general(r, pre, xyz(na), xyz(nb), tlo, thi)
                                              : slab_fast(r, pre, xyz(na), xyz(nb), tlo, thi);
                        if (CULL)
                        {
                            if (widewarp && wide) enter = enter && !cull_axes(r, pre, pad3, xyz(na), xyz(nb), r.dist * 1.0078125f + growth, growth);
                            else enter = enter && !(thi < -pre.cull_pad) && !(tlo > r.dist * 1.0078125f + pre.cull_pad);
                        }
                        if (!enter) i = escape;
                        else if (tri < 0) i = i + 1;
                        else
                        {
                            bool placed = false;
#pragma unroll
                            for (int k = 0; k < RT_LEAF_SLOTS; k++)
                                if (!placed && leaf[k] < 0) { leaf[k] = tri; placed = true; }
                            i = escape;
                        }
                    }
                }
            }
            // Triangle phase: the held leaves, in walk order
#pragma unroll
            for (int k = 0; k < RT_LEAF_SLOTS; k++)
            {
                const int lf = leaf[k];
                if (__ballot_sync(RT_FULL_MASK, lf >= 0) == 0) break;
                if (lf >= 0)
                {
                    const float4 t0 = __ldg(tris + 4 * (size_t)lf);
                    const float4 t1 = __ldg(tris + 4 * (size_t)lf + 1);
                    const float4 t2 = __ldg(tris + 4 * (size_t)lf + 2);
                    const float4 t3 = __ldg(tris + 4 * (size_t)lf + 3);
                    tris_seen++;
                    float3 hp; float hd;
                    if (triangle_test(r, xyz(t0), xyz(t1), xyz(t2), xyz(t3), hp, hd))
                    {
                        r.dist = hd;
                        bpos = hp;
                        best = lf;
                        if (CULL && any)
                        {
                            i = n;
#pragma unroll
                            for (int j = 0; j < RT_LEAF_SLOTS; j++) leaf[j] = -1;
                        }
                    }
                }
            }
            if (have && i >= n)
            {
                // walk complete: hand the result to the shade kernel
                int* cur = reinterpret_cast<int*>(w.pool.cur + id);
                if (best < 0 && sky_on_miss)
                {
                    // RayTrace's miss branch (RayTracerScene.cpp:90-94) for a camera ray: nothing to fold
                    const int4 pa = w.pool.pa[id];
                    const float3 L = sky_color(r.d);
                    a.samples[(size_t)pa.y * ((size_t)a.width * a.height) + pa.x] = make_float4(L.x, L.y, L.z, 0.0f);
                    cur[2] = ST_IDLE;           // the shade kernel skips it
                }
                else
                {
                    w.pool.ro[id].w = r.dist;
                    cur[1] = best;
                    cur[2] = ST_MESHDONE | (any ? 256 : 0);
                    w.pool.bp[id] = make_float4(bpos.x, bpos.y, bpos.z, 0.0f);
                }
                walk_max = max(walk_max, nodes_seen - walk_start);
                have = false;
            }
            else if (have && nodes_seen - walk_start > long_limit)
            {
                // A walk this long would hold the round: park it (cursor, best hit so far) for the
                // long-walk kernel, which spends a whole warp on it.  No leaf is pending here.
                w.pool.ro[id].w = r.dist;
                reinterpret_cast<int*>(w.pool.cur + id)[1] = best;
                w.pool.bp[id] = make_float4(bpos.x, bpos.y, bpos.z, __int_as_float(i));
                w.longq[atomicAdd(w.lcounts + round, 1u)] = id;
                have = false;
            }
            if (__popc(__ballot_sync(RT_FULL_MASK, have)) < min_lanes) break;
        }
    }
    cnt.node_visits = nodes_seen; cnt.tri_visits = tris_seen;
    flush_counters(cnt, a.counters, a.exact);
    // longest single walk of the batch (tooling: rt_gpu_debug_rounds)
    for (int o = 16; o > 0; o >>= 1) walk_max = max(walk_max, __shfl_xor_sync(RT_FULL_MASK, walk_max, o));
    if (lane == 0 && walk_max > 0) atomicMax(w.counts + RT_MAX_ROUNDS, walk_max);
}

// ---- kernel P: packet walk (coherent rounds) ---------------------------------------------------------------
// Round 0 holds camera rays in generation order: 32 consecutive entries come from one 8x4-pixel block, so
// their walks visit almost the same nodes.  Here a warp walks its 32 rays TOGETHER: one cursor per lane as
// before, but each step the warp visits the smallest cursor c of its lanes — the array is in visiting
// order, so every lane still meets exactly its own nodes, in its own order — loads node c ONCE (uniform
// address: one transaction instead of up to 32), and the lanes standing at c test it.  A leaf is tested on
// the spot by the lanes that entered its box (same triangle for all of them).  Per ray the tests, their
// order and their results are those of rt_walk_kernel; only the schedule differs.  The warp needs |union of
// the lanes' node sets| steps instead of sum/active-lanes, without divergence and with far fewer memory
// requests.  Lanes of other meshes wait their turn (one group per mesh); a packet that exceeds the step
// budget parks its unfinished lanes for the long-walk kernel.
template <bool CULL>
__global__ void __launch_bounds__(256, RT_WALK_BLOCKS)
rt_walk_packet_kernel(const DevScene sc, const RenderArgs a, const WaveArgs w, int round)
{
    const int lane = threadIdx.x & 31;
    const unsigned count = w.counts[round] < w.pool.cap ? w.counts[round] : w.pool.cap;
    const unsigned* __restrict__ queue = w.queue[round & 1];
    unsigned* head = w.heads + round;
    if (count == 0 || count < w.small_round) return;   // empty, or thin: the long-walk kernel takes all of it
    const unsigned step_limit = count < w.thin_count ? w.thin_limit : w.long_limit;
    Counters cnt = { 0, 0, 0, 0, 0, 0 };
    unsigned nodes_seen = 0, tris_seen = 0;
    for (;;)
    {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(head, 32u);
        base = __shfl_sync(RT_FULL_MASK, base, 0);
        if (base >= count) break;
        const unsigned item = base + (unsigned)lane;
        unsigned id = 0;
        bool active = false;
        int shape = -1;
        bool any = false, sky_on_miss = false;
        Ray r; r.o = V3(0, 0, 0); r.d = V3(0, 0, 1); r.dist = 0.0f;
        if (item < count)
        {
            id = queue[item];
            const int4 cur = w.pool.cur[id];
            if ((cur.z & 255) == ST_TRAVERSE)
            {
                const float4 ro = w.pool.ro[id], rd = w.pool.rd[id];
                r.o = xyz(ro); r.dist = ro.w; r.d = xyz(rd);
                any = (cur.z & 256) != 0;
                sky_on_miss = (cur.z & 512) != 0;
                shape = cur.x;
                active = true;
            }
        }
        RayPre pre = ray_pre(r);
        const bool weird = !(pre.ex && pre.ey && pre.ez && finite3(r.o) && finite3(r.d));
        unsigned todo = __ballot_sync(RT_FULL_MASK, active);
        while (todo != 0)
        {
            // one group per mesh (a scene with one mesh: one group)
            const int leader = __ffs((int)todo) - 1;
            const int gshape = __shfl_sync(RT_FULL_MASK, shape, leader);
            const bool mine = active && shape == gshape;
            todo &= ~__ballot_sync(RT_FULL_MASK, mine);
            const DevMesh* m = sc.meshes + sc.shapes[gshape].mesh;
            const float4* __restrict__ nodes = m->nodes;
            const float4* __restrict__ tris = m->tris;
            const int n = m->num_nodes;
            float3 pad3 = V3(0, 0, 0);
            float growth = 0.0f;
            bool wide = false;
            if (CULL)
            {
                pre.cull_pad = cull_pad_for(r, pre, m->cull_scale);
                growth = cull_growth(r, m->cull_scale);
                pad3.x = pre.ex ? growth * fabsf(pre.inv.x) + growth : FLT_MAX;
                pad3.y = pre.ey ? growth * fabsf(pre.inv.y) + growth : FLT_MAX;
                pad3.z = pre.ez ? growth * fabsf(pre.inv.z) + growth : FLT_MAX;
                const bool finite = finite3(r.o) && finite3(r.d) && pre.cull_pad < FLT_MAX;
                wide = finite && (pre.cull_pad > 4096.0f * growth || !(pre.ex && pre.ey && pre.ez));
            }
            const bool verbatim = __any_sync(RT_FULL_MASK, mine && weird);
            const bool widewarp = CULL && __any_sync(RT_FULL_MASK, mine && wide);
            int best = -1;
            float3 bpos = V3(0, 0, 0);
            unsigned cursor = mine ? 0u : 0xffffffffu;
            const unsigned group_lanes = (unsigned)__popc(__ballot_sync(RT_FULL_MASK, mine));
            unsigned steps = 0;
            bool parked = false, done = false;
            while (!done)
            {
                // a window of packet_probe steps, then the packet is judged
                const unsigned seen_before = nodes_seen;
                for (unsigned k = 0; k < w.packet_probe; k++)
                {
                    const unsigned c = __reduce_min_sync(RT_FULL_MASK, cursor);
                    if (c >= (unsigned)n) { done = true; break; }
                    const float4 na = __ldg(nodes + 2 * (size_t)c);
                    const float4 nb = __ldg(nodes + 2 * (size_t)c + 1);
                    const int escape = __float_as_int(na.w);
                    const int tri = __float_as_int(nb.w);
                    bool enter = false;
                    if (cursor == c)
                    {
                        nodes_seen++;
                        float tlo, thi;
                        enter = verbatim ? slab_general(r, pre, xyz(na), xyz(nb), tlo, thi)
                                         : slab_fast(r, pre, xyz(na), xyz(nb), tlo, thi);
                        if (CULL)
                        {
                            if (widewarp && wide) enter = enter && !cull_axes(r, pre, pad3, xyz(na), xyz(nb), r.dist * 1.0078125f + growth, growth);
                            else enter = enter && !(thi < -pre.cull_pad) && !(tlo > r.dist * 1.0078125f + pre.cull_pad);
                        }
                        cursor = (enter && tri < 0) ? c + 1u : (unsigned)escape;
                    }
                    if (tri >= 0 && __any_sync(RT_FULL_MASK, enter))
                    {
                        const float4 t0 = __ldg(tris + 4 * (size_t)tri);
                        const float4 t1 = __ldg(tris + 4 * (size_t)tri + 1);
                        const float4 t2 = __ldg(tris + 4 * (size_t)tri + 2);
                        const float4 t3 = __ldg(tris + 4 * (size_t)tri + 3);
                        if (enter)
                        {
                            tris_seen++;
                            float3 hp; float hd;
                            if (triangle_test(r, xyz(t0), xyz(t1), xyz(t2), xyz(t3), hp, hd))
                            {
                                r.dist = hd; bpos = hp; best = tri;
                                if (CULL && any) cursor = (unsigned)n;
                            }
                        }
                    }
                }
                if (done) break;
                steps += w.packet_probe;
                // Not a coherent packet after all (rays of one pixel block spread over many small triangles): the
                // steps of the last window were mostly other lanes' nodes (the top of the tree is common to all
                // rays; coherence shows, or ends, further down).  Its unfinished lanes go on one by one in
                // rt_walk_kernel, from where they are.  Likewise a packet that outlasts the step budget: those
                // lanes go on alone in the long-walk kernel.
                const unsigned tests = __reduce_add_sync(RT_FULL_MASK, nodes_seen - seen_before);
                const bool incoherent = tests * 32u < w.packet_probe * w.packet_min_lanes * group_lanes;
                if (incoherent || steps > step_limit)
                {
                    if (mine && cursor < (unsigned)n)
                    {
                        w.pool.ro[id].w = r.dist;
                        reinterpret_cast<int*>(w.pool.cur + id)[1] = best;
                        w.pool.bp[id] = make_float4(bpos.x, bpos.y, bpos.z, __int_as_float((int)cursor));
                        if (incoherent) w.slowq[atomicAdd(w.scounts + round, 1u)] = id;
                        else w.longq[atomicAdd(w.lcounts + round, 1u)] = id;
                        parked = true;
                    }
                    break;
                }
            }
            if (mine && !parked)
            {
                // walk complete: hand the result to the shade kernel
                int* curw = reinterpret_cast<int*>(w.pool.cur + id);
                if (best < 0 && sky_on_miss)
                {
                    const int4 pa = w.pool.pa[id];
                    const float3 L = sky_color(r.d);
                    a.samples[(size_t)pa.y * ((size_t)a.width * a.height) + pa.x] = make_float4(L.x, L.y, L.z, 0.0f);
                    curw[2] = ST_IDLE;
                }
                else
                {
                    w.pool.ro[id].w = r.dist;
                    curw[1] = best;
                    curw[2] = ST_MESHDONE | (any ? 256 : 0);
                    w.pool.bp[id] = make_float4(bpos.x, bpos.y, bpos.z, 0.0f);
                }
            }
        }
    }
    cnt.node_visits = nodes_seen; cnt.tri_visits = tris_seen;
    flush_counters(cnt, a.counters, a.exact);
}

// The device copy of an inner node keeps its right child in the `tri` field (-2 - index; any negative value
// still reads "inner node" to the sequential walks): right child = escape of the left child (k + 1).
__global__ void rt_patch_right_child(rt_bvh_node* nodes, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && nodes[k].tri < 0) nodes[k].tri = -2 - (k + 1 < n ? nodes[k + 1].escape : n);
}

// ---- kernel L: long walks ------------------------------------------------------------------------------------
// One WARP per walk the walk kernel parked (or per entry of a thin round).  Two facts make a walk parallel
// without changing a bit of its result:
//   * which nodes a walk visits does not depend on what it hits (the reference's box test is a line test,
//     KdTree.cpp:131; culling with the Distance of the moment the walk was parked only drops leaves that
//     would be rejected anyway, Distance only shrinks), and
//   * the array is in visiting order, so "in the reference's order" == "by ascending leaf slot".
// So the warp first expands the rest of the tree as a FRONTIER — 32 pending nodes per step from a stack in
// shared memory, each lane one slab test, children pushed back (an inner node's `tri` field holds its right
// child, patched at upload) — collecting the leaves reached; then sorts those leaf slots and REPLAYS the
// triangle tests one after the other (Distance shrinks exactly as in the reference; every lane computes the
// same test on shuffled operands).  A 600-node walk is ~40 memory round trips instead of 600.
// The walk resumes at a cursor: the rest of the traversal is the cursor's subtree, then its escape's, ...;
// that chain is followed by lane 31, one link per step.  A frontier or leaf list that outgrows its shared
// memory falls back to the sequential window replay below (nothing has been written by then).
// The G lanes of a GROUP share one walk (G = 32, 16 or 8: a warp runs 1, 2 or 4 walks as independent
// mini-warps, every collective masked to the group).  A walk's frontier is rarely 32 nodes wide, and the
// kernel is bound by memory round trips, so narrower groups keep more walks in flight per SM.
#ifndef RT_LONG_GROUP
#define RT_LONG_GROUP 32
#endif
#define RT_FW_INTS_PER_LANE 16              // stack and leaf list hold 16 x G entries each (32 KB per block together)

// Sequential fallback: the next G nodes i..i+G-1 tested at once, the cursor replayed through the results.
template <bool CULL, int G>
__device__ __forceinline__ void longwalk_windows(const float4* __restrict__ nodes, const float4* __restrict__ tris, int n, int gl, unsigned gmask,
                                                 Ray& r, const RayPre& pre, float3 pad3, float growth, bool cull, bool any,
                                                 int& i, int& best, float3& bpos, unsigned& nodes_seen, unsigned& tris_seen)
{
    while (i < n)
    {
        const int node = i + gl;
        bool enter = false;
        int escape = n, tri = -1;
        if (node < n)
        {
            const float4 na = __ldg(nodes + 2 * (size_t)node);
            const float4 nb = __ldg(nodes + 2 * (size_t)node + 1);
            escape = __float_as_int(na.w); tri = __float_as_int(nb.w);
            float tlo, thi;
            enter = slab_general(r, pre, xyz(na), xyz(nb), tlo, thi);
            if (CULL && cull) enter = enter && !cull_axes(r, pre, pad3, xyz(na), xyz(nb), r.dist * 1.0078125f + growth, growth);
        }
        const int wend = i + G < n ? i + G : n;
        int c = i;
        while (c < wend)
        {
            const int src = c - i;
            const bool en = __shfl_sync(gmask, (int)enter, src, G) != 0;
            const int es = __shfl_sync(gmask, escape, src, G);
            const int tr = __shfl_sync(gmask, tri, src, G);
            nodes_seen++;
            if (!en) c = es;
            else if (tr < 0) c = c + 1;
            else
            {
                const float4 t0 = __ldg(tris + 4 * (size_t)tr);
                const float4 t1 = __ldg(tris + 4 * (size_t)tr + 1);
                const float4 t2 = __ldg(tris + 4 * (size_t)tr + 2);
                const float4 t3 = __ldg(tris + 4 * (size_t)tr + 3);
                tris_seen++;
                float3 hp; float hd;
                c = es;
                if (triangle_test(r, xyz(t0), xyz(t1), xyz(t2), xyz(t3), hp, hd))
                {
                    r.dist = hd; bpos = hp; best = tr;
                    if (CULL && any) c = n;
                }
            }
        }
        i = c;
    }
}

template <bool CULL, int G>
__global__ void __launch_bounds__(256, RT_LONG_BLOCKS)
rt_longwalk_kernel(const DevScene sc, const RenderArgs a, const WaveArgs w, int round)
{
    constexpr int CAP = RT_FW_INTS_PER_LANE * G;
    __shared__ int s_stack[256 * RT_FW_INTS_PER_LANE];
    __shared__ int s_leaf[256 * RT_FW_INTS_PER_LANE];
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);                  // lane within the group
    const int gshift = lane - gl;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << gshift);
    const unsigned lt_mask = (1u << gl) - 1u;
    int* stk = s_stack + (threadIdx.x / G) * CAP;
    int* lst = s_leaf + (threadIdx.x / G) * CAP;
    // a thin round (the walk kernel skipped it) is taken whole from the round's queue; otherwise only the
    // walks that kernel parked
    const unsigned round_count = w.counts[round] < w.pool.cap ? w.counts[round] : w.pool.cap;
    const bool whole = round_count < w.small_round;
    const unsigned count = whole ? round_count : w.lcounts[round];
    const unsigned* __restrict__ src = whole ? w.queue[round & 1] : w.longq;
    if (count == 0) return;
    Counters cnt = { 0, 0, 0, 0, 0, 0 };
    unsigned nodes_seen = 0, tris_seen = 0;
    for (;;)
    {
        unsigned e = 0;
        if (gl == 0) e = atomicAdd(w.lheads + round, 1u);
        e = __shfl_sync(gmask, e, 0, G);
        if (e >= count) break;
        const unsigned id = src[e];
        const int4 cur = w.pool.cur[id];
        if ((cur.z & 255) != ST_TRAVERSE) continue;      // (round 0 may hold entries that need no walk)
        const float4 ro = w.pool.ro[id], rd = w.pool.rd[id], bp = w.pool.bp[id];
        Ray r; r.o = xyz(ro); r.dist = ro.w; r.d = xyz(rd);
        RayPre pre = ray_pre(r);
        const bool any = (cur.z & 256) != 0, sky_on_miss = (cur.z & 512) != 0;
        const DevMesh* m = sc.meshes + sc.shapes[cur.x].mesh;
        const float4* __restrict__ nodes = m->nodes;
        const float4* __restrict__ tris = m->tris;
        const int n = m->num_nodes;
        float3 pad3 = V3(FLT_MAX, FLT_MAX, FLT_MAX);
        float growth = 0.0f;
        bool cull = false;
        if (CULL)
        {
            growth = cull_growth(r, m->cull_scale);
            pad3.x = pre.ex ? growth * fabsf(pre.inv.x) + growth : FLT_MAX;
            pad3.y = pre.ey ? growth * fabsf(pre.inv.y) + growth : FLT_MAX;
            pad3.z = pre.ez ? growth * fabsf(pre.inv.z) + growth : FLT_MAX;
            cull = finite3(r.o) && finite3(r.d) && growth < FLT_MAX;
        }
        int i = __float_as_int(bp.w), best = cur.y;
        float3 bpos = xyz(bp);
        const unsigned walk_start_seen = nodes_seen;

        // ---- frontier: the leaves the rest of the walk reaches -------------------------------------------
        const float reach = r.dist * 1.0078125f + growth;        // Distance at parking time: only shrinks from here
        int size = 0, nleaf = 0, chain = i;
        unsigned frontier_nodes = 0;
        bool overflow = false;
        while (size > 0 || chain < n)
        {
            const bool has_chain = chain < n;
            const int room = CAP - size;
            if (room < 4) { overflow = true; break; }
            // a popped node nets at most one entry (two children pushed), the chain node two
            int k = size < G - 1 ? size : G - 1;
            if (k + 2 > room) k = room - 2;
            int node = -1;
            if (gl < k) node = stk[size - 1 - gl];
            else if (gl == G - 1 && has_chain) node = chain;
            if (node >= n) node = -1;
            size -= k;
            __syncwarp(gmask);
            bool enter = false;
            int escape = n, tri = -1;
            if (node >= 0)
            {
                const float4 na = __ldg(nodes + 2 * (size_t)node);
                const float4 nb = __ldg(nodes + 2 * (size_t)node + 1);
                escape = __float_as_int(na.w); tri = __float_as_int(nb.w);
                float tlo, thi;
                enter = slab_general(r, pre, xyz(na), xyz(nb), tlo, thi);
                if (CULL && cull) enter = enter && !cull_axes(r, pre, pad3, xyz(na), xyz(nb), reach, growth);
            }
            frontier_nodes += (unsigned)__popc(__ballot_sync(gmask, node >= 0));
            if (has_chain) chain = __shfl_sync(gmask, escape, G - 1, G);
            const unsigned leaves = __ballot_sync(gmask, enter && tri >= 0) >> gshift;
            if (leaves != 0)
            {
                if (nleaf + __popc(leaves) > CAP) { overflow = true; break; }
                if (enter && tri >= 0) lst[nleaf + __popc(leaves & lt_mask)] = tri;
                nleaf += __popc(leaves);
            }
            const unsigned inner = __ballot_sync(gmask, enter && tri < 0) >> gshift;
            if (enter && tri < 0)
            {
                const int right = -2 - tri;
                const int pos = size + 2 * __popc(inner & lt_mask);
                // (right child on the bottom: the left subtree is expanded first, which keeps the stack short)
                stk[pos] = right < escape ? right : n;
                stk[pos + 1] = node + 1 < escape ? node + 1 : n;
            }
            size += 2 * __popc(inner);
            __syncwarp(gmask);
            // entries that name no node (single-child nodes of a foreign tree) are dropped when popped
            while (size > 0 && stk[size - 1] >= n) size--;
        }
        if (overflow)
        {
            __syncwarp(gmask);
            longwalk_windows<CULL, G>(nodes, tris, n, gl, gmask, r, pre, pad3, growth, cull, any, i, best, bpos, nodes_seen, tris_seen);
        }
        else
        {
            nodes_seen += frontier_nodes;
            // ---- replay: sort the leaf slots (== visiting order), then the triangle tests in that order ----
            __syncwarp(gmask);
            for (int x = gl; x < nleaf; x += G)
            {
                const int v = lst[x];
                int rank = 0;
                for (int j = 0; j < nleaf; j++) rank += lst[j] < v ? 1 : 0;
                stk[rank] = v;
            }
            __syncwarp(gmask);
            bool stop = false;
            for (int base = 0; base < nleaf && !stop; base += G)
            {
                const int lf = base + gl < nleaf ? stk[base + gl] : -1;
                float4 t0 = make_float4(0, 0, 0, 0), t1 = t0, t2 = t0, t3 = t0;
                if (lf >= 0)
                {
                    t0 = __ldg(tris + 4 * (size_t)lf);
                    t1 = __ldg(tris + 4 * (size_t)lf + 1);
                    t2 = __ldg(tris + 4 * (size_t)lf + 2);
                    t3 = __ldg(tris + 4 * (size_t)lf + 3);
                }
                const int batch = nleaf - base < G ? nleaf - base : G;
                for (int k = 0; k < batch; k++)
                {
                    const float3 p0 = V3(__shfl_sync(gmask, t0.x, k, G), __shfl_sync(gmask, t0.y, k, G), __shfl_sync(gmask, t0.z, k, G));
                    const float3 p1 = V3(__shfl_sync(gmask, t1.x, k, G), __shfl_sync(gmask, t1.y, k, G), __shfl_sync(gmask, t1.z, k, G));
                    const float3 p2 = V3(__shfl_sync(gmask, t2.x, k, G), __shfl_sync(gmask, t2.y, k, G), __shfl_sync(gmask, t2.z, k, G));
                    const float3 nn = V3(__shfl_sync(gmask, t3.x, k, G), __shfl_sync(gmask, t3.y, k, G), __shfl_sync(gmask, t3.z, k, G));
                    const int slot = __shfl_sync(gmask, lf, k, G);
                    tris_seen++;
                    float3 hp; float hd;
                    if (triangle_test(r, p0, p1, p2, nn, hp, hd))
                    {
                        r.dist = hd; bpos = hp; best = slot;
                        if (CULL && any) { stop = true; break; }
                    }
                }
            }
            __syncwarp(gmask);
        }
        if (gl == 0)
        {
            atomicMax(w.counts + RT_MAX_ROUNDS, nodes_seen - walk_start_seen);      // tooling: longest walk
            int* curw = reinterpret_cast<int*>(w.pool.cur + id);
            if (best < 0 && sky_on_miss)
            {
                const int4 pa = w.pool.pa[id];
                const float3 L = sky_color(r.d);
                a.samples[(size_t)pa.y * ((size_t)a.width * a.height) + pa.x] = make_float4(L.x, L.y, L.z, 0.0f);
                curw[2] = ST_IDLE;
            }
            else
            {
                w.pool.ro[id].w = r.dist;
                curw[1] = best;
                curw[2] = ST_MESHDONE | (any ? 256 : 0);
                w.pool.bp[id] = make_float4(bpos.x, bpos.y, bpos.z, 0.0f);
            }
        }
    }
    __syncwarp();
    // every lane of a group saw the same walk: count it once
    if (gl == 0) { cnt.node_visits = nodes_seen; cnt.tri_visits = tris_seen; }
    flush_counters(cnt, a.counters, a.exact);
}

// ---- kernel S: shade ----------------------------------------------------------------------------------------
// One thread per entry of the round's queue (grid-stride).  Finishes the mesh hit (attributes, texture),
// runs the rest of the shape list; a query that reaches another mesh goes to the next round's queue,
// a completed query is shaded — material bounce, alpha test, light loop — and either ends the path
// (fold + sample) or begins the next segment, whose shape list runs here as well.
template <bool CULL, int MODE>
__global__ void __launch_bounds__(256, RT_SHADE_BLOCKS)
rt_shade_kernel(const DevScene sc, const RenderArgs a, const WaveArgs w, int round)
{
    const unsigned count = w.counts[round] < w.pool.cap ? w.counts[round] : w.pool.cap;
    const unsigned* __restrict__ queue = w.queue[round & 1];
    unsigned* next_queue = w.queue[(round + 1) & 1];
    unsigned* next_count = w.counts + round + 1;
    if (count == 0) return;                     // an empty round (or retry pass) costs a launch, nothing more
    Counters cnt = { 0, 0, 0, 0, 0, 0 };
    const unsigned stride = gridDim.x * blockDim.x;
    // whole warps iterate together so that the queue pushes see converged lanes
    const unsigned first = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned rounded = (count + 31u) & ~31u;
    for (unsigned e = first; e < rounded; e += stride)
    {
        bool push = false;
        unsigned id = 0;
        if (e < count)
        {
            id = queue[e];
        }
        // entries the walk kernel already retired (camera rays that saw the sky)
        if (e < count && (w.pool.cur[id].z & 255) != ST_IDLE)
        {
            Query q; PathState s; int state;
            pool_load<MODE>(w.pool, id, q, state, s);
            query_mesh_done(sc, q, state, cnt);
            for (;;)
            {
                query_shapes<CULL>(sc, q, state, cnt);
                if (state == ST_TRAVERSE) { push = true; break; }
                Ray next; next.o = V3(0, 0, 0); next.d = V3(0, 0, 0); next.dist = 0.0f;
                bool next_any = false;
                if (!shade_query<MODE>(sc, a, w.pool, id, q, s, next, next_any)) break;
                query_begin(q, next, next_any, cnt);
                s.seg_dist = next.dist;
                state = ST_SHAPES;
            }
            if (push) pool_store<MODE>(w.pool, id, q, state, s);
        }
        queue_push(next_queue, next_count, push, id);
    }
    flush_counters(cnt, a.counters, a.exact);
}

// ---- kernel F: finish ------------------------------------------------------------------------------------------
// After a few rounds only a percent of the paths is still alive, and a round costs the latency of its
// longest walk whatever its size.  This kernel takes everything that is left and runs each path to
// its end in ONE launch: a lane pops a path, then alternates walk (the resumable, warp-collective
// query_traverse of rt_device.cuh) and shade until the path ends, and pops the next.  Lane efficiency
// is poor and does not matter here; the critical path drops from (rounds left) x (longest walk) to
// one path's length.
template <bool CULL, int MODE>
__global__ void __launch_bounds__(128)
rt_finish_kernel(const DevScene sc, const RenderArgs a, const WaveArgs w, int round)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned count = w.counts[round] < w.pool.cap ? w.counts[round] : w.pool.cap;
    const unsigned* __restrict__ queue = w.queue[round & 1];
    unsigned* head = w.heads + round;
    Counters cnt = { 0, 0, 0, 0, 0, 0 };
    unsigned win_pos = 0, win_end = 0;
    bool exhausted = count == 0;

    int state = ST_IDLE;
    unsigned id = 0;
    Query q;
    q.r.o = V3(0, 0, 0); q.r.d = V3(0, 0, 1); q.r.dist = 0.0f; q.pre = ray_pre(q.r); q.weird = false;
    q.h.pos = V3(0, 0, 0); q.h.nrm = V3(0, 0, 0); q.h.dist = 0.0f; q.h.color = V3(1, 1, 1); q.h.alpha = 1.0f;
    q.bpos = V3(0, 0, 0); q.si = 0; q.node = 0; q.best = -1; q.hit_shape = -1; q.tri = -1; q.any = false;
    PathState s;
    s.pixel = 0; s.slot = 0; s.rng.key = 0; s.rng.n = 0; s.depth_left = 0; s.sp = 0; s.pass_mask = 0; s.light = 0; s.seg_dist = 0.0f;
    s.w_pos = s.w_nrm = s.w_surface = s.w_sum = V3(0, 0, 0);

    for (;;)
    {
        for (;;)
        {
            const unsigned idle = __ballot_sync(RT_FULL_MASK, state == ST_IDLE);
            if (idle == 0) break;
            if (win_pos >= win_end)
            {
                if (exhausted) break;
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(head, 32u);
                base = __shfl_sync(RT_FULL_MASK, base, 0);
                if (base >= count) { exhausted = true; break; }
                win_pos = base;
                win_end = count - base < 32u ? count : base + 32u;
            }
            const unsigned item = win_pos + (unsigned)__popc(idle & lt_mask);
            if (state == ST_IDLE && item < win_end)
            {
                id = queue[item];
                pool_load<MODE>(w.pool, id, q, state, s);
                if (state == ST_TRAVERSE)
                {
                    if (CULL) q.pre.cull_pad = cull_pad_for(q.r, q.pre, sc.meshes[sc.shapes[q.si].mesh].cull_scale);
                    q.node = 0; q.best = -1;
                }
                else if (state != ST_SHAPES && state != ST_SHADE && state != ST_MESHDONE) state = ST_IDLE;
            }
            const unsigned taken = win_pos + (unsigned)__popc(idle);
            win_pos = taken < win_end ? taken : win_end;
        }
        if (!__any_sync(RT_FULL_MASK, state != ST_IDLE)) break;

        query_traverse<CULL>(sc, q, state, exhausted ? 1 : 12, w.leaf_wait, cnt);
        query_mesh_done(sc, q, state, cnt);
        if (state == ST_SHAPES || state == ST_SHADE)
        {
            for (;;)
            {
                query_shapes<CULL>(sc, q, state, cnt);
                if (state == ST_TRAVERSE) break;
                Ray next; next.o = V3(0, 0, 0); next.d = V3(0, 0, 0); next.dist = 0.0f;
                bool next_any = false;
                if (!shade_query<MODE>(sc, a, w.pool, id, q, s, next, next_any)) { state = ST_IDLE; break; }
                query_begin(q, next, next_any, cnt);
                s.seg_dist = next.dist;
                state = ST_SHAPES;
            }
        }
    }
    flush_counters(cnt, a.counters, a.exact);
}

// ---- sample fold: AccumulatePixel::AddPixel + GetGammaSpacePixel ---------------------------------------
// (RayTracerProgram.cpp:57-71, :155-185).  One thread per pixel of the task; streaming.
__global__ void rt_resolve_kernel(const RenderArgs a, int pass_count)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int pixel = a.start + idx;
    if (pixel > a.end) return;
    const int x = pixel % a.width, y = pixel / a.width;
    if (!owns_pixel(a, x, y)) return;
    const size_t stride = (size_t)a.width * a.height;
    float4 acc = a.accum[pixel];
    float3 sum = V3(acc.x, acc.y, acc.z);
    int num = (int)acc.w;
    float3 last = V3(0, 0, 0);
    for (int p = 0; p < pass_count; p++)
    {
        float3 col;
        if (a.antialias)
        {
            col = V3(0, 0, 0);
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                const float4 s = a.samples[(size_t)(p * 4 + i) * stride + pixel];
                col = add3(col, V3(s.x, s.y, s.z));
            }
            col = V3(col.x / 4.0f, col.y / 4.0f, col.z / 4.0f);
        }
        else
        {
            const float4 s = a.samples[(size_t)p * stride + pixel];
            col = V3(s.x, s.y, s.z);
        }
        sum = add3(sum, col); num++;
        last = col;
    }
    a.accum[pixel] = make_float4(sum.x, sum.y, sum.z, (float)num);
    const float fn = (float)num;
    const float3 lin = a.mode == RT_MODE_PREVIEW ? last : V3(sum.x / fn, sum.y / fn, sum.z / fn);
    a.display[pixel] = make_pixel(lin);
}

__global__ void rt_display_kernel(const float4* accum, uint32_t* display, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 acc = accum[i];
    const int num = (int)acc.w;
    if (num <= 0) { display[i] = 0; return; }
    const float fn = (float)num;
    display[i] = make_pixel(V3(acc.x / fn, acc.y / fn, acc.z / fn));
}

// ---- multi-GPU tile exchange ---------------------------------------------------------------------------
// Dense order of the pixels a rank owns: owned tiles in tile-id order, row-major inside a tile,
// only the pixels inside the image.  Position = exclusive prefix, computed arithmetically.
struct TileArgs { int width, height, tile_size, tile_count, tile_rank, tiles_x, tiles_y; };

// one CTA per owned tile; dir 0: frame -> dense, 1: dense -> frame.  tile_offsets[k] precomputed on host.
__global__ void rt_tile_copy_kernel(float4* frame, float4* dense, const long long* tile_offsets, TileArgs t, int dir)
{
    const int k = blockIdx.x;
    const int tile = t.tile_rank + k * t.tile_count;
    const int tx = tile % t.tiles_x, ty = tile / t.tiles_x;
    const int ox = tx * t.tile_size, oy = ty * t.tile_size;
    const int w = min(t.tile_size, t.width - ox), h = min(t.tile_size, t.height - oy);
    const long long base = tile_offsets[k];
    for (int i = threadIdx.x; i < w * h; i += blockDim.x)
    {
        const int lx = i % w, ly = i / w;
        const size_t f = (size_t)(oy + ly) * t.width + (ox + lx);
        if (dir == 0) dense[base + i] = frame[f];
        else frame[f] = dense[base + i];
    }
}

// one CTA per owned tile: this rank's pixels written straight into the root GPU's frame over NVLink
// (16-byte stores to peer memory; no dense staging buffer, no collective)
__global__ void rt_tile_push_kernel(const float4* __restrict__ frame, float4* __restrict__ peer_frame, TileArgs t)
{
    const int tile = t.tile_rank + blockIdx.x * t.tile_count;
    const int tx = tile % t.tiles_x, ty = tile / t.tiles_x;
    const int ox = tx * t.tile_size, oy = ty * t.tile_size;
    const int w = min(t.tile_size, t.width - ox), h = min(t.tile_size, t.height - oy);
    for (int i = threadIdx.x; i < w * h; i += blockDim.x)
    {
        const int lx = i % w, ly = i / w;
        const size_t f = (size_t)(oy + ly) * t.width + (ox + lx);
        peer_frame[f] = frame[f];
    }
}

// ---- test hooks: arbitrary rays and primitive known-answer tests -------------------------------------------
template <bool CULL>
__global__ void rt_trace_rays_kernel(const DevScene sc, const float* rays, int n, int* shape_out, int* tri_out,
                                     float* hit11, unsigned long long* counters, int exact)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    Counters cnt = { 0, 0, 0, 0, 0, 0 };
    Ray r; r.o = V3(0, 0, 0); r.d = V3(0, 0, 1); r.dist = 0.0f;
    if (active)
    {
        const float* q = rays + 7 * (size_t)i;
        r.o = V3(q[0], q[1], q[2]); r.d = V3(q[3], q[4], q[5]); r.dist = q[6];
    }
    Hit h; h.pos = V3(0, 0, 0); h.nrm = V3(0, 0, 0); h.dist = 0.0f; h.color = V3(1.0f, 1.0f, 1.0f); h.alpha = 1.0f;
    int tri = -1;
    const int s = trace_scene<CULL>(sc, r, active, false, h, tri, cnt);
    if (active)
    {
        shape_out[i] = s; tri_out[i] = s >= 0 ? tri : -1;
        float* o = hit11 + 11 * (size_t)i;
        for (int k = 0; k < 11; k++) o[k] = 0.0f;
        if (s >= 0)
        {
            o[0] = h.pos.x; o[1] = h.pos.y; o[2] = h.pos.z; o[3] = h.nrm.x; o[4] = h.nrm.y; o[5] = h.nrm.z;
            o[6] = h.dist; o[7] = h.color.x; o[8] = h.color.y; o[9] = h.color.z; o[10] = h.alpha;
        }
    }
    flush_counters(cnt, counters, exact);
}

// kind: 0 aabb (prim 6 floats; out7[0] = tmin), 1 triangle (9), 2 sphere (4), 3 plane (6), 4 capsule (7),
//       5 q_rsqrt (rays unused; prim 1 float; out7[0]), 6 barycentric (prim 12: p,a,b,c; out7[0..2]),
//       7 display (prim 3: linear rgb; flags = ARGB)
__global__ void rt_kat_kernel(int kind, const float* rays, const float* prims, int n, int* flags, float* out7)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r; r.o = V3(0, 0, 0); r.d = V3(0, 0, 1); r.dist = 1.0f;
    if (rays)
    {
        const float* q = rays + 7 * (size_t)i;
        r.o = V3(q[0], q[1], q[2]); r.d = V3(q[3], q[4], q[5]); r.dist = q[6];
    }
    float* o = out7 + 7 * (size_t)i;
    for (int k = 0; k < 7; k++) o[k] = 0.0f;
    float3 pos = V3(0, 0, 0), nrm = V3(0, 0, 0); float dist = 0.0f;
    bool hit = false;
    if (kind == 0)
    {
        const float* b = prims + 6 * (size_t)i;
        RayPre pre = ray_pre(r);
        float tlo, thi;
        hit = slab_general(r, pre, ld3(b), ld3(b + 3), tlo, thi);
        const bool all_axes = pre.ex && pre.ey && pre.ez && finite3(r.o) && finite3(r.d);
        if (all_axes)
        {
            float tlo2, thi2;
            const bool hit2 = slab_fast(r, pre, ld3(b), ld3(b + 3), tlo2, thi2);
            if (hit2 != hit) hit = !hit;   // would surface as a mismatch against the oracle
            if (hit && __float_as_uint(tlo2) != __float_as_uint(tlo)) tlo = __uint_as_float(0x7fc00000u);
        }
        flags[i] = hit ? 1 : 0;
        o[0] = hit ? tlo : 0.0f;
        return;
    }
    if (kind == 1)
    {
        const float* t = prims + 9 * (size_t)i;
        const float3 p0 = ld3(t), p1 = ld3(t + 3), p2 = ld3(t + 6);
        const float3 nn = normalized3(cross3(sub3(p1, p0), sub3(p2, p0)));
        hit = triangle_test(r, p0, p1, p2, nn, pos, dist);
        nrm = nn;
    }
    else if (kind == 2) { const float* s = prims + 4 * (size_t)i; hit = sphere_test(r, ld3(s), s[3], pos, nrm, dist); }
    else if (kind == 3) { const float* s = prims + 6 * (size_t)i; hit = plane_test(r, ld3(s), ld3(s + 3), pos, nrm, dist); }
    else if (kind == 4)
    {
        const float* s = prims + 7 * (size_t)i;
        hit = cylinder_test(r, ld3(s), ld3(s + 3), s[6], pos, nrm, dist);
        if (!hit)
        {
            float3 p1, n1, p2, n2; float d1 = 0.0f, d2 = 0.0f;
            const bool b1 = sphere_test(r, ld3(s), s[6], p1, n1, d1);
            const bool b2 = sphere_test(r, ld3(s + 3), s[6], p2, n2, d2);
            hit = b1 || b2;
            if (hit) { const bool first = (b1 && b2) ? (d1 < d2) : b1; pos = first ? p1 : p2; nrm = first ? n1 : n2; dist = first ? d1 : d2; }
        }
    }
    else if (kind == 5) { o[0] = q_rsqrt(prims[i]); flags[i] = 1; return; }
    else if (kind == 6)
    {
        const float* q = prims + 12 * (size_t)i;
        barycentric(ld3(q), ld3(q + 3), ld3(q + 6), ld3(q + 9), o[0], o[1], o[2]);
        flags[i] = 1; return;
    }
    else if (kind == 7) { flags[i] = (int)make_pixel(ld3(prims + 3 * (size_t)i)); return; }
    flags[i] = hit ? 1 : 0;
    if (hit) { o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = nrm.x; o[4] = nrm.y; o[5] = nrm.z; o[6] = dist; }
}

__global__ void rt_kat_texture_kernel(cudaTextureObject_t atlas, DevTexture t, const float* uv, int n, float* out4)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = texture_sample(atlas, t, uv[2 * i], uv[2 * i + 1]);
    out4[4 * i] = c.x; out4[4 * i + 1] = c.y; out4[4 * i + 2] = c.z; out4[4 * i + 3] = c.w;
}

// =====================================================================================================
// host side of the boundary
// =====================================================================================================
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

    int width = 0, height = 0;
    float4* accum = nullptr;
    uint32_t* display = nullptr;
    int2* prim_ids = nullptr;
    float* prim_dist = nullptr;
    unsigned long long* counters = nullptr;     // 8 x u64 (rt_counters)
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
    unsigned tune_long_limit = RT_LONG_LIMIT;
    unsigned tune_small_round = RT_SMALL_ROUND;
    unsigned tune_thin_count = RT_THIN_COUNT;
    int tune_long_group = RT_LONG_GROUP;
    int tune_packet_rounds = -1;            // -1: by mode
    unsigned tune_packet_probe = RT_PACKET_PROBE;
    unsigned tune_packet_min_lanes = RT_PACKET_MIN_LANES;
    unsigned tune_thin_limit = RT_THIN_LIMIT;
};

static thread_local std::string g_create_error;

static int fail(rt_gpu_ctx* c, int code, const std::string& msg)
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

static void free_scene(rt_gpu_ctx* ctx)
{
    for (cudaTextureObject_t t : ctx->texobjs) cudaDestroyTextureObject(t);
    for (cudaArray_t a : ctx->arrays) cudaFreeArray(a);
    for (void* p : ctx->scene_allocs) cudaFree(p);
    ctx->texobjs.clear(); ctx->arrays.clear(); ctx->scene_allocs.clear(); ctx->host_textures.clear();
    ctx->has_scene = false; ctx->scene_bytes = 0;
    memset(&ctx->scene, 0, sizeof ctx->scene);
}

static void free_frame(rt_gpu_ctx* ctx)
{
    cudaFree(ctx->accum); cudaFree(ctx->display); cudaFree(ctx->prim_ids); cudaFree(ctx->prim_dist);
    ctx->accum = nullptr; ctx->display = nullptr; ctx->prim_ids = nullptr; ctx->prim_dist = nullptr;
    ctx->width = ctx->height = 0;
}

template <typename T>
static int upload(rt_gpu_ctx* ctx, const T* src, size_t count, T** out)
{
    *out = nullptr;
    if (count == 0) return RT_OK;
    void* p = nullptr;
    RT_CUDA(cudaMalloc(&p, count * sizeof(T)));
    ctx->scene_allocs.push_back(p);
    ctx->scene_bytes += count * sizeof(T);
    RT_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = (T*)p;
    return RT_OK;
}

template <bool CULL>
static cudaError_t launch_generate(int mode, unsigned grid, cudaStream_t st, const DevScene& sc, const RenderArgs& a, const WaveArgs& w)
{
    switch (mode)
    {
    case RT_MODE_PATH: rt_generate_kernel<CULL, RT_MODE_PATH><<<grid, 256, 0, st>>>(sc, a, w); break;
    case RT_MODE_PREVIEW: rt_generate_kernel<CULL, RT_MODE_PREVIEW><<<grid, 256, 0, st>>>(sc, a, w); break;
    case RT_MODE_WHITTED: rt_generate_kernel<CULL, RT_MODE_WHITTED><<<grid, 256, 0, st>>>(sc, a, w); break;
    case RT_MODE_PRIMARY: rt_generate_kernel<CULL, RT_MODE_PRIMARY><<<grid, 256, 0, st>>>(sc, a, w); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <bool CULL>
static cudaError_t launch_shade(int mode, unsigned grid, cudaStream_t st, const DevScene& sc, const RenderArgs& a, const WaveArgs& w, int round)
{
    switch (mode)
    {
    case RT_MODE_PATH: rt_shade_kernel<CULL, RT_MODE_PATH><<<grid, 256, 0, st>>>(sc, a, w, round); break;
    case RT_MODE_PREVIEW: rt_shade_kernel<CULL, RT_MODE_PREVIEW><<<grid, 256, 0, st>>>(sc, a, w, round); break;
    case RT_MODE_WHITTED: rt_shade_kernel<CULL, RT_MODE_WHITTED><<<grid, 256, 0, st>>>(sc, a, w, round); break;
    case RT_MODE_PRIMARY: rt_shade_kernel<CULL, RT_MODE_PRIMARY><<<grid, 256, 0, st>>>(sc, a, w, round); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <bool CULL>
static cudaError_t launch_finish(int mode, unsigned grid, cudaStream_t st, const DevScene& sc, const RenderArgs& a, const WaveArgs& w, int round)
{
    switch (mode)
    {
    case RT_MODE_PATH: rt_finish_kernel<CULL, RT_MODE_PATH><<<grid, 128, 0, st>>>(sc, a, w, round); break;
    case RT_MODE_PREVIEW: rt_finish_kernel<CULL, RT_MODE_PREVIEW><<<grid, 128, 0, st>>>(sc, a, w, round); break;
    case RT_MODE_WHITTED: rt_finish_kernel<CULL, RT_MODE_WHITTED><<<grid, 128, 0, st>>>(sc, a, w, round); break;
    case RT_MODE_PRIMARY: rt_finish_kernel<CULL, RT_MODE_PRIMARY><<<grid, 128, 0, st>>>(sc, a, w, round); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <typename T>
static cudaError_t grow(T** ptr, size_t* cap, size_t need)
{
    if (need <= *cap) return cudaSuccess;
    cudaFree(*ptr); *ptr = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc((void**)ptr, need * sizeof(T));
    if (e == cudaSuccess) *cap = need;
    return e;
}

extern "C" {

int rt_gpu_abi_version(void) { return RT_GPU_ABI_VERSION; }

int rt_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static void tuning_from_env(rt_gpu_ctx* ctx);
int rt_gpu_create(int device, rt_gpu_ctx** out_ctx)
{
    rt_gpu_ctx* ctx = nullptr;
    if (!out_ctx) return fail(nullptr, RT_ERR_INVALID, "out_ctx is null");
    *out_ctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, RT_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") +
                                              (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(nullptr, RT_ERR_INVALID, "device index out of range");
    RT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, RT_ERR_CUDA, "device is not sm_100 (this library carries sm_100a code only)");
    ctx = new rt_gpu_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    memset(&ctx->scene, 0, sizeof ctx->scene);
    cudaError_t e2 = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&ctx->ev0);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&ctx->ev1);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&ctx->counters, 8 * sizeof(unsigned long long));
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&ctx->fork, cudaEventDisableTiming);
    for (int k = 0; k < RT_PIPES && e2 == cudaSuccess; k++)
    {
        e2 = cudaStreamCreateWithFlags(&ctx->pipes[k].stream, cudaStreamNonBlocking);
        if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&ctx->pipes[k].done, cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&ctx->pipes[k].round_counters, (6 * RT_MAX_ROUNDS + 1) * sizeof(unsigned));
        if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&ctx->pipes[k].retry_counts, RT_MAX_RETRIES * sizeof(unsigned));
        memset(&ctx->pipes[k].pool, 0, sizeof(PathPool));
    }
    if (e2 == cudaSuccess) e2 = cudaMemsetAsync(ctx->counters, 0, 8 * sizeof(unsigned long long), ctx->stream);
    if (e2 != cudaSuccess)
    {
        std::string msg = std::string("context setup: ") + cudaGetErrorString(e2);
        delete ctx;
        return fail(nullptr, RT_ERR_CUDA, msg);
    }
    tuning_from_env(ctx);
    *out_ctx = ctx;
    return RT_OK;
}

int rt_gpu_destroy(rt_gpu_ctx* ctx)
{
    if (!ctx) return RT_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_scene(ctx);
    free_frame(ctx);
    cudaFree(ctx->counters);
    for (int k = 0; k < RT_PIPES; k++)
    {
        rt_gpu_ctx::Pipe& pp = ctx->pipes[k];
        if (pp.stream) cudaStreamSynchronize(pp.stream);
        for (void* q : pp.allocs) cudaFree(q);
        cudaFree(pp.round_counters); cudaFree(pp.retry_counts); cudaFree(pp.retry[0]); cudaFree(pp.retry[1]); cudaFree(pp.samples);
        if (pp.done) cudaEventDestroy(pp.done);
        if (pp.stream) cudaStreamDestroy(pp.stream);
    }
    if (ctx->fork) cudaEventDestroy(ctx->fork);
    for (rt_gpu_ctx::TileTable& tt : ctx->tile_tables) cudaFree(tt.offsets);
    cudaFree(ctx->gather_staging);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->kev) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RT_OK;
}

const char* rt_gpu_last_error(rt_gpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rt_gpu_upload_scene(rt_gpu_ctx* ctx, const rt_scene_desc* s)
{
    if (!ctx) return RT_ERR_INVALID;
    if (!s) return fail(ctx, RT_ERR_INVALID, "scene is null");
    if (s->abi_version != RT_GPU_ABI_VERSION) return fail(ctx, RT_ERR_INVALID, "rt_scene_desc.abi_version mismatch");
    if (s->num_shapes < 0 || s->num_materials < 0 || s->num_meshes < 0 || s->num_lights < 0)
        return fail(ctx, RT_ERR_INVALID, "negative count in scene");
    // validate indices before anything is copied
    bool needs_table = false;
    for (int i = 0; i < s->num_materials; i++)
    {
        const rt_material& m = s->materials[i];
        if (m.type < RT_MAT_DIFFUSE || m.type > RT_MAT_NULL) return fail(ctx, RT_ERR_INVALID, "unknown material type");
        if (m.type == RT_MAT_BLEND || m.type == RT_MAT_COMBINE)
            if (m.child_a < 0 || m.child_a >= s->num_materials || m.child_b < 0 || m.child_b >= s->num_materials)
                return fail(ctx, RT_ERR_INVALID, "material child index out of range");
        if (m.type == RT_MAT_DIFFUSE || m.type == RT_MAT_CHECKER) needs_table = true;
    }
    for (int i = 0; i < s->num_shapes; i++)
    {
        const rt_shape& sh = s->shapes[i];
        if (sh.type < RT_SHAPE_SPHERE || sh.type > RT_SHAPE_TRIANGLE) return fail(ctx, RT_ERR_INVALID, "unknown shape type");
        if (sh.material >= s->num_materials) return fail(ctx, RT_ERR_INVALID, "shape material index out of range");
        if (sh.type == RT_SHAPE_MESH && (sh.mesh < 0 || sh.mesh >= s->num_meshes)) return fail(ctx, RT_ERR_INVALID, "shape mesh index out of range");
    }
    for (int i = 0; i < s->num_meshes; i++)
    {
        const rt_mesh& m = s->meshes[i];
        if (m.num_nodes < 0 || m.num_tris < 0 || m.num_textures < 0) return fail(ctx, RT_ERR_INVALID, "negative count in mesh");
        if (m.num_nodes > 0 && (!m.nodes || !m.tris || !m.shade)) return fail(ctx, RT_ERR_INVALID, "mesh arrays missing");
        for (int k = 0; k < m.num_nodes; k++)
        {
            const rt_bvh_node& nd = m.nodes[k];
            if (nd.escape <= k || nd.escape > m.num_nodes || nd.tri >= m.num_tris)
                return fail(ctx, RT_ERR_INVALID, "malformed BVH node (escape/tri index)");
            if (nd.tri < 0)
            {
                // an inner node's range [k, escape) is its left subtree [k+1, r) followed by its right one [r, escape)
                if (k + 1 >= nd.escape) return fail(ctx, RT_ERR_INVALID, "malformed BVH node (inner node without children)");
                const int r = m.nodes[k + 1].escape;
                if (r > nd.escape || (r < nd.escape && m.nodes[r].escape != nd.escape))
                    return fail(ctx, RT_ERR_INVALID, "malformed BVH node (subtrees do not nest)");
            }
            else if (nd.escape != k + 1) return fail(ctx, RT_ERR_INVALID, "malformed BVH node (leaf with a subtree)");
        }
        for (int k = 0; k < m.num_tris; k++)
        {
            if (m.tris[k].index < 0 || m.tris[k].index >= m.num_tris) return fail(ctx, RT_ERR_INVALID, "triangle index out of range");
            if (m.shade[k].texture >= m.num_textures) return fail(ctx, RT_ERR_INVALID, "texture index out of range");
        }
        for (int k = 0; k < m.num_textures; k++)
            if (m.textures[k].rgba && (m.textures[k].width <= 0 || m.textures[k].height <= 0))
                return fail(ctx, RT_ERR_INVALID, "texture with non-positive size");
    }

    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    free_scene(ctx);

    DevScene d;
    memset(&d, 0, sizeof d);
    int rc;
    rt_shape* dshapes; rt_material* dmats; rt_light* dlights;
    if ((rc = upload(ctx, s->shapes, (size_t)s->num_shapes, &dshapes)) != RT_OK) return rc;
    if ((rc = upload(ctx, s->materials, (size_t)s->num_materials, &dmats)) != RT_OK) return rc;
    if ((rc = upload(ctx, s->lights, (size_t)s->num_lights, &dlights)) != RT_OK) return rc;
    d.shapes = dshapes; d.materials = dmats; d.lights = dlights;
    d.num_shapes = s->num_shapes; d.num_materials = s->num_materials; d.num_lights = s->num_lights;
    d.num_meshes = s->num_meshes;

    // ---- texture atlas: shelf-pack every decoded texture of the scene into one float4 cudaArray ----
    struct AtlasRect { int x, y, w, h; };
    std::vector<AtlasRect> atlas_rects;          // in (mesh, slot) order, textured slots only
    size_t atlas_next = 0;
    cudaArray_t atlas_array = nullptr;
    {
        int max_w = 0;
        for (int i = 0; i < s->num_meshes; i++)
            for (int k = 0; k < s->meshes[i].num_textures; k++)
                if (s->meshes[i].textures[k].rgba)
                {
                    const rt_texture& t = s->meshes[i].textures[k];
                    atlas_rects.push_back(AtlasRect{ 0, 0, t.width, t.height });
                    if (t.width > max_w) max_w = t.width;
                }
        if (!atlas_rects.empty())
        {
            const int shelf_w = max_w > 8192 ? max_w : 8192;
            std::vector<size_t> order(atlas_rects.size());
            for (size_t k = 0; k < order.size(); k++) order[k] = k;
            std::stable_sort(order.begin(), order.end(), [&](size_t l, size_t r) { return atlas_rects[l].h > atlas_rects[r].h; });
            int cx = 0, cy = 0, shelf_h = 0, used_w = 0;
            for (size_t k : order)
            {
                AtlasRect& r = atlas_rects[k];
                if (cx + r.w > shelf_w) { cy += shelf_h; cx = 0; shelf_h = 0; }
                r.x = cx; r.y = cy;
                cx += r.w;
                if (r.h > shelf_h) shelf_h = r.h;
                if (cx > used_w) used_w = cx;
            }
            const int atlas_h = cy + shelf_h;
            if (used_w > 131072 || atlas_h > 65536) return fail(ctx, RT_ERR_INVALID, "textures do not fit one atlas (131072 x 65536 texels)");
            cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float4>();
            RT_CUDA(cudaMallocArray(&atlas_array, &fmt, (size_t)used_w, (size_t)atlas_h));
            ctx->arrays.push_back(atlas_array);
            ctx->scene_bytes += (size_t)used_w * atlas_h * 16;
            cudaResourceDesc res; memset(&res, 0, sizeof res);
            res.resType = cudaResourceTypeArray; res.res.array.array = atlas_array;
            cudaTextureDesc td; memset(&td, 0, sizeof td);
            td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint;       // texels only; RTexture::Sample's lerps are done in fp32 by hand
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            cudaTextureObject_t obj = 0;
            RT_CUDA(cudaCreateTextureObject(&obj, &res, &td, nullptr));
            ctx->texobjs.push_back(obj);
            d.atlas = obj;
        }
    }

    std::vector<DevMesh> meshes((size_t)s->num_meshes);
    for (int i = 0; i < s->num_meshes; i++)
    {
        const rt_mesh& m = s->meshes[i];
        DevMesh& dm = meshes[i];
        memset(&dm, 0, sizeof dm);
        rt_bvh_node* dn; rt_tri* dt; rt_shade* dsh;
        if ((rc = upload(ctx, m.nodes, (size_t)m.num_nodes, &dn)) != RT_OK) return rc;
        if ((rc = upload(ctx, m.tris, (size_t)m.num_tris, &dt)) != RT_OK) return rc;
        if ((rc = upload(ctx, m.shade, (size_t)m.num_tris, &dsh)) != RT_OK) return rc;
        if (m.num_nodes > 0)
        {
            rt_patch_right_child<<<(unsigned)((m.num_nodes + 255) / 256), 256, 0, ctx->stream>>>(dn, m.num_nodes);
            RT_CUDA(cudaGetLastError());
        }
        dm.nodes = (const float4*)dn; dm.tris = (const float4*)dt; dm.shade = (const float4*)dsh;
        dm.num_nodes = m.num_nodes; dm.num_tris = m.num_tris; dm.num_textures = m.num_textures;
        float scale = 0.0f;
        if (m.num_nodes > 0)
            for (int k = 0; k < 3; k++)
            {
                scale = fmaxf(scale, fabsf(m.nodes[0].bmin[k]));
                scale = fmaxf(scale, fabsf(m.nodes[0].bmax[k]));
            }
        dm.cull_scale = scale;
        std::vector<DevTexture> texs((size_t)m.num_textures);
        for (int k = 0; k < m.num_textures; k++)
        {
            const rt_texture& t = m.textures[k];
            DevTexture& dt2 = texs[k];
            dt2.x0 = dt2.y0 = 0; dt2.width = t.width; dt2.height = t.height;
            if (!t.rgba) { dt2.width = dt2.height = 0; continue; }
            const AtlasRect& rc2 = atlas_rects[atlas_next++];
            dt2.x0 = rc2.x; dt2.y0 = rc2.y;
            RT_CUDA(cudaMemcpy2DToArrayAsync(atlas_array, (size_t)rc2.x * 16, (size_t)rc2.y, t.rgba, (size_t)t.width * 16,
                                             (size_t)t.width * 16, (size_t)t.height, cudaMemcpyHostToDevice, ctx->stream));
            ctx->host_textures.push_back(dt2);
        }
        // a shade record may only name a slot that holds pixels
        for (int k = 0; k < m.num_tris; k++)
            if (m.shade[k].texture >= 0 && !m.textures[m.shade[k].texture].rgba)
                return fail(ctx, RT_ERR_INVALID, "shade record names an empty texture slot");
        DevTexture* dtex;
        if ((rc = upload(ctx, texs.data(), texs.size(), &dtex)) != RT_OK) return rc;
        dm.textures = dtex;
        RT_CUDA(cudaStreamSynchronize(ctx->stream));   // texs is a local
    }
    DevMesh* dmeshes;
    if ((rc = upload(ctx, meshes.data(), meshes.size(), &dmeshes)) != RT_OK) return rc;
    d.meshes = dmeshes;

    if (s->num_unit_vectors > 0 && s->unit_vectors)
    {
        // pad xyz -> float4 so one 16-byte load fetches a direction
        const size_t n = s->num_unit_vectors;
        float4* dv = nullptr;
        RT_CUDA(cudaMalloc((void**)&dv, n * sizeof(float4)));
        ctx->scene_allocs.push_back(dv);
        ctx->scene_bytes += n * sizeof(float4);
        const size_t step = 1u << 20;
        std::vector<float4> stage(step < n ? step : n);
        for (size_t lo = 0; lo < n; lo += step)
        {
            const size_t cnt = (n - lo) < step ? (n - lo) : step;
            for (size_t k = 0; k < cnt; k++)
            {
                const float* v = s->unit_vectors + 3 * (lo + k);
                stage[k] = make_float4(v[0], v[1], v[2], 0.0f);
            }
            RT_CUDA(cudaMemcpy(dv + lo, stage.data(), cnt * sizeof(float4), cudaMemcpyHostToDevice));
        }
        d.unit_vectors = dv;
        d.num_unit_vectors = s->num_unit_vectors;
    }
    for (int k = 0; k < 3; k++) d.eye[k] = s->eye[k];
    d.dir_z = s->dir_z; d.ray_distance = s->ray_distance; d.bounce_offset = s->bounce_offset;
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->scene = d;
    ctx->has_scene = true;
    ctx->needs_table = needs_table;
    ctx->host_shape_is_mesh.assign((size_t)s->num_shapes, 0);
    for (int i = 0; i < s->num_shapes; i++) ctx->host_shape_is_mesh[i] = s->shapes[i].type == RT_SHAPE_MESH ? 1 : 0;
    ctx->all_bounded = true;
    for (int i = 0; i < s->num_shapes; i++) if (!s->shapes[i].has_bounds) ctx->all_bounded = false;
    return RT_OK;
}

int rt_gpu_reset_accum(rt_gpu_ctx* ctx, int32_t width, int32_t height)
{
    if (!ctx) return RT_ERR_INVALID;
    if (width <= 0 || height <= 0 || (long long)width * height > 0x7fffffffLL) return fail(ctx, RT_ERR_INVALID, "bad frame size");
    RT_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)width * height;
    if (width != ctx->width || height != ctx->height)
    {
        RT_CUDA(cudaStreamSynchronize(ctx->stream));
        free_frame(ctx);
        RT_CUDA(cudaMalloc((void**)&ctx->accum, n * sizeof(float4)));
        RT_CUDA(cudaMalloc((void**)&ctx->display, n * sizeof(uint32_t)));
        RT_CUDA(cudaMalloc((void**)&ctx->prim_ids, n * sizeof(int2)));
        RT_CUDA(cudaMalloc((void**)&ctx->prim_dist, n * sizeof(float)));
        ctx->width = width; ctx->height = height;
    }
    RT_CUDA(cudaMemsetAsync(ctx->accum, 0, n * sizeof(float4), ctx->stream));
    RT_CUDA(cudaMemsetAsync(ctx->display, 0, n * sizeof(uint32_t), ctx->stream));
    RT_CUDA(cudaMemsetAsync(ctx->prim_ids, 0xff, n * sizeof(int2), ctx->stream));
    RT_CUDA(cudaMemsetAsync(ctx->prim_dist, 0, n * sizeof(float), ctx->stream));
    return RT_OK;
}

int rt_gpu_render_tile(rt_gpu_ctx* ctx, const rt_render_params* p)
{
    if (!ctx) return RT_ERR_INVALID;
    if (!p) return fail(ctx, RT_ERR_INVALID, "params is null");
    if (!ctx->has_scene) return fail(ctx, RT_ERR_NO_SCENE, "rt_gpu_render_tile before rt_gpu_upload_scene");
    if (p->width <= 0 || p->height <= 0 || (long long)p->width * p->height > 0x7fffffffLL) return fail(ctx, RT_ERR_INVALID, "bad frame size");
    const int npix = p->width * p->height;
    if (p->start < 0 || p->end >= npix) return fail(ctx, RT_ERR_INVALID, "pixel range outside the frame");
    if (p->mode < RT_MODE_PATH || p->mode > RT_MODE_PRIMARY) return fail(ctx, RT_ERR_INVALID, "unknown mode");
    if (p->traverse != RT_TRAVERSE_EXACT && p->traverse != RT_TRAVERSE_CULLED) return fail(ctx, RT_ERR_INVALID, "unknown traverse");
    if (p->mode == RT_MODE_PATH && ctx->needs_table && ctx->scene.num_unit_vectors == 0)
        return fail(ctx, RT_ERR_INVALID, "scene has Diffuse materials but no unit-vector table (rt_host_set_unit_vectors)");
    if (p->max_bounce < 0 || p->max_bounce > RT_MAX_PATH_DEPTH) return fail(ctx, RT_ERR_INVALID, "max_bounce must be in [0, 32]");
    if (p->pass_count < 0 || p->pass_begin < 0) return fail(ctx, RT_ERR_INVALID, "negative pass range");
    if (p->tile_count > 1 && (p->tile_size <= 0 || p->tile_rank < 0 || p->tile_rank >= p->tile_count))
        return fail(ctx, RT_ERR_INVALID, "bad tile ownership fields");
    RT_CUDA(cudaSetDevice(ctx->device));
    if (p->width != ctx->width || p->height != ctx->height)
    {
        int rc = rt_gpu_reset_accum(ctx, p->width, p->height);
        if (rc != RT_OK) return rc;
    }
    RT_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    ctx->timed = true;
    ctx->kev_used = 0;
    if (p->end < p->start || (p->mode != RT_MODE_PRIMARY && p->pass_count == 0))
    {
        RT_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        return RT_OK;                                 // empty task
    }

    RenderArgs a;
    memset(&a, 0, sizeof a);
    a.width = p->width; a.height = p->height; a.start = p->start; a.end = p->end;
    a.mode = p->mode; a.max_bounce = p->max_bounce; a.antialias = p->antialias ? 1 : 0; a.seed = p->seed;
    a.spp = (a.antialias && p->mode != RT_MODE_PRIMARY) ? 4 : 1;
    if (p->mode == RT_MODE_PRIMARY) a.antialias = 0;
    a.tiled = (p->tile_count > 1 && p->tile_size > 0) ? 1 : 0;
    a.tile_size = p->tile_size; a.tile_count = p->tile_count; a.tile_rank = p->tile_rank;
    if (a.tiled)
    {
        a.tiles_x = (p->width + p->tile_size - 1) / p->tile_size;
        const int tiles_y = (p->height + p->tile_size - 1) / p->tile_size;
        const int ntiles = a.tiles_x * tiles_y;
        const int owned = p->tile_rank < ntiles ? (ntiles - p->tile_rank + p->tile_count - 1) / p->tile_count : 0;
        a.blocks_x = (p->tile_size + 7) / 8;
        a.blocks_per_tile = a.blocks_x * ((p->tile_size + 3) / 4);
        a.num_blocks = (unsigned)owned * (unsigned)a.blocks_per_tile;
    }
    else
    {
        a.row0 = p->start / p->width;
        a.rows = p->end / p->width - a.row0 + 1;
        a.blocks_x = (p->width + 7) / 8;
        a.blocks_per_tile = a.blocks_x * ((a.rows + 3) / 4);
        a.num_blocks = (unsigned)a.blocks_per_tile;
    }
    a.accum = ctx->accum; a.display = ctx->display; a.prim_ids = ctx->prim_ids; a.prim_dist = ctx->prim_dist;
    a.counters = ctx->counters;
    a.exact = p->traverse == RT_TRAVERSE_EXACT ? 1 : 0;
    a.all_bounded = ctx->all_bounded ? 1 : 0;
    if (a.num_blocks == 0)
    {
        RT_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        return RT_OK;
    }

    if (ctx->walk_blocks_per_sm == 0)
    {
        int b0 = 0, b1 = 0;
        RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, rt_walk_kernel<true>, 256, 0));
        RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, rt_walk_kernel<false>, 256, 0));
        ctx->walk_blocks_per_sm = b0 < b1 ? b0 : b1;
        if (ctx->walk_blocks_per_sm < 1) return fail(ctx, RT_ERR_CUDA, "walk kernel does not fit on an SM");
    }
    // rounds: one mesh walk per round; a path needs at most (segments) x (mesh shapes) walks
    int mesh_shapes = 0;
    for (int i = 0; i < ctx->scene.num_shapes; i++) mesh_shapes += ctx->host_shape_is_mesh[i] ? 1 : 0;
    const int segments = p->mode == RT_MODE_PATH ? (p->max_bounce > 0 ? p->max_bounce : 1)
                       : p->mode == RT_MODE_WHITTED ? 1 + ctx->scene.num_lights : 1;
    const int rounds = segments * (mesh_shapes > 0 ? mesh_shapes : 1);
    if (rounds >= RT_MAX_ROUNDS) return fail(ctx, RT_ERR_INVALID, "max_bounce x mesh shapes exceeds the round table");

    const int total_passes = p->mode == RT_MODE_PRIMARY ? 1 : p->pass_count;
    // sample buffer: whole frames of float4 per sample; split long calls into pass chunks
    // A call with fewer camera rays (one rank's share of a multi-GPU frame) is split in fewer, longer chunks:
    // every chunk pays its thin last rounds once, and there is less dense work to hide them behind.
    const unsigned long long call_items = (unsigned long long)a.num_blocks * 32ull * (unsigned long long)a.spp * (unsigned long long)total_passes;
    const size_t sample_budget = getenv("RT_SAMPLE_BUDGET_MB") ? ((size_t)atoi(getenv("RT_SAMPLE_BUDGET_MB")) << 20)
                               : call_items < RT_FEW_ITEMS ? RT_SAMPLE_BUDGET_FEW_BYTES : RT_SAMPLE_BUDGET_BYTES;
    size_t passes_per_chunk = sample_budget / ((size_t)npix * sizeof(float4) * (size_t)a.spp);
    if (!getenv("RT_SAMPLE_BUDGET_MB") && call_items < RT_FEW_ITEMS && passes_per_chunk > (size_t)(total_passes + 1) / 2)
        passes_per_chunk = (size_t)(total_passes + 1) / 2;         // two chunks all the same (two pipes overlap)
    if (passes_per_chunk < 1) passes_per_chunk = 1;
    if (passes_per_chunk > (size_t)total_passes) passes_per_chunk = (size_t)total_passes;
    {
        // the work list of one launch is indexed with 32 bits
        const unsigned long long per_pass = (unsigned long long)a.num_blocks * 32ull * (unsigned long long)a.spp;
        if (per_pass > 0xffffffffull) return fail(ctx, RT_ERR_INVALID, "frame too large for one launch");
        const size_t fit = (size_t)(0xffffffffull / per_pass);
        if (passes_per_chunk > fit) passes_per_chunk = fit;
    }
    {
        // equal-sized chunks (16 passes with room for 5 -> 4 x 4, not 5+5+5+1)
        const size_t nchunks = ((size_t)total_passes + passes_per_chunk - 1) / passes_per_chunk;
        passes_per_chunk = ((size_t)total_passes + nchunks - 1) / nchunks;
    }
    const int nchunks_total = (int)(((size_t)total_passes + passes_per_chunk - 1) / passes_per_chunk);
    if (p->mode != RT_MODE_PRIMARY)
    {
        const size_t need = passes_per_chunk * (size_t)a.spp * (size_t)npix;
        for (int k = 0; k < RT_PIPES && k < nchunks_total; k++)
        {
            rt_gpu_ctx::Pipe& pp = ctx->pipes[k];
            if (need > pp.samples_cap)
            {
                RT_CUDA(cudaStreamSynchronize(ctx->stream));
                RT_CUDA(cudaStreamSynchronize(pp.stream));
                cudaFree(pp.samples); pp.samples = nullptr; pp.samples_cap = 0;
                RT_CUDA(cudaMalloc((void**)&pp.samples, need * sizeof(float4)));
                pp.samples_cap = need;
            }
        }
    }

    // ---- batches, pools and pipes ------------------------------------------------------------------------
    // A batch is as large as possible (every batch pays the latency of its thin last rounds once); its
    // pool is smaller: most camera rays never become paths.  When a pool does fill up, the items it
    // turned away are generated again by retry passes.
    const unsigned long long items_per_chunk = (unsigned long long)passes_per_chunk * a.spp * a.num_blocks * 32ull;
    const int npipes = ctx->tune_pipes;
    const size_t batch = (size_t)((items_per_chunk + 255ull) & ~255ull);
    size_t pool_want = batch < ctx->max_pool_paths ? batch : ctx->max_pool_paths;
    const size_t levels_want = p->mode == RT_MODE_PATH ? (size_t)(p->max_bounce > 0 ? p->max_bounce : 1) : 1;
    if (pool_want > ctx->pool_cap || levels_want > ctx->pool_levels || (p->mode == RT_MODE_WHITTED && !ctx->pool_whitted))
    {
        // (only when the pools have to grow: the memory query costs a driver round trip)
        // keep all pools together within ~half of the device memory that is free right now (deep bounce
        // budgets make a path record large: 36 B per level); a smaller pool only costs retry passes
        const size_t levels_now = p->mode == RT_MODE_PATH ? (size_t)(p->max_bounce > 0 ? p->max_bounce : 1) : 1;
        const size_t lv = levels_now > ctx->pool_levels ? levels_now : ctx->pool_levels;
        const size_t per_path = 9 * 16 + ((p->mode == RT_MODE_WHITTED || ctx->pool_whitted) ? 64 : 0) + lv * 36 + 12;
        size_t free_b = 0, total_b = 0;
        RT_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t have_b = free_b + ctx->pool_cap * (9 * 16 + (ctx->pool_whitted ? 64 : 0) + ctx->pool_levels * 36 + 12) * RT_PIPES;
        const size_t fit = have_b / 2 / RT_PIPES / per_path;
        if (pool_want > fit) pool_want = fit > 255 ? (fit & ~(size_t)255) : fit;
        if (pool_want < (batch < 4096 ? batch : (size_t)4096)) return fail(ctx, RT_ERR_NOMEM, "not enough device memory for the path pools");
    }
    const int retries = (int)((batch + pool_want - 1) / pool_want) - 1;
    if (retries >= RT_MAX_RETRIES) return fail(ctx, RT_ERR_NOMEM, "path pool too small for this frame (raise the pool size)");
    {
        const size_t levels = p->mode == RT_MODE_PATH ? (size_t)(p->max_bounce > 0 ? p->max_bounce : 1) : 1;
        const bool whitted = p->mode == RT_MODE_WHITTED;
        if (pool_want > ctx->pool_cap || levels > ctx->pool_levels || (whitted && !ctx->pool_whitted))
        {
            RT_CUDA(cudaStreamSynchronize(ctx->stream));
            const bool same_shape = levels <= ctx->pool_levels && (!whitted || ctx->pool_whitted);
            const size_t cap = (same_shape && ctx->pool_cap > pool_want) ? ctx->pool_cap : pool_want;
            const size_t lv = levels > ctx->pool_levels ? levels : ctx->pool_levels;
            const bool wh = whitted || ctx->pool_whitted;
            for (int k = 0; k < RT_PIPES; k++)
            {
                rt_gpu_ctx::Pipe& pp = ctx->pipes[k];
                RT_CUDA(cudaStreamSynchronize(pp.stream));
                for (void* q : pp.allocs) cudaFree(q);
                pp.allocs.clear();
                auto alloc = [&](size_t bytes, void** out) -> cudaError_t {
                    cudaError_t e = cudaMalloc(out, bytes);
                    if (e == cudaSuccess) pp.allocs.push_back(*out);
                    return e;
                };
                PathPool& pl = pp.pool;
                memset(&pl, 0, sizeof pl);
                RT_CUDA(alloc(cap * 16, (void**)&pl.ro)); RT_CUDA(alloc(cap * 16, (void**)&pl.rd));
                RT_CUDA(alloc(cap * 16, (void**)&pl.cur)); RT_CUDA(alloc(cap * 16, (void**)&pl.bp));
                RT_CUDA(alloc(cap * 16, (void**)&pl.h0)); RT_CUDA(alloc(cap * 16, (void**)&pl.h1));
                RT_CUDA(alloc(cap * 16, (void**)&pl.h2)); RT_CUDA(alloc(cap * 16, (void**)&pl.pa));
                RT_CUDA(alloc(cap * 16, (void**)&pl.pb));
                if (wh)
                {
                    RT_CUDA(alloc(cap * 16, (void**)&pl.w0)); RT_CUDA(alloc(cap * 16, (void**)&pl.w1));
                    RT_CUDA(alloc(cap * 16, (void**)&pl.w2)); RT_CUDA(alloc(cap * 16, (void**)&pl.w3));
                }
                RT_CUDA(alloc(cap * lv * 16, (void**)&pl.st0)); RT_CUDA(alloc(cap * lv * 16, (void**)&pl.st1));
                RT_CUDA(alloc(cap * lv * 4, (void**)&pl.st2));
                RT_CUDA(alloc(cap * 4, (void**)&pp.queue[0])); RT_CUDA(alloc(cap * 4, (void**)&pp.queue[1]));
                RT_CUDA(alloc(cap * 4, (void**)&pp.longq));
                RT_CUDA(alloc(cap * 4, (void**)&pp.slowq));
                pl.cap = (unsigned)cap;
            }
            ctx->pool_cap = cap; ctx->pool_levels = lv; ctx->pool_whitted = wh;
        }
        for (int k = 0; k < RT_PIPES; k++)
        {
            rt_gpu_ctx::Pipe& pp = ctx->pipes[k];
            if (retries > 0 && batch > pp.retry_cap)
            {
                RT_CUDA(cudaStreamSynchronize(ctx->stream));
                RT_CUDA(cudaStreamSynchronize(pp.stream));
                cudaFree(pp.retry[0]); cudaFree(pp.retry[1]); pp.retry[0] = pp.retry[1] = nullptr; pp.retry_cap = 0;
                RT_CUDA(cudaMalloc((void**)&pp.retry[0], batch * 4));
                RT_CUDA(cudaMalloc((void**)&pp.retry[1], batch * 4));
                pp.retry_cap = batch;
            }
        }
    }
    const bool cull = p->traverse == RT_TRAVERSE_CULLED;
    // the round whose queue is still in camera order is walked as packets (bounce and shadow rays of later
    // rounds are not coherent enough: measured slower)
    const int packet_rounds = ctx->tune_packet_rounds >= 0 ? ctx->tune_packet_rounds : 1;
    const unsigned walk_grid = (unsigned)(ctx->num_sms * ctx->walk_blocks_per_sm);

    // Chunks alternate between the pipes.  Each pipe renders into its own sample buffer and folds it
    // itself; the only cross-pipe order is fold(k) before fold(k+1) (AddPixel sums in pass order), so
    // the thin, latency-bound last rounds of chunk k overlap the dense first rounds of chunk k+1.
    RT_CUDA(cudaEventRecord(ctx->fork, ctx->stream));
    bool used[RT_PIPES] = { false };
    int chunk_index = 0, last_pipe = -1;
    for (int done = 0; done < total_passes; done += (int)passes_per_chunk, chunk_index++)
    {
        const int chunk = (total_passes - done) < (int)passes_per_chunk ? (total_passes - done) : (int)passes_per_chunk;
        const int pipe = chunk_index % npipes;
        rt_gpu_ctx::Pipe& pp = ctx->pipes[pipe];
        if (!used[pipe]) { RT_CUDA(cudaStreamWaitEvent(pp.stream, ctx->fork, 0)); used[pipe] = true; }
        a.pass_begin = p->pass_begin + done;
        a.num_samples = chunk * a.spp;
        a.num_items = (unsigned)((unsigned long long)a.num_samples * a.num_blocks * 32ull);
        a.samples = pp.samples;
        {
            WaveArgs w;
            memset(&w, 0, sizeof w);
            w.pool = pp.pool;
            w.queue[0] = pp.queue[0]; w.queue[1] = pp.queue[1];
            w.counts = pp.round_counters; w.heads = pp.round_counters + RT_MAX_ROUNDS + 1;
            w.longq = pp.longq; w.lcounts = w.heads + RT_MAX_ROUNDS; w.lheads = w.lcounts + RT_MAX_ROUNDS;
            w.slowq = pp.slowq; w.scounts = w.lheads + RT_MAX_ROUNDS; w.sheads = w.scounts + RT_MAX_ROUNDS;
            w.packet_probe = ctx->tune_packet_probe; w.packet_min_lanes = ctx->tune_packet_min_lanes;
            w.long_limit = ctx->tune_long_limit; w.small_round = ctx->tune_small_round;
            w.thin_count = ctx->tune_thin_count; w.thin_limit = ctx->tune_thin_limit;
            w.packets = packet_rounds > 0 && mesh_shapes > 0 ? 1 : 0;
            w.min_lanes = ctx->tune_min_lanes; w.leaf_wait = ctx->tune_leaf_wait; w.window = ctx->tune_window;
            w.item_begin = 0u;
            w.item_count = a.num_items;
            unsigned gen_grid = (a.num_blocks * 32u + 255u) / 256u;
            if (gen_grid > (unsigned)ctx->num_sms * 8u) gen_grid = (unsigned)ctx->num_sms * 8u;
            // the shade grid strides over the round's queue; a few waves are enough
            unsigned shade_grid = (unsigned)((ctx->pool_cap < w.item_count ? ctx->pool_cap : w.item_count) + 255u) / 256u;
            const unsigned shade_max = (unsigned)ctx->num_sms * 16u;
            if (shade_grid > shade_max) shade_grid = shade_max;
            if (retries > 0) RT_CUDA(cudaMemsetAsync(pp.retry_counts, 0, RT_MAX_RETRIES * sizeof(unsigned), pp.stream));
            for (int pass = 0; pass <= retries; pass++)
            {
                // pass 0 generates the slice; pass k > 0 the items pass k-1 could not place
                w.retry_in = pass > 0 ? pp.retry[(pass - 1) & 1] : nullptr;
                w.retry_in_count = pass > 0 ? pp.retry_counts + (pass - 1) : nullptr;
                w.retry_out = retries > 0 ? pp.retry[pass & 1] : nullptr;
                w.retry_out_count = retries > 0 ? pp.retry_counts + pass : nullptr;
                RT_CUDA(cudaMemsetAsync(pp.round_counters, 0, (6 * RT_MAX_ROUNDS + 1) * sizeof(unsigned), pp.stream));
                RT_CUDA(cull ? launch_generate<true>(p->mode, gen_grid, pp.stream, ctx->scene, a, w)
                             : launch_generate<false>(p->mode, gen_grid, pp.stream, ctx->scene, a, w));
                ctx->launches++;
                const int wave_rounds = (ctx->tune_finish_round > 0 && ctx->tune_finish_round < rounds) ? ctx->tune_finish_round : rounds;
                for (int round = 0; round < wave_rounds; round++)
                {
                    if (mesh_shapes > 0)
                    {
                        if (ctx->time_walks)
                        {
                            while ((int)ctx->kev.size() < ctx->kev_used + 2)
                            {
                                cudaEvent_t e = nullptr;
                                RT_CUDA(cudaEventCreate(&e));
                                ctx->kev.push_back(e);
                            }
                            RT_CUDA(cudaEventRecord(ctx->kev[ctx->kev_used], pp.stream));
                        }
                        if (round < packet_rounds)
                        {
                            // packets first; what they hand back (incoherent ones) goes on lane by lane
                            if (cull) rt_walk_packet_kernel<true><<<walk_grid, 256, 0, pp.stream>>>(ctx->scene, a, w, round);
                            else rt_walk_packet_kernel<false><<<walk_grid, 256, 0, pp.stream>>>(ctx->scene, a, w, round);
                            RT_CUDA(cudaGetLastError());
                            ctx->launches++;
                            if (cull) rt_walk_kernel<true><<<walk_grid, 256, 0, pp.stream>>>(ctx->scene, a, w, round, 1);
                            else rt_walk_kernel<false><<<walk_grid, 256, 0, pp.stream>>>(ctx->scene, a, w, round, 1);
                        }
                        else if (cull) rt_walk_kernel<true><<<walk_grid, 256, 0, pp.stream>>>(ctx->scene, a, w, round, 0);
                        else rt_walk_kernel<false><<<walk_grid, 256, 0, pp.stream>>>(ctx->scene, a, w, round, 0);
                        RT_CUDA(cudaGetLastError());
                        static const bool time_long = getenv("RT_TIME_LONG") != nullptr;     // tooling: bracket walk + long walk
                        if (ctx->time_walks && !time_long)
                        {
                            RT_CUDA(cudaEventRecord(ctx->kev[ctx->kev_used + 1], pp.stream));
                            ctx->kev_used += 2;
                        }
                        ctx->launches++;
                        // the walks that kernel parked as too long, one warp each
                        {
                            const int group = ctx->tune_long_group;
                            const unsigned lgrid = (unsigned)ctx->num_sms * RT_LONG_BLOCKS;
#define RT_LAUNCH_LONG(G) (cull ? rt_longwalk_kernel<true, G><<<lgrid, 256, 0, pp.stream>>>(ctx->scene, a, w, round) \
                                : rt_longwalk_kernel<false, G><<<lgrid, 256, 0, pp.stream>>>(ctx->scene, a, w, round))
                            if (group == 32) RT_LAUNCH_LONG(32); else if (group == 16) RT_LAUNCH_LONG(16); else RT_LAUNCH_LONG(8);
#undef RT_LAUNCH_LONG
                        }
                        RT_CUDA(cudaGetLastError());
                        if (ctx->time_walks && time_long)
                        {
                            RT_CUDA(cudaEventRecord(ctx->kev[ctx->kev_used + 1], pp.stream));
                            ctx->kev_used += 2;
                        }
                        ctx->launches++;
                    }
                    RT_CUDA(cull ? launch_shade<true>(p->mode, shade_grid, pp.stream, ctx->scene, a, w, round)
                                 : launch_shade<false>(p->mode, shade_grid, pp.stream, ctx->scene, a, w, round));
                    ctx->launches++;
                }
                if (wave_rounds < rounds)
                {
                    // everything still alive after the wavefront rounds runs to its end in one launch
                    RT_CUDA(cull ? launch_finish<true>(p->mode, (unsigned)ctx->num_sms * 4u, pp.stream, ctx->scene, a, w, wave_rounds)
                                 : launch_finish<false>(p->mode, (unsigned)ctx->num_sms * 4u, pp.stream, ctx->scene, a, w, wave_rounds));
                    ctx->launches++;
                }
            }
        }
        if (p->mode != RT_MODE_PRIMARY)
        {
            if (last_pipe >= 0 && last_pipe != pipe) RT_CUDA(cudaStreamWaitEvent(pp.stream, ctx->pipes[last_pipe].done, 0));
            const int n = p->end - p->start + 1;
            rt_resolve_kernel<<<(n + 255) / 256, 256, 0, pp.stream>>>(a, chunk);
            RT_CUDA(cudaGetLastError());
            ctx->launches++;
        }
        RT_CUDA(cudaEventRecord(pp.done, pp.stream));
        last_pipe = pipe;
    }
    // join: the context stream continues after every pipe
    for (int k = 0; k < RT_PIPES; k++)
        if (used[k]) RT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pipes[k].done, 0));
    RT_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    return RT_OK;
}

int rt_gpu_synchronize(rt_gpu_ctx* ctx)
{
    if (!ctx) return RT_ERR_INVALID;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_gpu_readback(rt_gpu_ctx* ctx, int what, void* dst, size_t bytes)
{
    if (!ctx) return RT_ERR_INVALID;
    if (!dst) return fail(ctx, RT_ERR_INVALID, "dst is null");
    RT_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->width * ctx->height;
    const void* src = nullptr; size_t need = 0;
    switch (what)
    {
    case RT_READ_ACCUM_RGBN_F32: src = ctx->accum; need = n * sizeof(float4); break;
    case RT_READ_DISPLAY_ARGB8: src = ctx->display; need = n * sizeof(uint32_t); break;
    case RT_READ_PRIMARY_IDS_I32X2: src = ctx->prim_ids; need = n * sizeof(int2); break;
    case RT_READ_PRIMARY_DIST_F32: src = ctx->prim_dist; need = n * sizeof(float); break;
    case RT_READ_COUNTERS_U64: src = ctx->counters; need = sizeof(rt_counters); break;
    default: return fail(ctx, RT_ERR_INVALID, "unknown readback selector");
    }
    if (what != RT_READ_COUNTERS_U64 && n == 0) return fail(ctx, RT_ERR_NO_SCENE, "no frame buffers yet (render or reset_accum first)");
    if (bytes < need) return fail(ctx, RT_ERR_SIZE, "readback buffer too small");
    RT_CUDA(cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_gpu_last_render_ms(rt_gpu_ctx* ctx, float* out_ms)
{
    if (!ctx || !out_ms) return RT_ERR_INVALID;
    if (!ctx->timed) return fail(ctx, RT_ERR_INVALID, "no render has been enqueued");
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaEventSynchronize(ctx->ev1));
    RT_CUDA(cudaEventElapsedTime(out_ms, ctx->ev0, ctx->ev1));
    return RT_OK;
}

int rt_gpu_last_kernel_ms(rt_gpu_ctx* ctx, float* out_ms, int32_t* out_launches)
{
    if (!ctx || !out_ms) return RT_ERR_INVALID;
    if (!ctx->timed) return fail(ctx, RT_ERR_INVALID, "no render has been enqueued");
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaEventSynchronize(ctx->ev1));
    float total = 0.0f;
    for (int k = 0; k + 1 < ctx->kev_used; k += 2)
    {
        float ms = 0.0f;
        RT_CUDA(cudaEventElapsedTime(&ms, ctx->kev[k], ctx->kev[k + 1]));
        total += ms;
    }
    *out_ms = total;
    if (out_launches) *out_launches = ctx->kev_used / 2;
    return RT_OK;
}

/* tooling: entries and walk-kernel time of each round of the last batch / call */
int rt_gpu_debug_rounds(rt_gpu_ctx* ctx, uint32_t* counts, float* ms, int32_t max_rounds)
{
    if (!ctx || !counts || !ms || max_rounds <= 0) return RT_ERR_INVALID;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n = max_rounds < RT_MAX_ROUNDS ? max_rounds : RT_MAX_ROUNDS;
    RT_CUDA(cudaMemcpy(counts, ctx->pipes[0].round_counters, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; k++)
    {
        ms[k] = 0.0f;
        if (2 * k + 1 < ctx->kev_used) RT_CUDA(cudaEventElapsedTime(&ms[k], ctx->kev[2 * k], ctx->kev[2 * k + 1]));
    }
    unsigned longest = 0;
    RT_CUDA(cudaMemcpy(&longest, ctx->pipes[0].round_counters + RT_MAX_ROUNDS, sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (n > 0) counts[n - 1] = longest;        // last slot: longest single walk (nodes) of the batch
    return ctx->kev_used / 2;
}

/* tooling: begin/end of every timed walk bracket of the last call, in ms since the call began (launch order:
   chunk by chunk, round by round); returns the number of brackets */
int rt_gpu_debug_timeline(rt_gpu_ctx* ctx, float* begin_ms, float* end_ms, int32_t cap)
{
    if (!ctx || !begin_ms || !end_ms || cap <= 0) return RT_ERR_INVALID;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n = ctx->kev_used / 2 < cap ? ctx->kev_used / 2 : cap;
    for (int k = 0; k < n; k++)
    {
        RT_CUDA(cudaEventElapsedTime(&begin_ms[k], ctx->ev0, ctx->kev[2 * k]));
        RT_CUDA(cudaEventElapsedTime(&end_ms[k], ctx->ev0, ctx->kev[2 * k + 1]));
    }
    return n;
}

/* tooling: long-walk queue sizes per round of the last batch on pipe 0 */
int rt_gpu_debug_long(rt_gpu_ctx* ctx, uint32_t* lcounts, int32_t max_rounds)
{
    if (!ctx || !lcounts || max_rounds <= 0) return RT_ERR_INVALID;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n = max_rounds < RT_MAX_ROUNDS ? max_rounds : RT_MAX_ROUNDS;
    RT_CUDA(cudaMemcpy(lcounts, ctx->pipes[0].round_counters + 2 * RT_MAX_ROUNDS + 1, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_gpu_reset_counters(rt_gpu_ctx* ctx)
{
    if (!ctx) return RT_ERR_INVALID;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaMemsetAsync(ctx->counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    return RT_OK;
}

int64_t rt_gpu_owned_pixels(int32_t width, int32_t height, int32_t tile_size, int32_t tile_count, int32_t tile_rank)
{
    if (width <= 0 || height <= 0) return 0;
    if (tile_count <= 1 || tile_size <= 0) return (int64_t)width * height;
    const int tiles_x = (width + tile_size - 1) / tile_size, tiles_y = (height + tile_size - 1) / tile_size;
    int64_t total = 0;
    for (int t = tile_rank; t < tiles_x * tiles_y; t += tile_count)
    {
        const int tx = t % tiles_x, ty = t / tiles_x;
        const int w = tile_size < width - tx * tile_size ? tile_size : width - tx * tile_size;
        const int h = tile_size < height - ty * tile_size ? tile_size : height - ty * tile_size;
        total += (int64_t)w * h;
    }
    return total;
}

static int tile_copy(rt_gpu_ctx* ctx, const rt_render_params* p, int rank, float4* dense, size_t bytes, int dir)
{
    if (!ctx || !p || !dense) return RT_ERR_INVALID;
    if (p->width != ctx->width || p->height != ctx->height) return fail(ctx, RT_ERR_INVALID, "frame size mismatch");
    RT_CUDA(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)p->width * p->height;
    if (p->tile_count <= 1 || p->tile_size <= 0)
    {
        if (bytes < npix * sizeof(float4)) return fail(ctx, RT_ERR_SIZE, "dense buffer too small");
        if (dir == 0) RT_CUDA(cudaMemcpyAsync(dense, ctx->accum, npix * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
        else RT_CUDA(cudaMemcpyAsync(ctx->accum, dense, npix * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
        return RT_OK;
    }
    if (rank < 0 || rank >= p->tile_count) return fail(ctx, RT_ERR_INVALID, "rank out of range");
    TileArgs t;
    t.width = p->width; t.height = p->height; t.tile_size = p->tile_size; t.tile_count = p->tile_count; t.tile_rank = rank;
    t.tiles_x = (p->width + p->tile_size - 1) / p->tile_size; t.tiles_y = (p->height + p->tile_size - 1) / p->tile_size;
    const int ntiles = t.tiles_x * t.tiles_y;
    std::vector<long long> offs;
    long long total = 0;
    for (int tile = rank; tile < ntiles; tile += p->tile_count)
    {
        const int tx = tile % t.tiles_x, ty = tile / t.tiles_x;
        const int w = p->tile_size < p->width - tx * p->tile_size ? p->tile_size : p->width - tx * p->tile_size;
        const int h = p->tile_size < p->height - ty * p->tile_size ? p->tile_size : p->height - ty * p->tile_size;
        offs.push_back(total);
        total += (long long)w * h;
    }
    if (bytes < (size_t)total * sizeof(float4)) return fail(ctx, RT_ERR_SIZE, "dense buffer too small");
    if (offs.empty()) return RT_OK;
    // per-(frame, tiling, rank) offset tables are uploaded once and kept: the exchange then needs no
    // host synchronisation at all
    const long long* dev_offsets = nullptr;
    for (const rt_gpu_ctx::TileTable& tt : ctx->tile_tables)
        if (tt.width == p->width && tt.height == p->height && tt.tile_size == p->tile_size && tt.tile_count == p->tile_count && tt.rank == rank)
            dev_offsets = tt.offsets;
    if (!dev_offsets)
    {
        rt_gpu_ctx::TileTable tt;
        tt.width = p->width; tt.height = p->height; tt.tile_size = p->tile_size; tt.tile_count = p->tile_count; tt.rank = rank;
        RT_CUDA(cudaMalloc((void**)&tt.offsets, offs.size() * sizeof(long long)));
        RT_CUDA(cudaMemcpy(tt.offsets, offs.data(), offs.size() * sizeof(long long), cudaMemcpyHostToDevice));
        ctx->tile_tables.push_back(tt);
        dev_offsets = tt.offsets;
    }
    rt_tile_copy_kernel<<<(unsigned)offs.size(), 256, 0, ctx->stream>>>(ctx->accum, dense, dev_offsets, t, dir);
    RT_CUDA(cudaGetLastError());
    ctx->launches++;
    return RT_OK;
}

int rt_gpu_pack_owned(rt_gpu_ctx* ctx, const rt_render_params* p, void* dev_ptr, size_t bytes)
{
    if (!ctx || !p) return RT_ERR_INVALID;
    return tile_copy(ctx, p, p->tile_rank, (float4*)dev_ptr, bytes, 0);
}

int rt_gpu_unpack_owned(rt_gpu_ctx* ctx, const rt_render_params* p, int32_t src_rank, const void* dev_ptr, size_t bytes)
{
    if (!ctx || !p) return RT_ERR_INVALID;
    return tile_copy(ctx, p, src_rank, (float4*)dev_ptr, bytes, 1);
}

/* Peer-memory exchange: the root exports its accumulation buffer (CUDA IPC), every other rank maps it and
   writes its owned tiles into it directly. */
int rt_gpu_export_frame(rt_gpu_ctx* ctx, void* handle64, size_t bytes)
{
    if (!ctx || !handle64) return RT_ERR_INVALID;
    if (bytes < sizeof(cudaIpcMemHandle_t)) return fail(ctx, RT_ERR_SIZE, "handle buffer too small (64 bytes)");
    if (!ctx->accum) return fail(ctx, RT_ERR_NO_SCENE, "no frame buffers yet (rt_gpu_reset_accum first)");
    RT_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    RT_CUDA(cudaIpcGetMemHandle(&h, ctx->accum));
    memcpy(handle64, &h, sizeof h);
    return RT_OK;
}

int rt_gpu_open_peer_frame(rt_gpu_ctx* ctx, const void* handle64, size_t bytes, void** out_dev_ptr)
{
    if (!ctx || !handle64 || !out_dev_ptr) return RT_ERR_INVALID;
    if (bytes < sizeof(cudaIpcMemHandle_t)) return fail(ctx, RT_ERR_SIZE, "handle buffer too small (64 bytes)");
    RT_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    void* ptr = nullptr;
    RT_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    *out_dev_ptr = ptr;
    return RT_OK;
}

int rt_gpu_close_peer_frame(rt_gpu_ctx* ctx, void* dev_ptr)
{
    if (!ctx || !dev_ptr) return RT_ERR_INVALID;
    RT_CUDA(cudaSetDevice(ctx->device));
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    RT_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return RT_OK;
}

int rt_gpu_push_owned(rt_gpu_ctx* ctx, const rt_render_params* p, void* peer_frame)
{
    if (!ctx || !p || !peer_frame) return RT_ERR_INVALID;
    if (p->width != ctx->width || p->height != ctx->height) return fail(ctx, RT_ERR_INVALID, "frame size mismatch");
    RT_CUDA(cudaSetDevice(ctx->device));
    if (p->tile_count <= 1 || p->tile_size <= 0)
    {
        RT_CUDA(cudaMemcpyAsync(peer_frame, ctx->accum, (size_t)p->width * p->height * sizeof(float4), cudaMemcpyDefault, ctx->stream));
        return RT_OK;
    }
    if (p->tile_rank < 0 || p->tile_rank >= p->tile_count) return fail(ctx, RT_ERR_INVALID, "rank out of range");
    TileArgs t;
    t.width = p->width; t.height = p->height; t.tile_size = p->tile_size; t.tile_count = p->tile_count; t.tile_rank = p->tile_rank;
    t.tiles_x = (p->width + p->tile_size - 1) / p->tile_size; t.tiles_y = (p->height + p->tile_size - 1) / p->tile_size;
    const int ntiles = t.tiles_x * t.tiles_y;
    const int owned = p->tile_rank < ntiles ? (ntiles - p->tile_rank + p->tile_count - 1) / p->tile_count : 0;
    if (owned == 0) return RT_OK;
    rt_tile_push_kernel<<<(unsigned)owned, 256, 0, ctx->stream>>>(ctx->accum, (float4*)peer_frame, t);
    RT_CUDA(cudaGetLastError());
    ctx->launches++;
    return RT_OK;
}

int rt_gpu_gather(rt_gpu_ctx** ctxs, int n, int root, const rt_render_params* p)
{
    if (!ctxs || n <= 0 || root < 0 || root >= n || !p) return RT_ERR_INVALID;
    rt_gpu_ctx* ctx = ctxs[root];
    if (!ctx) return RT_ERR_INVALID;
    if (p->tile_count != n && n > 1) return fail(ctx, RT_ERR_INVALID, "tile_count must equal the number of contexts");
    if (n == 1) return RT_OK;
    for (int r = 0; r < n; r++)
    {
        if (r == root) continue;
        rt_gpu_ctx* src = ctxs[r];
        if (!src) return fail(ctx, RT_ERR_INVALID, "null context in gather");
        rt_render_params q = *p; q.tile_rank = r;
        const size_t count = (size_t)rt_gpu_owned_pixels(p->width, p->height, p->tile_size, p->tile_count, r);
        if (count == 0) continue;
        // pack on the source GPU
        if (count > src->gather_staging_cap)
        {
            cudaSetDevice(src->device);
            cudaStreamSynchronize(src->stream);
            cudaFree(src->gather_staging); src->gather_staging = nullptr; src->gather_staging_cap = 0;
            if (cudaMalloc((void**)&src->gather_staging, count * sizeof(float4)) != cudaSuccess)
                return fail(ctx, RT_ERR_NOMEM, "gather staging allocation failed");
            src->gather_staging_cap = count;
        }
        int rc = rt_gpu_pack_owned(src, &q, src->gather_staging, count * sizeof(float4));
        if (rc != RT_OK) return fail(ctx, rc, std::string("pack on source failed: ") + src->err);
        cudaSetDevice(src->device);
        if (cudaStreamSynchronize(src->stream) != cudaSuccess) return fail(ctx, RT_ERR_CUDA, "source stream sync failed");
        // move over NVLink into root staging, then scatter
        RT_CUDA(cudaSetDevice(ctx->device));
        if (count > ctx->gather_staging_cap)
        {
            RT_CUDA(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->gather_staging); ctx->gather_staging = nullptr; ctx->gather_staging_cap = 0;
            RT_CUDA(cudaMalloc((void**)&ctx->gather_staging, count * sizeof(float4)));
            ctx->gather_staging_cap = count;
        }
        RT_CUDA(cudaMemcpyPeerAsync(ctx->gather_staging, ctx->device, src->gather_staging, src->device, count * sizeof(float4), ctx->stream));
        rc = rt_gpu_unpack_owned(ctx, &q, r, ctx->gather_staging, count * sizeof(float4));
        if (rc != RT_OK) return rc;
        RT_CUDA(cudaStreamSynchronize(ctx->stream));    // staging is reused for the next rank
    }
    return RT_OK;
}

int rt_gpu_resolve_display(rt_gpu_ctx* ctx)
{
    if (!ctx) return RT_ERR_INVALID;
    const int n = ctx->width * ctx->height;
    if (n == 0) return fail(ctx, RT_ERR_NO_SCENE, "no frame buffers yet");
    RT_CUDA(cudaSetDevice(ctx->device));
    rt_display_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->accum, ctx->display, n);
    RT_CUDA(cudaGetLastError());
    ctx->launches++;
    return RT_OK;
}

void* rt_gpu_stream(rt_gpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

/* used by the other translation units of the library (rt_bvh_build.cu) */
int rt_gpu_device_of(rt_gpu_ctx* ctx) { return ctx ? ctx->device : -1; }
void rt_gpu_set_error(rt_gpu_ctx* ctx, const char* msg) { if (ctx && msg) ctx->err = msg; }

void* rt_gpu_accum_device_ptr(rt_gpu_ctx* ctx) { return ctx ? (void*)ctx->accum : nullptr; }

uint64_t rt_gpu_launch_count(rt_gpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

// experiment knobs (tools/, tests that force the long-walk path); the defaults are what ships
static void tuning_from_env(rt_gpu_ctx* ctx)
{
    if (getenv("RT_FINISH_ROUND")) ctx->tune_finish_round = atoi(getenv("RT_FINISH_ROUND"));
    if (getenv("RT_LONG_LIMIT")) ctx->tune_long_limit = (unsigned)atoi(getenv("RT_LONG_LIMIT"));
    if (getenv("RT_SMALL_ROUND")) ctx->tune_small_round = (unsigned)atoi(getenv("RT_SMALL_ROUND"));
    if (getenv("RT_THIN_COUNT")) ctx->tune_thin_count = (unsigned)atoi(getenv("RT_THIN_COUNT"));
    if (getenv("RT_LONG_GROUP_N")) ctx->tune_long_group = atoi(getenv("RT_LONG_GROUP_N"));
    if (getenv("RT_PACKET_ROUNDS")) ctx->tune_packet_rounds = atoi(getenv("RT_PACKET_ROUNDS"));
    if (getenv("RT_PACKET_PROBE")) ctx->tune_packet_probe = (unsigned)atoi(getenv("RT_PACKET_PROBE"));
    if (getenv("RT_PACKET_MIN_LANES")) ctx->tune_packet_min_lanes = (unsigned)atoi(getenv("RT_PACKET_MIN_LANES"));
    if (getenv("RT_THIN_LIMIT")) ctx->tune_thin_limit = (unsigned)atoi(getenv("RT_THIN_LIMIT"));
}

int rt_gpu_set_tuning(rt_gpu_ctx* ctx, int32_t window_items, int32_t min_lanes, int32_t leaf_wait, int32_t pool_kpaths)
{
    if (!ctx) return RT_ERR_INVALID;
    if (window_items < 32 || window_items % 32 != 0 || min_lanes < 1 || min_lanes > 32)
        return fail(ctx, RT_ERR_INVALID, "window_items must be a positive multiple of 32, min_lanes in [1, 32]");
    ctx->tune_window = (unsigned)window_items;
    ctx->tune_min_lanes = min_lanes;
    ctx->tune_leaf_wait = leaf_wait < 0 ? 0 : (leaf_wait > 32 ? 32 : leaf_wait);
    if (pool_kpaths > 0) ctx->max_pool_paths = (size_t)pool_kpaths << 10;
    tuning_from_env(ctx);
    return RT_OK;
}

int rt_gpu_time_kernels(rt_gpu_ctx* ctx, int32_t on)
{
    if (!ctx) return RT_ERR_INVALID;
    ctx->time_walks = on != 0;
    return RT_OK;
}

int rt_gpu_set_pipes(rt_gpu_ctx* ctx, int32_t pipes)
{
    if (!ctx) return RT_ERR_INVALID;
    if (pipes < 1 || pipes > RT_PIPES) return fail(ctx, RT_ERR_INVALID, "pipes out of range");
    ctx->tune_pipes = pipes;
    return RT_OK;
}

uint64_t rt_gpu_scene_bytes(rt_gpu_ctx* ctx) { return ctx ? (uint64_t)ctx->scene_bytes : 0; }

int rt_gpu_trace_rays(rt_gpu_ctx* ctx, const float* rays, int32_t n, int32_t traverse, int32_t* shape, int32_t* tri, float* hit11)
{
    if (!ctx) return RT_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, RT_ERR_NO_SCENE, "rt_gpu_trace_rays before rt_gpu_upload_scene");
    if (n < 0 || (n > 0 && (!rays || !shape || !tri || !hit11))) return fail(ctx, RT_ERR_INVALID, "bad arguments");
    if (n == 0) return RT_OK;
    RT_CUDA(cudaSetDevice(ctx->device));
    float* drays = nullptr; int* dshape = nullptr; int* dtri = nullptr; float* dhit = nullptr;
    RT_CUDA(cudaMalloc((void**)&drays, (size_t)n * 7 * sizeof(float)));
    cudaError_t e = cudaMalloc((void**)&dshape, (size_t)n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dtri, (size_t)n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dhit, (size_t)n * 11 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(drays, rays, (size_t)n * 7 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
    {
        const int exact = traverse == RT_TRAVERSE_EXACT ? 1 : 0;
        if (traverse == RT_TRAVERSE_CULLED)
            rt_trace_rays_kernel<true><<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, drays, n, dshape, dtri, dhit, ctx->counters, exact);
        else
            rt_trace_rays_kernel<false><<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene, drays, n, dshape, dtri, dhit, ctx->counters, exact);
        e = cudaGetLastError();
        ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(shape, dshape, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tri, dtri, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hit11, dhit, (size_t)n * 11 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(drays); cudaFree(dshape); cudaFree(dtri); cudaFree(dhit);
    if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, std::string("rt_gpu_trace_rays: ") + cudaGetErrorString(e));
    return RT_OK;
}

int rt_gpu_kat(rt_gpu_ctx* ctx, int32_t kind, const float* rays, const float* prims, int32_t prim_floats, int32_t n,
               int32_t* flags, float* out7)
{
    if (!ctx) return RT_ERR_INVALID;
    if (n <= 0 || !prims || !flags || !out7 || kind < 0 || kind > 7) return fail(ctx, RT_ERR_INVALID, "bad arguments");
    RT_CUDA(cudaSetDevice(ctx->device));
    float* drays = nullptr; float* dprims = nullptr; int* dflags = nullptr; float* dout = nullptr;
    cudaError_t e = cudaSuccess;
    if (rays) { e = cudaMalloc((void**)&drays, (size_t)n * 7 * sizeof(float)); if (e == cudaSuccess) e = cudaMemcpy(drays, rays, (size_t)n * 7 * sizeof(float), cudaMemcpyHostToDevice); }
    if (e == cudaSuccess) e = cudaMalloc((void**)&dprims, (size_t)n * prim_floats * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(dprims, prims, (size_t)n * prim_floats * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dflags, (size_t)n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dout, (size_t)n * 7 * sizeof(float));
    if (e == cudaSuccess)
    {
        rt_kat_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(kind, drays, dprims, n, dflags, dout);
        e = cudaGetLastError();
        ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy(flags, dflags, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out7, dout, (size_t)n * 7 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(drays); cudaFree(dprims); cudaFree(dflags); cudaFree(dout);
    if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, std::string("rt_gpu_kat: ") + cudaGetErrorString(e));
    return RT_OK;
}

int rt_gpu_kat_texture(rt_gpu_ctx* ctx, int32_t texture, const float* uv, int32_t n, float* out4)
{
    if (!ctx) return RT_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, RT_ERR_NO_SCENE, "no scene");
    if (texture < 0 || texture >= (int)ctx->host_textures.size() || n <= 0 || !uv || !out4) return fail(ctx, RT_ERR_INVALID, "bad arguments");
    RT_CUDA(cudaSetDevice(ctx->device));
    float* duv = nullptr; float* dout = nullptr;
    cudaError_t e = cudaMalloc((void**)&duv, (size_t)n * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dout, (size_t)n * 4 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(duv, uv, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        rt_kat_texture_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->scene.atlas, ctx->host_textures[texture], duv, n, dout);
        e = cudaGetLastError();
        ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out4, dout, (size_t)n * 4 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(duv); cudaFree(dout);
    if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, std::string("rt_gpu_kat_texture: ") + cudaGetErrorString(e));
    return RT_OK;
}

} // extern "C"
