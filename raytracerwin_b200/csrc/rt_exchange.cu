// rt_exchange.cu — the one exchange step of a multi-GPU frame: a rank's owned tiles are packed
// into a dense buffer (for an NCCL / peer-copy gather), unpacked on the root, or written straight into a
// peer's frame over NVLink.  Replaces nothing in the reference (it has one address space); the tiles play the
// role of its 10-row RenderThreadTasks (RayTracerProgram.cpp:294-302).
#include "rt_context.hpp"

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

// one CTA per owned tile: accuBuffer and bitcolor of this rank's pixels written into a HOST frame (registered,
// mapped memory) over this GPU's own PCIe link — every rank delivers its share of the frame in parallel
__global__ void rt_tile_deliver_kernel(const float4* __restrict__ accum, const uint32_t* __restrict__ display,
                                       float4* __restrict__ host_accum, uint32_t* __restrict__ host_display, TileArgs t)
{
    const int tile = t.tile_rank + blockIdx.x * t.tile_count;
    const int tx = tile % t.tiles_x, ty = tile / t.tiles_x;
    const int ox = tx * t.tile_size, oy = ty * t.tile_size;
    const int w = min(t.tile_size, t.width - ox), h = min(t.tile_size, t.height - oy);
    for (int i = threadIdx.x; i < w * h; i += blockDim.x)
    {
        const int lx = i % w, ly = i / w;
        const size_t f = (size_t)(oy + ly) * t.width + (ox + lx);
        if (host_accum) host_accum[f] = accum[f];
        if (host_display) host_display[f] = display[f];
    }
}

// stream-ordered "this much has been delivered": one word of a registered host frame's header
__global__ void rt_signal_host_kernel(volatile uint32_t* word, uint32_t value)
{
    __threadfence_system();
    *word = value;
}

extern "C" {

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
        // on the context's stream, which is non-blocking: a plain cudaMemcpy from pageable memory would not be
        // ordered before the kernel below; `offs` is a local, so wait once (tables are cached from here on)
        RT_CUDA(cudaMemcpyAsync(tt.offsets, offs.data(), offs.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        RT_CUDA(cudaStreamSynchronize(ctx->stream));
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
    return rt_guard(ctx, [&]() -> int {
        if (!ctx || !p) return RT_ERR_INVALID;
        return tile_copy(ctx, p, p->tile_rank, (float4*)dev_ptr, bytes, 0);
    });
}

int rt_gpu_unpack_owned(rt_gpu_ctx* ctx, const rt_render_params* p, int32_t src_rank, const void* dev_ptr, size_t bytes)
{
    return rt_guard(ctx, [&]() -> int {
        if (!ctx || !p) return RT_ERR_INVALID;
        return tile_copy(ctx, p, src_rank, (float4*)dev_ptr, bytes, 1);
    });
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
    return rt_guard(ctx, [&]() -> int {
        if (!ctx || !handle64 || !out_dev_ptr) return RT_ERR_INVALID;
        if (bytes < sizeof(cudaIpcMemHandle_t)) return fail(ctx, RT_ERR_SIZE, "handle buffer too small (64 bytes)");
        RT_CUDA(cudaSetDevice(ctx->device));
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, sizeof h);
        void* ptr = nullptr;
        RT_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        *out_dev_ptr = ptr;
        return RT_OK;
    });
}

int rt_gpu_close_peer_frame(rt_gpu_ctx* ctx, void* dev_ptr)
{
    return rt_guard(ctx, [&]() -> int {
        if (!ctx || !dev_ptr) return RT_ERR_INVALID;
        RT_CUDA(cudaSetDevice(ctx->device));
        RT_CUDA(cudaStreamSynchronize(ctx->stream));
        RT_CUDA(cudaIpcCloseMemHandle(dev_ptr));
        return RT_OK;
    });
}

int rt_gpu_push_owned(rt_gpu_ctx* ctx, const rt_render_params* p, void* peer_frame)
{
    return rt_guard(ctx, [&]() -> int {
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
    });
}

int rt_gpu_register_host_frame(rt_gpu_ctx* ctx, void* host, size_t bytes, void** out_dev_ptr)
{
    return rt_guard(ctx, [&]() -> int {
        if (!ctx || !host || bytes == 0 || !out_dev_ptr) return RT_ERR_INVALID;
        RT_CUDA(cudaSetDevice(ctx->device));
        RT_CUDA(cudaHostRegister(host, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
        void* dev = nullptr;
        const cudaError_t e = cudaHostGetDevicePointer(&dev, host, 0);
        if (e != cudaSuccess) { cudaHostUnregister(host); RT_CUDA(e); }
        *out_dev_ptr = dev;
        return RT_OK;
    });
}

int rt_gpu_unregister_host_frame(rt_gpu_ctx* ctx, void* host)
{
    return rt_guard(ctx, [&]() -> int {
        if (!ctx || !host) return RT_ERR_INVALID;
        RT_CUDA(cudaSetDevice(ctx->device));
        RT_CUDA(cudaStreamSynchronize(ctx->stream));
        RT_CUDA(cudaHostUnregister(host));
        return RT_OK;
    });
}

int rt_gpu_deliver_owned(rt_gpu_ctx* ctx, const rt_render_params* p, void* host_accum, void* host_display)
{
    return rt_guard(ctx, [&]() -> int {
        if (!ctx || !p || (!host_accum && !host_display)) return RT_ERR_INVALID;
        if (p->width != ctx->width || p->height != ctx->height) return fail(ctx, RT_ERR_INVALID, "frame size mismatch");
        RT_CUDA(cudaSetDevice(ctx->device));
        const size_t npix = (size_t)p->width * p->height;
        if (p->tile_count <= 1 || p->tile_size <= 0)
        {
            if (host_accum) RT_CUDA(cudaMemcpyAsync(host_accum, ctx->accum, npix * sizeof(float4), cudaMemcpyDefault, ctx->stream));
            if (host_display) RT_CUDA(cudaMemcpyAsync(host_display, ctx->display, npix * sizeof(uint32_t), cudaMemcpyDefault, ctx->stream));
            return RT_OK;
        }
        if (p->tile_rank < 0 || p->tile_rank >= p->tile_count) return fail(ctx, RT_ERR_INVALID, "rank out of range");
        TileArgs t;
        t.width = p->width; t.height = p->height; t.tile_size = p->tile_size; t.tile_count = p->tile_count; t.tile_rank = p->tile_rank;
        t.tiles_x = (p->width + p->tile_size - 1) / p->tile_size; t.tiles_y = (p->height + p->tile_size - 1) / p->tile_size;
        const int ntiles = t.tiles_x * t.tiles_y;
        const int owned = p->tile_rank < ntiles ? (ntiles - p->tile_rank + p->tile_count - 1) / p->tile_count : 0;
        if (owned == 0) return RT_OK;
        rt_tile_deliver_kernel<<<(unsigned)owned, 256, 0, ctx->stream>>>(ctx->accum, ctx->display, (float4*)host_accum, (uint32_t*)host_display, t);
        RT_CUDA(cudaGetLastError());
        ctx->launches++;
        return RT_OK;
    });
}

int rt_gpu_signal_host(rt_gpu_ctx* ctx, void* host_word_dev, uint32_t value)
{
    return rt_guard(ctx, [&]() -> int {
        if (!ctx || !host_word_dev) return RT_ERR_INVALID;
        RT_CUDA(cudaSetDevice(ctx->device));
        rt_signal_host_kernel<<<1, 1, 0, ctx->stream>>>((volatile uint32_t*)host_word_dev, value);
        RT_CUDA(cudaGetLastError());
        ctx->launches++;
        return RT_OK;
    });
}

int rt_gpu_gather(rt_gpu_ctx** ctxs, int n, int root, const rt_render_params* p)
{
    return rt_guard(ctxs && n > 0 && root >= 0 && root < n ? ctxs[root] : nullptr, [&]() -> int {
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
    });
}

} // extern "C"
