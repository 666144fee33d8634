// ort_b200.cu -- the extern "C" boundary (include/ort_b200.h) over the sm_100a
// kernels: scene upload, render entry points, batched ray cast.
// There is no CPU path in this file: without a CUDA device every entry point
// that computes returns ORT_ERR_CUDA.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <chrono>
#include <algorithm>
#include <mutex>

#include "ort_internal.h"
#include "wavefront.cuh"
#include "bvh_build.cuh"
#include "scene_flatten.h"

using namespace ort;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if(e__ != cudaSuccess)                                                             \
            return fail(ORT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while(0)

} // namespace

#ifndef WF_BATCH
#define WF_BATCH 8
#endif
#define WF_MAX_POOLS 4

// A pool = a contiguous share of the slots with its own stream, sort scratch and counters.  Two
// pools run the same loop half a period apart, so that one pool's SHADE (latency-bound) overlaps the
// other's EXTEND (issue-bound) and every kernel's tail is filled by the other stream's blocks.
struct WfPool
{
    WfBuffers wf;
    uint32_t *d_sort;                       // hist[512] | cursor[512] | live | chunk counter
    unsigned int *d_active, *h_active;      // 2 x WF_BATCH "slots still active" counters
    cudaStream_t stream;
    cudaEvent_t done[2], ev[2][WF_BATCH][4];
    int cur;
    bool finished;
};

struct OrtScene
{
    int device;
    OrtSceneInfo info;
    OrtBuildStats build_stats;
    uint32_t main_root, tri_root;
    uint32_t *d_rank_to_prim;
    // device arrays (layout: bvh.h, scene_flatten.h)
    q4 *d_nodes, *d_prims, *d_cyl, *d_materials;
    uint8_t *d_light_is_sphere;
    uint32_t node_count, prim_count, light_count;
    // scratch owned by the handle
    unsigned long long *d_stats;        // STAT_COUNT counters
    long long *d_accum; size_t accum_pixels;
    float *d_rgb; size_t rgb_pixels;
    uint32_t *d_rgbe = 0; size_t rgbe_capacity = 0;      // RGBE words in file order (ort_render_rgbe)
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int sm_count;
    int mega_blocks_per_sm, mega_blocks_per_sm_count;
    int wf_extend_blocks;
    uint32_t extend_launches;           // EXTEND launches of the last wavefront render
    uint32_t stack_rows;                // wide-tree depth + 1: rows of the shared-memory traversal stack
    float extend_ms, shade_ms, sort_ms; // summed stage times of the last wavefront render
    // wavefront path pool(s)
    WfBuffers wf;                       // the whole allocation; pools[] view shares of it
    unsigned int *d_active;             // per-iteration "slots still active" counters
    unsigned int *h_active;             // pinned mirror
    uint32_t *d_sort;                   // per pool: hist[512] | cursor[512] | live | chunk counter
    WfPool pools[WF_MAX_POOLS];
    int wf_ready;
    cudaEvent_t wf_start;
    // The handle's scratch (framebuffers, counters, slot pools, streams, events) is shared by all
    // entry points, so every entry point that touches it holds this lock for the whole call: the
    // handle may be used from any number of host threads -- e.g. the reference's nine tile workers
    // (code/macos_main.mm:574-598) -- and their calls are serialised (the GPU would serialise them anyway).
    std::recursive_mutex *mu;

    SceneView view() const
    {
        SceneView v;
        v.nodes = d_nodes; v.prims = d_prims; v.cyl = d_cyl;
        v.node_count = node_count; v.prim_count = prim_count; v.main_root = main_root; v.tri_root = tri_root;
        return v;
    }
};

namespace {

struct SceneLock
{
    std::lock_guard<std::recursive_mutex> guard;
    explicit SceneLock(OrtScene *s) : guard(*s->mu) {}
};

template <typename T>
int upload(T **dst, const void *src, size_t bytes, uint64_t *total)
{
    *dst = 0;
    size_t alloc = bytes ? bytes : 16;
    CUDA_TRY(cudaMalloc((void **)dst, alloc));
    if(bytes) CUDA_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    *total += alloc;
    return ORT_OK;
}

int ensure_accum(OrtScene *s, size_t pixels)
{
    if(s->accum_pixels >= pixels) return ORT_OK;
    if(s->d_accum) cudaFree(s->d_accum);
    s->d_accum = 0; s->accum_pixels = 0;
    CUDA_TRY(cudaMalloc((void **)&s->d_accum, pixels * 4 * sizeof(long long)));
    s->accum_pixels = pixels;
    return ORT_OK;
}

int ensure_rgb(OrtScene *s, size_t pixels)
{
    if(s->rgb_pixels >= pixels) return ORT_OK;
    if(s->d_rgb) cudaFree(s->d_rgb);
    s->d_rgb = 0; s->rgb_pixels = 0;
    CUDA_TRY(cudaMalloc((void **)&s->d_rgb, pixels * 3 * sizeof(float)));
    s->rgb_pixels = pixels;
    return ORT_OK;
}

int check_params(const OrtRenderParams *P)
{
    if(!P) return fail(ORT_ERR_ARG, "null params");
    if(P->output_width <= 0 || P->output_height <= 0) return fail(ORT_ERR_ARG, "output size must be positive");
    if((int64_t)P->output_width * P->output_height >= (int64_t)WF_PIXEL_MASK) return fail(ORT_ERR_ARG, "image too large (2^29 - 1 pixels at most)");
    if(P->tile_min_x < 0 || P->tile_min_y < 0 || P->tile_one_past_max_x > P->output_width ||
       P->tile_one_past_max_y > P->output_height || P->tile_min_x > P->tile_one_past_max_x ||
       P->tile_min_y > P->tile_one_past_max_y)
        return fail(ORT_ERR_ARG, "tile rect outside the image");
    if(P->ray_per_pixel_count == 0) return fail(ORT_ERR_ARG, "ray_per_pixel_count must be > 0");
    if(P->kernel > ORT_KERNEL_WAVEFRONT) return fail(ORT_ERR_ARG, "unknown kernel id");
    return ORT_OK;
}

struct ChunkPlan { uint32_t chunk_spp, n_chunks, begin, count; };

ChunkPlan plan_chunks(const OrtRenderParams *P)
{
    ChunkPlan c;
    uint32_t spp = P->ray_per_pixel_count;
    c.chunk_spp = P->chunk_spp ? P->chunk_spp : spp;
    if(c.chunk_spp > spp) c.chunk_spp = spp;
    c.n_chunks = (spp + c.chunk_spp - 1) / c.chunk_spp;
    uint32_t b = P->chunk_begin, e = P->chunk_end;
    if(b == 0 && e == 0) e = c.n_chunks;
    if(e > c.n_chunks) e = c.n_chunks;
    if(b > e) b = e;
    c.begin = b; c.count = e - b;
    return c;
}

PathConsts make_consts(const OrtScene *s, const OrtCamera *cam, const OrtRenderParams *P)
{
    PathConsts c;
    c.cam_p = mk3(cam->p.x, cam->p.y, cam->p.z);
    c.cam_x = mk3(cam->x_axis.x, cam->x_axis.y, cam->x_axis.z);
    c.cam_y = mk3(cam->y_axis.x, cam->y_axis.y, cam->y_axis.z);
    c.cam_z = mk3(cam->z_axis.x, cam->z_axis.y, cam->z_axis.z);
    // ray.cpp:1198, evaluated once on the host with the same IEEE operations
    c.focal_length = length(c.cam_p - mk3(P->focus_target[0], P->focus_target[1], P->focus_target[2]));
    c.aperture_radius = P->aperture_radius;
    c.lens_z_offset = P->lens_z_offset;
    c.roughness = P->roughness;
    c.eps = P->dont_get_too_close_epsilon;
    c.rr = P->russian_roulette_value;
    c.width = P->output_width; c.height = P->output_height;
    c.light_count = s->light_count;
    c.light_is_sphere = s->d_light_is_sphere;
    c.materials = s->d_materials;
    return c;
}

int ensure_wavefront(OrtScene *s, uint32_t capacity)
{
    if(!s->wf_ready)
    {
        CUDA_TRY(cudaMalloc((void **)&s->d_sort, WF_MAX_POOLS * (2 * WF_KEY_BINS + 4) * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc((void **)&s->d_active, WF_MAX_POOLS * 2 * WF_BATCH * sizeof(unsigned int)));
        CUDA_TRY(cudaMallocHost((void **)&s->h_active, WF_MAX_POOLS * 2 * WF_BATCH * sizeof(unsigned int)));
        CUDA_TRY(cudaEventCreateWithFlags(&s->wf_start, cudaEventDisableTiming));
        for(int p = 0; p < WF_MAX_POOLS; ++p)
        {
            WfPool &pl = s->pools[p];
            CUDA_TRY(cudaStreamCreateWithFlags(&pl.stream, cudaStreamNonBlocking));
            for(int b = 0; b < 2; ++b)
            {
                CUDA_TRY(cudaEventCreateWithFlags(&pl.done[b], cudaEventDisableTiming));
                for(int it = 0; it < WF_BATCH; ++it)
                    for(int k = 0; k < 4; ++k) CUDA_TRY(cudaEventCreate(&pl.ev[b][it][k]));
            }
            pl.d_sort = s->d_sort + p * (2 * WF_KEY_BINS + 4);
            pl.d_active = s->d_active + p * 2 * WF_BATCH;
            pl.h_active = s->h_active + p * 2 * WF_BATCH;
        }
        s->wf_ready = 1;
    }
    if(s->wf.capacity >= capacity) return ORT_OK;
    cudaFree(s->wf.rec); cudaFree(s->wf.key); cudaFree(s->wf.perm); cudaFree(s->wf.cold);
    memset(&s->wf, 0, sizeof(s->wf));
    CUDA_TRY(cudaMalloc((void **)&s->wf.rec, (size_t)capacity * WF_REC_QUADS * sizeof(float4)));
#if WF_SPLIT_COLD
    CUDA_TRY(cudaMalloc((void **)&s->wf.cold, (size_t)capacity * 2 * sizeof(float4)));
#endif
    CUDA_TRY(cudaMalloc((void **)&s->wf.key, (size_t)capacity * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc((void **)&s->wf.perm, (size_t)capacity * sizeof(uint32_t)));
    s->wf.capacity = capacity;
    return ORT_OK;
}

// Wavefront driver: per pool, reset -> shade (generates the first rays) -> { extend -> sort -> shade }*
// until no slot of the pool is active.  The "still active" counts are read back once per WF_BATCH
// iterations, one batch behind the enqueue, so the device never drains while the host decides.
int launch_wavefront(OrtScene *s, const RenderArgs &a, cudaStream_t stream, uint32_t *launches)
{
    // 6 Mi slots x 104 B = 650 MB in two pools (measured at 1080p x 256 spp, final kernels, two pools:
    // 4 Mi 1053, 6 Mi 1066, 8 Mi 1062 Msamples/s; three pools of 2 Mi 1039, four of 2 Mi 1026)
    uint32_t capacity = 6u << 20;
    if(const char *e = getenv("ORT_WF_SLOTS")) { long v = atol(e); if(v >= 1024 && v <= (1l << 26)) capacity = (uint32_t)v; }
    unsigned long long items128 = (a.total_items + 127ull) & ~127ull;
    if(items128 < capacity) capacity = (uint32_t)items128;
    int n_pools = 2;
    if(const char *e = getenv("ORT_WF_POOLS")) { int v = atoi(e); if(v >= 1 && v <= WF_MAX_POOLS) n_pools = v; }
    if(capacity < (1u << 18)) n_pools = 1;
    int rc = ensure_wavefront(s, capacity);
    if(rc != ORT_OK) return rc;
    const int sorted = getenv("ORT_WF_NOSORT") ? 0 : 1;
    const size_t stack_bytes = (size_t)s->stack_rows * 128 * sizeof(uint2);
    // persistent EXTEND grid: as many 4-warp blocks as stay resident
    if(s->wf_extend_blocks == 0)
    {
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_wf_extend<false>, 128, stack_bytes));
        s->wf_extend_blocks = (nb > 0 ? nb : 1) * s->sm_count;
    }
    const int extend_blocks_per_sm = getenv("ORT_WF_EXTEND_BPSM") ? atoi(getenv("ORT_WF_EXTEND_BPSM")) : 0;
    s->extend_ms = s->shade_ms = s->sort_ms = 0.f;
    s->extend_launches = 0;
    const bool pipelined = a.total_items > 4ull * capacity;
    // per-stage CUDA events (OrtRenderStats.extend_ms / sort_ms / shade_ms) are opt-in: four timed event records per
    // iteration keep the next kernel from being queued behind the running one -- measured on B200, two-pool frame
    // ms with / without: C3 1080p x 64 spp 137.7 / 134.8, C4 4K x 16 spp 126.2 / 123.6, 4.4 M-triangle grid 139.3 / 136.0
    const char *timing_env = getenv("ORT_WF_TIMING");
    const bool stage_timing = timing_env && atoi(timing_env) != 0;

    // the pools start after whatever the caller queued on its stream (e.g. zeroing the framebuffer)
    CUDA_TRY(cudaEventRecord(s->wf_start, stream));
    const uint32_t per_pool = ((capacity / (uint32_t)n_pools) + 127u) & ~127u;
    for(int p = 0; p < n_pools; ++p)
    {
        WfPool &pl = s->pools[p];
        uint32_t lo = (uint32_t)p * per_pool, hi = lo + per_pool;
        if(hi > capacity || p == n_pools - 1) hi = capacity;
        if(lo > hi) lo = hi;
        pl.wf.rec = s->wf.rec + (size_t)WF_REC_QUADS * lo;
        pl.wf.cold = s->wf.cold ? s->wf.cold + 2ull * lo : 0;
        pl.wf.key = s->wf.key + lo;
        pl.wf.perm = s->wf.perm + lo;
        pl.wf.capacity = hi - lo;
        pl.cur = 0; pl.finished = pl.wf.capacity == 0;
        CUDA_TRY(cudaStreamWaitEvent(pl.stream, s->wf_start, 0));
    }

    auto enqueue = [&](WfPool &pl, int b) -> int
    {
        const uint32_t cap = pl.wf.capacity;
        const unsigned grid = (cap + 127u) / 128u;
        unsigned egrid = (unsigned)s->wf_extend_blocks;
        if(extend_blocks_per_sm > 0 && (unsigned)(extend_blocks_per_sm * s->sm_count) < egrid) egrid = (unsigned)(extend_blocks_per_sm * s->sm_count);
        if(egrid > grid) egrid = grid;
        uint32_t *hist = pl.d_sort, *cursor = pl.d_sort + WF_KEY_BINS, *live = pl.d_sort + 2 * WF_KEY_BINS;
        uint32_t *chunk_counter = pl.d_sort + 2 * WF_KEY_BINS + 1;
        unsigned int *act = pl.d_active + b * WF_BATCH;
        cudaStream_t st = pl.stream;
        CUDA_TRY(cudaMemsetAsync(act, 0, WF_BATCH * sizeof(unsigned int), st));
        for(int it = 0; it < WF_BATCH; ++it)
        {
#if !ORT_EXTEND_FUSED_SCAN
            CUDA_TRY(cudaMemsetAsync(hist, 0, WF_KEY_BINS * sizeof(uint32_t), st));
            CUDA_TRY(cudaMemsetAsync(chunk_counter, 0, sizeof(uint32_t), st));
#endif
            if(stage_timing) CUDA_TRY(cudaEventRecord(pl.ev[b][it][0], st));
#ifdef ORT_COUNTERS
            k_wf_extend<true><<<egrid, 128, stack_bytes, st>>>(a.scene, pl.wf, chunk_counter, a.stats, hist);
#else
            k_wf_extend<false><<<egrid, 128, stack_bytes, st>>>(a.scene, pl.wf, chunk_counter, a.stats, hist);
#endif
            if(stage_timing) CUDA_TRY(cudaEventRecord(pl.ev[b][it][1], st));
            if(sorted)
            {
#if !ORT_EXTEND_FUSED_SCAN
                k_wf_scan<<<1, WF_KEY_BINS, 0, st>>>(hist, cursor, live);
                *launches += 1;
#endif
                k_wf_scatter<<<(cap + 1024u * ORT_SCATTER_ITEMS - 1u) / (1024u * ORT_SCATTER_ITEMS), 1024, 0, st>>>(pl.wf, cursor);
                *launches += 1;
            }
            if(stage_timing) CUDA_TRY(cudaEventRecord(pl.ev[b][it][2], st));
            k_wf_shade<<<(cap + ORT_SHADE_THREADS - 1u) / ORT_SHADE_THREADS, ORT_SHADE_THREADS, 0, st>>>(a, pl.wf, act + it, live, sorted);
            if(stage_timing) CUDA_TRY(cudaEventRecord(pl.ev[b][it][3], st));
        }
        *launches += 2 * WF_BATCH;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(pl.h_active + b * WF_BATCH, act, WF_BATCH * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaEventRecord(pl.done[b], st));
        return ORT_OK;
    };
    auto collect = [&](WfPool &pl, int b, bool *finished) -> int
    {
        CUDA_TRY(cudaEventSynchronize(pl.done[b]));
        for(int it = 0; it < WF_BATCH; ++it)
        {
            s->extend_launches += 1;
            if(!stage_timing) continue;
            float e = 0.f, so = 0.f, sh = 0.f;
            cudaEventElapsedTime(&e, pl.ev[b][it][0], pl.ev[b][it][1]);
            cudaEventElapsedTime(&so, pl.ev[b][it][1], pl.ev[b][it][2]);
            cudaEventElapsedTime(&sh, pl.ev[b][it][2], pl.ev[b][it][3]);
            s->extend_ms += e; s->sort_ms += so; s->shade_ms += sh;
        }
        *finished = pl.h_active[b * WF_BATCH + WF_BATCH - 1] == 0;
        return ORT_OK;
    };

    for(int p = 0; p < n_pools; ++p)
    {
        WfPool &pl = s->pools[p];
        if(pl.finished) continue;
        const unsigned grid = (pl.wf.capacity + 127u) / 128u;
        k_wf_reset<<<grid, 128, 0, pl.stream>>>(pl.wf);
#if ORT_EXTEND_FUSED_SCAN
        // histogram, chunk counter and arrival counter start at zero; EXTEND's last block leaves them so
        CUDA_TRY(cudaMemsetAsync(pl.d_sort, 0, (2 * WF_KEY_BINS + 4) * sizeof(uint32_t), pl.stream));
#endif
        CUDA_TRY(cudaMemsetAsync(pl.d_active, 0, 2 * WF_BATCH * sizeof(unsigned int), pl.stream));
        k_wf_shade<<<(pl.wf.capacity + ORT_SHADE_THREADS - 1u) / ORT_SHADE_THREADS, ORT_SHADE_THREADS, 0, pl.stream>>>(a, pl.wf, pl.d_active, pl.d_sort + 2 * WF_KEY_BINS, 0);
        *launches += 2;
        rc = enqueue(pl, 0);
        if(rc != ORT_OK) return rc;
    }
    for(;;)
    {
        bool any = false;
        for(int p = 0; p < n_pools; ++p)
        {
            WfPool &pl = s->pools[p];
            if(pl.finished) continue;
            any = true;
            if(pipelined)
            {
                rc = enqueue(pl, pl.cur ^ 1);
                if(rc != ORT_OK) return rc;
                rc = collect(pl, pl.cur, &pl.finished);
                if(rc != ORT_OK) return rc;
                pl.cur ^= 1;
                if(pl.finished) { bool dummy; rc = collect(pl, pl.cur, &dummy); if(rc != ORT_OK) return rc; }
            }
            else
            {
                rc = collect(pl, pl.cur, &pl.finished);
                if(rc != ORT_OK) return rc;
                if(!pl.finished) { rc = enqueue(pl, pl.cur); if(rc != ORT_OK) return rc; }
            }
        }
        if(!any) break;
    }
    for(int p = 0; p < n_pools; ++p) CUDA_TRY(cudaStreamSynchronize(s->pools[p].stream));
    return ORT_OK;
}

// launches the render kernels for the chunk range of P into accum (fixed point)
// or rgb (single-chunk float mode); exactly one of the two is non-null
int launch_render(OrtScene *s, const OrtCamera *cam, const OrtRenderParams *P, const ChunkPlan &cp,
                  long long *accum, float *rgb, cudaStream_t stream, uint32_t *launches)
{
    int tw = P->tile_one_past_max_x - P->tile_min_x, th = P->tile_one_past_max_y - P->tile_min_y;
    if(tw <= 0 || th <= 0 || cp.count == 0) return ORT_OK;
    RenderArgs a;
    a.scene = s->view();
    a.pc = make_consts(s, cam, P);
    a.tile_min_x = P->tile_min_x; a.tile_min_y = P->tile_min_y; a.tile_w = tw; a.tile_h = th;
    a.blocks_x = (uint32_t)(tw + 7) / 8u; a.blocks_y = (uint32_t)(th + 3) / 4u;
    a.spp = P->ray_per_pixel_count; a.chunk_spp = cp.chunk_spp; a.n_chunks = cp.n_chunks;
    a.chunk_begin = cp.begin; a.chunk_count = cp.count;
    a.base_seed = P->base_seed;
    a.total_items = (unsigned long long)a.blocks_x * a.blocks_y * 32ull * cp.count;
    a.work_counter = s->d_stats + STAT_WORK_COUNTER;
    a.stats = s->d_stats;
    a.accum = accum; a.rgb = rgb;

    CUDA_TRY(cudaMemsetAsync(s->d_stats + STAT_WORK_COUNTER, 0, sizeof(unsigned long long), stream));
    uint32_t kernel = P->kernel;
    if(kernel == ORT_KERNEL_DEFAULT)
    {
        // The two forms give bit-identical images (tests/test_gpu_render.py).  The wavefront needs a few
        // hundred launches to fill and drain its slot pool whatever the job size, so small jobs go to the
        // single-launch megakernel -- measured on B200, testscene.scn, samples in the call:
        //   2.1 M: 5.0 vs 12.3 ms   8.3 M: 16.5 vs 16.3   14.7 M: 28.6 vs 20.7 (mega vs wavefront, final round-2 kernels;
        //   before the scan moved into EXTEND and the stage events became opt-in: 16.6, 22.7 and 28.8 ms, threshold 16 M)
        const char *e = getenv("ORT_KERNEL");
        const unsigned long long samples = (unsigned long long)tw * th * (unsigned long long)a.chunk_spp * cp.count;
        if(e && atoi(e) != 0) kernel = atoi(e) == ORT_KERNEL_MEGAKERNEL ? ORT_KERNEL_MEGAKERNEL : ORT_KERNEL_WAVEFRONT;
        else kernel = samples < 8000000ull ? ORT_KERNEL_MEGAKERNEL : ORT_KERNEL_WAVEFRONT;
    }
    if(kernel == ORT_KERNEL_WAVEFRONT) return launch_wavefront(s, a, stream, launches);
    if(s->mega_blocks_per_sm == 0)
    {
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_render_mega<false>, 128, 0));
        s->mega_blocks_per_sm = nb > 0 ? nb : 1;
    }
    unsigned long long want = (a.total_items + 127ull) / 128ull;
    unsigned long long grid = (unsigned long long)s->sm_count * s->mega_blocks_per_sm;
    if(grid > want) grid = want;
    if(grid == 0) grid = 1;
#ifdef ORT_COUNTERS
    k_render_mega<true><<<(unsigned)grid, 128, 0, stream>>>(a);
#else
    k_render_mega<false><<<(unsigned)grid, 128, 0, stream>>>(a);
#endif
    CUDA_TRY(cudaGetLastError());
    *launches += 1;
    return ORT_OK;
}

int read_stats(OrtScene *s, cudaStream_t stream, OrtRenderStats *stats, float ms, uint32_t launches)
{
    if(!stats) return ORT_OK;
    unsigned long long h[STAT_COUNT];
    CUDA_TRY(cudaMemcpyAsync(h, s->d_stats, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    stats->samples = h[STAT_SAMPLES]; stats->rays = h[STAT_RAYS];
    stats->node_visits = h[STAT_NODE_VISITS]; stats->box_tests = h[STAT_BOX_TESTS]; stats->shape_tests = h[STAT_SHAPE_TESTS];
    stats->device_ms = ms; stats->kernel_launches = launches;
    stats->extend_ms = s->extend_ms; stats->shade_ms = s->shade_ms; stats->sort_ms = s->sort_ms;
    stats->extend_launches = s->extend_launches;
    return ORT_OK;
}

} // namespace

extern "C" {

const char *ort_last_error(void) { return g_last_error.c_str(); }
void ort_set_last_error_(const char *msg) { g_last_error = msg ? msg : ""; }   // used by host_scene.cpp

int ort_device_count(int *count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if(count) *count = (e == cudaSuccess) ? n : 0;
    if(e != cudaSuccess || n == 0) return fail(ORT_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    return ORT_OK;
}

int ort_selftest_div3(int device, uint64_t triples, uint32_t seed, uint64_t *tested, uint64_t *mismatches)
{
    if(!tested || !mismatches) return fail(ORT_ERR_ARG, "null argument");
    int n = 0;
    if(ort_device_count(&n) != ORT_OK) return ORT_ERR_CUDA;
    if(device < 0 || device >= n) return fail(ORT_ERR_ARG, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
    const int blocks = 148 * 8, threads = 256;
    uint64_t per_thread = (triples + (uint64_t)blocks * threads - 1) / ((uint64_t)blocks * threads);
    if(per_thread == 0) per_thread = 1;
    if(per_thread > (1u << 24)) return fail(ORT_ERR_ARG, "too many triples for one launch");
    unsigned long long *d_bad = 0;
    CUDA_TRY(cudaMalloc((void **)&d_bad, 2 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(d_bad, 0, 2 * sizeof(unsigned long long)));
    k_selftest_div3<<<blocks, threads>>>(seed, (int)per_thread, d_bad);
    CUDA_TRY(cudaGetLastError());
    unsigned long long bad = 0;
    CUDA_TRY(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
    cudaFree(d_bad);
    *tested = per_thread * (uint64_t)blocks * threads;
    *mismatches = bad;
    return ORT_OK;
}

int ort_measure_fp32_peak(int device, float *tflops_non_fma, float *sm_clock_mhz_unused)
{
    (void)sm_clock_mhz_unused;
    if(!tflops_non_fma) return fail(ORT_ERR_ARG, "null argument");
    int n = 0;
    if(ort_device_count(&n) != ORT_OK) return ORT_ERR_CUDA;
    if(device < 0 || device >= n) return fail(ORT_ERR_ARG, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float *d_out = 0;
    const int blocks = sms * 8, threads = 256, iters = 1 << 16;
    CUDA_TRY(cudaMalloc((void **)&d_out, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    float best = 0.f;
    for(int rep = 0; rep < 5; ++rep)
    {
        CUDA_TRY(cudaEventRecord(e0, 0));
        k_fp32_peak<<<blocks, threads>>>(d_out, iters, 0.999f, 0.001f);
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double flops = (double)blocks * threads * (double)iters * 16.0;
        float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if(rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *tflops_non_fma = best;
    return ORT_OK;
}

static double seconds_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int ort_scene_create(const OrtWorld *world, const OrtBVHOctreeNode *top_most_node, int device, OrtScene **scene_out)
{
    // ORT_BVH_BUILD = host | device; unset: the host's binned-SAH builder up to a million records (well
    // under a second, slightly better trees), the CUDA builder beyond
    uint32_t flags = ORT_BUILD_AUTO;
    if(const char *e = getenv("ORT_BVH_BUILD")) flags = !strcmp(e, "device") ? ORT_BUILD_ON_DEVICE : 0u;
    return ort_scene_create_ex(world, top_most_node, device, flags, scene_out);
}

static int make_scene_handle(FlatScene &flat, build::DeviceBuildResult &built, bool on_device, OrtBuildStats bs, int device, OrtScene **scene_out);

// everything after the records are known: build the tree (host or device), upload, create the handle
static int create_scene_from_records(std::vector<HostPrim> &prims, FlatScene &flat, int device, uint32_t flags,
                                     OrtBuildStats bs, OrtScene **scene_out)
{
    std::string err;
    int rc = ORT_OK;
    BuildOptions opt;
    if(const char *e = getenv("ORT_BVH_TRAVERSAL_COST")) opt.traversal_cost = (float)atof(e);
    if(const char *e = getenv("ORT_BVH_PAD_REL")) opt.pad_rel = (float)atof(e);
    if(const char *e = getenv("ORT_BVH_MERGE_SHAPES")) opt.merge_shapes = atoi(e) != 0;
    if(const char *e = getenv("ORT_BVH_OPTIMAL")) opt.optimal_collapse = atoi(e) != 0;
    if(const char *e = getenv("ORT_BVH_NODE_COST")) opt.wide_node_cost = (float)atof(e);
    const bool on_device = (flags & ORT_BUILD_ON_DEVICE) != 0u || ((flags & ORT_BUILD_AUTO) != 0u && prims.size() >= ORT_BUILD_AUTO_RECORDS);
    bs.on_device = on_device ? 1u : 0u;
    build::DeviceBuildResult built; memset(&built, 0, sizeof(built));
    auto t0 = std::chrono::steady_clock::now();
    if(on_device)
    {
        // Morton sort, PLOC clustering and the collapse to 8-wide run as CUDA kernels
        // (bvh_build.cuh); the host only pads the boxes and packs the records
        ParallelBuildInput in;
        rc = prepare_parallel_build(prims, opt, &flat, &in, &err);
        if(rc != ORT_OK) return fail(rc, err);
        bs.prepare_s = (float)seconds_since(t0);
        t0 = std::chrono::steady_clock::now();
        uint32_t radius = ORT_PLOC_RADIUS;
        if(const char *e = getenv("ORT_PLOC_RADIUS")) radius = (uint32_t)atoi(e);
        if(radius < 1u) radius = 1u;
        if(!build::build_on_device(flat, in, radius, &built, &err)) return fail(ORT_ERR_CUDA, err);
        flat.wide_depth = std::max(in.sphere_depth, built.depth);
        if(flat.wide_depth + 2 > ORT_STACK_SIZE)
        {
            cudaFree(built.d_nodes); cudaFree(built.d_prims); cudaFree(built.d_rank_to_prim);
            return fail(ORT_ERR_LIMIT, "wide BVH deeper than the traversal stack");
        }
        flat.info.bvh_node_count = built.node_count;
        flat.info.bvh_node_bytes = (uint32_t)sizeof(WideNode);
        bs.build_s = (float)seconds_since(t0);
        bs.device_build_ms = built.build_ms;
        bs.ploc_iterations = built.ploc_iterations;
    }
    else
    {
        // the optimal collapse can give a deeper tree than the greedy one; should that ever exceed the
        // traversal stack (ORT_STACK_SIZE), fall back to the greedy collapse
        std::vector<HostPrim> backup;
        if(opt.optimal_collapse && prims.size() < 2000000u) backup = prims;
        rc = build_wide_bvh(prims, opt, &flat, &err);
        if(rc == ORT_ERR_LIMIT && !backup.empty())
        {
            BuildOptions greedy = opt; greedy.optimal_collapse = false;
            err.clear();
            rc = build_wide_bvh(backup, greedy, &flat, &err);
        }
        if(rc != ORT_OK) return fail(rc, err);
        bs.build_s = (float)seconds_since(t0);
    }
    bs.wide_depth = flat.wide_depth;
    return make_scene_handle(flat, built, on_device, bs, device, scene_out);
}

static int make_scene_handle(FlatScene &flat, build::DeviceBuildResult &built, bool on_device, OrtBuildStats bs, int device, OrtScene **scene_out)
{
    int rc = ORT_OK;
    OrtScene *s = new OrtScene();
    memset(s, 0, sizeof(*s));
    s->mu = new std::recursive_mutex();
    s->device = device;
    s->info = flat.info;
    s->build_stats = bs;
    s->main_root = flat.main_root;
    s->tri_root = flat.tri_root;
    s->stack_rows = flat.wide_depth + 2u;
    s->light_count = (uint32_t)flat.light_is_sphere.size();
    uint64_t total = 0;
    if(on_device)
    {
        s->node_count = built.node_count; s->prim_count = built.prim_count;
        s->d_nodes = (q4 *)built.d_nodes; s->d_prims = (q4 *)built.d_prims; s->d_rank_to_prim = built.d_rank_to_prim;
        total += (uint64_t)built.node_count * sizeof(WideNode) + (uint64_t)built.prim_count * sizeof(PrimRec)
               + (uint64_t)flat.info.record_count * sizeof(uint32_t);
        rc = ORT_OK;
    }
    else
    {
        s->node_count = (uint32_t)flat.nodes.size();
        s->prim_count = (uint32_t)flat.prims.size();
        rc = upload(&s->d_nodes, flat.nodes.data(), flat.nodes.size() * sizeof(WideNode), &total);
        if(rc == ORT_OK) rc = upload(&s->d_prims, flat.prims.data(), flat.prims.size() * sizeof(PrimRec), &total);
        if(rc == ORT_OK) rc = upload(&s->d_rank_to_prim, flat.rank_to_prim.data(), flat.rank_to_prim.size() * sizeof(uint32_t), &total);
    }
    if(rc == ORT_OK) rc = upload(&s->d_cyl, flat.cylinders.data(), flat.cylinders.size() * sizeof(CylinderAux), &total);
    if(rc == ORT_OK) rc = upload(&s->d_materials, flat.materials.data(), flat.materials.size() * sizeof(DevMaterial), &total);
    if(rc == ORT_OK) rc = upload(&s->d_light_is_sphere, flat.light_is_sphere.data(), flat.light_is_sphere.size(), &total);
    if(rc == ORT_OK) rc = upload(&s->d_stats, (const void *)0, 0, &total);
    if(rc != ORT_OK) { ort_scene_destroy(s); return rc; }
    cudaFree(s->d_stats); s->d_stats = 0;
    CUDA_TRY(cudaMalloc((void **)&s->d_stats, STAT_COUNT * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(s->d_stats, 0, STAT_COUNT * sizeof(unsigned long long)));
    CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&s->ev0));
    CUDA_TRY(cudaEventCreate(&s->ev1));
    CUDA_TRY(cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device));
    s->info.device_bytes = total;
    *scene_out = s;
    return ORT_OK;
}

static int check_device(int device, OrtScene **scene_out)
{
    if(!scene_out) return fail(ORT_ERR_ARG, "scene_out is null");
    *scene_out = 0;
    int n = 0;
    if(ort_device_count(&n) != ORT_OK) return ORT_ERR_CUDA;
    if(device < 0 || device >= n) return fail(ORT_ERR_ARG, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
    return ORT_OK;
}

int ort_scene_create_ex(const OrtWorld *world, const OrtBVHOctreeNode *top_most_node, int device, uint32_t flags, OrtScene **scene_out)
{
    ORT_GUARD_BEGIN
    int rc = check_device(device, scene_out);
    if(rc != ORT_OK) return rc;
    FlatScene flat;
    std::string err;
    OrtBuildStats bs; memset(&bs, 0, sizeof(bs));
    auto t0 = std::chrono::steady_clock::now();
    std::vector<HostPrim> prims;
    rc = collect_records(world, top_most_node, &prims, &flat, &err);
    if(rc != ORT_OK) return fail(rc, err);
    bs.collect_s = (float)seconds_since(t0);
    return create_scene_from_records(prims, flat, device, flags, bs, scene_out);
    ORT_GUARD_END
}

int ort_scene_create_from_lists(const OrtWorld *world, const OrtShapeLists *lists, int device, uint32_t flags, OrtScene **scene_out)
{
    ORT_GUARD_BEGIN
    int rc = check_device(device, scene_out);
    if(rc != ORT_OK) return rc;
    FlatScene flat;
    std::string err;
    OrtBuildStats bs; memset(&bs, 0, sizeof(bs));
    auto t0 = std::chrono::steady_clock::now();
    if(!world || !lists) return fail(ORT_ERR_ARG, "null world / shape lists");
    if(flags & ORT_BUILD_AUTO)
    {
        uint64_t records = (uint64_t)lists->cylinder_count + lists->box_count + lists->sphere_count;
        for(uint32_t m = 0; m < lists->mesh_count; ++m) records += lists->meshes[m].index_count / 3u;
        if(records >= ORT_BUILD_AUTO_RECORDS) flags |= ORT_BUILD_ON_DEVICE;
    }
    if(flags & ORT_BUILD_ON_DEVICE)
    {
        // records, ranks and the tree all on the device: the host touches the meshes only to copy them
        BuildOptions opt;
        if(const char *e = getenv("ORT_BVH_TRAVERSAL_COST")) opt.traversal_cost = (float)atof(e);
        if(const char *e = getenv("ORT_BVH_PAD_REL")) opt.pad_rel = (float)atof(e);
        ParallelBuildInput in;
        build::DeviceRecords dr;
        if(!build::records_from_lists_on_device(world, lists, opt, &flat, &in, &dr, &err)) return fail(ORT_ERR_CUDA, err);
        bs.on_device = 1u;
        bs.collect_s = (float)seconds_since(t0);
        bs.prepare_s = dr.records_ms * 1e-3f;
        t0 = std::chrono::steady_clock::now();
        uint32_t radius = ORT_PLOC_RADIUS;
        if(const char *e = getenv("ORT_PLOC_RADIUS")) radius = (uint32_t)atoi(e);
        if(radius < 1u) radius = 1u;
        build::DeviceBuildResult built; memset(&built, 0, sizeof(built));
        bool ok = build::build_on_device(flat, in, radius, &built, &err, dr.d_boxes, dr.d_recs, dr.n_rest);
        cudaFree(dr.d_boxes); cudaFree(dr.d_recs);
        if(!ok) return fail(ORT_ERR_CUDA, err);
        flat.wide_depth = std::max(in.sphere_depth, built.depth);
        if(flat.wide_depth + 2 > ORT_STACK_SIZE)
        {
            cudaFree(built.d_nodes); cudaFree(built.d_prims); cudaFree(built.d_rank_to_prim);
            return fail(ORT_ERR_LIMIT, "wide BVH deeper than the traversal stack");
        }
        flat.info.bvh_node_count = built.node_count;
        flat.info.bvh_node_bytes = (uint32_t)sizeof(WideNode);
        bs.build_s = (float)seconds_since(t0);
        bs.device_build_ms = built.build_ms;
        bs.ploc_iterations = built.ploc_iterations;
        bs.wide_depth = flat.wide_depth;
        return make_scene_handle(flat, built, true, bs, device, scene_out);
    }
    std::vector<HostPrim> prims;
    rc = collect_records_from_lists(world, lists, &prims, &flat, &err);
    if(rc != ORT_OK) return fail(rc, err);
    bs.collect_s = (float)seconds_since(t0);
    return create_scene_from_records(prims, flat, device, flags, bs, scene_out);
    ORT_GUARD_END
}

int ort_scene_build_stats(const OrtScene *s, OrtBuildStats *out)
{
    if(!s || !out) return fail(ORT_ERR_ARG, "null argument");
    *out = s->build_stats;
    return ORT_OK;
}

int ort_scene_download(OrtScene *s, void *nodes, uint64_t nodes_bytes, void *prims, uint64_t prims_bytes)
{
    if(!s) return fail(ORT_ERR_ARG, "null scene");
    CUDA_TRY(cudaSetDevice(s->device));
    if(nodes)
    {
        if(nodes_bytes != (uint64_t)s->node_count * sizeof(WideNode)) return fail(ORT_ERR_ARG, "nodes buffer size mismatch");
        CUDA_TRY(cudaMemcpy(nodes, s->d_nodes, nodes_bytes, cudaMemcpyDeviceToHost));
    }
    if(prims)
    {
        if(prims_bytes != (uint64_t)s->prim_count * sizeof(PrimRec)) return fail(ORT_ERR_ARG, "records buffer size mismatch");
        CUDA_TRY(cudaMemcpy(prims, s->d_prims, prims_bytes, cudaMemcpyDeviceToHost));
    }
    return ORT_OK;
}

int ort_scene_destroy(OrtScene *s)
{
    if(!s) return ORT_OK;
    cudaSetDevice(s->device);
    if(s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->d_rank_to_prim);
    cudaFree(s->d_nodes); cudaFree(s->d_prims); cudaFree(s->d_cyl); cudaFree(s->d_materials);
    cudaFree(s->d_light_is_sphere); cudaFree(s->d_stats); cudaFree(s->d_accum); cudaFree(s->d_rgb); cudaFree(s->d_rgbe);
    cudaFree(s->wf.rec); cudaFree(s->d_active);
    if(s->wf_ready)
    {
        cudaEventDestroy(s->wf_start);
        for(int p = 0; p < WF_MAX_POOLS; ++p)
        {
            cudaStreamDestroy(s->pools[p].stream);
            for(int b = 0; b < 2; ++b)
            {
                cudaEventDestroy(s->pools[p].done[b]);
                for(int it = 0; it < WF_BATCH; ++it) for(int k = 0; k < 4; ++k) cudaEventDestroy(s->pools[p].ev[b][it][k]);
            }
        }
    }
    cudaFree(s->wf.key); cudaFree(s->wf.perm); cudaFree(s->wf.cold); cudaFree(s->d_sort);
    if(s->h_active) cudaFreeHost(s->h_active);
    if(s->ev0) cudaEventDestroy(s->ev0);
    if(s->ev1) cudaEventDestroy(s->ev1);
    if(s->stream) cudaStreamDestroy(s->stream);
    delete s->mu;
    delete s;
    return ORT_OK;
}

int ort_scene_device(const OrtScene *s, int *device)
{
    if(!s || !device) return fail(ORT_ERR_ARG, "null argument");
    *device = s->device;
    return ORT_OK;
}

int ort_scene_info(const OrtScene *s, OrtSceneInfo *info)
{
    if(!s || !info) return fail(ORT_ERR_ARG, "null argument");
    *info = s->info;
    return ORT_OK;
}

void ort_render_params_default(OrtRenderParams *p, int32_t width, int32_t height, uint32_t ray_per_pixel_count)
{
    memset(p, 0, sizeof(*p));
    p->output_width = width; p->output_height = height;
    p->tile_one_past_max_x = width; p->tile_one_past_max_y = height;
    p->ray_per_pixel_count = ray_per_pixel_count;
    p->russian_roulette_value = 0.8f;          // macos_main.mm:656
    p->base_seed = 1234567u;
    p->kernel = ORT_KERNEL_DEFAULT;
    p->roughness = 0.01f;                      // ray.cpp:1194
    p->dont_get_too_close_epsilon = 0.0001f;   // ray.cpp:1196
    p->aperture_radius = 0.1f;                 // ray.cpp:1199
    p->lens_z_offset = 0.1f;                   // ray.cpp:1234
    p->focus_target[0] = 0.0f; p->focus_target[1] = 0.0f; p->focus_target[2] = 0.2f;   // ray.cpp:1198
}

// the device part of ort_render: leaves the image as float3 pixels in s->d_rgb (and, when the
// job has more than one chunk, the fixed-point sums in s->d_accum); *fixed_out tells which
static int render_to_device(OrtScene *s, const OrtCamera *camera, const OrtRenderParams *P, bool *fixed_out, uint32_t *launches)
{
    ChunkPlan cp = plan_chunks(P);
    size_t pixels = (size_t)P->output_width * P->output_height;
    int tw = P->tile_one_past_max_x - P->tile_min_x, th = P->tile_one_past_max_y - P->tile_min_y;
    int rc = ensure_rgb(s, pixels);
    if(rc != ORT_OK) return rc;
    const bool fixed = cp.n_chunks > 1;
    if(fixed) { rc = ensure_accum(s, pixels); if(rc != ORT_OK) return rc; }
    cudaStream_t st = s->stream;
    CUDA_TRY(cudaMemsetAsync(s->d_stats, 0, STAT_COUNT * sizeof(unsigned long long), st));
    CUDA_TRY(cudaEventRecord(s->ev0, st));
    if(fixed)
    {
        // only the tile's rows need clearing, but rows are contiguous: clear [min_y, max_y)
        size_t off = (size_t)P->tile_min_y * P->output_width * 4;
        size_t cnt = (size_t)th * P->output_width * 4;
        if(cnt) CUDA_TRY(cudaMemsetAsync(s->d_accum + off, 0, cnt * sizeof(long long), st));
    }
    rc = launch_render(s, camera, P, cp, fixed ? s->d_accum : 0, fixed ? 0 : s->d_rgb, st, launches);
    if(rc != ORT_OK) return rc;
    if(fixed && tw > 0 && th > 0)
    {
        int n = tw * th;
        k_resolve_fixed<<<(n + 255) / 256, 256, 0, st>>>(s->d_accum, s->d_rgb, P->output_width, P->tile_min_x, P->tile_min_y,
                                                          tw, th, P->ray_per_pixel_count);
        CUDA_TRY(cudaGetLastError());
        (*launches)++;
    }
    *fixed_out = fixed;
    return ORT_OK;
}

int ort_render(OrtScene *s, const OrtCamera *camera, const OrtRenderParams *P, ort_v3 *output_buffer, OrtRenderStats *stats)
{
    if(!s || !camera || !output_buffer) return fail(ORT_ERR_ARG, "null argument");
    int rc = check_params(P);
    if(rc != ORT_OK) return rc;
    SceneLock lock(s);
    CUDA_TRY(cudaSetDevice(s->device));
    int tw = P->tile_one_past_max_x - P->tile_min_x, th = P->tile_one_past_max_y - P->tile_min_y;
    uint32_t launches = 0;
    bool fixed = false;
    cudaStream_t st = s->stream;
    rc = render_to_device(s, camera, P, &fixed, &launches);
    if(rc != ORT_OK) return rc;
    CUDA_TRY(cudaEventRecord(s->ev1, st));
    if(tw > 0 && th > 0)
    {
        // device -> host, tile rows only (the reference writes only the tile's pixels, ray.cpp:1201-1436)
        CUDA_TRY(cudaMemcpy2DAsync(output_buffer + (size_t)P->tile_min_y * P->output_width + P->tile_min_x,
                                   (size_t)P->output_width * sizeof(ort_v3),
                                   s->d_rgb + 3 * ((size_t)P->tile_min_y * P->output_width + P->tile_min_x),
                                   (size_t)P->output_width * 3 * sizeof(float),
                                   (size_t)tw * sizeof(ort_v3), (size_t)th, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    return read_stats(s, st, stats, ms, launches);
}

int ort_rgbe_encode_device(OrtScene *s, const void *rgb_device, int32_t width, int32_t height, void *rgbe_device, void *stream)
{
    if(!s || !rgb_device || !rgbe_device || width <= 0 || height <= 0) return fail(ORT_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(s->device));
    int n = width * height;
    k_rgbe_from_rgb<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float *)rgb_device, (uint32_t *)rgbe_device, width, height);
    CUDA_TRY(cudaGetLastError());
    return ORT_OK;
}

int ort_accum_resolve_rgbe_device(OrtScene *s, const void *accum_device, int32_t width, int32_t height,
                                  uint32_t ray_per_pixel_count, void *rgbe_device, void *stream)
{
    if(!s || !accum_device || !rgbe_device || width <= 0 || height <= 0 || ray_per_pixel_count == 0)
        return fail(ORT_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(s->device));
    int n = width * height;
    k_rgbe_from_fixed<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const long long *)accum_device, (uint32_t *)rgbe_device,
                                                                        width, height, ray_per_pixel_count);
    CUDA_TRY(cudaGetLastError());
    return ORT_OK;
}

int ort_render_rgbe(OrtScene *s, const OrtCamera *camera, const OrtRenderParams *P, uint32_t *rgbe_out, OrtRenderStats *stats)
{
    if(!s || !camera || !rgbe_out) return fail(ORT_ERR_ARG, "null argument");
    int rc = check_params(P);
    if(rc != ORT_OK) return rc;
    if(P->tile_min_x != 0 || P->tile_min_y != 0 || P->tile_one_past_max_x != P->output_width || P->tile_one_past_max_y != P->output_height)
        return fail(ORT_ERR_ARG, "ort_render_rgbe renders whole images: the tile must cover the output");
    SceneLock lock(s);
    CUDA_TRY(cudaSetDevice(s->device));
    size_t pixels = (size_t)P->output_width * P->output_height;
    if(s->rgbe_capacity < pixels)
    {
        cudaFree(s->d_rgbe); s->d_rgbe = 0; s->rgbe_capacity = 0;
        CUDA_TRY(cudaMalloc((void **)&s->d_rgbe, pixels * sizeof(uint32_t)));
        s->rgbe_capacity = pixels;
    }
    uint32_t launches = 0;
    bool fixed = false;
    cudaStream_t st = s->stream;
    rc = render_to_device(s, camera, P, &fixed, &launches);
    if(rc != ORT_OK) return rc;
    if(fixed)
        k_rgbe_from_fixed<<<(unsigned)((pixels + 255) / 256), 256, 0, st>>>(s->d_accum, s->d_rgbe, P->output_width, P->output_height,
                                                                              P->ray_per_pixel_count);
    else
        k_rgbe_from_rgb<<<(unsigned)((pixels + 255) / 256), 256, 0, st>>>(s->d_rgb, s->d_rgbe, P->output_width, P->output_height);
    CUDA_TRY(cudaGetLastError());
    launches++;
    CUDA_TRY(cudaEventRecord(s->ev1, st));
    CUDA_TRY(cudaMemcpyAsync(rgbe_out, s->d_rgbe, pixels * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    return read_stats(s, st, stats, ms, launches);
}

int ort_render_hdr(OrtScene *s, const OrtCamera *camera, const OrtRenderParams *P, const char *path, OrtRenderStats *stats)
{
    if(!path || !P) return fail(ORT_ERR_ARG, "null argument");
    if(P->output_width <= 0 || P->output_height <= 0) return fail(ORT_ERR_ARG, "bad image size");
    ORT_GUARD_BEGIN
    std::vector<uint32_t> words((size_t)P->output_width * P->output_height);
    int rc = ort_render_rgbe(s, camera, P, words.data(), stats);
    if(rc != ORT_OK) return rc;
    return ort_write_hdr_rgbe(path, words.data(), P->output_width, P->output_height);
    ORT_GUARD_END
}

int ort_tiled_raytrace_bvh(OrtScene *scene, const OrtCamera *camera, ort_v3 *output_buffer,
                           int32_t output_width, int32_t output_height,
                           int32_t tile_min_x, int32_t tile_min_y, int32_t tile_one_past_max_x, int32_t tile_one_past_max_y,
                           OrtRandomSeries *series, uint32_t ray_per_pixel_count, float russian_roulette_value,
                           uint64_t *test_shape_count)
{
    if(!series) return fail(ORT_ERR_ARG, "series is null");
    if(!scene) return fail(ORT_ERR_ARG, "null scene");
    SceneLock lock(scene);
    OrtRenderParams P;
    ort_render_params_default(&P, output_width, output_height, ray_per_pixel_count);
    P.tile_min_x = tile_min_x; P.tile_min_y = tile_min_y;
    P.tile_one_past_max_x = tile_one_past_max_x; P.tile_one_past_max_y = tile_one_past_max_y;
    P.russian_roulette_value = russian_roulette_value;
    P.base_seed = series->next_random;
    OrtRenderStats st; memset(&st, 0, sizeof(st));
    int rc = ort_render(scene, camera, &P, output_buffer, &st);
    if(rc != ORT_OK) return rc;
    xor_shift_32(&series->next_random);
    if(test_shape_count) *test_shape_count = st.shape_tests;     // primitive tests executed, as ray.cpp:661-715, 1173
    return ORT_OK;
}

int ort_accum_zero_device(OrtScene *s, void *accum_device, int32_t width, int32_t height, void *stream)
{
    if(!s || !accum_device || width <= 0 || height <= 0) return fail(ORT_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaMemsetAsync(accum_device, 0, (size_t)width * height * 4 * sizeof(long long), (cudaStream_t)stream));
    return ORT_OK;
}

int ort_render_accumulate_device(OrtScene *s, const OrtCamera *camera, const OrtRenderParams *P, void *accum_device,
                                 void *stream, OrtRenderStats *stats)
{
    if(!s || !camera || !accum_device) return fail(ORT_ERR_ARG, "null argument");
    int rc = check_params(P);
    if(rc != ORT_OK) return rc;
    SceneLock lock(s);
    CUDA_TRY(cudaSetDevice(s->device));
    ChunkPlan cp = plan_chunks(P);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t launches = 0;
    CUDA_TRY(cudaMemsetAsync(s->d_stats, 0, STAT_COUNT * sizeof(unsigned long long), st));
    CUDA_TRY(cudaEventRecord(s->ev0, st));
    rc = launch_render(s, camera, P, cp, (long long *)accum_device, 0, st, &launches);
    if(rc != ORT_OK) return rc;
    CUDA_TRY(cudaEventRecord(s->ev1, st));
    // The call returns when the chunk range is done, whichever kernel family ran it (the wavefront loop
    // has host-synchronised its pools already; the single-launch megakernel is waited for here): the
    // sums are visible to the caller's next launch on any stream, and the handle's counters are free
    // for the next call.
    CUDA_TRY(cudaStreamSynchronize(st));
    if(stats)
    {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        return read_stats(s, st, stats, ms, launches);
    }
    return ORT_OK;
}

int ort_accum_resolve_device(OrtScene *s, const void *accum_device, int32_t width, int32_t height,
                             uint32_t ray_per_pixel_count, void *rgb_device, void *stream)
{
    if(!s || !accum_device || !rgb_device || width <= 0 || height <= 0 || ray_per_pixel_count == 0)
        return fail(ORT_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(s->device));
    int n = width * height;
    k_resolve_fixed<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const long long *)accum_device, (float *)rgb_device,
                                                                      width, 0, 0, width, height, ray_per_pixel_count);
    CUDA_TRY(cudaGetLastError());
    return ORT_OK;
}

int ort_raycast_batch_device(OrtScene *s, uint64_t n, const float *origins, const float *dirs,
                             float *hit_t, uint32_t *prim_rank, uint32_t *mat_index, float *hit_normal, void *stream)
{
    if(!s || (n && (!origins || !dirs))) return fail(ORT_ERR_ARG, "null argument");
    if(n == 0) return ORT_OK;
    CUDA_TRY(cudaSetDevice(s->device));
    unsigned long long want = (n + 127ull) / 128ull;
    unsigned long long grid = (unsigned long long)s->sm_count * 8ull;
    if(grid > want) grid = want;
    k_raycast<false><<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(s->view(), n, origins, dirs, hit_t, prim_rank,
                                                                        mat_index, hit_normal, s->d_stats);
    CUDA_TRY(cudaGetLastError());
    return ORT_OK;
}

int ort_raycast_brute_device(OrtScene *s, uint64_t n, const float *origins, const float *dirs,
                             float *hit_t, uint32_t *prim_rank, uint32_t *mat_index, void *stream)
{
    if(!s || (n && (!origins || !dirs))) return fail(ORT_ERR_ARG, "null argument");
    if(n == 0) return ORT_OK;
    CUDA_TRY(cudaSetDevice(s->device));
    unsigned long long grid = (n + 255ull) / 256ull;
    k_raycast_brute<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(s->view(), n, origins, dirs, hit_t, prim_rank, mat_index);
    CUDA_TRY(cudaGetLastError());
    return ORT_OK;
}

int ort_raycast_counters_device(OrtScene *s, uint64_t n, const float *origins, const float *dirs,
                                uint64_t *node_visits, uint64_t *box_tests, uint64_t *shape_tests)
{
    if(!s || (n && (!origins || !dirs))) return fail(ORT_ERR_ARG, "null argument");
    SceneLock lock(s);
    CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    CUDA_TRY(cudaMemsetAsync(s->d_stats, 0, STAT_COUNT * sizeof(unsigned long long), st));
    if(n)
    {
        unsigned long long want = (n + 127ull) / 128ull;
        unsigned long long grid = (unsigned long long)s->sm_count * 8ull;
        if(grid > want) grid = want;
        k_raycast<true><<<(unsigned)grid, 128, 0, st>>>(s->view(), n, origins, dirs, 0, 0, 0, 0, s->d_stats);
        CUDA_TRY(cudaGetLastError());
    }
    unsigned long long h[STAT_COUNT];
    CUDA_TRY(cudaMemcpyAsync(h, s->d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if(node_visits) *node_visits = h[STAT_NODE_VISITS];
    if(box_tests) *box_tests = h[STAT_BOX_TESTS];
    if(shape_tests) *shape_tests = h[STAT_SHAPE_TESTS];
    return ORT_OK;
}

int ort_raycast_batch(OrtScene *s, uint64_t n, const float *origins, const float *dirs,
                      float *hit_t, uint32_t *prim_rank, uint32_t *mat_index, float *hit_normal, OrtRenderStats *stats)
{
    if(!s || (n && (!origins || !dirs))) return fail(ORT_ERR_ARG, "null argument");
    SceneLock lock(s);
    CUDA_TRY(cudaSetDevice(s->device));
    if(stats) memset(stats, 0, sizeof(*stats));
    if(n == 0) return ORT_OK;
    cudaStream_t st = s->stream;
    float *d_o = 0, *d_d = 0, *d_t = 0, *d_n = 0; uint32_t *d_r = 0, *d_m = 0;
    int rc = ORT_OK;
    auto cleanup = [&]() { cudaFree(d_o); cudaFree(d_d); cudaFree(d_t); cudaFree(d_n); cudaFree(d_r); cudaFree(d_m); };
#define TRY_OR_CLEAN(expr) do { cudaError_t e__ = (expr); if(e__ != cudaSuccess) { cleanup(); return fail(ORT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while(0)
    TRY_OR_CLEAN(cudaMalloc((void **)&d_o, n * 3 * sizeof(float)));
    TRY_OR_CLEAN(cudaMalloc((void **)&d_d, n * 3 * sizeof(float)));
    if(hit_t) TRY_OR_CLEAN(cudaMalloc((void **)&d_t, n * sizeof(float)));
    if(prim_rank) TRY_OR_CLEAN(cudaMalloc((void **)&d_r, n * sizeof(uint32_t)));
    if(mat_index) TRY_OR_CLEAN(cudaMalloc((void **)&d_m, n * sizeof(uint32_t)));
    if(hit_normal) TRY_OR_CLEAN(cudaMalloc((void **)&d_n, n * 3 * sizeof(float)));
    TRY_OR_CLEAN(cudaMemcpyAsync(d_o, origins, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    TRY_OR_CLEAN(cudaMemcpyAsync(d_d, dirs, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    TRY_OR_CLEAN(cudaEventRecord(s->ev0, st));
    rc = ort_raycast_batch_device(s, n, d_o, d_d, d_t, d_r, d_m, d_n, st);
    if(rc != ORT_OK) { cleanup(); return rc; }
    TRY_OR_CLEAN(cudaEventRecord(s->ev1, st));
    if(hit_t) TRY_OR_CLEAN(cudaMemcpyAsync(hit_t, d_t, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if(prim_rank) TRY_OR_CLEAN(cudaMemcpyAsync(prim_rank, d_r, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if(mat_index) TRY_OR_CLEAN(cudaMemcpyAsync(mat_index, d_m, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if(hit_normal) TRY_OR_CLEAN(cudaMemcpyAsync(hit_normal, d_n, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    TRY_OR_CLEAN(cudaStreamSynchronize(st));
    if(stats)
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s->ev0, s->ev1);
        stats->rays = n; stats->device_ms = ms; stats->kernel_launches = 1;
    }
    cleanup();
#undef TRY_OR_CLEAN
    return ORT_OK;
}

} // extern "C"
