// kernels.cuh -- sm_100a kernels of the hot path.  One ray per thread, FP32 CUDA
// cores only (no stage of this path is a dense contraction, so tensor cores are
// deliberately unused), contraction off (-fmad=false; see core_math.h).
//
//   k_raycast        N explicit rays -> (t, rank, mat, normal)      replaces raycast_top_most_node, ray.cpp:1165
//   k_raycast_brute  the same answer by exhaustive search (validation)
//   k_render_mega    persistent-thread path tracer with path regeneration:
//                    generate + extend + shade + accumulate in registers       replaces tiled_raytrace_bvh, ray.cpp:1178
//   k_resolve_*      accumulate: fixed-point / float sums -> v3 pixels          replaces ray.cpp:1428
#pragma once

#include <cuda_runtime.h>

#include "path.h"

namespace ort {

// stats slots (uint64 each) in the per-scene device stats buffer
enum { STAT_SAMPLES = 0, STAT_RAYS = 1, STAT_NODE_VISITS = 2, STAT_BOX_TESTS = 3, STAT_SHAPE_TESTS = 4,
       STAT_WORK_COUNTER = 5, STAT_COUNT = 8 };

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
    for(int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// raycast_batch: persistent grid-stride loop, one ray per thread per trip.
// ---------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(128)
k_raycast(SceneView scene, unsigned long long n, const float *__restrict__ origins, const float *__restrict__ dirs,
          float *__restrict__ hit_t, uint32_t *__restrict__ prim_rank, uint32_t *__restrict__ mat_index,
          float *__restrict__ hit_normal, unsigned long long *__restrict__ stats)
{
    unsigned long long nodes = 0, boxes = 0, shapes = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for(unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        f3 o = mk3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        f3 d = mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        TraceHit hit; TraceCounters cnt; cnt.node_visits = cnt.box_tests = cnt.shape_tests = 0;
        trace<COUNT ? 1 : 0>(scene, o, d, &hit, &cnt);
        if(hit_t) hit_t[i] = hit.t;
        if(prim_rank) prim_rank[i] = hit.rank;
        if(mat_index || hit_normal)
        {
            uint32_t mat; f3 nrm;
            finish_hit(scene, hit, o, d, &mat, &nrm);
            if(mat_index) mat_index[i] = mat;
            if(hit_normal) { hit_normal[3 * i] = nrm.x; hit_normal[3 * i + 1] = nrm.y; hit_normal[3 * i + 2] = nrm.z; }
        }
        if(COUNT) { nodes += cnt.node_visits; boxes += cnt.box_tests; shapes += cnt.shape_tests; }
    }
    if(COUNT)
    {
        nodes = warp_sum(nodes); boxes = warp_sum(boxes); shapes = warp_sum(shapes);
        if((threadIdx.x & 31) == 0)
        {
            atomicAdd(&stats[STAT_NODE_VISITS], nodes);
            atomicAdd(&stats[STAT_BOX_TESTS], boxes);
            atomicAdd(&stats[STAT_SHAPE_TESTS], shapes);
        }
    }
}

// exhaustive closest hit: every thread walks ALL records (staged through shared
// memory, one tile of 256 records at a time) with the same intersectors and the
// same (t, rank) ordering -- the structure-free definition of the answer.
__global__ void __launch_bounds__(256)
k_raycast_brute(SceneView scene, unsigned long long n, const float *__restrict__ origins, const float *__restrict__ dirs,
                float *__restrict__ hit_t, uint32_t *__restrict__ prim_rank, uint32_t *__restrict__ mat_index)
{
    __shared__ q4 tile[3 * 256];
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    f3 o = mk3(0.f, 0.f, 0.f), d = mk3(1.f, 0.f, 0.f);
    if(active)
    {
        o = mk3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        d = mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    }
    float best_t = FLT_MAX; uint32_t best_rank = 0xFFFFFFFFu, best_mat = 0u;
    for(uint32_t base = 0; base < scene.prim_count; base += 256u)
    {
        uint32_t count = min(256u, scene.prim_count - base);
        __syncthreads();
        for(uint32_t k = threadIdx.x; k < 3u * count; k += blockDim.x) tile[k] = scene.prims[3u * base + k];
        __syncthreads();
        if(active)
        {
            for(uint32_t k = 0; k < count; ++k)
            {
                uint32_t rank, mat;
                const q4 *p = tile + 3u * k;
                q4 A = p[0], B = p[1], C = p[2];
                rank = f2u(A.w); mat = f2u(B.w);
                uint32_t kind = f2u(C.w);
                exact::Hit h;
                if((kind & 0xFFu) == PRIM_TRIANGLE) h = exact::triangle(q3(A), q3(B), q3(C), o, d);
                else if((kind & 0xFFu) == PRIM_AAB) h = exact::aab(q3(A), q3(B), o, d);
                else if((kind & 0xFFu) == PRIM_SPHERE) { int inner; h = exact::sphere(q3(A), B.x, o, d, &inner); }
                else
                {
                    const q4 *c = scene.cyl + 4u * (kind >> 8);
                    q4 c0 = ldq(c), c1 = ldq(c + 1), c2 = ldq(c + 2), c3 = ldq(c + 3);
                    exact::m3 rot; rot.r0 = q3(c1); rot.r1 = q3(c2); rot.r2 = q3(c3);
                    h = exact::cylinder_pre(q3(c0), rot, c0.w, c1.w, o, d);
                }
                if(h.t >= ORT_HIT_T_THRESHOLD && (h.t < best_t || (h.t == best_t && rank < best_rank)))
                {
                    best_t = h.t; best_rank = rank; best_mat = mat;
                }
            }
        }
    }
    if(active)
    {
        if(hit_t) hit_t[i] = best_t;
        if(prim_rank) prim_rank[i] = best_rank;
        if(mat_index) mat_index[i] = best_mat;
    }
}

// ---------------------------------------------------------------------------
// Render megakernel.
//
// Work item = one sample stream = (pixel, chunk): chunk_spp samples of one pixel
// drawn sequentially from the stream's own xorshift state (include/ort_b200.h).
// Items are numbered so that 32 consecutive items are an 8x4 pixel block of the
// same chunk; every thread pulls its next item from a global counter the moment
// its stream ends (path regeneration), so no lane idles while work remains.
// The loop body extends exactly ONE ray per trip -- primary or bounce alike -- so
// the lanes of a warp reconverge at the traversal, which is >95% of the work.
// ---------------------------------------------------------------------------
struct RenderArgs
{
    SceneView scene;
    PathConsts pc;
    int32_t tile_min_x, tile_min_y, tile_w, tile_h;     // tile rect
    uint32_t blocks_x, blocks_y;                        // 8x4 pixel blocks covering the tile
    uint32_t spp, chunk_spp, n_chunks, chunk_begin, chunk_count;
    uint32_t base_seed;
    unsigned long long total_items;                     // blocks_x*blocks_y*32*chunk_count
    unsigned long long *work_counter;                   // global item counter (zeroed before launch)
    unsigned long long *stats;
    long long *accum;                                   // int64[H*W*4] fixed-point sums, or null
    float *rgb;                                         // float3[H*W] final pixels (single-chunk mode), or null
};

template <bool COUNT>
__global__ void __launch_bounds__(128)
k_render_mega(const RenderArgs a)
{
    Path p;
    f3 color = mk3(0.f, 0.f, 0.f), focal_point = mk3(0.f, 0.f, 0.f);
    uint32_t samples_left = 0, pixel_index = 0xFFFFFFFFu;
    bool have_ray = false, primary = false;
    unsigned long long n_rays = 0, n_samples = 0, nodes = 0, boxes = 0, shapes = 0;

    for(;;)
    {
        if(!have_ray)
        {
            if(samples_left == 0)
            {
                // ---- accumulate the finished stream (ray.cpp:1428) ----
                if(pixel_index != 0xFFFFFFFFu)
                {
                    if(a.accum)
                    {
                        unsigned long long *dst = (unsigned long long *)(a.accum + 4ull * pixel_index);
                        atomicAdd(dst + 0, (unsigned long long)to_fixed(color.x));
                        atomicAdd(dst + 1, (unsigned long long)to_fixed(color.y));
                        atomicAdd(dst + 2, (unsigned long long)to_fixed(color.z));
                    }
                    else
                    {
                        f3 px = color / (float)a.spp;
                        a.rgb[3ull * pixel_index + 0] = px.x;
                        a.rgb[3ull * pixel_index + 1] = px.y;
                        a.rgb[3ull * pixel_index + 2] = px.z;
                    }
                    pixel_index = 0xFFFFFFFFu;
                }
                // ---- next stream ----
                int x, y; uint32_t chunk;
                for(;;)
                {
                    unsigned long long item = atomicAdd(a.work_counter, 1ull);
                    if(item >= a.total_items) { x = -1; break; }
                    uint32_t lane = (uint32_t)(item & 31ull);
                    unsigned long long blk = item >> 5;
                    uint32_t bx = (uint32_t)(blk % a.blocks_x); blk /= a.blocks_x;
                    uint32_t by = (uint32_t)(blk % a.blocks_y); blk /= a.blocks_y;
                    chunk = a.chunk_begin + (uint32_t)blk;
                    int lx = (int)(bx * 8u + (lane & 7u)), ly = (int)(by * 4u + (lane >> 3));
                    if(lx < a.tile_w && ly < a.tile_h) { x = a.tile_min_x + lx; y = a.tile_min_y + ly; break; }
                }
                if(x < 0) break;
                pixel_index = (uint32_t)(y * a.pc.width + x);
                samples_left = a.chunk_spp;
                if((chunk + 1u) * a.chunk_spp > a.spp) samples_left = a.spp - chunk * a.chunk_spp;
                p.series = ort_stream_seed(a.base_seed, pixel_index, chunk);
                color = mk3(0.f, 0.f, 0.f);
                focal_point = pixel_focal_point(a.pc, x, y);
            }
            // ---- generate (ray.cpp:1232-1246) ----
            generate_primary(a.pc, focal_point, &p);
            --samples_left;
            ++n_samples;
            primary = true;
        }
        // ---- extend (ray.cpp:1249, 1352) ----
        TraceHit hit; TraceCounters cnt; cnt.node_visits = cnt.box_tests = cnt.shape_tests = 0;
        trace<COUNT ? 1 : 2>(a.scene, p.origin, p.dir, &hit, &cnt);
        ++n_rays;
        shapes += cnt.shape_tests;           // always: the reference's entry point returns this tally (ray.cpp:1173)
        if(COUNT) { nodes += cnt.node_visits; boxes += cnt.box_tests; }
        uint32_t mat; f3 nrm;
        finish_hit(a.scene, hit, p.origin, p.dir, &mat, &nrm);
        // ---- shade (ray.cpp:1251-1277, 1355-1421) + head of the next bounce (ray.cpp:1280-1349) ----
        bool alive = primary ? shade_primary(a.pc, &p, hit.t, mat, nrm, &color)
                             : shade_bounce(a.pc, &p, hit.t, mat, nrm, &color);
        primary = false;
        have_ray = alive && next_bounce(a.pc, &p);
    }

    n_rays = warp_sum(n_rays); n_samples = warp_sum(n_samples); shapes = warp_sum(shapes);
    if(COUNT) { nodes = warp_sum(nodes); boxes = warp_sum(boxes); }
    if((threadIdx.x & 31) == 0)
    {
        atomicAdd(&a.stats[STAT_RAYS], n_rays);
        atomicAdd(&a.stats[STAT_SAMPLES], n_samples);
        atomicAdd(&a.stats[STAT_SHAPE_TESTS], shapes);
        if(COUNT)
        {
            atomicAdd(&a.stats[STAT_NODE_VISITS], nodes);
            atomicAdd(&a.stats[STAT_BOX_TESTS], boxes);
        }
    }
}

// accumulate: fixed-point sums -> mean radiance (include/ort_b200.h)
__global__ void k_resolve_fixed(const long long *__restrict__ accum, float *__restrict__ rgb,
                                int32_t width, int32_t min_x, int32_t min_y, int32_t w, int32_t h, uint32_t spp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= w * h) return;
    int x = min_x + i % w, y = min_y + i / w;
    size_t pix = (size_t)y * width + x;
    const double inv = 1.0 / (double)(1 << ORT_ACCUM_FRAC_BITS);
    for(int k = 0; k < 3; ++k)
        rgb[3 * pix + k] = (float)((double)accum[4 * pix + k] * inv) / (float)spp;
}

// Radiance RGBE pixel, v3_to_rgbe (macos_main.mm:242-261) with the same float operations as the
// host form (host_scene.cpp): frexpf is exact, the product and the quotient are IEEE (no
// contraction), roundf rounds halves away from zero on both sides.  Defined for the non-negative
// radiance the integrator produces (a negative component converts to 0 here).
__device__ __forceinline__ uint32_t rgbe_encode(float x, float y, float z)
{
    float m = ref_max(ref_max(x, y), z);
    if(!(m >= 1e-32f)) return 0u;
    int e;
    float denom = frexpf(m, &e) * 255.0f / m;
    return ((uint32_t)roundf(x * denom) << 0) | ((uint32_t)roundf(y * denom) << 8) |
           ((uint32_t)roundf(z * denom) << 16) | ((uint32_t)(e + 128) << 24);
}

// float3 pixels (buffer row 0 = bottom) -> RGBE words in FILE order: the writer emits buffer rows
// h-1 -> 0 (macos_main.mm:686-705), so output row r is buffer row h-1-r
__global__ void k_rgbe_from_rgb(const float *__restrict__ rgb, uint32_t *__restrict__ rgbe, int32_t w, int32_t h)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= w * h) return;
    int x = i % w, r = i / w;
    size_t pix = (size_t)(h - 1 - r) * w + x;
    rgbe[i] = rgbe_encode(rgb[3 * pix], rgb[3 * pix + 1], rgb[3 * pix + 2]);
}

// the same fused after the fixed-point resolve (ray.cpp:1428 + macos_main.mm:682-707): one pass
// from the int64 sums -- e.g. straight out of the NCCL reduce -- to the bytes of the .hdr file
__global__ void k_rgbe_from_fixed(const long long *__restrict__ accum, uint32_t *__restrict__ rgbe, int32_t w, int32_t h, uint32_t spp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= w * h) return;
    int x = i % w, r = i / w;
    size_t pix = (size_t)(h - 1 - r) * w + x;
    const double inv = 1.0 / (double)(1 << ORT_ACCUM_FRAC_BITS);
    float c[3];
    for(int k = 0; k < 3; ++k) c[k] = (float)((double)accum[4 * pix + k] * inv) / (float)spp;
    rgbe[i] = rgbe_encode(c[0], c[1], c[2]);
}

// FP32-pipe roofline denominator (MEASURED_PEAKS.json has no FP32 figure).
// Eight independent multiply-add chains per thread; with -fmad=false each `a*b+c` is
// an FMUL and an FADD, i.e. the same non-FMA instruction mix the intersectors issue.
// flops = threads * iters * 8 chains * 2.
__global__ void __launch_bounds__(256)
k_fp32_peak(float *out, int iters, float b, float c)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
#pragma unroll 4
    for(int i = 0; i < iters; ++i)
    {
        a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
        a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
    }
    float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if(s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true; keeps the chains alive
}

// Self-test of core_math.h's div3_shared (three IEEE quotients from one reciprocal) against the
// compiler's own division: each thread draws `iters` random (x, y, z, s) -- random sign and mantissa,
// exponent uniform in a window that straddles the fast path's range guard, and now and then a
// zero, a subnormal, an infinity or a NaN -- and counts the quotients whose bits differ.
__device__ __forceinline__ uint32_t st_next(uint32_t &x) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; return x; }
__device__ __forceinline__ float st_float(uint32_t &x)
{
    uint32_t r = st_next(x), m = st_next(x);
    uint32_t special = (r >> 16) & 63u;
    if(special == 0u) return __uint_as_float(r & 0x80000000u);                      // +-0
    if(special == 1u) return __uint_as_float((r & 0x80000000u) | (m & 0x007FFFFFu)); // subnormal
    if(special == 2u) return __uint_as_float((r & 0x80000000u) | 0x7F800000u);       // +-inf
    if(special == 3u) return __uint_as_float(0x7FC00000u | (m & 0x3FFFFFu));         // NaN
    uint32_t e = 127u - 70u + (r & 0xFFFFu) % 141u;                                  // 2^-70 .. 2^70
    return __uint_as_float((r & 0x80000000u) | (e << 23) | (m & 0x007FFFFFu));
}
__global__ void __launch_bounds__(256)
k_selftest_div3(uint32_t seed, int iters, unsigned long long *mismatches)
{
    uint32_t x = seed ^ ((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u);
    if(x == 0u) x = 1u;
    unsigned int bad = 0;
    for(int i = 0; i < iters; ++i)
    {
        f3 a = mk3(st_float(x), st_float(x), st_float(x));
        float s = st_float(x);
        f3 q = div3_shared(a, s);
        float rx = a.x / s, ry = a.y / s, rz = a.z / s;
        // NaNs compare as "both NaN": payloads of invalid operations are not part of the contract
        bad += ((q.x != q.x) && (rx != rx)) ? 0u : (__float_as_uint(q.x) != __float_as_uint(rx));
        bad += ((q.y != q.y) && (ry != ry)) ? 0u : (__float_as_uint(q.y) != __float_as_uint(ry));
        bad += ((q.z != q.z) && (rz != rz)) ? 0u : (__float_as_uint(q.z) != __float_as_uint(rz));
#ifdef ORT_SELFTEST_PRINT
        if(!((q.x != q.x) && (rx != rx)) && __float_as_uint(q.x) != __float_as_uint(rx) && atomicAdd(mismatches + 1, 1ull) < 12ull)
            printf("x %08x s %08x got %08x want %08x\n", __float_as_uint(a.x), __float_as_uint(s), __float_as_uint(q.x), __float_as_uint(rx));
#endif
    }
    bad = __reduce_add_sync(0xFFFFFFFFu, bad);
    if((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, (unsigned long long)bad);
}

__global__ void k_zero_u64(unsigned long long *p, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) p[i] = 0ull;
}

} // namespace ort
