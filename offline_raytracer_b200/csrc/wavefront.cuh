// wavefront.cuh -- the wavefront form of the radiance loop (code/ray.cpp:1178-1466):
// separate GENERATE / EXTEND / SHADE / ACCUMULATE stages over a pool of path slots that
// lives in HBM, instead of one loop per pixel.
//
//   k_wf_reset    marks every slot FRESH
//   k_wf_shade    per slot: SHADE the hit of the ray just extended (ray.cpp:1251-1277,
//                 1355-1421), draw the roulette / light pick / BSDF sample of the next bounce
//                 (ray.cpp:1280-1349); when the path ends: ACCUMULATE into the stream's sum
//                 (ray.cpp:1257,1364), and GENERATE the next camera ray of the slot's sample
//                 stream (ray.cpp:1215-1246) -- or pull the next (pixel, chunk) stream from the
//                 global work counter.  A slot therefore always leaves this kernel holding a
//                 ray (path regeneration), so the wavefront stays full until the work runs out
//                 and no pool-wide compaction pass is needed while it does.
//   k_wf_extend   per slot: closest hit of its ray (bvh.h) -> (t, primitive), the slot's shading key and the
//                 key histogram; the last block to finish turns the histogram into the sort cursors
//   k_wf_scatter  counting sort: perm[] = the slots grouped by shading key, for SHADE
//
// Why split: the single-kernel form is ~6000 SASS instructions (96 KB); ncu shows its
// warps starved by instruction-cache misses (icc hit rate 64 %, "no_instruction" the top
// stall) with 8 of 32 lanes active on average.  EXTEND alone is a 1600-instruction loop
// that stays cache resident and runs at 8-10 Grays/s on the same rays.
//
// Slot state: one 96-byte record per slot = 6 x 16 B = exactly three 32-byte sectors, so the
// SHADE stage -- which visits slots in material order, i.e. at random addresses -- moves only
// bytes it uses (a structure-of-arrays layout fetched a 32-byte sector for every 16 bytes and
// made SHADE HBM-bound: its time fell 23 % just by shrinking the pool into L2).  The fields are
// grouped by who touches them, because after the instruction diet of SHADE the stage moves
// ~3 TB/s of these records:
//   sector 0   q0  origin.xyz | hit t                          EXTEND reads it and writes t, prim
//              q1  direction.xyz | flags, then hit primitive   back into the SAME sector (no fill)
//   sector 1   q2  wo.xyz | xorshift state of the stream
//              q3  throughput weight.xyz | pixel index (29 bits) + state (2 bits) + primary (1 bit)
//   sector 2   q4  radiance sum of the stream so far .xyz | samples left     read and written only
//              q5  chunk index of the stream | - | - | -                     when a path ENDS
// A continuing path (4 of 5) therefore costs SHADE two sectors in and two out, not three.
// q1.w carries the flags (state, primary) from SHADE to EXTEND, which overwrites it with the hit
// primitive; q3.w keeps them for SHADE.
// plus key[] (shading key per slot) and perm[] (slots in key order), 4 B each, coalesced.
#pragma once

#include "kernels.cuh"

namespace ort {

enum { WF_DEAD = 0u, WF_ACTIVE = 1u, WF_FRESH = 2u };
#define WF_PIXEL_MASK 0x1FFFFFFFu           // also "no pixel yet"
__device__ __forceinline__ uint32_t wf_flags(uint32_t state, bool primary) { return (state << 29) | (primary ? 0x80000000u : 0u); }
__device__ __forceinline__ uint32_t wf_state_of(uint32_t word) { return (word >> 29) & 3u; }

// WF_SPLIT_COLD: the two quads only ending paths touch (q4, q5) live in an array of their own, so that the
// four hot quads are one 64-byte-aligned block -- exactly one L2 <- DRAM fetch at the default 64-byte
// fetch granularity (measured variants: profiles/README.md)
#ifndef WF_SPLIT_COLD
#define WF_SPLIT_COLD 1
#endif
#ifndef WF_REC_QUADS
#if WF_SPLIT_COLD
#define WF_REC_QUADS 4u
#else
#define WF_REC_QUADS 6u       // 8u pads the record to 128 bytes
#endif
#endif

struct WfBuffers
{
    float4 *rec;         // WF_REC_QUADS quads per slot
    float4 *cold;        // WF_SPLIT_COLD: q4, q5 of every slot (2 quads per slot); else unused
    uint32_t *key;       // shading key written by EXTEND: primary << 8 | min(material, 254); 511 = dead (so 256 | 255 must never be a live key)
    uint32_t *perm;      // slots grouped by key (counting sort)
    uint32_t capacity;
};

// 256-bit global loads / stores (sm_100: LDG.E.256 / STG.E.256): a whole 32-byte sector of a slot record per
// instruction -- half the memory instructions of SHADE's record traffic.  p must be 32-byte aligned (it is: hot
// blocks are 64-byte aligned, cold pairs 32-byte aligned).  Measured on B200 (SHADE ms, 128-bit vs 256-bit):
// C3 1080p x 64 spp 51.1 -> 49.1, C4 4K x 16 spp 49.4 -> 44.1, 4.4 M-triangle grid 28.5 -> 26.7; EXTEND unchanged.
#ifndef ORT_LDST256
#define ORT_LDST256 1
#endif
__device__ __forceinline__ void wf_ld2(const float4 *p, float4 &a, float4 &b)
{
#if ORT_LDST256
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
    a = p[0]; b = p[1];
#endif
}
__device__ __forceinline__ void wf_st2(float4 *p, float4 a, float4 b)
{
#if ORT_LDST256
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
#else
    p[0] = a; p[1] = b;
#endif
}

__device__ __forceinline__ float4 *wf_cold(const WfBuffers &wf, uint32_t slot)
{
#if WF_SPLIT_COLD
    return wf.cold + 2ull * slot;
#else
    return wf.rec + (size_t)WF_REC_QUADS * slot + 4u;
#endif
}

__global__ void k_wf_reset(WfBuffers wf)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < wf.capacity)
    {
        float4 *rec = wf.rec + (size_t)WF_REC_QUADS * i;
        rec[1] = make_float4(0.f, 0.f, 0.f, __uint_as_float(wf_flags(WF_FRESH, false)));
        rec[3] = make_float4(0.f, 0.f, 0.f, __uint_as_float(WF_PIXEL_MASK | wf_flags(WF_FRESH, false)));
        float4 *cold = wf_cold(wf, i);
        cold[0] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
        cold[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// shared-memory traversal stack: entry k of thread t lives at column t of row k, so the 32
// lanes of a warp touch 32 consecutive 8-byte words (no bank conflicts) and the stack never
// competes with BVH nodes for L1 lines
struct SharedStack
{
    uint2 *col;      // &smem[threadIdx.x]; rows of 128 threads
    __device__ __forceinline__ void put(int i, uint32_t a, uint32_t b) { col[i * 128] = make_uint2(a, b); }
    __device__ __forceinline__ void get(int i, uint32_t &a, uint32_t &b) const { uint2 v = col[i * 128]; a = v.x; b = v.y; }
};

// EXTEND's result goes back into the sector the ray came from: q0.w = t, q1.w = primitive
// (tried, B200: also storing the material here and loading the whole record + material up front in
//  SHADE, to shorten its chain of dependent loads -- SHADE 133.1 vs 130.7 ms, the extra live registers
//  cost more than the shorter chain saves)
__device__ __forceinline__ void wf_store_hit(const WfBuffers &wf, uint32_t slot, float t, uint32_t prim, uint32_t mat)
{
    (void)mat;
    float *r = reinterpret_cast<float *>(wf.rec + (size_t)WF_REC_QUADS * slot);
    r[3] = t;
    r[7] = __uint_as_float(prim);
}

#define WF_KEY_DEAD 511u
#define WF_KEY_BINS 512u

// EXTEND.  Persistent warps: each warp owns a contiguous range of slots and keeps its 32 lanes
// busy by handing the next rays of the range to lanes whose traversal has finished (dynamic
// fetch), instead of letting them idle until the slowest ray of a fixed batch of 32 is done --
// ncu: 8.5 of 32 lanes active without it.  Lanes then advance in lock step, one wide-node visit
// per trip.  Also emits the slot's shading key (primary flag, material of the hit) and counts it,
// for the counting sort that groups SHADE by material.
// resident 4-warp blocks per SM the register allocation must allow (measured on B200, 1080p:
// extend 8 / shade 6 = 116 ms; 6/6 = 119 ms; unconstrained (96 / 89 registers) = 127 ms;
// after the 96-byte slot records: shade 6 / 7 / 8 / 10 blocks = 139.8 / 135.1 / 131.3 / 138.1 ms per 1080p x 128 spp)
#ifndef ORT_EXTEND_MIN_BLOCKS
#define ORT_EXTEND_MIN_BLOCKS 8
#endif
#ifndef ORT_SHADE_MIN_BLOCKS
#define ORT_SHADE_MIN_BLOCKS 8
#endif
// measured on B200 (SHADE ms, C3 1080p x 64 spp / C4 4K x 16 spp / 4.4 M-triangle grid; 0 = off 49.0 / 44.1 / 26.8):
// 148: 48.7 / 43.9 / 26.6, 296: 48.3 / 43.3 / 26.4, 592: 47.4 / 42.0 / 26.2, 888: 47.4 / 42.2 / 26.2, 1184: 47.8 / 42.4 / 26.4,
// 2368: 49.1 / 43.9 / 27.1, 4736: 51.4 / 46.1 / 28.0; both sectors of the hot block (ORT_SHADE_PREFETCH_BOTH) 48.3 / 42.4 / 26.4 at 592
#ifndef ORT_SHADE_PREFETCH_BLOCKS
#define ORT_SHADE_PREFETCH_BLOCKS 592
#endif
// measured on B200 (two-pool frame ms, C3 1080p x 64 spp / C4 4K x 16 spp / 4.4 M-triangle grid): separate k_wf_scan + two
// memsets per iteration 144.2 / 132.6 / 143.8, fused 142.7 / 130.9 / 143.2
#ifndef ORT_EXTEND_FUSED_SCAN
#define ORT_EXTEND_FUSED_SCAN 1
#endif
// slots per thread in k_wf_scatter (the kernel is a chain of latencies per block -- keys, shared atomics, one global
// atomic per key present, the scattered store -- so more independent slots per chain is what speeds it up)
// measured on B200 (SORT ms per C3 1080p x 64 spp / C4 4K x 16 spp / grid, one pool): 1 slot per thread 12.7 / 12.5 / 7.7,
// 2: 8.7 / 8.3 / 5.4, 4: 6.9 / 6.6 / 4.4, 6: 6.6 / 6.2 / 4.3, 8: 8.5 / 8.0 / 5.3
#ifndef ORT_SCATTER_ITEMS
#define ORT_SCATTER_ITEMS 4
#endif
// threads per SHADE block = the domain of the regeneration compaction (measured on B200, steady state, profiles/README.md)
#ifndef ORT_SHADE_THREADS
#define ORT_SHADE_THREADS 128
#endif
// measured on B200 (steady state, two-pool frame ms, C3 x 256 spp / C4 4K x 128 spp / grid x 128 spp): 460.0 vs 463.7,
// 845.7 vs 855.0, 473.1 vs 476.0, images bit-identical
#ifndef ORT_SHADE_AAB_FAST
#define ORT_SHADE_AAB_FAST 1
#endif
#ifndef ORT_SHADE_PREFETCH_BOTH
#define ORT_SHADE_PREFETCH_BOTH 0
#endif
// dynamic shared memory: (wide-tree depth + 1) stack rows of 128 uint2 -- sized per scene, so a
// shallow tree does not pay for ORT_STACK_SIZE rows of occupancy
#ifndef ORT_EXTEND_FULL_STORE
#define ORT_EXTEND_FULL_STORE 0
#endif
#ifndef ORT_FETCH_MIN
#define ORT_FETCH_MIN 8      // refill when at least this many lanes are idle (or none has a ray)
#endif
#ifndef WF_CHUNK
#define WF_CHUNK 128u        // slots a warp takes from the global counter at a time (measured on B200: 64 / 128 / 256 / 512, profiles/README.md)
#endif
// one primitive test per trip instead of all the records a visit yielded (measured on B200, EXTEND ms:
// C3 170.0 -> 167.9, testscene 71.3 -> 70.5, 4.4 M-triangle grid 229.6 -> 208.7)
#ifndef ORT_EXTEND_ONE_PRIM
#define ORT_EXTEND_ONE_PRIM 1
#endif

template <bool COUNT>
__global__ void __launch_bounds__(128, ORT_EXTEND_MIN_BLOCKS)
k_wf_extend(SceneView scene, WfBuffers wf, uint32_t *chunk_counter, unsigned long long *stats, uint32_t *hist)
{
    extern __shared__ uint2 smem_stack[];
    __shared__ uint32_t sh_hist[WF_KEY_BINS];
    __shared__ uint32_t sh_done;
    if(threadIdx.x == 0) sh_done = 0u;
    for(uint32_t k = threadIdx.x; k < WF_KEY_BINS; k += blockDim.x) sh_hist[k] = 0u;
    __syncthreads();
    SharedStack st; st.col = smem_stack + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    // slot ranges are handed out in chunks from a global counter, so that all warps of the grid
    // finish together (static ranges left SMs half empty at the end of every launch)
    uint32_t next = 0u, end = 0u;
    bool exhausted = false;

    Trav t;
    t.ng_x = t.ng_y = 0u; t.sp = 0;
    bool has_ray = false;
    uint32_t slot = 0u, is_primary = 0u;
#if ORT_EXTEND_ONE_PRIM
    uint32_t pg_x = 0u, pg_y = 0u;
#endif
    unsigned long long nodes = 0, boxes = 0;
    uint32_t shapes = 0, rays = 0;          // per thread and launch: far below 2^32

    for(;;)
    {
        uint32_t idle_mask = __ballot_sync(0xFFFFFFFFu, !has_ray);
        if(!exhausted && (__popc(idle_mask) >= ORT_FETCH_MIN || idle_mask == 0xFFFFFFFFu))
        {
            if(next >= end)
            {
                uint32_t base = 0u;
                if(lane == 0) base = atomicAdd(chunk_counter, WF_CHUNK);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                next = base; end = base + WF_CHUNK;
                if(end > wf.capacity) end = wf.capacity;
                if(base >= wf.capacity) { exhausted = true; next = end = 0u; }
            }
            // hand consecutive slots of the chunk to the idle lanes
            uint32_t my = next + __popc(idle_mask & ((1u << lane) - 1u));
            if(!has_ray && my < end)
            {
                float4 ro, rd;
                wf_ld2(wf.rec + (size_t)WF_REC_QUADS * my, ro, rd);                                // one sector, one round trip
                if(wf_state_of(__float_as_uint(rd.w)) == WF_ACTIVE)
                {
                    trav_init(scene, t, st, mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z));
                    is_primary = __float_as_uint(rd.w) >> 31;
                    slot = my;
                    has_ray = true;
                    ++rays;
                }
                else { wf.key[my] = WF_KEY_DEAD; atomicAdd(&sh_hist[WF_KEY_DEAD], 1u); }
            }
            next += __popc(idle_mask);
            if(next > end) next = end;
        }
        if(__ballot_sync(0xFFFFFFFFu, has_ray) == 0u)
        {
            if(exhausted) break;
            continue;
        }
        {
            TraceCounters cnt; cnt.node_visits = cnt.box_tests = cnt.shape_tests = 0;
            // (tried and rejected, B200, C3: postponing the primitive tests of a step until >= 8 lanes
            //  have some, after Ylitie et al. 2017 -- EXTEND 349 ms vs 179 ms: the hit bound arrives late
            //  and far more nodes are visited)
#if ORT_EXTEND_ONE_PRIM
            // at most ONE primitive test per trip: a lane with records pending takes no node visit until
            // they are done (so its hit bound is as fresh as ever and it visits the same nodes), while the
            // other lanes go on visiting instead of idling through the longest record list of the warp
            bool done = false;
            if(has_ray)
            {
                if(pg_y == 0u && (t.ng_y & 0xFF000000u))
                {
                    uint32_t node_index;
                    trav_visit<COUNT ? 1 : 0>(scene, t, st, t.best_t, &cnt, &pg_x, &pg_y, &node_index);
                }
                if(pg_y != 0u)
                {
                    uint32_t prim = pg_x + lsb32(pg_y), rank, mat;
                    pg_y &= pg_y - 1u;
                    exact::Hit h = intersect_prim(scene, prim, t.o, t.d, t.inv, &rank, &mat);
                    (void)mat;
                    ++shapes;                // always: ort_tiled_raytrace_bvh returns this tally (ray.cpp:1173)
                    if(h.t >= ORT_HIT_T_THRESHOLD && (h.t < t.best_t || (h.t == t.best_t && rank < t.best_rank)))
                    {
                        t.best_t = h.t; t.best_prim = prim; t.best_rank = rank;
                    }
                }
                if(pg_y == 0u) done = !trav_next(t, st);
            }
#else
            bool done = has_ray && trav_step<COUNT ? 1 : 2>(scene, t, st, &cnt);
            shapes += cnt.shape_tests;
#endif
            if(COUNT) { nodes += cnt.node_visits; boxes += cnt.box_tests; }
            if(done)
            {
                uint32_t mat = 0u;
                if(t.best_prim != 0xFFFFFFFFu) mat = f2u(ldq(scene.prims + 3u * t.best_prim + 1u).w);
#if ORT_EXTEND_FULL_STORE
                // the whole sector (origin | t, direction | primitive) as two 128-bit stores instead of two 32-bit ones
                wf_st2(wf.rec + (size_t)WF_REC_QUADS * slot, make_float4(t.o.x, t.o.y, t.o.z, t.best_t),
                       make_float4(t.d.x, t.d.y, t.d.z, __uint_as_float(t.best_prim)));
#else
                wf_store_hit(wf, slot, t.best_t, t.best_prim, mat);
#endif
                uint32_t key = (is_primary ? 256u : 0u) | (mat < 254u ? mat : 254u);
                wf.key[slot] = key;
                atomicAdd(&sh_hist[key], 1u);
                has_ray = false;
            }
        }
    }
    rays = __reduce_add_sync(0xFFFFFFFFu, rays);
    shapes = __reduce_add_sync(0xFFFFFFFFu, shapes);
    if(COUNT) { nodes = warp_sum(nodes); boxes = warp_sum(boxes); }
    if(lane == 0 && rays)
    {
        atomicAdd(&stats[STAT_RAYS], (unsigned long long)rays);
        atomicAdd(&stats[STAT_SHAPE_TESTS], (unsigned long long)shapes);
        if(COUNT)
        {
            atomicAdd(&stats[STAT_NODE_VISITS], nodes);
            atomicAdd(&stats[STAT_BOX_TESTS], boxes);
        }
    }
    // the block's share of the key histogram (first pass of the counting sort), flushed by
    // whichever of its warps finishes last -- no barrier for the others to wait at
    __syncwarp();
    uint32_t finished = 0u;
    if(lane == 0) { __threadfence_block(); finished = atomicAdd(&sh_done, 1u); }
    finished = __shfl_sync(0xFFFFFFFFu, finished, 0);
    if(finished == 3u)
    {
        __threadfence_block();
        for(uint32_t k = lane; k < WF_KEY_BINS; k += 32u)
        {
            uint32_t v = ((volatile uint32_t *)sh_hist)[k];
            if(v) atomicAdd(&hist[k], v);
        }
#if ORT_EXTEND_FUSED_SCAN
        // The last block of the grid to get here turns the histogram into the cursors of the counting sort
        // (what k_wf_scan did in a launch of its own) and leaves the histogram, the chunk counter and the
        // arrival counter zeroed for the next EXTEND of the pool: one launch and two memsets less per iteration.
        // sort[] = hist[512] | cursor[512] | live | chunk counter | arrival counter (hist == sort).
        __threadfence();
        __syncwarp();                      // every lane's histogram atomics are performed before lane 0 announces the block
        uint32_t last = 0u;
        if(lane == 0) last = atomicAdd(hist + 2u * WF_KEY_BINS + 2u, 1u) == gridDim.x - 1u ? 1u : 0u;
        last = __shfl_sync(0xFFFFFFFFu, last, 0);
        if(last)
        {
            __threadfence();
            uint32_t v[WF_KEY_BINS / 32u], sum = 0u;
#pragma unroll
            for(uint32_t k = 0; k < WF_KEY_BINS / 32u; ++k) { v[k] = __ldcg(hist + (WF_KEY_BINS / 32u) * lane + k); sum += v[k]; }
            uint32_t incl = sum;
#pragma unroll
            for(uint32_t o = 1; o < 32u; o <<= 1)
            {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if(lane >= o) incl += u;
            }
            uint32_t run = incl - sum;
#pragma unroll
            for(uint32_t k = 0; k < WF_KEY_BINS / 32u; ++k)
            {
                uint32_t bin = (WF_KEY_BINS / 32u) * lane + k;
                hist[WF_KEY_BINS + bin] = run;                                   // cursor
                if(bin == WF_KEY_DEAD) hist[2u * WF_KEY_BINS] = run;             // live: everything before the dead bin
                run += v[k];
                hist[bin] = 0u;
            }
            if(lane == 0) { hist[2u * WF_KEY_BINS + 1u] = 0u; hist[2u * WF_KEY_BINS + 2u] = 0u; }
        }
#endif
    }
}

// Two other schedulings of this kernel were built, measured on B200 and removed (profiles/README.md,
// "tried and rejected"; the code is in the history: commits a82c651 and 1bcd9df): primitive tests
// redistributed over the lanes of a warp through shared-memory queues (238 vs 202 ms of EXTEND, 213 when
// drained at once) and a warp-wide vote on the next action -- node visit, triangle test or any test --
// with lanes holding their pending records (212-224 vs 168 ms).  Both were bit-identical to this kernel;
// in both the extra trips, ballots and shuffles cost more than the fuller primitive tests saved.

// ---- counting sort of the slots by shading key -------------------------------------------
// hist[k] = number of slots with key k; offsets = exclusive scan; perm = slots grouped by key.
// one block: exclusive scan of the 512 bins into cursor[], and live = slots that are not dead
__global__ void __launch_bounds__(512)
k_wf_scan(const uint32_t *hist, uint32_t *cursor, uint32_t *live)
{
    __shared__ uint32_t sh[WF_KEY_BINS];
    uint32_t v = hist[threadIdx.x];
    sh[threadIdx.x] = v;
    __syncthreads();
    for(uint32_t o = 1; o < WF_KEY_BINS; o <<= 1)
    {
        uint32_t add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    cursor[threadIdx.x] = sh[threadIdx.x] - v;
    if(threadIdx.x == WF_KEY_DEAD) *live = sh[threadIdx.x] - v;     // everything before the dead bin
}

// Two-level append: lanes of a warp with the same key are ranked with one match_any, warps of a
// block reserve their ranges in shared-memory counters, and the block then reserves its range of
// each key with ONE global atomic per key present -- a handful per 1024 slots instead of one per
// warp (ncu: the per-warp version spent 58 us per pass serialised on a few hot counters).
__global__ void __launch_bounds__(1024)
k_wf_scatter(WfBuffers wf, uint32_t *cursor)
{
    __shared__ uint32_t cnt[WF_KEY_BINS], base[WF_KEY_BINS];
    for(uint32_t k = threadIdx.x; k < WF_KEY_BINS; k += blockDim.x) cnt[k] = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    // the block takes ORT_SCATTER_ITEMS stripes of 1024 consecutive slots, thread t slot t of each: the 32 slots a warp
    // ranks at a time stay neighbours in perm[], which is what lets SHADE's warps share record lines (consecutive
    // slots PER THREAD measured 3 ... 7 % slower in SHADE)
    const uint32_t i0 = blockIdx.x * (1024u * ORT_SCATTER_ITEMS) + threadIdx.x;
    uint32_t key[ORT_SCATTER_ITEMS], local[ORT_SCATTER_ITEMS];
#pragma unroll
    for(int r = 0; r < ORT_SCATTER_ITEMS; ++r)
        key[r] = i0 + 1024u * r < wf.capacity ? wf.key[i0 + 1024u * r] : WF_KEY_BINS;      // padding: a key of its own, never counted
#pragma unroll
    for(int r = 0; r < ORT_SCATTER_ITEMS; ++r)
    {
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, key[r]);
        uint32_t leader = __ffs(peers) - 1u;
        uint32_t l = 0u;
        if(lane == leader && key[r] < WF_KEY_BINS) l = atomicAdd(&cnt[key[r]], (uint32_t)__popc(peers));
        local[r] = __shfl_sync(0xFFFFFFFFu, l, leader) + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    for(uint32_t k = threadIdx.x; k < WF_KEY_BINS; k += blockDim.x)
        if(cnt[k]) base[k] = atomicAdd(&cursor[k], cnt[k]);
    __syncthreads();
#pragma unroll
    for(int r = 0; r < ORT_SCATTER_ITEMS; ++r)
        if(key[r] < WF_KEY_BINS) wf.perm[base[key[r]] + local[r]] = i0 + 1024u * r;
}

// material + normalised normal of the winning record (ray.cpp:817).  Triangles -- almost all
// hits -- need only cross(e1, e2) of the stored vertices, the very value the intersector
// returns (ray.cpp:110); the analytic shapes re-run their intersector.
__device__ __forceinline__ void wf_finish_hit(const SceneView &s, uint32_t prim, f3 o, f3 d, uint32_t *mat, f3 *normal)
{
    if(prim == 0xFFFFFFFFu) { *mat = 0u; *normal = mk3(0.f, 0.f, 0.f); return; }
    const q4 *p = s.prims + 3u * prim;
    q4 A = ldq(p), B = ldq(p + 1), C = ldq(p + 2);
    *mat = f2u(B.w);
    if((f2u(C.w) & 0xFFu) == PRIM_TRIANGLE)
    {
        *normal = normalize(cross(q3(B) - q3(A), q3(C) - q3(A)));
        return;
    }
#if ORT_SHADE_AAB_FAST
    // a box returns one of (+-1, 0, 0), (0, +-1, 0), (0, 0, +-1) (core_math.h: aab_inv); normalize() of that is
    // the vector itself, bit for bit (length = sqrtf(1) = 1, x / 1 = x), so its square root and quotients are skipped
    if((f2u(C.w) & 0xFFu) == PRIM_AAB)
    {
        *normal = exact::aab_inv(q3(A), q3(B), o, mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z)).n;
        return;
    }
#endif
    TraceHit h; h.t = 0.f; h.prim = prim; h.rank = 0u;
    finish_hit(s, h, o, d, mat, normal);
}

// `sorted`: thread j handles slot perm[j] for j < *live (slots grouped by material, so the lanes
// of a warp mostly take the same branches of the BSDF code); else thread j handles slot j.
//
// Two phases per block.  Phase 1, every thread: shade its slot's hit and sample the next bounce.
// Phase 2: the slots whose path ended -- roulette, a light, a miss: about one in five, scattered
// over the warps -- are compacted through shared memory and the first threads of the block, with
// full warps, ACCUMULATE the finished stream, pull the next one and GENERATE the camera ray.  (ncu:
// done in place by the thread that owned the slot, these ~300 instructions ran with 6.8 of 32 lanes
// and every warp paid for them.)
struct WfRegen           // what phase 2 needs to know about a slot whose path ended, 32 B
{
    float dx, dy, dz;            // radiance the path adds to its stream's sum
    uint32_t series, pixel_index, slot, pad0, pad1;
};

__global__ void __launch_bounds__(ORT_SHADE_THREADS, ORT_SHADE_MIN_BLOCKS)
k_wf_shade(const RenderArgs a, WfBuffers wf, unsigned int *active_out, const uint32_t *live, int sorted)
{
    __shared__ WfRegen sh_regen[ORT_SHADE_THREADS];
    __shared__ uint32_t sh_count;
    if(threadIdx.x == 0) sh_count = 0u;
    __syncthreads();

    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n_samples = 0;
    unsigned int still_active = 0;
    uint32_t i = j;
    bool in_range = j < wf.capacity;
    if(sorted)
    {
        in_range = j < *live;
        if(in_range) i = wf.perm[j];
    }
    // ---- phase 1: shade + next bounce ----
    bool regen = false;
    WfRegen rg;
    rg.dx = rg.dy = rg.dz = 0.f; rg.series = 0u; rg.pixel_index = WF_PIXEL_MASK; rg.slot = i; rg.pad0 = rg.pad1 = 0u;
    if(in_range)
    {
        float4 *rec = wf.rec + (size_t)WF_REC_QUADS * i;
        float4 ro, rd, swo, sw;
        wf_ld2(rec, ro, rd); wf_ld2(rec + 2, swo, sw);
#if ORT_SHADE_PREFETCH_BLOCKS
        // SHADE waits on the record of a slot it only learns from perm[] (ncu: 20 % of its warp samples): ask L2
        // for the record of the thread ORT_SHADE_PREFETCH_BLOCKS blocks ahead -- about the blocks resident at
        // once -- while this thread waits for its own
        if(sorted)
        {
            uint32_t jp = j + 128u * ORT_SHADE_PREFETCH_BLOCKS;      // in units of 128 slots whatever the block size
            if(jp < *live)
            {
                const float4 *pp = wf.rec + (size_t)WF_REC_QUADS * wf.perm[jp];
                asm volatile("prefetch.global.L2 [%0];" :: "l"(pp));
#if ORT_SHADE_PREFETCH_BOTH
                asm volatile("prefetch.global.L2 [%0];" :: "l"(pp + 2));
#endif
            }
        }
#endif
        const uint32_t word = __float_as_uint(sw.w);
        const uint32_t state = wf_state_of(word);
        if(state == WF_ACTIVE)
        {
            Path p;
            p.origin = mk3(ro.x, ro.y, ro.z); p.dir = mk3(rd.x, rd.y, rd.z);
            p.wo = mk3(swo.x, swo.y, swo.z); p.series = __float_as_uint(swo.w);
            p.weight = mk3(sw.x, sw.y, sw.z);
            const uint32_t pixel_index = word & WF_PIXEL_MASK;
            const bool primary = (word >> 31) != 0u;
            const float hit_t = ro.w;
            const uint32_t hit_prim = __float_as_uint(rd.w);
            // what this hit adds to the stream's radiance sum: `sum = sum + delta` below is the very
            // addition ray.cpp:1257 / 1364 perform (a path adds at most once, when it ends on a light)
            f3 delta = mk3(0.f, 0.f, 0.f);
            p.normal = mk3(0.f, 0.f, 0.f); p.mat = 0u;
            uint32_t mat; f3 nrm;
            wf_finish_hit(a.scene, hit_prim, p.origin, p.dir, &mat, &nrm);
            bool alive = primary ? shade_primary(a.pc, &p, hit_t, mat, nrm, &delta)
                                 : shade_bounce(a.pc, &p, hit_t, mat, nrm, &delta);
            if(alive && next_bounce(a.pc, &p))
            {
                wf_st2(rec, make_float4(p.origin.x, p.origin.y, p.origin.z, 0.f),
                       make_float4(p.dir.x, p.dir.y, p.dir.z, __uint_as_float(wf_flags(WF_ACTIVE, false))));
                wf_st2(rec + 2, make_float4(p.wo.x, p.wo.y, p.wo.z, __uint_as_float(p.series)),
                       make_float4(p.weight.x, p.weight.y, p.weight.z, __uint_as_float(pixel_index | wf_flags(WF_ACTIVE, false))));
                still_active = 1;
            }
            else
            {
                regen = true;
                rg.dx = delta.x; rg.dy = delta.y; rg.dz = delta.z;
                rg.series = p.series; rg.pixel_index = pixel_index;
            }
        }
        else if(state == WF_FRESH) regen = true;      // no stream yet: samples left 0, no pixel
    }
    // ---- compaction of the slots to regenerate ----
    {
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t m = __ballot_sync(0xFFFFFFFFu, regen);
        uint32_t base = 0u;
        if(lane == 0 && m) base = atomicAdd(&sh_count, (uint32_t)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if(regen) sh_regen[base + __popc(m & ((1u << lane) - 1u))] = rg;
    }
    __syncthreads();
    // ---- phase 2: accumulate / next stream / generate, with full warps ----
    if(threadIdx.x < sh_count)
    {
        rg = sh_regen[threadIdx.x];
        float4 *rec = wf.rec + (size_t)WF_REC_QUADS * rg.slot;
        float4 *cold = wf_cold(wf, rg.slot);
        float4 scol, sch;                                   // the third sector: only paths that end touch it
        wf_ld2(cold, scol, sch);
        f3 color = mk3(scol.x, scol.y, scol.z) + mk3(rg.dx, rg.dy, rg.dz);
        uint32_t samples_left = __float_as_uint(scol.w), pixel_index = rg.pixel_index, chunk = __float_as_uint(sch.x);
        Path p;
        p.series = rg.series;
        bool dead = false;
        if(samples_left == 0)
        {
            // ---- accumulate the finished stream (ray.cpp:1428) ----
            if(pixel_index != WF_PIXEL_MASK)
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
            }
            // ---- next stream from the global work counter ----
            int x = -1, y = -1;
            for(;;)
            {
                unsigned long long item = atomicAdd(a.work_counter, 1ull);
                if(item >= a.total_items) break;
                uint32_t lane = (uint32_t)(item & 31ull);
                unsigned long long blk = item >> 5;
                uint32_t bx = (uint32_t)(blk % a.blocks_x); blk /= a.blocks_x;
                uint32_t by = (uint32_t)(blk % a.blocks_y); blk /= a.blocks_y;
                int lx = (int)(bx * 8u + (lane & 7u)), ly = (int)(by * 4u + (lane >> 3));
                if(lx < a.tile_w && ly < a.tile_h)
                {
                    x = a.tile_min_x + lx; y = a.tile_min_y + ly;
                    chunk = a.chunk_begin + (uint32_t)blk;
                    break;
                }
            }
            if(x < 0) dead = true;
            else
            {
                pixel_index = (uint32_t)(y * a.pc.width + x);
                samples_left = a.chunk_spp;
                if((chunk + 1u) * a.chunk_spp > a.spp) samples_left = a.spp - chunk * a.chunk_spp;
                p.series = ort_stream_seed(a.base_seed, pixel_index, chunk);
                color = mk3(0.f, 0.f, 0.f);
            }
        }
        if(dead)
        {
            rec[1] = make_float4(0.f, 0.f, 0.f, __uint_as_float(wf_flags(WF_DEAD, false)));
            rec[3] = make_float4(0.f, 0.f, 0.f, __uint_as_float(WF_PIXEL_MASK | wf_flags(WF_DEAD, false)));
        }
        else
        {
            // ---- generate (ray.cpp:1215-1246) ----
            int x = (int)(pixel_index % (uint32_t)a.pc.width), y = (int)(pixel_index / (uint32_t)a.pc.width);
            generate_primary(a.pc, pixel_focal_point(a.pc, x, y), &p);
            --samples_left;
            n_samples = 1;
            wf_st2(rec, make_float4(p.origin.x, p.origin.y, p.origin.z, 0.f),
                   make_float4(p.dir.x, p.dir.y, p.dir.z, __uint_as_float(wf_flags(WF_ACTIVE, true))));
            wf_st2(rec + 2, make_float4(p.wo.x, p.wo.y, p.wo.z, __uint_as_float(p.series)),
                   make_float4(p.weight.x, p.weight.y, p.weight.z, __uint_as_float(pixel_index | wf_flags(WF_ACTIVE, true))));
            wf_st2(cold, make_float4(color.x, color.y, color.z, __uint_as_float(samples_left)),
                   make_float4(__uint_as_float(chunk), 0.f, 0.f, 0.f));
            still_active += 1;
        }
    }
    n_samples = warp_sum(n_samples);
    still_active = __reduce_add_sync(0xFFFFFFFFu, still_active);
    if((threadIdx.x & 31) == 0)
    {
        if(n_samples) atomicAdd(&a.stats[STAT_SAMPLES], n_samples);
        if(still_active) atomicAdd(active_out, still_active);
    }
}

} // namespace ort
