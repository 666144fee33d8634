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
//   k_wf_extend   per slot: closest hit of its ray (bvh.h) -> (t, primitive)
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
// plus key[]: one shading-class byte per slot, written by EXTEND and sorted tile by tile inside SHADE.
#pragma once

#include "kernels.cuh"

namespace ort {

enum { WF_DEAD = 0u, WF_ACTIVE = 1u, WF_FRESH = 2u };
#define WF_PIXEL_MASK 0x1FFFFFFFu           // also "no pixel yet"
__device__ __forceinline__ uint32_t wf_flags(uint32_t state, bool primary) { return (state << 29) | (primary ? 0x80000000u : 0u); }
__device__ __forceinline__ uint32_t wf_state_of(uint32_t word) { return (word >> 29) & 3u; }

#define WF_REC_QUADS 6u

struct WfBuffers
{
    float4 *rec;         // WF_REC_QUADS quads per slot
    uint8_t *key;        // shading class written by EXTEND (WF_KEY_*), one byte per slot; 255 = dead slot
    uint32_t capacity;
};

__global__ void k_wf_reset(WfBuffers wf)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < wf.capacity)
    {
        float4 *rec = wf.rec + (size_t)WF_REC_QUADS * i;
        rec[1] = make_float4(0.f, 0.f, 0.f, __uint_as_float(wf_flags(WF_FRESH, false)));
        rec[3] = make_float4(0.f, 0.f, 0.f, __uint_as_float(WF_PIXEL_MASK | wf_flags(WF_FRESH, false)));
        rec[4] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
        rec[5] = make_float4(0.f, 0.f, 0.f, 0.f);
        wf.key[i] = (uint8_t)64u;       // WF_KEY_FRESH
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

// The ray's own constants live in shared memory too, three float4 per thread in column layout:
//   row 0  origin.xyz | sign octant (7 - oct) + bit 8: "some |d| component is tiny: exact 1/d differs from the clamped one"
//   row 1  clamped reciprocal direction of the slab tests (bvh.h, SlabConsts)
//   row 2  direction.xyz | slot index
// A node visit reads rows 0-1, a primitive test rows 0 and 2; what stays in registers is the traversal
// state proper (node group, stack pointer, best hit, pending records): the per-visit re-derivation of
// the clamped reciprocal and the octant (18 instructions of 290) becomes two LDS.128, and the nine
// registers that held origin / direction / reciprocal are free for the scheduler.
struct TravS
{
    float4 *ray;                  // &smem_ray[threadIdx.x]; rows of 128 threads
    uint32_t ng_x, ng_y;
    int sp;
    float best_t;
    uint32_t best_prim, best_rank;
};
__device__ __forceinline__ SlabConsts slab_consts(const TravS &t)
{
    float4 a = t.ray[0], b = t.ray[128];
    SlabConsts c;
    c.ox = a.x; c.oy = a.y; c.oz = a.z; c.octinv = __float_as_uint(a.w) & 7u;
    c.idx = b.x; c.idy = b.y; c.idz = b.z;
    return c;
}
__device__ __forceinline__ void travs_set_ray(TravS &t, f3 o, f3 d, uint32_t slot)
{
    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);         // IEEE, as ray.cpp:210 forms it for box shapes
    SlabConsts c = make_slab_consts(o, d, inv);
    uint32_t inexact = (c.idx != inv.x || c.idy != inv.y || c.idz != inv.z) ? 8u : 0u;
    t.ray[0] = make_float4(o.x, o.y, o.z, __uint_as_float(c.octinv | inexact));
    t.ray[128] = make_float4(c.idx, c.idy, c.idz, 0.0f);
    t.ray[256] = make_float4(d.x, d.y, d.z, __uint_as_float(slot));
}
// exact intersection of one record with the ray kept in shared memory (bvh.h: intersect_prim)
__device__ __forceinline__ exact::Hit travs_intersect(const SceneView &s, const TravS &t, uint32_t prim, uint32_t *rank)
{
    const q4 *p = s.prims + 3u * prim;
    q4 A = ldq(p), B = ldq(p + 1), C = ldq(p + 2);
    *rank = f2u(A.w);
    const uint32_t kind = f2u(C.w);
    float4 ro = t.ray[0];
    f3 o = mk3(ro.x, ro.y, ro.z);
    if((kind & 0xFFu) == PRIM_AAB)
    {
        float4 ri = t.ray[128];
        f3 inv = mk3(ri.x, ri.y, ri.z);
        if(__float_as_uint(ro.w) & 8u)
        {
            float4 rd = t.ray[256];
            inv = mk3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
        }
        return exact::aab_inv(q3(A), q3(B), o, inv);
    }
    float4 rd = t.ray[256];
    f3 d = mk3(rd.x, rd.y, rd.z);
    if((kind & 0xFFu) == PRIM_TRIANGLE) return exact::triangle(q3(A), q3(B), q3(C), o, d);
    if((kind & 0xFFu) == PRIM_SPHERE) { int inner; return exact::sphere(q3(A), B.x, o, d, &inner); }
    const q4 *c = s.cyl + 4u * (kind >> 8);
    q4 c0 = ldq(c), c1 = ldq(c + 1), c2 = ldq(c + 2), c3 = ldq(c + 3);
    exact::m3 rot; rot.r0 = q3(c1); rot.r1 = q3(c2); rot.r2 = q3(c3);
    return exact::cylinder_pre(q3(c0), rot, c0.w, c1.w, o, d);
}

// EXTEND's result goes back into the sector the ray came from: q0.w = t, q1.w = primitive
__device__ __forceinline__ void wf_store_hit(const WfBuffers &wf, uint32_t slot, float t, uint32_t prim)
{
    float *r = reinterpret_cast<float *>(wf.rec + (size_t)WF_REC_QUADS * slot);
    r[3] = t;
    r[7] = __uint_as_float(prim);
}

// Shading class of a hit = what decides the code path SHADE takes for it: primary or bounce ray, triangle
// or analytic shape (the latter re-runs its intersector for the normal), emitter (the path ends) and
// which BSDF lobes the material has.  EXTEND writes one byte per slot; SHADE sorts each tile of slots by it.
#define WF_KEY_PRIMARY 32u
#define WF_KEY_TRIANGLE 16u
#define WF_KEY_CLASSES 64u          // classes 0..63
#define WF_KEY_FRESH 64u            // slot without a stream yet (after k_wf_reset): regeneration only
#define WF_KEY_BINS 65u
#define WF_KEY_DEAD 255u

// EXTEND.  Persistent warps: each warp owns a contiguous range of slots and keeps its 32 lanes
// busy by handing the next rays of the range to lanes whose traversal has finished (dynamic
// fetch), instead of letting them idle until the slowest ray of a fixed batch of 32 is done --
// ncu: 8.5 of 32 lanes active without it.  Lanes then advance in lock step, one wide-node visit
// per trip.  Also emits the slot's shading class for the tile sort of SHADE.
// resident 4-warp blocks per SM the register allocation must allow (measured on B200, 1080p:
// extend 8 / shade 6 = 116 ms; 6/6 = 119 ms; unconstrained (96 / 89 registers) = 127 ms;
// after the 96-byte slot records: shade 6 / 7 / 8 / 10 blocks = 139.8 / 135.1 / 131.3 / 138.1 ms per 1080p x 128 spp)
#ifndef ORT_EXTEND_MIN_BLOCKS
#define ORT_EXTEND_MIN_BLOCKS 8
#endif
#ifndef ORT_SHADE_MIN_BLOCKS
#define ORT_SHADE_MIN_BLOCKS 8
#endif
// dynamic shared memory: 3 rows of ray constants (float4) + (wide-tree depth + 2) stack rows of 128 uint2 --
// sized per scene, so a shallow tree does not pay for ORT_STACK_SIZE rows of occupancy
#define WF_RAY_ROWS 3u
#ifndef ORT_FETCH_MIN
#define ORT_FETCH_MIN 8      // refill when at least this many lanes are idle (or none has a ray)
#endif
#define WF_CHUNK 256u        // slots a warp takes from the global counter at a time

template <bool COUNT>
__global__ void __launch_bounds__(128, ORT_EXTEND_MIN_BLOCKS)
k_wf_extend(SceneView scene, WfBuffers wf, uint32_t *chunk_counter, unsigned long long *stats)
{
    extern __shared__ float4 smem_dyn[];
    float4 *smem_ray = smem_dyn;                                                   // WF_RAY_ROWS x 128 float4
    uint2 *smem_stack = reinterpret_cast<uint2 *>(smem_dyn + WF_RAY_ROWS * 128u);  // rows of 128 uint2
    SharedStack st; st.col = smem_stack + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    // slot ranges are handed out in chunks from a global counter, so that all warps of the grid
    // finish together (static ranges left SMs half empty at the end of every launch)
    uint32_t next = 0u, end = 0u;
    bool exhausted = false;

    TravS t;
    t.ray = smem_ray + threadIdx.x;
    t.ng_x = t.ng_y = 0u; t.sp = 0;
    t.best_t = FLT_MAX; t.best_prim = t.best_rank = 0xFFFFFFFFu;
    bool has_ray = false;
    uint32_t is_primary = 0u;
    uint32_t pg_x = 0u, pg_y = 0u;
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
                float4 ro = wf.rec[WF_REC_QUADS * my], rd = wf.rec[WF_REC_QUADS * my + 1u];      // one sector, one round trip
                if(wf_state_of(__float_as_uint(rd.w)) == WF_ACTIVE)
                {
                    travs_set_ray(t, mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z), my);
                    trav_init_state(scene, t, st);
                    is_primary = __float_as_uint(rd.w) >> 31;
                    has_ray = true;
                    ++rays;
                }
                else wf.key[my] = (uint8_t)WF_KEY_DEAD;
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
            // at most ONE primitive test per trip: a lane with records pending takes no node visit until
            // they are done (so its hit bound is as fresh as ever and it visits the same nodes), while the
            // other lanes go on visiting instead of idling through the longest record list of the warp
            // (measured on B200, EXTEND ms: C3 170.0 -> 167.9, testscene 71.3 -> 70.5, 4.4 M-triangle grid 229.6 -> 208.7)
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
                    uint32_t prim = pg_x + lsb32(pg_y), rank;
                    pg_y &= pg_y - 1u;
                    exact::Hit h = travs_intersect(scene, t, prim, &rank);
                    ++shapes;                // always: ort_tiled_raytrace_bvh returns this tally (ray.cpp:1173)
                    if(h.t >= ORT_HIT_T_THRESHOLD && (h.t < t.best_t || (h.t == t.best_t && rank < t.best_rank)))
                    {
                        t.best_t = h.t; t.best_prim = prim; t.best_rank = rank;
                    }
                }
                if(pg_y == 0u) done = !trav_next(t, st);
            }
            if(COUNT) { nodes += cnt.node_visits; boxes += cnt.box_tests; }
            if(done)
            {
                const uint32_t slot = __float_as_uint(t.ray[256].w);
                uint32_t key = is_primary ? WF_KEY_PRIMARY : 0u;
                if(t.best_prim != 0xFFFFFFFFu)
                {
                    const q4 *p = scene.prims + 3u * t.best_prim;
                    uint32_t mat = f2u(ldq(p + 1u).w), kind = f2u(ldq(p + 2u).w) & 0xFFu;
                    key |= (uint32_t)scene.mat_class[mat] | (kind == PRIM_TRIANGLE ? WF_KEY_TRIANGLE : 0u);
                }
                wf_store_hit(wf, slot, t.best_t, t.best_prim);
                wf.key[slot] = (uint8_t)key;
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
}

// Two other schedulings of this kernel were built, measured on B200 and removed (profiles/README.md,
// "tried and rejected"; the code is in the history: commits a82c651 and 1bcd9df): primitive tests
// redistributed over the lanes of a warp through shared-memory queues (238 vs 202 ms of EXTEND, 213 when
// drained at once) and a warp-wide vote on the next action -- node visit, triangle test or any test --
// with lanes holding their pending records (212-224 vs 168 ms).  Both were bit-identical to this kernel;
// in both the extra trips, ballots and shuffles cost more than the fuller primitive tests saved.

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
    TraceHit h; h.t = 0.f; h.prim = prim; h.rank = 0u;
    finish_hit(s, h, o, d, mat, normal);
}

// SHADE.  Persistent blocks take TILES of WF_TILE consecutive slots from a global counter.  Per tile:
//   1. the tile's WF_TILE class bytes are read (coalesced) and counting-sorted IN SHARED MEMORY into a local
//      permutation; every class starts at a multiple of 32, so that no warp ever mixes two classes (the gaps
//      are idle lanes, which cost a latency-bound kernel nothing);
//   2. rounds of 128 permutation entries: every thread shades its slot's hit and samples the next bounce;
//   3. the slots whose path ended -- roulette, a light, a miss: about one in five -- are collected in shared
//      memory and, whenever 128 of them wait (and at the end of the tile), full warps ACCUMULATE the finished
//      stream, pull the next one and GENERATE its camera ray (done in place by the owning thread these ~300
//      instructions ran with 6.8 of 32 lanes and every warp paid for them).
// This replaces the global counting sort of round 1 (histogram in EXTEND, k_wf_scan, k_wf_scatter, a 4-byte key
// and a 4-byte permutation entry per slot): the gathers of a block now stay inside one contiguous
// WF_TILE x 96-byte region of the pool, which the memory system serves at close to streaming efficiency, where
// the global permutation made every 64-byte record access a random DRAM access (ncu, round 1: long-scoreboard
// stall 7.1 cycles per issue, 2.9 TB/s of sector traffic at 45 % of the HBM copy peak).
#ifndef WF_TILE
#define WF_TILE 2048u
#endif
#define WF_PERM_MAX (WF_TILE + 32u * WF_KEY_BINS + 32u)
#define WF_REGEN_MAX 256u
struct WfRegen           // what the regeneration pass needs to know about a slot whose path ended, 16 B
{
    float dx, dy, dz;            // radiance the path adds to its stream's sum
    uint32_t slot;               // slot index in the pool
};

__global__ void __launch_bounds__(128, ORT_SHADE_MIN_BLOCKS)
k_wf_shade(const RenderArgs a, WfBuffers wf, unsigned int *active_out, uint32_t *tile_counter, int sorted)
{
    __shared__ __align__(16) uint16_t sh_perm[WF_PERM_MAX];
    __shared__ uint32_t sh_cnt[WF_KEY_BINS + 3], sh_start[WF_KEY_BINS + 3];
    __shared__ WfRegen sh_regen[WF_REGEN_MAX];
    __shared__ uint32_t sh_regen_series[WF_REGEN_MAX], sh_regen_pixel[WF_REGEN_MAX];
    __shared__ uint32_t sh_count, sh_tile, sh_total;

    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long n_samples = 0;
    unsigned int still_active = 0;
    uint32_t waiting = 0u;                                 // entries of sh_regen in use: the same in every thread
    if(threadIdx.x == 0) sh_count = 0u;

    for(;;)
    {
        __syncthreads();                                   // previous tile completely done
        if(threadIdx.x == 0) sh_tile = atomicAdd(tile_counter, 1u);
        for(uint32_t k = threadIdx.x; k < WF_KEY_BINS; k += blockDim.x) sh_cnt[k] = 0u;
        for(uint32_t k = threadIdx.x; k < WF_PERM_MAX / 2u; k += blockDim.x) reinterpret_cast<uint32_t *>(sh_perm)[k] = 0xFFFFFFFFu;
        __syncthreads();
        const uint32_t base = sh_tile * WF_TILE;
        if(base >= wf.capacity) break;
        const uint32_t n = min(WF_TILE, wf.capacity - base);

        // ---- 1. counting sort of the tile by shading class ----
        uint32_t keys[WF_TILE / 128u / 4u];                // 4 class bytes per word
        uint32_t pos[WF_TILE / 128u / 2u];                 // 2 positions (within the class) per word
#pragma unroll
        for(uint32_t r = 0; r < WF_TILE / 128u; ++r)
        {
            const uint32_t i = r * 128u + threadIdx.x;
            uint32_t key = i < n ? (uint32_t)wf.key[base + i] : WF_KEY_DEAD;
            if(!sorted && key != WF_KEY_DEAD) key = 0u;    // evidence runs: one class, i.e. slot order
            if(key > WF_KEY_FRESH) key = WF_KEY_DEAD;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
            const uint32_t leader = __ffs(peers) - 1u;
            uint32_t p = 0u;
            if(lane == leader && key != WF_KEY_DEAD) p = atomicAdd(&sh_cnt[key], (uint32_t)__popc(peers));
            p = __shfl_sync(0xFFFFFFFFu, p, leader) + __popc(peers & ((1u << lane) - 1u));
            if((r & 3u) == 0u) keys[r / 4u] = 0u;
            keys[r / 4u] |= key << (8u * (r & 3u));
            if((r & 1u) == 0u) pos[r / 2u] = 0u;
            pos[r / 2u] |= p << (16u * (r & 1u));
        }
        __syncthreads();
        if(threadIdx.x < 32u)
        {
            // exclusive scan of the class counts, each rounded up to a multiple of 32 (3 bins per lane: 65 <= 96)
            uint32_t c0 = 0u, c1 = 0u, c2 = 0u;
            const uint32_t b0 = 3u * lane;
            if(b0 < WF_KEY_BINS) c0 = (sh_cnt[b0] + 31u) & ~31u;
            if(b0 + 1u < WF_KEY_BINS) c1 = (sh_cnt[b0 + 1u] + 31u) & ~31u;
            if(b0 + 2u < WF_KEY_BINS) c2 = (sh_cnt[b0 + 2u] + 31u) & ~31u;
            uint32_t sum = c0 + c1 + c2, incl = sum;
            for(uint32_t o = 1u; o < 32u; o <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if(lane >= o) incl += v; }
            uint32_t excl = incl - sum;
            if(b0 < WF_KEY_BINS) sh_start[b0] = excl;
            if(b0 + 1u < WF_KEY_BINS) sh_start[b0 + 1u] = excl + c0;
            if(b0 + 2u < WF_KEY_BINS) sh_start[b0 + 2u] = excl + c0 + c1;
            if(lane == 31u) sh_total = incl;
        }
        __syncthreads();
#pragma unroll
        for(uint32_t r = 0; r < WF_TILE / 128u; ++r)
        {
            const uint32_t key = (keys[r / 4u] >> (8u * (r & 3u))) & 0xFFu;
            const uint32_t p = (pos[r / 2u] >> (16u * (r & 1u))) & 0xFFFFu;
            if(key != WF_KEY_DEAD) sh_perm[sh_start[key] + p] = (uint16_t)(r * 128u + threadIdx.x);
        }
        __syncthreads();
        const uint32_t total = sh_total;                   // multiple of 32

        // ---- 2. rounds of 128 entries of the local permutation ----
        for(uint32_t j0 = 0u; j0 < total; j0 += 128u)
        {
            const uint32_t j = j0 + threadIdx.x;
            const uint32_t e = j < total ? (uint32_t)sh_perm[j] : 0xFFFFu;
            bool regen = false;
            WfRegen rg; rg.dx = rg.dy = rg.dz = 0.f; rg.slot = 0u;
            uint32_t rg_series = 0u, rg_pixel = WF_PIXEL_MASK;
            if(e != 0xFFFFu)
            {
                const uint32_t i = base + e;
                rg.slot = i;
                float4 *rec = wf.rec + (size_t)WF_REC_QUADS * i;
                float4 ro = rec[0], rd = rec[1], swo = rec[2], sw = rec[3];
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
                        rec[0] = make_float4(p.origin.x, p.origin.y, p.origin.z, 0.f);
                        rec[1] = make_float4(p.dir.x, p.dir.y, p.dir.z, __uint_as_float(wf_flags(WF_ACTIVE, false)));
                        rec[2] = make_float4(p.wo.x, p.wo.y, p.wo.z, __uint_as_float(p.series));
                        rec[3] = make_float4(p.weight.x, p.weight.y, p.weight.z, __uint_as_float(pixel_index | wf_flags(WF_ACTIVE, false)));
                        still_active += 1;
                    }
                    else
                    {
                        regen = true;
                        rg.dx = delta.x; rg.dy = delta.y; rg.dz = delta.z;
                        rg_series = p.series; rg_pixel = pixel_index;
                    }
                }
                else if(state == WF_FRESH) regen = true;      // no stream yet: samples left 0, no pixel
            }
            // ---- collect the slots to regenerate ----
            {
                uint32_t m = __ballot_sync(0xFFFFFFFFu, regen);
                uint32_t at = 0u;
                if(lane == 0 && m) at = atomicAdd(&sh_count, (uint32_t)__popc(m));
                at = __shfl_sync(0xFFFFFFFFu, at, 0) + __popc(m & ((1u << lane) - 1u));
                if(regen) { sh_regen[at] = rg; sh_regen_series[at] = rg_series; sh_regen_pixel[at] = rg_pixel; }
            }
            waiting += (uint32_t)__syncthreads_count(regen ? 1 : 0);   // barrier + block-wide count in one
            if(waiting < 128u && j0 + 128u < total) continue;        // uniform: keep collecting
            // ---- 3. accumulate / next stream / generate, with full warps ----
            for(uint32_t k = threadIdx.x; k < waiting; k += 128u)
            {
                rg = sh_regen[k];
                float4 *rec = wf.rec + (size_t)WF_REC_QUADS * rg.slot;
                float4 scol = rec[4], sch = rec[5];                 // the third sector: only paths that end touch it
                f3 color = mk3(scol.x, scol.y, scol.z) + mk3(rg.dx, rg.dy, rg.dz);
                uint32_t samples_left = __float_as_uint(scol.w), pixel_index = sh_regen_pixel[k], chunk = __float_as_uint(sch.x);
                Path p;
                p.series = sh_regen_series[k];
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
                        uint32_t ln = (uint32_t)(item & 31ull);
                        unsigned long long blk = item >> 5;
                        uint32_t bx = (uint32_t)(blk % a.blocks_x); blk /= a.blocks_x;
                        uint32_t by = (uint32_t)(blk % a.blocks_y); blk /= a.blocks_y;
                        int lx = (int)(bx * 8u + (ln & 7u)), ly = (int)(by * 4u + (ln >> 3));
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
                    n_samples += 1;
                    rec[0] = make_float4(p.origin.x, p.origin.y, p.origin.z, 0.f);
                    rec[1] = make_float4(p.dir.x, p.dir.y, p.dir.z, __uint_as_float(wf_flags(WF_ACTIVE, true)));
                    rec[2] = make_float4(p.wo.x, p.wo.y, p.wo.z, __uint_as_float(p.series));
                    rec[3] = make_float4(p.weight.x, p.weight.y, p.weight.z, __uint_as_float(pixel_index | wf_flags(WF_ACTIVE, true)));
                    rec[4] = make_float4(color.x, color.y, color.z, __uint_as_float(samples_left));
                    rec[5] = make_float4(__uint_as_float(chunk), 0.f, 0.f, 0.f);
                    still_active += 1;
                }
            }
            __syncthreads();
            if(threadIdx.x == 0) sh_count = 0u;
            waiting = 0u;
            __syncthreads();
        }
    }
    n_samples = warp_sum(n_samples);
    still_active = __reduce_add_sync(0xFFFFFFFFu, still_active);
    if(lane == 0)
    {
        if(n_samples) atomicAdd(&a.stats[STAT_SAMPLES], n_samples);
        if(still_active) atomicAdd(active_out, still_active);
    }
}

} // namespace ort
