// bvh.h -- flattened scene layout in HBM and the closest-hit traversal ("extend"),
// shared by every kernel.  Replaces raycast_top_most_node / raycast_bvh
// (code/ray.cpp:624-822, 1165-1176) and the pointer-based BVHOctreeNode octree
// they walk (code/ray.h:115-133).
//
// LAYOUT (all 16-byte aligned, read with 128-bit loads):
//   WideNode, 80 B = 5 x 16 B: an 8-wide BVH node whose child boxes are
//   quantised to 8 bits per plane on a per-node power-of-two grid
//   (origin p, biased exponents e), rounded OUTWARD so the quantised box always
//   contains the child's float box (after Ylitie, Karras, Laine 2017,
//   "Efficient incoherent ray traversal on GPUs through compressed wide BVHs").
//     q0: p.x p.y p.z | ex ey ez imask            imask bit s = slot s is an inner node
//     q1: child_base | prim_base | meta[0..3] | meta[4..7]
//     q2: qlo_x[0..7] | qlo_y[0..7]
//     q3: qlo_z[0..7] | qhi_x[0..7]
//     q4: qhi_y[0..7] | qhi_z[0..7]
//   meta[s]: 0 = empty; inner child: 0b001_11sss (low 5 bits = 24 + s);
//            leaf child: (unary prim count 1/3/7) << 5 | offset of its first
//            primitive from prim_base (<= 24 primitives per node, <= 3 per leaf).
//   Inner children are stored contiguously from child_base in slot order.
//   Slots are assigned so that slot s lies towards octant s of the node
//   (bit0 = +x, bit1 = +y, bit2 = +z); a ray with sign-octant `oct` visits hit
//   children in order of (s ^ (7 - oct)), highest first = front to back.
//
//   PrimRec, 48 B = 3 x 16 B: one record of the reference's leaf push buffers.
//     q0: a.xyz | rank      q1: b.xyz | mat_index      q2: c.xyz | kind + (aux << 8)
//     triangle: a,b,c = v0,v1,v2 UNTRANSFORMED -- the reference subtracts at test
//               time (ray.cpp:87-92) and bit-exact t requires the same operands;
//     sphere: a = center, b.x = r;   box: a = min, b = max;
//     cylinder: aux indexes CylinderAux (base, rotation rows, |axis|, r).
//   `rank` is the record's position in the order the reference's raycast_bvh
//   would test records with no culling; exact-t ties go to the lowest rank,
//   which reproduces "first tested wins" (strict `<` at ray.cpp:653-708).
//
// The traversal is closest-hit over the primitives with the reference's own
// intersectors (core_math.h, namespace exact); the result is therefore the
// argmin over ALL records of (t, rank) -- independent of the acceleration
// structure -- provided the node boxes are conservative.  They are: every
// primitive box is padded before quantisation (bvh_build.cpp) by more than the
// rounding of the slab arithmetic below, and culling is inclusive (tmin <= tmax).
#pragma once

#include <string.h>

#include "core_math.h"

namespace ort {

struct alignas(16) q4 { float x, y, z, w; };

#ifndef ORT_NODE_PAD
#define ORT_NODE_PAD 0
#endif
struct alignas(16) WideNode
{
    float px, py, pz;
    uint8_t ex, ey, ez, imask;
    uint32_t child_base, prim_base;       // prim_base bit 31 (ORT_NODE_MIXED_KINDS): some leaf record of the node is not a triangle
    uint8_t meta[8];
    uint8_t qlo_x[8], qlo_y[8];
    uint8_t qlo_z[8], qhi_x[8];
    uint8_t qhi_y[8], qhi_z[8];
#if ORT_NODE_PAD
    uint8_t pad[16];                      // measured variant: 96-byte nodes = three aligned 32-byte sectors, fetched as 3 x 256 bits
#endif
};
#if ORT_NODE_PAD
static_assert(sizeof(WideNode) == 96, "padded WideNode is 6 x 16 B");
#else
static_assert(sizeof(WideNode) == 80, "WideNode is 5 x 16 B");
#endif
#define ORT_NODE_QUADS ((uint32_t)(sizeof(WideNode) / 16u))

struct alignas(16) PrimRec
{
    float ax, ay, az; uint32_t rank;
    float bx, by, bz; uint32_t mat;
    float cx, cy, cz; uint32_t kind;
};
static_assert(sizeof(PrimRec) == 48, "PrimRec is 3 x 16 B");

#define ORT_NODE_MIXED_KINDS 0x80000000u
enum { PRIM_TRIANGLE = 0, PRIM_SPHERE = 1, PRIM_AAB = 2, PRIM_CYLINDER = 3 };

struct alignas(16) CylinderAux
{
    float base[3]; float axis_len;
    float r0[3]; float radius;
    float r1[3]; float pad0;
    float r2[3]; float pad1;
};
static_assert(sizeof(CylinderAux) == 64, "CylinderAux is 4 x 16 B");

struct SceneView
{
    const q4 *nodes;           // WideNode[node_count] viewed as q4[5*node_count]
    const q4 *prims;           // PrimRec[prim_count] viewed as q4[3*prim_count]
    const q4 *cyl;             // CylinderAux[] viewed as q4[4*n]
    uint32_t node_count;
    uint32_t prim_count;
    // Two trees share the node array.  Nodes [0, main_root) hold the SPHERES and
    // are traversed WITHOUT clipping to the best hit so far: a ray tangent to a
    // sphere is reported by the reference at t = -b/(2a), half way to the sphere
    // (code/ray.cpp:174-183), so a nearer hit must not cull the sphere's box.
    // The clipped trees follow: boxes + cylinders in [main_root, tri_root) (absent if equal),
    // triangles from tri_root.
    uint32_t main_root;
    uint32_t tri_root;
};

struct TraceCounters { uint32_t node_visits, box_tests, shape_tests; };

struct TraceHit
{
    float t;          // FLT_MAX on a miss (ray.cpp:627)
    uint32_t prim;    // index into prims, 0xFFFFFFFF on a miss
    uint32_t rank;    // ORT_MISS_RANK on a miss
};

// ---- portable bit helpers --------------------------------------------------
#if defined(__CUDA_ARCH__)
ORT_HD q4 ldq(const q4 *p) { float4 v = __ldg(reinterpret_cast<const float4 *>(p)); q4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
ORT_HD uint32_t f2u(float f) { return __float_as_uint(f); }
ORT_HD float u2f(uint32_t u) { return __uint_as_float(u); }
ORT_HD uint32_t popc32(uint32_t v) { return (uint32_t)__popc(v); }
ORT_HD uint32_t msb32(uint32_t v) { return 31u - (uint32_t)__clz((int)v); }
ORT_HD uint32_t lsb32(uint32_t v) { return (uint32_t)__ffs((int)v) - 1u; }
#else
ORT_HD q4 ldq(const q4 *p) { return *p; }
ORT_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
ORT_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
ORT_HD uint32_t popc32(uint32_t v) { return (uint32_t)__builtin_popcount(v); }
ORT_HD uint32_t msb32(uint32_t v) { return 31u - (uint32_t)__builtin_clz(v); }
ORT_HD uint32_t lsb32(uint32_t v) { return (uint32_t)__builtin_ctz(v); }
#endif

ORT_HD f3 q3(q4 v) { return mk3(v.x, v.y, v.z); }

// exact intersection of one record (dispatch on kind); returns t (-1 = miss) and
// the UNNORMALISED normal, as the reference's intersectors do.
ORT_HD exact::Hit intersect_prim(const SceneView &s, uint32_t prim, f3 o, f3 d, f3 inv, uint32_t *rank, uint32_t *mat)
{
    const q4 *p = s.prims + 3u * prim;
    q4 A = ldq(p), B = ldq(p + 1), C = ldq(p + 2);
    *rank = f2u(A.w);
    *mat = f2u(B.w);
    uint32_t kind = f2u(C.w);
    if((kind & 0xFFu) == PRIM_TRIANGLE) return exact::triangle(q3(A), q3(B), q3(C), o, d);
    if((kind & 0xFFu) == PRIM_AAB) return exact::aab_inv(q3(A), q3(B), o, inv);
    if((kind & 0xFFu) == PRIM_SPHERE) { int inner; return exact::sphere(q3(A), B.x, o, d, &inner); }
    const q4 *c = s.cyl + 4u * (kind >> 8);
    q4 c0 = ldq(c), c1 = ldq(c + 1), c2 = ldq(c + 2), c3 = ldq(c + 3);
    exact::m3 rot; rot.r0 = q3(c1); rot.r1 = q3(c2); rot.r2 = q3(c3);
    return exact::cylinder_pre(q3(c0), rot, c0.w, c1.w, o, d);
}

#ifndef ORT_STACK_SIZE
#define ORT_STACK_SIZE 32
#endif

// byte k (0..3) of w as a float, without the quarter-rate I2F conversion unit: PRMT builds
// the bit pattern of 2^23 + byte, and subtracting 2^23 is exact.
#if defined(__CUDA_ARCH__)
ORT_HD float byte_f(uint32_t w, uint32_t k) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + k)) - 8388608.0f; }
// the same value through the conversion unit (XU pipe): used for a share of the planes so that the
// ALU pipe, the busiest in the node test, is relieved (ORT_I2F_PLANES of the 6 planes per child)
ORT_HD float byte_f_xu(uint32_t w, uint32_t k) { return (float)((w >> (8u * k)) & 0xFFu); }
#else
ORT_HD float byte_f(uint32_t w, uint32_t k) { return (float)((w >> (8u * k)) & 0xFFu); }
ORT_HD float byte_f_xu(uint32_t w, uint32_t k) { return (float)((w >> (8u * k)) & 0xFFu); }
#endif
// measured on B200 (extend ms per 1080p x 128 spp, C3): 0 planes 179.1, 2: 174.0, 4: 169.8, 5: 168.4, 6: 168.9
#ifndef ORT_I2F_PLANES
#define ORT_I2F_PLANES 5
#endif

// Packed f32x2 FMA for the slab test (sm_100: fma.rn.f32x2, __ffma2_rn) -- built, bit-identical, measured on
// B200 and left OFF: 24 FFMA2 instead of 48 FFMA per node visit (SASS: 1344 -> 1264 instructions), yet EXTEND
// is ~1 % SLOWER (C3 1080p x 64 spp 90.3 vs 89.2 ms, C4 83.9 vs 82.8, 4.4 M-triangle grid 115.3 vs 113.9): the
// FMA pipe was never the busy one (ncu: fma 25 %, alu 58 %, xu 52 %) and the register pairs the packed form
// needs cost more than the saved issue slots.  -DORT_SLAB_FMA2=1 selects it (device only).
#ifndef ORT_SLAB_FMA2
#define ORT_SLAB_FMA2 0
#endif
#if ORT_SLAB_FMA2 && !(defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000))
#undef ORT_SLAB_FMA2
#define ORT_SLAB_FMA2 0
#endif

// traversal stack of pending node groups: a per-thread array here; the wavefront extend
// kernel substitutes a shared-memory column (wavefront.cuh)
struct LocalStack
{
    uint32_t x[ORT_STACK_SIZE], y[ORT_STACK_SIZE];
    ORT_HD void put(int i, uint32_t a, uint32_t b) { x[i] = a; y[i] = b; }
    ORT_HD void get(int i, uint32_t &a, uint32_t &b) const { a = x[i]; b = y[i]; }
};

// state of one ray's traversal, so that a kernel can interleave rays step by step
struct Trav
{
    f3 o, d;
    f3 inv;                       // (1/d.x, 1/d.y, 1/d.z), IEEE, as ray.cpp:210 forms it for box shapes
    uint32_t ng_x, ng_y;          // current node group: base index | hit bits << 24 | imask
    int sp;
    float best_t;
    uint32_t best_prim, best_rank;
};

// what a node visit needs to know about the ray: origin, the reciprocal direction of the (conservative)
// slab tests -- the exact one, except that zero and denormal-small components are clamped so that
// 0 * inf never appears and (p - o) * id cannot overflow -- and the sign octant
struct SlabConsts { float ox, oy, oz, idx, idy, idz; uint32_t octinv; };
#define ORT_SLAB_TINY 1e-20f
#define ORT_SLAB_HUGE 1e20f
ORT_HD SlabConsts make_slab_consts(f3 o, f3 d, f3 inv)
{
    SlabConsts c;
    c.ox = o.x; c.oy = o.y; c.oz = o.z;
    c.idx = fabsf(d.x) > ORT_SLAB_TINY ? inv.x : (f2u(d.x) >> 31 ? -ORT_SLAB_HUGE : ORT_SLAB_HUGE);
    c.idy = fabsf(d.y) > ORT_SLAB_TINY ? inv.y : (f2u(d.y) >> 31 ? -ORT_SLAB_HUGE : ORT_SLAB_HUGE);
    c.idz = fabsf(d.z) > ORT_SLAB_TINY ? inv.z : (f2u(d.z) >> 31 ? -ORT_SLAB_HUGE : ORT_SLAB_HUGE);
    c.octinv = 7u - ((c.idx < 0.0f ? 1u : 0u) | (c.idy < 0.0f ? 2u : 0u) | (c.idz < 0.0f ? 4u : 0u));
    return c;
}
ORT_HD SlabConsts slab_consts(const Trav &t) { return make_slab_consts(t.o, t.d, t.inv); }

template <class T, class Stack>
ORT_HD void trav_init_state(const SceneView &s, T &t, Stack &st)
{
    t.best_t = FLT_MAX; t.best_prim = 0xFFFFFFFFu; t.best_rank = 0xFFFFFFFFu;
    // node group: a root as the single hit child of a virtual parent (imask 0 => relative index 0).
    // Order: analytic shapes (their hit clips what follows), triangles, spheres (unclipped).
    t.ng_x = s.main_root; t.ng_y = 0x80000000u;
    t.sp = 0;
    if(s.main_root != 0u) { st.put(t.sp, 0u, 0x80000000u); ++t.sp; }
    if(s.tri_root != s.main_root) { st.put(t.sp, s.tri_root, 0x80000000u); ++t.sp; }
}

template <class Stack>
ORT_HD void trav_init(const SceneView &s, Trav &t, Stack &st, f3 o, f3 d)
{
    t.o = o; t.d = d;
    t.best_t = FLT_MAX; t.best_prim = 0xFFFFFFFFu; t.best_rank = 0xFFFFFFFFu;
    t.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    // node group: a root as the single hit child of a virtual parent (imask 0 => relative index 0).
    // Order: analytic shapes (their hit clips what follows), triangles, spheres (unclipped).
    t.ng_x = s.main_root; t.ng_y = 0x80000000u;
    t.sp = 0;
    if(s.main_root != 0u) { st.put(t.sp, 0u, 0x80000000u); ++t.sp; }
    if(s.tri_root != s.main_root) { st.put(t.sp, s.tri_root, 0x80000000u); ++t.sp; }
}

// Visits the nearest pending child of the current node group (which must have node bits set):
// tests its 8 children against the ray clipped to [0, clip_t], leaves the hit inner children in
// the node group and returns the hit leaf primitives as a group (base index, bit mask) plus the
// index of the node visited (its tree decides the kind of the primitives and whether clip_t
// applies: nodes below main_root -- the sphere tree -- are never clipped).
// COUNT: 0 = no counters, 1 = all three (the counters build), 2 = primitive tests only (the product
// build: the reference's entry point returns that tally, ray.cpp:661-715, 1173)
template <int COUNT, class T, class Stack>
ORT_HD void trav_visit(const SceneView &s, T &t, Stack &st, float clip_t, TraceCounters *cnt,
                       uint32_t *tg_x_out, uint32_t *tg_y_out, uint32_t *node_index_out, uint32_t *mixed_out = 0)
{
    const SlabConsts rc = slab_consts(t);
    const float idx = rc.idx, idy = rc.idy, idz = rc.idz;
    const uint32_t octinv = rc.octinv;
    const bool nx = ((7u - octinv) & 1u) != 0u, ny = ((7u - octinv) & 2u) != 0u, nz = ((7u - octinv) & 4u) != 0u;

    uint32_t bit = msb32(t.ng_y);
    uint32_t hits_imask = t.ng_y;
    t.ng_y &= ~(1u << bit);
    if(t.ng_y & 0xFF000000u)
    {
        if(t.sp < ORT_STACK_SIZE) { st.put(t.sp, t.ng_x, t.ng_y); ++t.sp; }
    }
    uint32_t slot = (bit - 24u) ^ octinv;
    uint32_t rel = popc32(hits_imask & ~(0xFFFFFFFFu << slot) & 0xFFu);
    const uint32_t node_index = t.ng_x + rel;
    const float t_clip = (node_index >= s.main_root) ? clip_t : FLT_MAX;
    const q4 *np = s.nodes + ORT_NODE_QUADS * node_index;
#if ORT_NODE_PAD && defined(__CUDA_ARCH__)
    q4 n0, n1, n2, n3, n4, n5;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w) : "l"(np));
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=f"(n3.x), "=f"(n3.y), "=f"(n3.z), "=f"(n3.w) : "l"(np + 2));
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(n4.x), "=f"(n4.y), "=f"(n4.z), "=f"(n4.w), "=f"(n5.x), "=f"(n5.y), "=f"(n5.z), "=f"(n5.w) : "l"(np + 4));
    (void)n5;
#else
    q4 n0 = ldq(np), n1 = ldq(np + 1), n2 = ldq(np + 2), n3 = ldq(np + 3), n4 = ldq(np + 4);
#endif
    if(COUNT == 1) cnt->node_visits++;

    uint32_t e_imask = f2u(n0.w);
    float ax = u2f((e_imask & 0xFFu) << 23) * idx;
    float ay = u2f(((e_imask >> 8) & 0xFFu) << 23) * idy;
    float az = u2f(((e_imask >> 16) & 0xFFu) << 23) * idz;
    float bx = (n0.x - rc.ox) * idx;
    float by = (n0.y - rc.oy) * idy;
    float bz = (n0.z - rc.oz) * idz;

    t.ng_x = f2u(n1.x);
    uint32_t hitmask = 0u;
    const uint32_t octinv4 = octinv * 0x01010101u;
#if ORT_SLAB_FMA2
    // Blackwell's packed FP32 pipe: the six plane distances of TWO children are three fma.rn.f32x2
    // pairs per side -- (x, y) of child k, (x, y) of child k + 1, (z of k, z of k + 1) -- i.e. 24
    // issue slots per node instead of 48.  Each half of an f32x2 FMA rounds exactly like the scalar
    // fmaf, so the hit masks (and the host build, which keeps the scalar form) are unchanged.
    const float2 a_xy = make_float2(ax, ay), b_xy = make_float2(bx, by);
    const float2 a_zz = make_float2(az, az), b_zz = make_float2(bz, bz);
#endif
#pragma unroll
    for(int half = 0; half < 2; ++half)
    {
        uint32_t meta4 = half ? f2u(n1.w) : f2u(n1.z);
        uint32_t lox = half ? f2u(n2.y) : f2u(n2.x);
        uint32_t loy = half ? f2u(n2.w) : f2u(n2.z);
        uint32_t loz = half ? f2u(n3.y) : f2u(n3.x);
        uint32_t hix = half ? f2u(n3.w) : f2u(n3.z);
        uint32_t hiy = half ? f2u(n4.y) : f2u(n4.x);
        uint32_t hiz = half ? f2u(n4.w) : f2u(n4.z);
        uint32_t nearx = nx ? hix : lox, farx = nx ? lox : hix;
        uint32_t neary = ny ? hiy : loy, fary = ny ? loy : hiy;
        uint32_t nearz = nz ? hiz : loz, farz = nz ? loz : hiz;
        // per byte: inner children (low 5 bits >= 24, i.e. bits 3 and 4 set) get their slot
        // XORed with octinv so that the highest hit bit is the nearest child
        uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        uint32_t inner_mask4 = (is_inner4 >> 4) * 0xFFu;
        // per byte m: bits 5-7 = what to set (unary primitive count, or 1 for an inner child), bits 0-4 = where;
        // a child's contribution to the hit mask is (m >> 5) << (m & 31) -- the shifter wraps at 32 on its own
        uint32_t mx4 = meta4 ^ (octinv4 & inner_mask4 & 0x07070707u);
#if ORT_SLAB_FMA2
#pragma unroll
        for(uint32_t k = 0; k < 4; k += 2)
        {
            float2 n0 = __ffma2_rn(make_float2(ORT_I2F_PLANES >= 6 ? byte_f_xu(nearx, k) : byte_f(nearx, k),
                                               ORT_I2F_PLANES >= 5 ? byte_f_xu(neary, k) : byte_f(neary, k)), a_xy, b_xy);
            float2 n1_ = __ffma2_rn(make_float2(ORT_I2F_PLANES >= 6 ? byte_f_xu(nearx, k + 1) : byte_f(nearx, k + 1),
                                                ORT_I2F_PLANES >= 5 ? byte_f_xu(neary, k + 1) : byte_f(neary, k + 1)), a_xy, b_xy);
            float2 nz2 = __ffma2_rn(make_float2(ORT_I2F_PLANES >= 4 ? byte_f_xu(nearz, k) : byte_f(nearz, k),
                                                ORT_I2F_PLANES >= 4 ? byte_f_xu(nearz, k + 1) : byte_f(nearz, k + 1)), a_zz, b_zz);
            float2 f0 = __ffma2_rn(make_float2(ORT_I2F_PLANES >= 3 ? byte_f_xu(farx, k) : byte_f(farx, k),
                                               ORT_I2F_PLANES >= 2 ? byte_f_xu(fary, k) : byte_f(fary, k)), a_xy, b_xy);
            float2 f1 = __ffma2_rn(make_float2(ORT_I2F_PLANES >= 3 ? byte_f_xu(farx, k + 1) : byte_f(farx, k + 1),
                                               ORT_I2F_PLANES >= 2 ? byte_f_xu(fary, k + 1) : byte_f(fary, k + 1)), a_xy, b_xy);
            float2 fz2 = __ffma2_rn(make_float2(ORT_I2F_PLANES >= 1 ? byte_f_xu(farz, k) : byte_f(farz, k),
                                                ORT_I2F_PLANES >= 1 ? byte_f_xu(farz, k + 1) : byte_f(farz, k + 1)), a_zz, b_zz);
            float tmin0 = fmaxf(fmaxf(n0.x, n0.y), fmaxf(nz2.x, 0.0f)), tmax0 = fminf(fminf(f0.x, f0.y), fminf(fz2.x, t_clip));
            float tmin1 = fmaxf(fmaxf(n1_.x, n1_.y), fmaxf(nz2.y, 0.0f)), tmax1 = fminf(fminf(f1.x, f1.y), fminf(fz2.y, t_clip));
            if(COUNT == 1) { if((meta4 >> (8u * k)) & 0xFFu) cnt->box_tests++; if((meta4 >> (8u * (k + 1u))) & 0xFFu) cnt->box_tests++; }
            uint32_t m0 = (mx4 >> (8u * k)) & 0xFFu, m1 = (mx4 >> (8u * (k + 1u))) & 0xFFu;
            hitmask |= (tmin0 <= tmax0) ? ((m0 >> 5) << (m0 & 31u)) : 0u;
            hitmask |= (tmin1 <= tmax1) ? ((m1 >> 5) << (m1 & 31u)) : 0u;
        }
#else
#pragma unroll
        for(uint32_t k = 0; k < 4; ++k)
        {
            float t0x = fmaf(ORT_I2F_PLANES >= 6 ? byte_f_xu(nearx, k) : byte_f(nearx, k), ax, bx);
            float t0y = fmaf(ORT_I2F_PLANES >= 5 ? byte_f_xu(neary, k) : byte_f(neary, k), ay, by);
            float t0z = fmaf(ORT_I2F_PLANES >= 4 ? byte_f_xu(nearz, k) : byte_f(nearz, k), az, bz);
            float t1x = fmaf(ORT_I2F_PLANES >= 3 ? byte_f_xu(farx, k) : byte_f(farx, k), ax, bx);
            float t1y = fmaf(ORT_I2F_PLANES >= 2 ? byte_f_xu(fary, k) : byte_f(fary, k), ay, by);
            float t1z = fmaf(ORT_I2F_PLANES >= 1 ? byte_f_xu(farz, k) : byte_f(farz, k), az, bz);
            float tmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
            float tmax = fminf(fminf(t1x, t1y), fminf(t1z, t_clip));
            if(COUNT == 1) { if((meta4 >> (8u * k)) & 0xFFu) cnt->box_tests++; }
            uint32_t m = (mx4 >> (8u * k)) & 0xFFu;
            hitmask |= (tmin <= tmax) ? ((m >> 5) << (m & 31u)) : 0u;
        }
#endif
    }
    t.ng_y = (hitmask & 0xFF000000u) | (e_imask >> 24);
    *tg_x_out = f2u(n1.y) & ~ORT_NODE_MIXED_KINDS;
    if(mixed_out) *mixed_out = f2u(n1.y) >> 31;
    *tg_y_out = hitmask & 0x00FFFFFFu;
    *node_index_out = node_index;
}

// Pops the next pending node group when the current one is exhausted; false = traversal done.
template <class T, class Stack>
ORT_HD bool trav_next(T &t, Stack &st)
{
    if(t.ng_y & 0xFF000000u) return true;
    if(t.sp == 0) return false;
    --t.sp;
    st.get(t.sp, t.ng_x, t.ng_y);
    return true;
}

// One step: visit one wide node, test the primitives it yielded, pop.  Returns true when the
// traversal is complete.
template <int COUNT, class Stack>
ORT_HD bool trav_step(const SceneView &s, Trav &t, Stack &st, TraceCounters *cnt)
{
    uint32_t tg_x = 0u, tg_y = 0u, node_index = 0u;
    if(t.ng_y & 0xFF000000u) trav_visit<COUNT>(s, t, st, t.best_t, cnt, &tg_x, &tg_y, &node_index);
    while(tg_y)
    {
        uint32_t bit = lsb32(tg_y);
        tg_y &= tg_y - 1u;
        uint32_t prim = tg_x + bit, rank, mat;
        exact::Hit h = intersect_prim(s, prim, t.o, t.d, t.inv, &rank, &mat);
        (void)mat;
        if(COUNT != 0) cnt->shape_tests++;
        // ray.cpp:653,670,686,708: t >= 1e-6 && t < best; exact ties -> lowest rank
        if(h.t >= ORT_HIT_T_THRESHOLD && (h.t < t.best_t || (h.t == t.best_t && rank < t.best_rank)))
        {
            t.best_t = h.t; t.best_prim = prim; t.best_rank = rank;
        }
    }
    return !trav_next(t, st);
}

// Closest hit.  COUNT adds work counters (the counters build of the same code,
// SURVEY.md 8d).  o/d as in raycast_top_most_node; d need not be unit.
template <int COUNT>
ORT_HD void trace(const SceneView &s, f3 o, f3 d, TraceHit *hit, TraceCounters *cnt)
{
    Trav t; LocalStack st;
    trav_init(s, t, st, o, d);
    while(!trav_step<COUNT>(s, t, st, cnt)) { }
    hit->t = t.best_t; hit->prim = t.best_prim; hit->rank = t.best_rank;
}

// Re-evaluates the winning record to obtain the material and the normalised
// normal (ray.cpp:817).  Same intersector, same operands => same bits as the
// value the reference keeps from its own (single) evaluation.
ORT_HD void finish_hit(const SceneView &s, const TraceHit &hit, f3 o, f3 d, uint32_t *mat, f3 *normal)
{
    if(hit.prim == 0xFFFFFFFFu)
    {
        *mat = 0u;
        *normal = mk3(0.0f, 0.0f, 0.0f);
        return;
    }
    uint32_t rank;
    exact::Hit h = intersect_prim(s, hit.prim, o, d, mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z), &rank, mat);
    *normal = normalize(h.n);
}

} // namespace ort
