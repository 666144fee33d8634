// bvh_build.h -- data-parallel construction of the 8-wide BVH ("BVH build / flatten
// on device").  The reference builds its octree by single-threaded pointer-chasing inserts
// (push_shape_inside_node + validate_nodes_and_reallocate_shapes, code/ray.cpp:1799-2045; 0.46 s for
// 139 k triangles => minutes for the 50 M of BASELINE config 5).  Here every step is a map over an
// array, written ONCE as a host/device function of the element index:
//
//   1. Morton code (63 bits) of each padded primitive box centre; radix sort.
//   2. PLOC (parallel locally-ordered clustering, Meister & Bittner 2018): every cluster looks
//      `radius` entries left and right in the Morton order for the neighbour whose union with it
//      has the smallest surface area; mutual choices merge into a binary node; a prefix sum
//      compacts the cluster list; repeat until few clusters are left (at most 65536, at most n / 64).
//      Agglomerative, so the quality of the bottom of the tree is that of a SAH build rather than of
//      a spatial-median LBVH; its weak spot is the top -- it only ever looks at Morton neighbours --
//      so the last clusters are joined top-down by the host's binned-SAH builder (a few thousand
//      boxes, milliseconds): node visits per ray on the 4.4 M-triangle grid 9.76 -> 9.13 (SAH 8.73).
//   3. Collapse to 8-wide, level by level, following the optimal-collapse cost tables that were filled
//      in as the binary nodes were created (which subtrees become leaves of <= max_leaf primitives,
//      which share a wide node).  Children are
//      assigned to octant slots, their boxes quantised outward to the node's 8-bit grid
//      (bvh.h), and the level's child nodes / primitive records are placed by prefix sums, so the
//      result is deterministic: the CUDA kernels (bvh_build.cuh) and the host loops over the same
//      functions (tests/sim) produce the same bytes.
//
// The tree only decides WHICH records a ray tests; the closest hit -- (t, rank) -- is whatever the
// exact intersectors say (bvh.h), so a scene built here returns bit-identical hits to one built by
// the host SAH builder (scene_flatten.cpp).
#pragma once

#include "bvh.h"

namespace ort {
namespace build {

// binary node: leaves are [0, n) in Morton order, inner nodes follow
struct B2
{
    float lo[3], hi[3];
    uint32_t left, right;       // inner: children; leaf: left = B2_LEAF, right = index of the input primitive
};
#define B2_LEAF 0xFFFFFFFFu

struct Item { uint32_t b2, wide; };      // binary subtree to turn into the wide node `wide`

struct Kids { uint32_t id[8]; uint32_t nk; uint32_t n_inner, n_prims; };

// ---- 0. records and ranks straight from the shape lists (include/ort_b200.h, OrtShapeLists) ---------
// the ten octant digits of a shape centre, first level in the top 3 bits of 30
// (get_bvh_octree_node_child_info, ray.cpp:1476-1522: bit0 = +x, bit1 = +y, bit2 = +z; the child's
// centre is the parent's -+ half of the parent's half dimension, in float)
ORT_HD uint32_t octant_path(f3 c, f3 node_center, f3 node_half)
{
    uint32_t code = 0;
    for(int level = 0; level < 10; ++level)
    {
        f3 h = 0.5f * node_half;
        uint32_t idx = 0;
        if(c.x >= node_center.x) { idx |= 1u; node_center.x += h.x; } else node_center.x -= h.x;
        if(c.y >= node_center.y) { idx |= 2u; node_center.y += h.y; } else node_center.y -= h.y;
        if(c.z >= node_center.z) { idx |= 4u; node_center.z += h.z; } else node_center.z -= h.z;
        node_half = h;
        code = (code << 3) | idx;
    }
    return code;
}
ORT_HD int common_levels(uint32_t a, uint32_t b)      // leading 3-bit digits two 30-bit paths share
{
    uint32_t x = a ^ b;
    if(x == 0) return 10;
    int top = (int)msb32(x);                           // highest differing bit, 0..29
    return (29 - top) / 3;
}
// centre of a triangle's box as get_shape_aabb forms it (ray.cpp:1675-1746)
ORT_HD f3 triangle_centre(f3 a, f3 b, f3 c)
{
    f3 mn = mk3(ref_min(ref_min(a.x, b.x), c.x), ref_min(ref_min(a.y, b.y), c.y), ref_min(ref_min(a.z, b.z), c.z));
    f3 mx = mk3(ref_max(ref_max(a.x, b.x), c.x), ref_max(ref_max(a.y, b.y), c.y), ref_max(ref_max(a.z, b.z), c.z));
    return 0.5f * (mn + mx);
}
// conservative padding of a primitive box (bvh.h): relative to its own diagonal plus an absolute term
ORT_HD void pad_box(float *lo, float *hi, double pad_rel, double pad_abs)
{
    double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
    double pad = pad_rel * sqrt(dx * dx + dy * dy + dz * dz) + pad_abs;
    for(int k = 0; k < 3; ++k)
    {
        lo[k] = nextafterf((float)((double)lo[k] - pad), -INFINITY);
        hi[k] = nextafterf((float)((double)hi[k] + pad), INFINITY);
    }
}

// ---- 1. Morton codes ---------------------------------------------------------------------------
ORT_HD uint64_t spread21(uint64_t v)
{
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
ORT_HD uint64_t morton63(const float *lo, const float *hi, const double *scene_lo, const double *scene_scale)
{
    uint64_t q[3];
    for(int k = 0; k < 3; ++k)
    {
        double c = 0.5 * ((double)lo[k] + (double)hi[k]);
        double u = (c - scene_lo[k]) * scene_scale[k];           // [0, 2^21)
        if(!(u > 0.0)) u = 0.0;
        if(u > 2097151.0) u = 2097151.0;
        q[k] = (uint64_t)u;
    }
    return spread21(q[0]) | (spread21(q[1]) << 1) | (spread21(q[2]) << 2);
}

// ---- 2. PLOC -------------------------------------------------------------------------------------
ORT_HD double union_half_area(const B2 &a, const B2 &b)
{
    double d[3];
    for(int k = 0; k < 3; ++k)
    {
        float l = a.lo[k] < b.lo[k] ? a.lo[k] : b.lo[k];
        float h = a.hi[k] > b.hi[k] ? a.hi[k] : b.hi[k];
        d[k] = (double)h - (double)l;
    }
    return d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
}
ORT_HD double half_area(const B2 &a) { return union_half_area(a, a); }

// nearest neighbour of cluster i among [i - radius, i + radius]; ties go to the smallest index, which
// guarantees that the globally closest pair always chooses each other (the loop makes progress)
ORT_HD uint32_t ploc_nearest(uint32_t i, uint32_t n, const uint32_t *cluster, const B2 *nodes, uint32_t radius)
{
    const B2 me = nodes[cluster[i]];
    uint32_t first = i > radius ? i - radius : 0u;
    uint32_t last = i + radius < n - 1u ? i + radius : n - 1u;
    double best = 1e300;
    uint32_t bj = i;
    for(uint32_t j = first; j <= last; ++j)
    {
        if(j == i) continue;
        double a = union_half_area(me, nodes[cluster[j]]);
        if(a < best) { best = a; bj = j; }
    }
    return bj;
}

// what becomes of entry i: 0 = absorbed by its partner, 1 = kept, 2 = kept and merged with nn[i]
ORT_HD uint32_t ploc_fate(uint32_t i, const uint32_t *nn)
{
    uint32_t j = nn[i];
    if(j == i || nn[j] != i) return 1u;
    return i < j ? 2u : 0u;
}

// sizes[id] = primitives below node id, bit 31 = "this subtree is a leaf child": it holds <= max_leaf
// primitives and testing them all whenever its box is hit costs no more (surface-area heuristic, cost
// tables below) than giving it a wide node of its own
#define B2_LEAF_FLAG 0x80000000u
ORT_HD uint32_t subtree_size(const uint32_t *sizes, uint32_t id) { return sizes[id] & ~B2_LEAF_FLAG; }

// Optimal collapse to 8-wide (Ylitie, Karras, Laine 2017, section 3), evaluated bottom-up as the binary
// nodes are created: C[i] = least SAH cost of representing the subtree with at most i child slots of a
// wide node,
//   C[1]     = min( A * P  [a leaf of P <= max_leaf primitives],  A * c_node + D(8)  [a wide node of its own] )
//   C[i > 1] = min( D(i), C[i - 1] ),    D(j) = min_{0<k<j} C_left[k] + C_right[j - k],   k[j] = the minimising k
struct Dp { float C[8]; uint8_t k[8]; };           // C[1..7]; k[j - 1] for j = 2..8
ORT_HD void dp_leaf(Dp *d, double area)
{
    for(int i = 0; i < 8; ++i) { d->C[i] = (float)area; d->k[i] = 0; }
}

// binary node `id` = parent of l and r: box, size, the cost table and the leaf decision
ORT_HD void merge_nodes(uint32_t id, uint32_t l, uint32_t r, B2 *nodes, uint32_t *sizes, Dp *cost, uint32_t max_leaf, float node_cost)
{
    B2 a = nodes[l], b = nodes[r], m;
    for(int k = 0; k < 3; ++k)
    {
        m.lo[k] = a.lo[k] < b.lo[k] ? a.lo[k] : b.lo[k];
        m.hi[k] = a.hi[k] > b.hi[k] ? a.hi[k] : b.hi[k];
    }
    m.left = l; m.right = r;
    nodes[id] = m;
    uint32_t count = subtree_size(sizes, l) + subtree_size(sizes, r);
    double area = half_area(m);
    const Dp cl = cost[l], cr = cost[r];
    Dp d;
    double D[9];
    d.k[0] = 0;
    for(int j = 2; j <= 8; ++j)
    {
        double best = 1e300; int bk = 1;
        for(int k = 1; k < j; ++k)
        {
            int kl = k > 7 ? 7 : k, kr = (j - k) > 7 ? 7 : (j - k);
            double v = (double)cl.C[kl] + (double)cr.C[kr];
            if(v < best) { best = v; bk = k; }
        }
        D[j] = best; d.k[j - 1] = (uint8_t)bk;
    }
    double as_leaf = count <= max_leaf ? area * (double)count : 1e300;
    double as_inner = (double)node_cost * area + D[8];
    bool leaf = as_leaf <= as_inner;
    sizes[id] = count | (leaf ? B2_LEAF_FLAG : 0u);
    d.C[0] = 0.f;
    d.C[1] = (float)(leaf ? as_leaf : as_inner);
    for(int i = 2; i <= 7; ++i) { float di = (float)D[i]; d.C[i] = di < d.C[i - 1] ? di : d.C[i - 1]; }
    cost[id] = d;
}

// writes the new cluster entry of i (pos = its index in the compacted list, mid = how many merges
// precede it)
ORT_HD void ploc_apply(uint32_t i, const uint32_t *nn, const uint32_t *cluster, uint32_t fate, uint32_t pos, uint32_t mid,
                       uint32_t next_node, B2 *nodes, uint32_t *sizes, Dp *cost, uint32_t *new_cluster,
                       uint32_t max_leaf, float traversal_cost)
{
    if(fate == 0u) return;
    if(fate == 1u) { new_cluster[pos] = cluster[i]; return; }
    uint32_t l = cluster[i], r = cluster[nn[i]];
    uint32_t id = next_node + mid;
    merge_nodes(id, l, r, nodes, sizes, cost, max_leaf, traversal_cost);
    new_cluster[pos] = id;
}

// ---- 3. collapse to 8-wide -------------------------------------------------------------------------
ORT_HD bool is_leaf_child(uint32_t id, const uint32_t *sizes, uint32_t max_leaf) { (void)max_leaf; return (sizes[id] & B2_LEAF_FLAG) != 0u; }

// children of the wide node made from binary subtree `root`: the subtrees the cost tables chose to share it
ORT_HD void gather_kids(uint32_t root, const B2 *nodes, const uint32_t *sizes, const Dp *cost, uint32_t max_leaf, Kids *out)
{
    uint32_t *kids = out->id;
    uint32_t nk = 0;
    if(is_leaf_child(root, sizes, max_leaf)) kids[nk++] = root;      // degenerate: the whole tree is one leaf
    else
    {
        uint32_t st_node[16]; int st_budget[16]; int sp = 0;
        int k = cost[root].k[7];
        st_node[sp] = nodes[root].right; st_budget[sp++] = 8 - k;
        st_node[sp] = nodes[root].left; st_budget[sp++] = k;
        while(sp > 0)
        {
            --sp;
            uint32_t c = st_node[sp]; int i = st_budget[sp] > 7 ? 7 : st_budget[sp];
            const Dp dc = cost[c];
            while(i > 1 && dc.C[i] == dc.C[i - 1]) --i;
            if(i == 1 || nodes[c].left == B2_LEAF) { kids[nk++] = c; continue; }
            int kk = dc.k[i - 1];
            st_node[sp] = nodes[c].right; st_budget[sp++] = i - kk;
            st_node[sp] = nodes[c].left; st_budget[sp++] = kk;
        }
    }
    out->nk = nk;
    out->n_inner = 0; out->n_prims = 0;
    for(uint32_t i = 0; i < nk; ++i)
    {
        if(is_leaf_child(kids[i], sizes, max_leaf)) out->n_prims += subtree_size(sizes, kids[i]);
        else out->n_inner++;
    }
}

// the <= 3 input primitives below a leaf child, left to right
ORT_HD uint32_t leaf_prims(uint32_t id, const B2 *nodes, uint32_t *out3)
{
    uint32_t stack[4]; int sp = 0; uint32_t n = 0;
    stack[sp++] = id;
    while(sp > 0)
    {
        uint32_t cur = stack[--sp];
        if(nodes[cur].left == B2_LEAF) { if(n < 3u) out3[n] = nodes[cur].right; ++n; }
        else { stack[sp++] = nodes[cur].right; stack[sp++] = nodes[cur].left; }
    }
    return n;
}

// smallest e with 255 * 2^e >= extent, from the exponent of extent / 255 alone (no log2, whose last
// bit differs between libms: the host and CUDA executions must agree)
ORT_HD int grid_exponent_hd(double extent)
{
    if(!(extent > 0.0)) return -100;
    int ex;
    double m = frexp(extent / 255.0, &ex);      // extent / 255 = m * 2^ex, m in [0.5, 1)
    int e = (m == 0.5) ? ex - 1 : ex;
    if(e < -100) e = -100;
    if(e > 120) e = 120;
    return e;
}

// Fills wide node `self`: octant slots, quantised child boxes, child / primitive bases; copies the
// records of its leaf children to prims_out[prim_base ...] and appends its inner children, in slot
// order, to next_items[item_base ...] as the nodes child_base + 0, 1, ...
//   node_offset / prim_offset: where this tree starts in the scene's node / record arrays
ORT_HD void emit_wide(const Kids &kd, uint32_t self, uint32_t child_base, uint32_t prim_base, uint32_t item_base,
                      const B2 *nodes, const uint32_t *sizes, uint32_t max_leaf,
                      const PrimRec *recs_in, WideNode *wide_out, PrimRec *prims_out, Item *next_items)
{
    const uint32_t nk = kd.nk;
    float nlo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, nhi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for(uint32_t i = 0; i < nk; ++i)
        for(int k = 0; k < 3; ++k)
        {
            const B2 &c = nodes[kd.id[i]];
            if(c.lo[k] < nlo[k]) nlo[k] = c.lo[k];
            if(c.hi[k] > nhi[k]) nhi[k] = c.hi[k];
        }
    // octant slots: greedy assignment maximising (centroid - centre) . octant direction
    int slot_of[8]; bool slot_used[8], kid_done[8];
    for(int s = 0; s < 8; ++s) { slot_used[s] = false; kid_done[s] = false; slot_of[s] = -1; }
    double ctr[3]; for(int k = 0; k < 3; ++k) ctr[k] = 0.5 * ((double)nlo[k] + nhi[k]);
    for(uint32_t round = 0; round < nk; ++round)
    {
        double bestc = -1e300; int bi = -1, bs = -1;
        for(uint32_t i = 0; i < nk; ++i)
        {
            if(kid_done[i]) continue;
            const B2 &cb = nodes[kd.id[i]];
            double dc[3]; for(int k = 0; k < 3; ++k) dc[k] = 0.5 * ((double)cb.lo[k] + cb.hi[k]) - ctr[k];
            for(int s = 0; s < 8; ++s)
            {
                if(slot_used[s]) continue;
                double c = ((s & 1) ? dc[0] : -dc[0]) + ((s & 2) ? dc[1] : -dc[1]) + ((s & 4) ? dc[2] : -dc[2]);
                if(c > bestc) { bestc = c; bi = (int)i; bs = s; }
            }
        }
        slot_of[bi] = bs; slot_used[bs] = true; kid_done[bi] = true;
    }
    int kid_at[8]; for(int s = 0; s < 8; ++s) kid_at[s] = -1;
    for(uint32_t i = 0; i < nk; ++i) kid_at[slot_of[i]] = (int)i;

    WideNode n;
    {
        uint32_t *w = reinterpret_cast<uint32_t *>(&n);
        for(uint32_t k = 0; k < sizeof(WideNode) / 4u; ++k) w[k] = 0u;
    }
    n.px = nlo[0]; n.py = nlo[1]; n.pz = nlo[2];
    int e[3];
    for(int k = 0; k < 3; ++k) e[k] = grid_exponent_hd((double)nhi[k] - nlo[k]);
    uint8_t qlo[3][8], qhi[3][8];
    for(int k = 0; k < 3; ++k)
    {
        for(;;)       // quantise outward; if rounding pushes a plane past 255, coarsen the grid
        {
            double step = ldexp(1.0, e[k]);
            bool ok = true;
            for(int s = 0; s < 8 && ok; ++s)
            {
                if(kid_at[s] < 0) { qlo[k][s] = 255; qhi[k][s] = 0; continue; }
                const B2 &cb = nodes[kd.id[kid_at[s]]];
                double lo = floor(((double)cb.lo[k] - (double)nlo[k]) / step);
                double hi = ceil(((double)cb.hi[k] - (double)nlo[k]) / step);
                if(lo < 0.0) lo = 0.0;
                if(hi > 255.0) { ok = false; break; }
                if(hi < lo) hi = lo;
                qlo[k][s] = (uint8_t)lo; qhi[k][s] = (uint8_t)hi;
            }
            if(ok) break;
            e[k]++;
        }
    }
    n.ex = (uint8_t)(e[0] + 127); n.ey = (uint8_t)(e[1] + 127); n.ez = (uint8_t)(e[2] + 127);
    for(int s = 0; s < 8; ++s)
    {
        n.qlo_x[s] = qlo[0][s]; n.qlo_y[s] = qlo[1][s]; n.qlo_z[s] = qlo[2][s];
        n.qhi_x[s] = qhi[0][s]; n.qhi_y[s] = qhi[1][s]; n.qhi_z[s] = qhi[2][s];
    }
    n.child_base = child_base;
    n.prim_base = prim_base;
    uint32_t prim_off = 0, inner = 0;
    bool only_triangles = true;
    for(int s = 0; s < 8; ++s)
    {
        if(kid_at[s] < 0) continue;
        uint32_t id = kd.id[kid_at[s]];
        if(!is_leaf_child(id, sizes, max_leaf))
        {
            n.imask |= (uint8_t)(1u << s);
            n.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            Item it; it.b2 = id; it.wide = child_base + inner;
            next_items[item_base + inner] = it;
            ++inner;
        }
        else
        {
            uint32_t src[3];
            uint32_t count = leaf_prims(id, nodes, src);
            uint32_t unary = (1u << count) - 1u;
            n.meta[s] = (uint8_t)((unary << 5) | prim_off);
            for(uint32_t i = 0; i < count; ++i)
            {
                PrimRec r = recs_in[src[i]];
                if((r.kind & 0xFFu) != PRIM_TRIANGLE) only_triangles = false;
                prims_out[prim_base + prim_off + i] = r;
            }
            prim_off += count;
        }
    }
    if(!only_triangles) n.prim_base |= ORT_NODE_MIXED_KINDS;
    wide_out[self] = n;
}

} // namespace build
} // namespace ort
