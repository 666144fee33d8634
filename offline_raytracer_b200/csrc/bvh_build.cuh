// bvh_build.cuh -- CUDA execution of the data-parallel BVH builder (bvh_build.h).
// Every kernel is a map over an array calling the per-element function of bvh_build.h; ordering
// (Morton sort) and placement (cluster compaction, child / record bases of a level) are CUB radix
// sort and prefix sums, so the tree is byte-identical to the host execution of the same functions
// (scene_flatten.cpp, build_wide_bvh_parallel_host -- checked by tests/test_gpu_build.py).
#pragma once

#include <cub/cub.cuh>

#include "bvh_build.h"
#include "scene_flatten.h"

namespace ort {
namespace build {

struct D3 { double v[3]; };

__global__ void k_morton(uint32_t n, const float *__restrict__ boxes, D3 lo, D3 scale, uint64_t *codes, uint32_t *idx)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    codes[i] = morton63(boxes + 6 * (size_t)i, boxes + 6 * (size_t)i + 3, lo.v, scale.v);
    idx[i] = i;
}

__global__ void k_init_leaves(uint32_t n, const uint32_t *__restrict__ order, const float *__restrict__ boxes,
                              B2 *nodes, uint32_t *sizes, Dp *cost, uint32_t *cluster)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    B2 l; uint32_t src = order[i];
    for(int k = 0; k < 3; ++k) { l.lo[k] = boxes[6 * (size_t)src + k]; l.hi[k] = boxes[6 * (size_t)src + 3 + k]; }
    l.left = B2_LEAF; l.right = src;
    nodes[i] = l; sizes[i] = 1u | B2_LEAF_FLAG; dp_leaf(&cost[i], half_area(l)); cluster[i] = i;
}

__global__ void k_ploc_nearest(uint32_t n, const uint32_t *__restrict__ cluster, const B2 *__restrict__ nodes, uint32_t radius, uint32_t *nn)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) nn[i] = ploc_nearest(i, n, cluster, nodes, radius);
}

__global__ void k_ploc_fate(uint32_t n, const uint32_t *__restrict__ nn, uint32_t *fate, uint32_t *keep, uint32_t *merge)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    uint32_t f = ploc_fate(i, nn);
    fate[i] = f; keep[i] = f != 0u ? 1u : 0u; merge[i] = f == 2u ? 1u : 0u;
}

__global__ void k_ploc_apply(uint32_t n, const uint32_t *__restrict__ nn, const uint32_t *__restrict__ cluster,
                             const uint32_t *__restrict__ fate, const uint32_t *__restrict__ pos, const uint32_t *__restrict__ mid,
                             uint32_t next_node, B2 *nodes, uint32_t *sizes, Dp *cost, uint32_t *new_cluster,
                             uint32_t max_leaf, float traversal_cost)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) ploc_apply(i, nn, cluster, fate[i], pos[i], mid[i], next_node, nodes, sizes, cost, new_cluster, max_leaf, traversal_cost);
}

// the clusters still alive: their nodes, sizes and costs, packed for the host's top-tree build
// the four numbers the host needs after a round (the last element of two exclusive scans and of their inputs),
// gathered into one 16-byte record so that a round costs ONE device-to-host copy instead of four blocking ones
__global__ void k_tail4(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, const uint32_t *__restrict__ c,
                        const uint32_t *__restrict__ d, uint32_t last, uint32_t *__restrict__ out)
{
    if(threadIdx.x == 0) { out[0] = a[last]; out[1] = b[last]; out[2] = c[last]; out[3] = d[last]; }
}

__global__ void k_pack_clusters(uint32_t count, const uint32_t *__restrict__ cluster, const B2 *__restrict__ nodes,
                                const uint32_t *__restrict__ sizes, const Dp *__restrict__ cost, B2 *out_nodes, uint32_t *out_sizes, Dp *out_cost)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= count) return;
    uint32_t id = cluster[i];
    out_nodes[i] = nodes[id]; out_sizes[i] = sizes[id]; out_cost[i] = cost[id];
}

__global__ void k_gather(uint32_t n_items, const Item *__restrict__ items, const B2 *__restrict__ nodes, const uint32_t *__restrict__ sizes,
                         const Dp *__restrict__ cost, uint32_t max_leaf, Kids *kids, uint32_t *n_inner, uint32_t *n_prims)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_items) return;
    Kids k;
    gather_kids(items[i].b2, nodes, sizes, cost, max_leaf, &k);
    kids[i] = k; n_inner[i] = k.n_inner; n_prims[i] = k.n_prims;
}

__global__ void k_emit(uint32_t n_items, const Item *__restrict__ items, const Kids *__restrict__ kids,
                       const uint32_t *__restrict__ inner_off, const uint32_t *__restrict__ prim_off,
                       uint32_t node_count, uint32_t prim_count,
                       const B2 *__restrict__ nodes, const uint32_t *__restrict__ sizes, uint32_t max_leaf,
                       const PrimRec *__restrict__ recs_in, WideNode *wide_out, PrimRec *prims_out, Item *next_items)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_items) return;
    emit_wide(kids[i], items[i].wide, node_count + inner_off[i], prim_count + prim_off[i], inner_off[i],
              nodes, sizes, max_leaf, recs_in, wide_out, prims_out, next_items);
}

__global__ void k_rank_to_prim(uint32_t n, const PrimRec *__restrict__ prims, uint32_t n_ranks, uint32_t *rank_to_prim)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n && prims[i].rank < n_ranks) rank_to_prim[prims[i].rank] = i;
}

struct DeviceBuildResult
{
    WideNode *d_nodes;      // [0, node_count): sphere tree (host-built) then the tree built here
    PrimRec *d_prims;
    uint32_t *d_rank_to_prim;
    uint32_t node_count, prim_count, depth;
    float build_ms;         // device time of the build proper (sort + PLOC + collapse)
    uint32_t ploc_iterations;
};

#define BUILD_TRY(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { *err = std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #expr; ok = false; goto done; } } while(0)

// `flat` holds the sphere tree (nodes / prims) and main_root; `in` the remaining primitives -- on the
// host (in.boxes / in.recs), or already on the device (d_boxes_in / d_recs_in, n_in of them; not freed here).
inline bool build_on_device(const FlatScene &flat, const ParallelBuildInput &in, uint32_t radius, DeviceBuildResult *res, std::string *err,
                            float *d_boxes_in = 0, PrimRec *d_recs_in = 0, uint32_t n_in = 0)
{
    bool ok = true;
    const uint32_t n = d_boxes_in ? n_in : (uint32_t)in.recs.size();
    const uint32_t node_offset = (uint32_t)flat.nodes.size(), prim_offset = (uint32_t)flat.prims.size();
    const uint32_t T = 256;
    auto G = [&](uint32_t m) { return (m + T - 1) / T; };
    memset(res, 0, sizeof(*res));

    float *d_boxes = 0; PrimRec *d_recs = 0;
    uint64_t *d_codes = 0, *d_codes2 = 0; uint32_t *d_idx = 0, *d_order = 0;
    B2 *d_b2 = 0; uint32_t *d_sizes = 0; Dp *d_cost = 0;
    uint32_t *d_cluster = 0, *d_cluster2 = 0, *d_nn = 0, *d_fate = 0, *d_keep = 0, *d_merge = 0, *d_pos = 0, *d_mid = 0;
    WideNode *d_wide = 0; Item *d_items = 0, *d_items2 = 0; Kids *d_kids = 0;
    void *d_temp = 0; size_t temp_bytes = 0;
    uint32_t *d_tail = 0;
    cudaEvent_t e0 = 0, e1 = 0;
    uint32_t wide_cap = 0, node_count = 0, prim_count = 0, depth = 0, iterations = 0;
    D3 lo, sc;
    for(int k = 0; k < 3; ++k) { lo.v[k] = in.scene_lo[k]; sc.v[k] = in.scene_scale[k]; }

    BUILD_TRY(cudaMalloc((void **)&res->d_prims, (size_t)(prim_offset + n + 1) * sizeof(PrimRec)));
    if(prim_offset) BUILD_TRY(cudaMemcpy(res->d_prims, flat.prims.data(), (size_t)prim_offset * sizeof(PrimRec), cudaMemcpyHostToDevice));
    BUILD_TRY(cudaMalloc((void **)&res->d_rank_to_prim, (size_t)(flat.info.record_count + 1) * sizeof(uint32_t)));
    BUILD_TRY(cudaMemset(res->d_rank_to_prim, 0xFF, (size_t)(flat.info.record_count + 1) * sizeof(uint32_t)));
    BUILD_TRY(cudaEventCreate(&e0)); BUILD_TRY(cudaEventCreate(&e1));

    if(n == 0)
    {
        WideNode e; memset(&e, 0, sizeof(e));
        e.ex = e.ey = e.ez = 127;
        for(int s = 0; s < 8; ++s) { e.qlo_x[s] = e.qlo_y[s] = e.qlo_z[s] = 255; }
        BUILD_TRY(cudaMalloc((void **)&res->d_nodes, (size_t)(node_offset + 1) * sizeof(WideNode)));
        if(node_offset) BUILD_TRY(cudaMemcpy(res->d_nodes, flat.nodes.data(), (size_t)node_offset * sizeof(WideNode), cudaMemcpyHostToDevice));
        BUILD_TRY(cudaMemcpy(res->d_nodes + node_offset, &e, sizeof(e), cudaMemcpyHostToDevice));
        node_count = node_offset + 1; prim_count = prim_offset; depth = 1;
    }
    else
    {
        wide_cap = n + 2u;           // every wide node but the root is one of the n - 1 binary inner nodes
        if(d_boxes_in) { d_boxes = d_boxes_in; d_recs = d_recs_in; }
        else
        {
            BUILD_TRY(cudaMalloc((void **)&d_boxes, (size_t)n * 6 * sizeof(float)));
            BUILD_TRY(cudaMalloc((void **)&d_recs, (size_t)n * sizeof(PrimRec)));
        }
        BUILD_TRY(cudaMalloc((void **)&d_codes, (size_t)n * 8)); BUILD_TRY(cudaMalloc((void **)&d_codes2, (size_t)n * 8));
        BUILD_TRY(cudaMalloc((void **)&d_idx, (size_t)n * 4)); BUILD_TRY(cudaMalloc((void **)&d_order, (size_t)n * 4));
        BUILD_TRY(cudaMalloc((void **)&d_b2, (size_t)n * 2 * sizeof(B2)));
        BUILD_TRY(cudaMalloc((void **)&d_sizes, (size_t)n * 2 * 4)); BUILD_TRY(cudaMalloc((void **)&d_cost, (size_t)n * 2 * sizeof(Dp)));
        BUILD_TRY(cudaMalloc((void **)&d_cluster, (size_t)n * 4)); BUILD_TRY(cudaMalloc((void **)&d_cluster2, (size_t)n * 4));
        BUILD_TRY(cudaMalloc((void **)&d_nn, (size_t)n * 4)); BUILD_TRY(cudaMalloc((void **)&d_fate, (size_t)n * 4));
        BUILD_TRY(cudaMalloc((void **)&d_keep, (size_t)n * 4)); BUILD_TRY(cudaMalloc((void **)&d_merge, (size_t)n * 4));
        BUILD_TRY(cudaMalloc((void **)&d_pos, (size_t)n * 4)); BUILD_TRY(cudaMalloc((void **)&d_mid, (size_t)n * 4));
        BUILD_TRY(cudaMalloc((void **)&d_wide, (size_t)(wide_cap + 1) * sizeof(WideNode)));
        BUILD_TRY(cudaMalloc((void **)&d_items, (size_t)wide_cap * sizeof(Item))); BUILD_TRY(cudaMalloc((void **)&d_items2, (size_t)wide_cap * sizeof(Item)));
        BUILD_TRY(cudaMalloc((void **)&d_kids, (size_t)wide_cap * sizeof(Kids)));
        {
            size_t a = 0, b = 0;
            cub::DeviceRadixSort::SortPairs((void *)0, a, d_codes, d_codes2, d_idx, d_order, (int)n, 0, 63);
            cub::DeviceScan::ExclusiveSum((void *)0, b, d_keep, d_pos, (int)n);
            temp_bytes = (a > b ? a : b) + 256;
            BUILD_TRY(cudaMalloc(&d_temp, temp_bytes));
            BUILD_TRY(cudaMalloc((void **)&d_tail, 16));
        }
        if(!d_boxes_in)
        {
            BUILD_TRY(cudaMemcpy(d_boxes, in.boxes.data(), (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice));
            BUILD_TRY(cudaMemcpy(d_recs, in.recs.data(), (size_t)n * sizeof(PrimRec), cudaMemcpyHostToDevice));
        }

        BUILD_TRY(cudaEventRecord(e0, 0));
        // 1. Morton order
        k_morton<<<G(n), T>>>(n, d_boxes, lo, sc, d_codes, d_idx);
        { size_t tb = temp_bytes; BUILD_TRY(cub::DeviceRadixSort::SortPairs(d_temp, tb, d_codes, d_codes2, d_idx, d_order, (int)n, 0, 63)); }
        // 2. PLOC
        k_init_leaves<<<G(n), T>>>(n, d_order, d_boxes, d_b2, d_sizes, d_cost, d_cluster);
        {
            uint32_t count = n, next_node = n;
            const uint32_t top_k = parallel_top_clusters(n);
            while(count > std::max(1u, top_k))
            {
                k_ploc_nearest<<<G(count), T>>>(count, d_cluster, d_b2, radius, d_nn);
                k_ploc_fate<<<G(count), T>>>(count, d_nn, d_fate, d_keep, d_merge);
                size_t tb = temp_bytes;
                BUILD_TRY(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_keep, d_pos, (int)count));
                tb = temp_bytes;
                BUILD_TRY(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_merge, d_mid, (int)count));
                k_ploc_apply<<<G(count), T>>>(count, d_nn, d_cluster, d_fate, d_pos, d_mid, next_node, d_b2, d_sizes, d_cost, d_cluster2,
                                               in.max_leaf, in.node_cost);
                uint32_t last[4];
                k_tail4<<<1, 32>>>(d_pos, d_keep, d_mid, d_merge, count - 1, d_tail);
                BUILD_TRY(cudaMemcpy(last, d_tail, 16, cudaMemcpyDeviceToHost));
                uint32_t kept = last[0] + last[1], merged = last[2] + last[3];
                if(merged == 0 || kept >= count) { *err = "internal: PLOC made no progress"; ok = false; goto done; }
                next_node += merged; count = kept;
                std::swap(d_cluster, d_cluster2);
                ++iterations;
            }
            if(count > 1)
            {
                // the top of the tree: binned SAH over the remaining clusters, on the host (scene_flatten.cpp)
                std::vector<B2> hn(2 * (size_t)count); std::vector<uint32_t> hs(2 * (size_t)count, 0u), hc(count); std::vector<Dp> hcost(2 * (size_t)count);
                B2 *d_pn = (B2 *)d_kids;                                   // scratch: the collapse has not started yet
                uint32_t *d_ps = d_keep; Dp *d_pc = (Dp *)d_items;          // scratch (count <= 65536 entries)
                k_pack_clusters<<<G(count), T>>>(count, d_cluster, d_b2, d_sizes, d_cost, d_pn, d_ps, d_pc);
                BUILD_TRY(cudaMemcpy(hn.data(), d_pn, (size_t)count * sizeof(B2), cudaMemcpyDeviceToHost));
                BUILD_TRY(cudaMemcpy(hs.data(), d_ps, (size_t)count * 4, cudaMemcpyDeviceToHost));
                BUILD_TRY(cudaMemcpy(hcost.data(), d_pc, (size_t)count * sizeof(Dp), cudaMemcpyDeviceToHost));
                BUILD_TRY(cudaMemcpy(hc.data(), d_cluster, (size_t)count * 4, cudaMemcpyDeviceToHost));
                std::vector<uint32_t> local(count);
                for(uint32_t c = 0; c < count; ++c) local[c] = c;
                uint32_t local_next = count, local_root = 0;
                BuildOptions o; o.traversal_cost = in.traversal_cost; o.max_leaf = in.max_leaf;
                if(build_top_tree(local, o, in.max_leaf, in.node_cost, hn.data(), hs.data(), hcost.data(), &local_next, &local_root, err) != ORT_OK)
                { ok = false; goto done; }
                // local ids -> global ids: clusters keep theirs, new nodes follow next_node
                auto global_id = [&](uint32_t l) { return l < count ? hc[l] : next_node + (l - count); };
                const uint32_t added = local_next - count;
                for(uint32_t j = count; j < local_next; ++j) { hn[j].left = global_id(hn[j].left); hn[j].right = global_id(hn[j].right); }
                BUILD_TRY(cudaMemcpy(d_b2 + next_node, hn.data() + count, (size_t)added * sizeof(B2), cudaMemcpyHostToDevice));
                BUILD_TRY(cudaMemcpy(d_sizes + next_node, hs.data() + count, (size_t)added * 4, cudaMemcpyHostToDevice));
                BUILD_TRY(cudaMemcpy(d_cost + next_node, hcost.data() + count, (size_t)added * sizeof(Dp), cudaMemcpyHostToDevice));
                uint32_t root_global = global_id(local_root);
                BUILD_TRY(cudaMemcpy(d_cluster, &root_global, 4, cudaMemcpyHostToDevice));
                next_node += added; count = 1;
            }
        }
        // 3. collapse to 8-wide, level by level
        {
            Item root;
            BUILD_TRY(cudaMemcpy(&root.b2, d_cluster, 4, cudaMemcpyDeviceToHost));
            root.wide = node_offset;
            BUILD_TRY(cudaMemcpy(d_items, &root, sizeof(root), cudaMemcpyHostToDevice));
            uint32_t n_items = 1;
            node_count = node_offset + 1; prim_count = prim_offset;
            uint32_t *d_ninner = d_keep, *d_nprims = d_merge, *d_ioff = d_pos, *d_poff = d_mid;     // reuse (n >= items)
            WideNode *wide_abs = d_wide - node_offset;          // emit_wide indexes nodes absolutely
            while(n_items > 0)
            {
                ++depth;
                if(n_items > wide_cap || node_count - node_offset > wide_cap) { *err = "internal: wide node bound exceeded"; ok = false; goto done; }
                k_gather<<<G(n_items), T>>>(n_items, d_items, d_b2, d_sizes, d_cost, in.max_leaf, d_kids, d_ninner, d_nprims);
                size_t tb = temp_bytes;
                BUILD_TRY(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_ninner, d_ioff, (int)n_items));
                tb = temp_bytes;
                BUILD_TRY(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_nprims, d_poff, (int)n_items));
                uint32_t last[4];
                k_tail4<<<1, 32>>>(d_ioff, d_ninner, d_poff, d_nprims, n_items - 1, d_tail);
                BUILD_TRY(cudaMemcpy(last, d_tail, 16, cudaMemcpyDeviceToHost));
                uint32_t total_inner = last[0] + last[1], total_prims = last[2] + last[3];
                if(node_count - node_offset + total_inner > wide_cap + 1u || total_inner > wide_cap) { *err = "internal: wide node bound exceeded"; ok = false; goto done; }
                k_emit<<<G(n_items), T>>>(n_items, d_items, d_kids, d_ioff, d_poff, node_count, prim_count, d_b2, d_sizes, in.max_leaf,
                                          d_recs, wide_abs, res->d_prims, d_items2);
                node_count += total_inner; prim_count += total_prims;
                n_items = total_inner;
                std::swap(d_items, d_items2);
            }
        }
        BUILD_TRY(cudaEventRecord(e1, 0));
        BUILD_TRY(cudaEventSynchronize(e1));
        BUILD_TRY(cudaGetLastError());
        BUILD_TRY(cudaEventElapsedTime(&res->build_ms, e0, e1));
        if(prim_count != prim_offset + n) { *err = "internal: parallel build lost records"; ok = false; goto done; }
        // final node array: sphere tree, then the nodes built here
        BUILD_TRY(cudaMalloc((void **)&res->d_nodes, (size_t)node_count * sizeof(WideNode)));
        if(node_offset) BUILD_TRY(cudaMemcpy(res->d_nodes, flat.nodes.data(), (size_t)node_offset * sizeof(WideNode), cudaMemcpyHostToDevice));
        BUILD_TRY(cudaMemcpy(res->d_nodes + node_offset, d_wide, (size_t)(node_count - node_offset) * sizeof(WideNode), cudaMemcpyDeviceToDevice));
    }
    if(prim_count) k_rank_to_prim<<<G(prim_count), T>>>(prim_count, res->d_prims, flat.info.record_count, res->d_rank_to_prim);
    BUILD_TRY(cudaDeviceSynchronize());
    res->node_count = node_count; res->prim_count = prim_count; res->depth = depth; res->ploc_iterations = iterations;
done:
    if(!d_boxes_in) { cudaFree(d_boxes); cudaFree(d_recs); }
    cudaFree(d_codes); cudaFree(d_codes2); cudaFree(d_idx); cudaFree(d_order);
    cudaFree(d_b2); cudaFree(d_sizes); cudaFree(d_cost); cudaFree(d_cluster); cudaFree(d_cluster2); cudaFree(d_nn); cudaFree(d_fate);
    cudaFree(d_keep); cudaFree(d_merge); cudaFree(d_pos); cudaFree(d_mid); cudaFree(d_wide); cudaFree(d_items); cudaFree(d_items2);
    cudaFree(d_kids); cudaFree(d_temp); cudaFree(d_tail);
    if(e0) cudaEventDestroy(e0);
    if(e1) cudaEventDestroy(e1);
    if(!ok)
    {
        cudaFree(res->d_nodes); cudaFree(res->d_prims); cudaFree(res->d_rank_to_prim);
        memset(res, 0, sizeof(*res));
    }
    return ok;
}

// =====================================================================================================
// Records and ranks straight from the shape lists, on the device (include/ort_b200.h, OrtShapeLists):
// the CUDA execution of scene_flatten.cpp's collect_records_from_lists + prepare_parallel_build for the
// triangles -- the millions of records; the handful of analytic shapes stay on the host.
// =====================================================================================================
struct MeshTable          // per mesh, on the device
{
    const unsigned long long *tri_first;   // [mesh_count + 1] prefix of triangle counts
    const unsigned long long *vert_base;   // first vertex of the mesh in the concatenated vertex array
    const uint32_t *mat;
    uint32_t mesh_count;
};

__device__ __forceinline__ void load_triangle(const MeshTable &mt, const float *__restrict__ verts, const uint32_t *__restrict__ idx,
                                              uint32_t i, f3 *a, f3 *b, f3 *c, uint32_t *mat)
{
    uint32_t lo = 0, hi = mt.mesh_count;              // last mesh whose first triangle is <= i
    while(hi - lo > 1u) { uint32_t mid = (lo + hi) >> 1; if(mt.tri_first[mid] <= (unsigned long long)i) lo = mid; else hi = mid; }
    const unsigned long long vb = mt.vert_base[lo];
    const uint32_t *t = idx + 3ull * i;
    const float *p0 = verts + 3ull * (vb + t[0]), *p1 = verts + 3ull * (vb + t[1]), *p2 = verts + 3ull * (vb + t[2]);
    *a = mk3(p0[0], p0[1], p0[2]); *b = mk3(p1[0], p1[1], p1[2]); *c = mk3(p2[0], p2[1], p2[2]);
    *mat = mt.mat[lo];
}

__global__ void k_tri_codes(uint32_t n_tri, MeshTable mt, const float *__restrict__ verts, const uint32_t *__restrict__ idx,
                            f3 root_center, f3 root_half, uint32_t *keys, uint32_t *who, uint32_t *absmax_bits)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if(i < n_tri)
    {
        f3 a, b, c; uint32_t mat;
        load_triangle(mt, verts, idx, i, &a, &b, &c, &mat);
        keys[i] = octant_path(triangle_centre(a, b, c), root_center, root_half);
        who[i] = i;
        // fmaxf drops NaNs: map them to +Inf first, so that the host sees a non-finite maximum and rejects the mesh
        auto mag = [](float v) { return v == v ? fabsf(v) : __int_as_float(0x7F800000); };
        m = fmaxf(fmaxf(fmaxf(mag(a.x), mag(a.y)), mag(a.z)), fmaxf(fmaxf(mag(b.x), mag(b.y)), mag(b.z)));
        m = fmaxf(m, fmaxf(fmaxf(mag(c.x), mag(c.y)), mag(c.z)));
    }
    for(int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xFFFFFFFFu, m, o));
    if((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(absmax_bits, __float_as_uint(m));      // non-negative floats order like their bits
}

__global__ void k_iota(uint32_t n, uint32_t first, uint32_t *who)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) who[i] = first + i;
}

// counters[0] = inner octree nodes, [1] = nodes holding records, [2] = max depth
__global__ void k_octree_depth(uint32_t n, const uint32_t *__restrict__ keys_sorted, const uint32_t *__restrict__ who_sorted,
                               uint8_t *depth_of, unsigned long long *counters)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int inner = 0, d = 0;
    if(p < n)
    {
        uint32_t me = keys_sorted[p];
        int shared = 0;
        if(p > 0) { int s2 = common_levels(me, keys_sorted[p - 1]); if(s2 > shared) shared = s2; }
        if(p + 1 < n) { int s2 = common_levels(me, keys_sorted[p + 1]); if(s2 > shared) shared = s2; }
        d = n == 1u ? 0u : (unsigned)(shared + 1 < 10 ? shared + 1 : 10);
        depth_of[who_sorted[p]] = (uint8_t)d;
        // every prefix (levels 0..9) shared by two or more records is an inner node: count it at the first record of its run
        if(n >= 2u)
            for(int level = 0; level < 10; ++level)
            {
                int shift = 3 * (10 - level);
                bool first = p == 0 || (level != 0 && (me >> shift) != (keys_sorted[p - 1] >> shift));
                bool more = p + 1 < n && (level == 0 || (me >> shift) == (keys_sorted[p + 1] >> shift));
                if(first && more) ++inner;
            }
    }
    inner = __reduce_add_sync(0xFFFFFFFFu, inner);
    d = __reduce_max_sync(0xFFFFFFFFu, d);
    if((threadIdx.x & 31) == 0)
    {
        if(inner) atomicAdd(&counters[0], (unsigned long long)inner);
        atomicMax(&counters[2], (unsigned long long)d);
    }
}

__global__ void k_rank_keys(uint32_t n, const uint32_t *__restrict__ keys, const uint8_t *__restrict__ depth_of, unsigned long long *key2, uint32_t *who2)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    uint32_t d = depth_of[i];
    key2[i] = ((unsigned long long)d << 30) | (d == 0u ? 0u : (keys[i] >> (3u * (10u - d))));
    who2[i] = i;
}

__global__ void k_rank_scatter(uint32_t n, const unsigned long long *__restrict__ key2_sorted, const uint32_t *__restrict__ order,
                               uint32_t *rank_of, unsigned long long *counters)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int holding = 0;
    if(r < n)
    {
        rank_of[order[r]] = r;
        holding = (r == 0 || key2_sorted[r] != key2_sorted[r - 1]) ? 1u : 0u;
    }
    holding = __reduce_add_sync(0xFFFFFFFFu, holding);
    if((threadIdx.x & 31) == 0 && holding) atomicAdd(&counters[1], (unsigned long long)holding);
}

__device__ __forceinline__ unsigned long long order_bits(double v)        // doubles -> integers with the same order
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
inline double order_bits_inverse(unsigned long long b)
{
    b = (b >> 63) ? (b & 0x7FFFFFFFFFFFFFFFull) : ~b;
    double v; memcpy(&v, &b, 8); return v;
}

// triangle records + padded boxes, and the bounds of the box centres: bounds[0..2] = min, [3..5] = max (order_bits)
__global__ void k_tri_records(uint32_t n_tri, MeshTable mt, const float *__restrict__ verts, const uint32_t *__restrict__ idx,
                              const uint32_t *__restrict__ rank_of, double pad_rel, double pad_abs,
                              float *boxes, PrimRec *recs, unsigned long long *bounds)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    double c[3] = { 0, 0, 0 };
    bool valid = i < n_tri;
    if(valid)
    {
        f3 a, b, cc; uint32_t mat;
        load_triangle(mt, verts, idx, i, &a, &b, &cc, &mat);
        PrimRec r;
        r.ax = a.x; r.ay = a.y; r.az = a.z; r.rank = rank_of[i];
        r.bx = b.x; r.by = b.y; r.bz = b.z; r.mat = mat;
        r.cx = cc.x; r.cy = cc.y; r.cz = cc.z; r.kind = PRIM_TRIANGLE;
        recs[i] = r;
        float lo[3] = { fminf(fminf(a.x, b.x), cc.x), fminf(fminf(a.y, b.y), cc.y), fminf(fminf(a.z, b.z), cc.z) };
        float hi[3] = { fmaxf(fmaxf(a.x, b.x), cc.x), fmaxf(fmaxf(a.y, b.y), cc.y), fmaxf(fmaxf(a.z, b.z), cc.z) };
        pad_box(lo, hi, pad_rel, pad_abs);
        for(int k = 0; k < 3; ++k) { boxes[6ull * i + k] = lo[k]; boxes[6ull * i + 3 + k] = hi[k]; c[k] = 0.5 * ((double)lo[k] + (double)hi[k]); }
    }
    for(int k = 0; k < 3; ++k)
    {
        unsigned long long mn = valid ? order_bits(c[k]) : ~0ull, mx = valid ? order_bits(c[k]) : 0ull;
        for(int o = 16; o > 0; o >>= 1)
        {
            unsigned long long a2 = __shfl_down_sync(0xFFFFFFFFu, mn, o), b2 = __shfl_down_sync(0xFFFFFFFFu, mx, o);
            mn = a2 < mn ? a2 : mn; mx = b2 > mx ? b2 : mx;
        }
        if((threadIdx.x & 31) == 0) { atomicMin(&bounds[k], mn); atomicMax(&bounds[3 + k], mx); }
    }
}

struct DeviceRecords
{
    float *d_boxes; PrimRec *d_recs;      // tris [0, n_tri), then the boxes / cylinders the host appended
    uint32_t n_tri, n_rest;
    float records_ms;
};

#define BUILD_TRY2(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { *err = std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #expr; ok = false; goto done; } } while(0)

// Fills `flat` (info, world tables, sphere tree, cylinder table), `in` (centre bounds, options) and the
// device arrays of the non-sphere records.  Host work is O(meshes + analytic shapes).
inline bool records_from_lists_on_device(const OrtWorld *world, const OrtShapeLists *L, const BuildOptions &opt,
                                         FlatScene *flat, ParallelBuildInput *in, DeviceRecords *out, std::string *err)
{
    bool ok = true;
    memset(out, 0, sizeof(*out));
    if(opt.max_leaf < 1 || opt.max_leaf > 3) { *err = "max_leaf must be 1..3"; return false; }
    const uint32_t T = 256;
    auto G = [&](uint64_t m) { return (unsigned)((m + T - 1) / T); };
    // ---- host: tables over the meshes ----
    std::vector<unsigned long long> tri_first(L->mesh_count + 1, 0), vert_base(L->mesh_count + 1, 0);
    std::vector<uint32_t> mesh_mat(L->mesh_count ? L->mesh_count : 1, 0);
    for(uint32_t m = 0; m < L->mesh_count; ++m)
    {
        const OrtMesh &mesh = L->meshes[m];
        if(mesh.index_count && (!mesh.vertices || !mesh.indices)) { *err = "mesh without vertices / indices"; return false; }
        if(mesh.mat_index >= world->mat_count) { *err = "record with material index out of range"; return false; }
        for(uint32_t k = 0; k + 2 < mesh.index_count; k += 3)
            if(mesh.indices[k] >= mesh.vertex_count || mesh.indices[k + 1] >= mesh.vertex_count || mesh.indices[k + 2] >= mesh.vertex_count)
            { *err = "mesh index out of range"; return false; }
        tri_first[m + 1] = tri_first[m] + mesh.index_count / 3u;
        vert_base[m + 1] = vert_base[m] + mesh.vertex_count;
        mesh_mat[m] = mesh.mat_index;
    }
    const uint64_t n_tri64 = tri_first[L->mesh_count];
    const uint64_t first_cyl = n_tri64, first_box = first_cyl + L->cylinder_count, first_sph = first_box + L->box_count;
    const uint64_t first_csg = first_sph + L->sphere_count, n64 = first_csg + (L->csg ? 1u : 0u);
    if(n64 >= 0x7FFFFFFFull) { *err = "too many records"; return false; }
    const uint32_t n_tri = (uint32_t)n_tri64, n = (uint32_t)n64, n_analytic = n - n_tri;
    const f3 root_min = mk3(L->root_min.x, L->root_min.y, L->root_min.z), root_max = mk3(L->root_max.x, L->root_max.y, L->root_max.z);
    const f3 root_center = 0.5f * (root_min + root_max);            // macos_main.mm:470-471
    const f3 root_half = root_max - root_center;

    // analytic records: insertion index, octant path (host)
    std::vector<HostPrim> analytic;          // cylinders, boxes, spheres in insertion order (CSG: rank only)
    std::vector<uint32_t> analytic_keys(n_analytic ? n_analytic : 1, 0);
    if(collect_analytic(world, L, root_center, root_half, &analytic, analytic_keys.data(), err) != ORT_OK) return false;

    float *d_verts = 0; uint32_t *d_idx = 0;
    unsigned long long *d_tri_first = 0, *d_vert_base = 0, *d_key2 = 0, *d_key2s = 0, *d_counters = 0;
    uint32_t *d_mat = 0, *d_keys = 0, *d_keys_s = 0, *d_who = 0, *d_who_s = 0, *d_order = 0, *d_rank_of = 0, *d_absmax = 0;
    uint8_t *d_depth = 0;
    void *d_temp = 0; size_t temp_bytes = 0;
    cudaEvent_t e0 = 0, e1 = 0;
    MeshTable mt;
    unsigned long long counters[9];
    std::vector<uint32_t> analytic_rank(n_analytic ? n_analytic : 1, 0);
    uint32_t absmax_bits = 0;
    double scene_abs = 0.0;
    std::vector<HostPrim> spheres, shapes, none;
    uint32_t n_rest = 0;

    BUILD_TRY2(cudaEventCreate(&e0)); BUILD_TRY2(cudaEventCreate(&e1));
    BUILD_TRY2(cudaMalloc((void **)&d_verts, (size_t)(vert_base[L->mesh_count] + 1) * 3 * sizeof(float)));
    BUILD_TRY2(cudaMalloc((void **)&d_idx, (size_t)(n_tri64 + 1) * 3 * sizeof(uint32_t)));
    for(uint32_t m = 0; m < L->mesh_count; ++m)
    {
        const OrtMesh &mesh = L->meshes[m];
        if(mesh.vertex_count) BUILD_TRY2(cudaMemcpy(d_verts + 3 * vert_base[m], mesh.vertices, (size_t)mesh.vertex_count * 3 * sizeof(float), cudaMemcpyHostToDevice));
        uint64_t cnt = mesh.index_count / 3u;
        if(cnt) BUILD_TRY2(cudaMemcpy(d_idx + 3 * tri_first[m], mesh.indices, (size_t)cnt * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    BUILD_TRY2(cudaMalloc((void **)&d_tri_first, tri_first.size() * 8)); BUILD_TRY2(cudaMalloc((void **)&d_vert_base, vert_base.size() * 8));
    BUILD_TRY2(cudaMalloc((void **)&d_mat, mesh_mat.size() * 4));
    BUILD_TRY2(cudaMemcpy(d_tri_first, tri_first.data(), tri_first.size() * 8, cudaMemcpyHostToDevice));
    BUILD_TRY2(cudaMemcpy(d_vert_base, vert_base.data(), vert_base.size() * 8, cudaMemcpyHostToDevice));
    BUILD_TRY2(cudaMemcpy(d_mat, mesh_mat.data(), mesh_mat.size() * 4, cudaMemcpyHostToDevice));
    mt.tri_first = d_tri_first; mt.vert_base = d_vert_base; mt.mat = d_mat; mt.mesh_count = L->mesh_count ? L->mesh_count : 1;
    BUILD_TRY2(cudaMalloc((void **)&d_keys, (size_t)(n + 1) * 4)); BUILD_TRY2(cudaMalloc((void **)&d_keys_s, (size_t)(n + 1) * 4));
    BUILD_TRY2(cudaMalloc((void **)&d_who, (size_t)(n + 1) * 4)); BUILD_TRY2(cudaMalloc((void **)&d_who_s, (size_t)(n + 1) * 4));
    BUILD_TRY2(cudaMalloc((void **)&d_order, (size_t)(n + 1) * 4)); BUILD_TRY2(cudaMalloc((void **)&d_rank_of, (size_t)(n + 1) * 4));
    BUILD_TRY2(cudaMalloc((void **)&d_key2, (size_t)(n + 1) * 8)); BUILD_TRY2(cudaMalloc((void **)&d_key2s, (size_t)(n + 1) * 8));
    BUILD_TRY2(cudaMalloc((void **)&d_depth, (size_t)(n + 1)));
    BUILD_TRY2(cudaMalloc((void **)&d_counters, 9 * 8)); BUILD_TRY2(cudaMalloc((void **)&d_absmax, 4));
    {
        size_t a = 0, b = 0;
        cub::DeviceRadixSort::SortPairs((void *)0, a, d_keys, d_keys_s, d_who, d_who_s, (int)n, 0, 30);
        cub::DeviceRadixSort::SortPairs((void *)0, b, d_key2, d_key2s, d_who, d_order, (int)n, 0, 34);
        temp_bytes = (a > b ? a : b) + 256;
        BUILD_TRY2(cudaMalloc(&d_temp, temp_bytes));
    }
    for(int k = 0; k < 9; ++k) counters[k] = 0;
    counters[3] = counters[4] = counters[5] = ~0ull;          // bounds: min
    BUILD_TRY2(cudaMemcpy(d_counters, counters, sizeof(counters), cudaMemcpyHostToDevice));
    BUILD_TRY2(cudaMemset(d_absmax, 0, 4));

    BUILD_TRY2(cudaEventRecord(e0, 0));
    // ---- ranks: sort by octant path, depth from the neighbours, sort by (depth, prefix) ----
    if(n_tri) k_tri_codes<<<G(n_tri), T>>>(n_tri, mt, d_verts, d_idx, root_center, root_half, d_keys, d_who, d_absmax);
    if(n_analytic)
    {
        BUILD_TRY2(cudaMemcpy(d_keys + n_tri, analytic_keys.data(), (size_t)n_analytic * 4, cudaMemcpyHostToDevice));
        k_iota<<<G(n_analytic), T>>>(n_analytic, n_tri, d_who + n_tri);
    }
    if(n)
    {
        size_t tb = temp_bytes;
        BUILD_TRY2(cub::DeviceRadixSort::SortPairs(d_temp, tb, d_keys, d_keys_s, d_who, d_who_s, (int)n, 0, 30));
        k_octree_depth<<<G(n), T>>>(n, d_keys_s, d_who_s, d_depth, d_counters);
        k_rank_keys<<<G(n), T>>>(n, d_keys, d_depth, d_key2, d_who);
        tb = temp_bytes;
        BUILD_TRY2(cub::DeviceRadixSort::SortPairs(d_temp, tb, d_key2, d_key2s, d_who, d_order, (int)n, 0, 34));
        k_rank_scatter<<<G(n), T>>>(n, d_key2s, d_order, d_rank_of, d_counters);
    }
    BUILD_TRY2(cudaMemcpy(counters, d_counters, 3 * 8, cudaMemcpyDeviceToHost));
    BUILD_TRY2(cudaMemcpy(&absmax_bits, d_absmax, 4, cudaMemcpyDeviceToHost));
    if(n_analytic) BUILD_TRY2(cudaMemcpy(analytic_rank.data(), d_rank_of + n_tri, (size_t)n_analytic * 4, cudaMemcpyDeviceToHost));

    // ---- host: the analytic shapes get their ranks, padding, records; the spheres their tree ----
    memset(&flat->info, 0, sizeof(flat->info));
    flat->nodes.clear(); flat->prims.clear(); flat->cylinders.clear();
    flat->info.triangle_count = n_tri;
    flat->info.cylinder_count = L->cylinder_count; flat->info.box_count = L->box_count; flat->info.sphere_count = L->sphere_count;
    flat->info.csg_count = L->csg ? 1u : 0u;
    flat->info.record_count = n;
    flat->info.octree_node_count = (uint32_t)(counters[0] + counters[1]);
    flat->info.octree_max_depth = (uint32_t)counters[2];
    flat->info.root_min[0] = L->root_min.x; flat->info.root_min[1] = L->root_min.y; flat->info.root_min[2] = L->root_min.z;
    flat->info.root_max[0] = L->root_max.x; flat->info.root_max[1] = L->root_max.y; flat->info.root_max[2] = L->root_max.z;
    if(fill_world_tables_public(world, flat, err) != ORT_OK) { ok = false; goto done; }
    // |v| of every mesh vertex was folded with an unsigned max of its bits: Inf and NaN (exponent all ones) end up on top
    if(absmax_bits >= 0x7F800000u) { *err = "mesh with non-finite vertices (NaN / Inf)"; ok = false; goto done; }
    { float m; memcpy(&m, &absmax_bits, 4); scene_abs = (double)m; }
    for(size_t k = 0; k < analytic.size(); ++k)
    {
        analytic[k].rank = analytic_rank[analytic[k].rank];          // .rank held the index among the analytic records
        for(int c = 0; c < 3; ++c)
        {
            scene_abs = std::max(scene_abs, fabs((double)analytic[k].lo[c]));
            scene_abs = std::max(scene_abs, fabs((double)analytic[k].hi[c]));
        }
    }
    {
        BuildOptions o = opt; o.merge_shapes = true;
        if(pad_and_split_public(analytic, o, scene_abs, &spheres, &shapes, &none, err) != ORT_OK) { ok = false; goto done; }
        in->sphere_depth = 0;
        if(!spheres.empty() && emit_tree_public(spheres, o, flat, &in->sphere_depth, err) != ORT_OK) { ok = false; goto done; }
        flat->main_root = flat->tri_root = (uint32_t)flat->nodes.size();
        in->max_leaf = o.max_leaf; in->traversal_cost = o.traversal_cost; in->node_cost = o.wide_node_cost;
    }
    n_rest = n_tri + (uint32_t)none.size();          // merge_shapes: boxes and cylinders come back in `none` (the "tris" list)
    BUILD_TRY2(cudaMalloc((void **)&out->d_boxes, (size_t)(n_rest + 1) * 6 * sizeof(float)));
    BUILD_TRY2(cudaMalloc((void **)&out->d_recs, (size_t)(n_rest + 1) * sizeof(PrimRec)));
    if(n_tri) k_tri_records<<<G(n_tri), T>>>(n_tri, mt, d_verts, d_idx, d_rank_of, (double)opt.pad_rel, (double)opt.pad_scene * scene_abs,
                                              out->d_boxes, out->d_recs, d_counters + 3);
    BUILD_TRY2(cudaMemcpy(counters + 3, d_counters + 3, 6 * 8, cudaMemcpyDeviceToHost));
    {
        double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
        if(n_tri) for(int k = 0; k < 3; ++k) { lo[k] = order_bits_inverse(counters[3 + k]); hi[k] = order_bits_inverse(counters[6 + k]); }
        std::vector<float> hb(6 * none.size() + 1);
        std::vector<PrimRec> hr(none.size() + 1);
        for(size_t k = 0; k < none.size(); ++k)
        {
            for(int c = 0; c < 3; ++c)
            {
                hb[6 * k + c] = none[k].lo[c]; hb[6 * k + 3 + c] = none[k].hi[c];
                double cc = 0.5 * ((double)none[k].lo[c] + (double)none[k].hi[c]);
                lo[c] = std::min(lo[c], cc); hi[c] = std::max(hi[c], cc);
            }
            hr[k] = make_record_public(none[k], flat);
        }
        if(!none.empty())
        {
            BUILD_TRY2(cudaMemcpy(out->d_boxes + 6ull * n_tri, hb.data(), none.size() * 6 * sizeof(float), cudaMemcpyHostToDevice));
            BUILD_TRY2(cudaMemcpy(out->d_recs + n_tri, hr.data(), none.size() * sizeof(PrimRec), cudaMemcpyHostToDevice));
        }
        for(int k = 0; k < 3; ++k)
        {
            in->scene_lo[k] = n_rest ? lo[k] : 0.0;
            double ext = n_rest ? hi[k] - lo[k] : 0.0;
            in->scene_scale[k] = ext > 0.0 ? 2097152.0 / ext : 0.0;
        }
    }
    BUILD_TRY2(cudaEventRecord(e1, 0));
    BUILD_TRY2(cudaEventSynchronize(e1));
    BUILD_TRY2(cudaGetLastError());
    BUILD_TRY2(cudaEventElapsedTime(&out->records_ms, e0, e1));
    out->n_tri = n_tri; out->n_rest = n_rest;
done:
    cudaFree(d_verts); cudaFree(d_idx); cudaFree(d_tri_first); cudaFree(d_vert_base); cudaFree(d_mat);
    cudaFree(d_keys); cudaFree(d_keys_s); cudaFree(d_who); cudaFree(d_who_s); cudaFree(d_order); cudaFree(d_rank_of);
    cudaFree(d_key2); cudaFree(d_key2s); cudaFree(d_depth); cudaFree(d_counters); cudaFree(d_absmax); cudaFree(d_temp);
    if(e0) cudaEventDestroy(e0);
    if(e1) cudaEventDestroy(e1);
    if(!ok) { cudaFree(out->d_boxes); cudaFree(out->d_recs); memset(out, 0, sizeof(*out)); }
    return ok;
}
#undef BUILD_TRY2

#undef BUILD_TRY

} // namespace build
} // namespace ort
