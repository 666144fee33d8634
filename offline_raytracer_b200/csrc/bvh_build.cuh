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
                              B2 *nodes, uint32_t *sizes, float *cost, uint32_t *cluster)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    B2 l; uint32_t src = order[i];
    for(int k = 0; k < 3; ++k) { l.lo[k] = boxes[6 * (size_t)src + k]; l.hi[k] = boxes[6 * (size_t)src + 3 + k]; }
    l.left = B2_LEAF; l.right = src;
    nodes[i] = l; sizes[i] = 1u | B2_LEAF_FLAG; cost[i] = (float)half_area(l); cluster[i] = i;
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
                             uint32_t next_node, B2 *nodes, uint32_t *sizes, float *cost, uint32_t *new_cluster,
                             uint32_t max_leaf, float traversal_cost)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) ploc_apply(i, nn, cluster, fate[i], pos[i], mid[i], next_node, nodes, sizes, cost, new_cluster, max_leaf, traversal_cost);
}

__global__ void k_gather(uint32_t n_items, const Item *__restrict__ items, const B2 *__restrict__ nodes, const uint32_t *__restrict__ sizes,
                         uint32_t max_leaf, Kids *kids, uint32_t *n_inner, uint32_t *n_prims)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_items) return;
    Kids k;
    gather_kids(items[i].b2, nodes, sizes, max_leaf, &k);
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

// `flat` holds the sphere tree (nodes / prims) and main_root; `in` the remaining primitives.
inline bool build_on_device(const FlatScene &flat, const ParallelBuildInput &in, uint32_t radius, DeviceBuildResult *res, std::string *err)
{
    bool ok = true;
    const uint32_t n = (uint32_t)in.recs.size();
    const uint32_t node_offset = (uint32_t)flat.nodes.size(), prim_offset = (uint32_t)flat.prims.size();
    const uint32_t T = 256;
    auto G = [&](uint32_t m) { return (m + T - 1) / T; };
    memset(res, 0, sizeof(*res));

    float *d_boxes = 0; PrimRec *d_recs = 0;
    uint64_t *d_codes = 0, *d_codes2 = 0; uint32_t *d_idx = 0, *d_order = 0;
    B2 *d_b2 = 0; uint32_t *d_sizes = 0; float *d_cost = 0;
    uint32_t *d_cluster = 0, *d_cluster2 = 0, *d_nn = 0, *d_fate = 0, *d_keep = 0, *d_merge = 0, *d_pos = 0, *d_mid = 0;
    WideNode *d_wide = 0; Item *d_items = 0, *d_items2 = 0; Kids *d_kids = 0;
    void *d_temp = 0; size_t temp_bytes = 0;
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
        BUILD_TRY(cudaMalloc((void **)&d_boxes, (size_t)n * 6 * sizeof(float)));
        BUILD_TRY(cudaMalloc((void **)&d_recs, (size_t)n * sizeof(PrimRec)));
        BUILD_TRY(cudaMalloc((void **)&d_codes, (size_t)n * 8)); BUILD_TRY(cudaMalloc((void **)&d_codes2, (size_t)n * 8));
        BUILD_TRY(cudaMalloc((void **)&d_idx, (size_t)n * 4)); BUILD_TRY(cudaMalloc((void **)&d_order, (size_t)n * 4));
        BUILD_TRY(cudaMalloc((void **)&d_b2, (size_t)n * 2 * sizeof(B2)));
        BUILD_TRY(cudaMalloc((void **)&d_sizes, (size_t)n * 2 * 4)); BUILD_TRY(cudaMalloc((void **)&d_cost, (size_t)n * 2 * 4));
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
        }
        BUILD_TRY(cudaMemcpy(d_boxes, in.boxes.data(), (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice));
        BUILD_TRY(cudaMemcpy(d_recs, in.recs.data(), (size_t)n * sizeof(PrimRec), cudaMemcpyHostToDevice));

        BUILD_TRY(cudaEventRecord(e0, 0));
        // 1. Morton order
        k_morton<<<G(n), T>>>(n, d_boxes, lo, sc, d_codes, d_idx);
        { size_t tb = temp_bytes; BUILD_TRY(cub::DeviceRadixSort::SortPairs(d_temp, tb, d_codes, d_codes2, d_idx, d_order, (int)n, 0, 63)); }
        // 2. PLOC
        k_init_leaves<<<G(n), T>>>(n, d_order, d_boxes, d_b2, d_sizes, d_cost, d_cluster);
        {
            uint32_t count = n, next_node = n;
            while(count > 1)
            {
                k_ploc_nearest<<<G(count), T>>>(count, d_cluster, d_b2, radius, d_nn);
                k_ploc_fate<<<G(count), T>>>(count, d_nn, d_fate, d_keep, d_merge);
                size_t tb = temp_bytes;
                BUILD_TRY(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_keep, d_pos, (int)count));
                tb = temp_bytes;
                BUILD_TRY(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_merge, d_mid, (int)count));
                k_ploc_apply<<<G(count), T>>>(count, d_nn, d_cluster, d_fate, d_pos, d_mid, next_node, d_b2, d_sizes, d_cost, d_cluster2,
                                               in.max_leaf, in.traversal_cost);
                uint32_t last[4];
                BUILD_TRY(cudaMemcpy(&last[0], d_pos + (count - 1), 4, cudaMemcpyDeviceToHost));
                BUILD_TRY(cudaMemcpy(&last[1], d_keep + (count - 1), 4, cudaMemcpyDeviceToHost));
                BUILD_TRY(cudaMemcpy(&last[2], d_mid + (count - 1), 4, cudaMemcpyDeviceToHost));
                BUILD_TRY(cudaMemcpy(&last[3], d_merge + (count - 1), 4, cudaMemcpyDeviceToHost));
                uint32_t kept = last[0] + last[1], merged = last[2] + last[3];
                if(merged == 0 || kept >= count) { *err = "internal: PLOC made no progress"; ok = false; goto done; }
                next_node += merged; count = kept;
                std::swap(d_cluster, d_cluster2);
                ++iterations;
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
                k_gather<<<G(n_items), T>>>(n_items, d_items, d_b2, d_sizes, in.max_leaf, d_kids, d_ninner, d_nprims);
                size_t tb = temp_bytes;
                BUILD_TRY(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_ninner, d_ioff, (int)n_items));
                tb = temp_bytes;
                BUILD_TRY(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_nprims, d_poff, (int)n_items));
                uint32_t last[4];
                BUILD_TRY(cudaMemcpy(&last[0], d_ioff + (n_items - 1), 4, cudaMemcpyDeviceToHost));
                BUILD_TRY(cudaMemcpy(&last[1], d_ninner + (n_items - 1), 4, cudaMemcpyDeviceToHost));
                BUILD_TRY(cudaMemcpy(&last[2], d_poff + (n_items - 1), 4, cudaMemcpyDeviceToHost));
                BUILD_TRY(cudaMemcpy(&last[3], d_nprims + (n_items - 1), 4, cudaMemcpyDeviceToHost));
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
    cudaFree(d_boxes); cudaFree(d_recs); cudaFree(d_codes); cudaFree(d_codes2); cudaFree(d_idx); cudaFree(d_order);
    cudaFree(d_b2); cudaFree(d_sizes); cudaFree(d_cost); cudaFree(d_cluster); cudaFree(d_cluster2); cudaFree(d_nn); cudaFree(d_fate);
    cudaFree(d_keep); cudaFree(d_merge); cudaFree(d_pos); cudaFree(d_mid); cudaFree(d_wide); cudaFree(d_items); cudaFree(d_items2);
    cudaFree(d_kids); cudaFree(d_temp);
    if(e0) cudaEventDestroy(e0);
    if(e1) cudaEventDestroy(e1);
    if(!ok)
    {
        cudaFree(res->d_nodes); cudaFree(res->d_prims); cudaFree(res->d_rank_to_prim);
        memset(res, 0, sizeof(*res));
    }
    return ok;
}

#undef BUILD_TRY

} // namespace build
} // namespace ort
