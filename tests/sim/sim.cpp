// tests/sim/sim.cpp -- TEST INFRASTRUCTURE.  Host build (g++, -ffp-contract=off)
// of the product's shared host/device headers (csrc/bvh.h, csrc/path.h) so that
// the flattener and the traversal / shading logic can be checked against the
// oracle on a box without a GPU.  Never linked into libort_b200.so and not
// reachable from the product API.
#include <thread>
#include <vector>
#include <atomic>
#include <string>
#include <string.h>

#include "scene_flatten.h"
#include "path.h"

using namespace ort;

struct SimScene { FlatScene flat; SceneView view; };

extern "C" {

void *sim_scene_create(const OrtWorld *world, const OrtBVHOctreeNode *root, float traversal_cost, float pad_rel, float pad_scene)
{
    SimScene *s = new SimScene();
    BuildOptions opt;
    if(traversal_cost > 0) opt.traversal_cost = traversal_cost;
    if(pad_rel >= 0) opt.pad_rel = pad_rel;
    if(pad_scene >= 0) opt.pad_scene = pad_scene;
    if(const char *e = getenv("ORT_BVH_OPTIMAL")) opt.optimal_collapse = atoi(e) != 0;
    if(const char *e = getenv("ORT_BVH_NODE_COST")) opt.wide_node_cost = (float)atof(e);
    std::string err;
    if(flatten_scene(world, root, opt, &s->flat, &err) != ORT_OK)
    {
        fprintf(stderr, "sim_scene_create: %s\n", err.c_str());
        delete s; return 0;
    }
    s->view.nodes = (const q4 *)s->flat.nodes.data();
    s->view.prims = (const q4 *)s->flat.prims.data();
    s->view.cyl = (const q4 *)s->flat.cylinders.data();
    s->view.node_count = (uint32_t)s->flat.nodes.size();
    s->view.prim_count = (uint32_t)s->flat.prims.size();
    s->view.main_root = s->flat.main_root;
    s->view.tri_root = s->flat.tri_root;
    return s;
}
// the same scene through the data-parallel builder (bvh_build.h), executed on the host
void *sim_scene_create_parallel(const OrtWorld *world, const OrtBVHOctreeNode *root, uint32_t radius)
{
    SimScene *s = new SimScene();
    BuildOptions opt;
    std::string err;
    std::vector<HostPrim> prims;
    int rc = collect_records(world, root, &prims, &s->flat, &err);
    if(rc == ORT_OK) rc = build_wide_bvh_parallel_host(prims, opt, radius ? radius : ORT_PLOC_RADIUS, &s->flat, &err);
    if(rc != ORT_OK)
    {
        fprintf(stderr, "sim_scene_create_parallel: %s\n", err.c_str());
        delete s; return 0;
    }
    s->view.nodes = (const q4 *)s->flat.nodes.data();
    s->view.prims = (const q4 *)s->flat.prims.data();
    s->view.cyl = (const q4 *)s->flat.cylinders.data();
    s->view.node_count = (uint32_t)s->flat.nodes.size();
    s->view.prim_count = (uint32_t)s->flat.prims.size();
    s->view.main_root = s->flat.main_root;
    s->view.tri_root = s->flat.tri_root;
    return s;
}
// records + ranks WITHOUT the octree (collect_records_from_lists), then the usual host SAH build
void *sim_scene_create_from_lists(const OrtWorld *world, const OrtShapeLists *lists)
{
    SimScene *s = new SimScene();
    BuildOptions opt;
    std::string err;
    std::vector<HostPrim> prims;
    int rc = collect_records_from_lists(world, lists, &prims, &s->flat, &err);
    if(rc == ORT_OK) rc = build_wide_bvh(prims, opt, &s->flat, &err);
    if(rc != ORT_OK)
    {
        fprintf(stderr, "sim_scene_create_from_lists: %s\n", err.c_str());
        delete s; return 0;
    }
    s->view.nodes = (const q4 *)s->flat.nodes.data();
    s->view.prims = (const q4 *)s->flat.prims.data();
    s->view.cyl = (const q4 *)s->flat.cylinders.data();
    s->view.node_count = (uint32_t)s->flat.nodes.size();
    s->view.prim_count = (uint32_t)s->flat.prims.size();
    s->view.main_root = s->flat.main_root;
    s->view.tri_root = s->flat.tri_root;
    return s;
}
// raw bytes of the flattened tree, for byte-for-byte comparison with the CUDA execution
uint64_t sim_scene_nodes(void *h, const void **p) { SimScene *s = (SimScene *)h; *p = s->flat.nodes.data(); return s->flat.nodes.size() * sizeof(WideNode); }
uint64_t sim_scene_prims(void *h, const void **p) { SimScene *s = (SimScene *)h; *p = s->flat.prims.data(); return s->flat.prims.size() * sizeof(PrimRec); }
void sim_scene_destroy(void *h) { delete (SimScene *)h; }
void sim_scene_info(void *h, OrtSceneInfo *info, uint32_t *wide_depth) { *info = ((SimScene *)h)->flat.info; *wide_depth = ((SimScene *)h)->flat.wide_depth; }

void sim_raycast_batch(void *h, uint64_t n, const float *origins, const float *dirs,
                       float *hit_t, uint32_t *prim_rank, uint32_t *mat_index, float *hit_normal,
                       uint64_t *counters, int n_threads)
{
    SimScene *s = (SimScene *)h;
    if(n_threads < 1) n_threads = 1;
    std::atomic<unsigned long long> c0(0), c1(0), c2(0);
    auto work = [&](int tid)
    {
        TraceCounters cnt = { 0, 0, 0 };
        uint64_t a = 0, b = 0, c = 0;
        for(uint64_t i = (uint64_t)tid; i < n; i += (uint64_t)n_threads)
        {
            f3 o = mk3(origins[3*i], origins[3*i+1], origins[3*i+2]);
            f3 d = mk3(dirs[3*i], dirs[3*i+1], dirs[3*i+2]);
            TraceHit hit; cnt.node_visits = cnt.box_tests = cnt.shape_tests = 0;
            trace<true>(s->view, o, d, &hit, &cnt);
            uint32_t mat; f3 nrm;
            finish_hit(s->view, hit, o, d, &mat, &nrm);
            if(hit_t) hit_t[i] = hit.t;
            if(prim_rank) prim_rank[i] = hit.rank;
            if(mat_index) mat_index[i] = mat;
            if(hit_normal) { hit_normal[3*i] = nrm.x; hit_normal[3*i+1] = nrm.y; hit_normal[3*i+2] = nrm.z; }
            a += cnt.node_visits; b += cnt.box_tests; c += cnt.shape_tests;
        }
        c0 += a; c1 += b; c2 += c;
    };
    std::vector<std::thread> threads;
    for(int i = 1; i < n_threads; ++i) threads.emplace_back(work, i);
    work(0);
    for(auto &t : threads) t.join();
    if(counters) { counters[0] += c0.load(); counters[1] += c1.load(); counters[2] += c2.load(); }
}


// same semantics as ort_render / oracle_render, on the host, from the product's path.h
void sim_render(void *h, const OrtCamera *cam, const OrtRenderParams *P, ort_v3 *out, uint64_t *counters, int n_threads)
{
    SimScene *s = (SimScene *)h;
    PathConsts c;
    c.cam_p = mk3(cam->p.x, cam->p.y, cam->p.z);
    c.cam_x = mk3(cam->x_axis.x, cam->x_axis.y, cam->x_axis.z);
    c.cam_y = mk3(cam->y_axis.x, cam->y_axis.y, cam->y_axis.z);
    c.cam_z = mk3(cam->z_axis.x, cam->z_axis.y, cam->z_axis.z);
    c.focal_length = length(c.cam_p - mk3(P->focus_target[0], P->focus_target[1], P->focus_target[2]));
    c.aperture_radius = P->aperture_radius; c.lens_z_offset = P->lens_z_offset;
    c.roughness = P->roughness; c.eps = P->dont_get_too_close_epsilon; c.rr = P->russian_roulette_value;
    c.width = P->output_width; c.height = P->output_height;
    c.light_count = (uint32_t)s->flat.light_is_sphere.size();
    c.light_is_sphere = s->flat.light_is_sphere.data();
    c.materials = (const q4 *)s->flat.materials.data();
    uint32_t spp = P->ray_per_pixel_count;
    uint32_t chunk_spp = P->chunk_spp ? P->chunk_spp : spp;
    if(chunk_spp > spp) chunk_spp = spp;
    uint32_t n_chunks = spp ? (spp + chunk_spp - 1) / chunk_spp : 0;
    uint32_t c_begin = P->chunk_begin, c_end = P->chunk_end;
    if(c_begin == 0 && c_end == 0) c_end = n_chunks;
    if(c_end > n_chunks) c_end = n_chunks;
    if(n_threads < 1) n_threads = 1;
    std::atomic<int> next_row(P->tile_min_y);
    std::atomic<unsigned long long> rays(0);
    auto work = [&]()
    {
        uint64_t nrays = 0;
        for(;;)
        {
            int y = next_row.fetch_add(1);
            if(y >= P->tile_one_past_max_y) break;
            for(int x = P->tile_min_x; x < P->tile_one_past_max_x; ++x)
            {
                uint32_t pixel_index = (uint32_t)(y * P->output_width + x);
                f3 focal_point = pixel_focal_point(c, x, y);
                long long acc[3] = { 0, 0, 0 };
                f3 color = mk3(0, 0, 0);
                for(uint32_t ch = (n_chunks == 1 ? 0 : c_begin); ch < (n_chunks == 1 ? 1 : c_end); ++ch)
                {
                    uint32_t n = chunk_spp;
                    if((ch + 1) * chunk_spp > spp) n = spp - ch * chunk_spp;
                    Path p; p.series = ort_stream_seed(P->base_seed, pixel_index, ch);
                    color = mk3(0, 0, 0);
                    for(uint32_t si = 0; si < n; ++si)
                    {
                        generate_primary(c, focal_point, &p);
                        TraceHit hit; uint32_t mat; f3 nrm;
                        trace<false>(s->view, p.origin, p.dir, &hit, 0); nrays++;
                        finish_hit(s->view, hit, p.origin, p.dir, &mat, &nrm);
                        bool alive = shade_primary(c, &p, hit.t, mat, nrm, &color);
                        while(alive && next_bounce(c, &p))
                        {
                            trace<false>(s->view, p.origin, p.dir, &hit, 0); nrays++;
                            finish_hit(s->view, hit, p.origin, p.dir, &mat, &nrm);
                            alive = shade_bounce(c, &p, hit.t, mat, nrm, &color);
                        }
                    }
                    acc[0] += to_fixed(color.x); acc[1] += to_fixed(color.y); acc[2] += to_fixed(color.z);
                }
                ort_v3 *pixel = out + (size_t)y * P->output_width + x;
                if(n_chunks == 1)
                {
                    f3 r = color / (float)spp;
                    pixel->x = r.x; pixel->y = r.y; pixel->z = r.z;
                }
                else
                {
                    const double inv = 1.0 / (double)(1 << ORT_ACCUM_FRAC_BITS);
                    pixel->x = (float)((double)acc[0] * inv) / (float)spp;
                    pixel->y = (float)((double)acc[1] * inv) / (float)spp;
                    pixel->z = (float)((double)acc[2] * inv) / (float)spp;
                }
            }
        }
        rays += nrays;
    };
    std::vector<std::thread> threads;
    for(int i = 1; i < n_threads; ++i) threads.emplace_back(work);
    work();
    for(auto &t : threads) t.join();
    if(counters) counters[0] += rays.load();
}

} // extern "C"
