// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Linux driver shim around the UNMODIFIED reference sources.  The reference
// (gyuhyun-lee/offline_raytracer) is a unity build whose only translation unit
// is the macOS-only code/macos_main.mm.  This file plays the role of that
// driver: it #includes the reference headers and ray.cpp / parser.cpp *where
// they lie* under /root/reference/code (nothing is copied into this repo; the
// build recipe is oracle/Makefile, output oracle/_ref/libref.so) and exports a
// small C API so tests can run the reference's own code:
//
//   * scene assembly, a step-for-step restatement of main()
//     (code/macos_main.mm:296-562): parse .scn, load + bake meshes, octree
//     insert in the order triangles -> cylinders -> boxes -> spheres -> CSG,
//     compaction, camera axes;
//   * tiled_raytrace_bvh (code/ray.cpp:1178) on arbitrary tile rects / seeds;
//   * raycast_top_most_node (code/ray.cpp:1165) on explicit ray buffers;
//   * the intersectors, BSDF functions, RNG and loaders one by one, for
//     differential tests of the restated oracle and of the CUDA device code.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the resulting library.
//
// Include order follows code/macos_main.mm:14-22.  platform.h is taken from a
// build-time patched copy (its `extern "C" {` wrapper, code/platform.h:10,342,
// encloses two overloaded static push_size functions, which g++ rejects).

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <thread>
#include <vector>
#include <atomic>

#undef internal
#undef assert

#include "types.h"
#include "math.h"
#include "platform.h"
#include "intrinsic.h"
#include "random.h"

#include "ray.cpp"
#include "parser.cpp"

namespace {

struct RefScene
{
    ParseSceneResult scene;
    World world;
    Camera camera;
    Mesh meshes[100];
    u32 mesh_count;
    CSG csgs[10];
    u32 csg_count;
    BVHOctreeNode *top_most_node;

    void *light_mem;
    void *node_mem;
    void *shape_mem;
    MemoryArena light_memory_arena;
    MemoryArena bvh_node_arena;
    MemoryArena shape_arena;
    ValidateNodesResult validate;
    int width, height;
};

static u8 *
read_whole_file(const char *path, u64 *size_out)
{
    FILE *f = fopen(path, "rb");
    if(!f) return 0;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    u8 *mem = (u8 *)malloc((size_t)sz + 16);
    if(fread(mem, 1, (size_t)sz, f) != (size_t)sz) { free(mem); fclose(f); return 0; }
    memset(mem + sz, 0, 16);
    fclose(f);
    *size_out = (u64)sz;
    return mem;
}

// seed of the (pixel, chunk) sample stream -- must equal ort_stream_seed() in
// include/ort_b200.h (the boundary's per-pixel seed mode).
static inline u32
stream_seed(u32 base, u32 pixel_index, u32 chunk)
{
    u32 h = base ^ (pixel_index * 0x9E3779B1u) ^ (chunk * 0x85EBCA77u);
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    if(h == 0) h = 0x6D2B79F5u;
    return h;
}

} // namespace

extern "C" {

// ---------------------------------------------------------------------------
// scene assembly: code/macos_main.mm:310-562, literally.
// node_arena_mb / shape_arena_mb default to the reference literals 512 / 8
// (macos_main.mm:418, 540) when 0 is passed.
// ---------------------------------------------------------------------------
void *
ref_scene_load(const char *scn_path, const char *base_dir, int width, int height,
               int with_csg, unsigned node_arena_mb, unsigned shape_arena_mb)
{
    if(node_arena_mb == 0) node_arena_mb = 512;
    if(shape_arena_mb == 0) shape_arena_mb = 8;

    RefScene *rs = (RefScene *)calloc(1, sizeof(RefScene));
    rs->width = width;
    rs->height = height;

    rs->light_mem = calloc(1, megabytes(2));
    rs->light_memory_arena = start_memory_arena(rs->light_mem, megabytes(2), false);
    rs->scene.light_push_buffer = start_temp_memory(&rs->light_memory_arena, rs->light_memory_arena.total_size, false);

    u64 scene_size = 0;
    u8 *scene_file = read_whole_file(scn_path, &scene_size);
    if(!scene_file) { free(rs->light_mem); free(rs); return 0; }
    parse_scene(&rs->scene, scene_file, (u32)scene_size, (char *)base_dir);
    ParseSceneResult &scene = rs->scene;
    scene.output_width = width;     // macos_main.mm:319-320 overrides the .scn
    scene.output_height = height;

    // hard-coded CSG, macos_main.mm:322-332 (inert record, ray.cpp:718-767)
    if(with_csg)
    {
        rs->csgs[0].sphere.center = v3_(0, 0, 0.8f);
        rs->csgs[0].sphere.r = 0.35f;
        rs->csgs[0].aab.min = v3_(0, 0, 0.8f) - v3_(0.3f, 0.3f, 0.3f);
        rs->csgs[0].aab.max = v3_(0, 0, 0.8f) + v3_(0.3f, 0.3f, 0.3f);
        rs->csgs[0].mat_index = 5;
        rs->csg_count = 1;
    }

    World &world = rs->world;
    world.ambient = scene.ambient;
    world.materials = scene.materials;
    world.mat_count = scene.mat_count;
    world.light_push_buffer = scene.light_push_buffer;
    world.light_count = scene.light_count;

    Mesh *meshes = rs->meshes;
    u32 mesh_count = 0;
    for(u32 mesh_index = 0; mesh_index < scene.mesh_count; mesh_index++)
    {
        MeshInfo *mesh_info = scene.mesh_infos + mesh_index;
        Mesh *mesh = meshes + mesh_count++;

        char extension_buffer[16] = {};
        get_extension(extension_buffer, mesh_info->file_path);
        u64 fsize = 0;
        u8 *fmem = read_whole_file(mesh_info->file_path, &fsize);
        if(!fmem) { fprintf(stderr, "ref_scene_load: cannot read %s\n", mesh_info->file_path); return 0; }
        if(string_compare(extension_buffer, (char *)"ply"))
        {
            ParsePlyHeaderResult ply_header = parse_ply_header(fmem, (u32)fsize);
            mesh->vertex_count = ply_header.vertex_count;
            mesh->index_count = ply_header.index_count;
            mesh->vertices = (v3 *)malloc(sizeof(v3) * mesh->vertex_count);
            mesh->indices = (u32 *)malloc(sizeof(u32) * mesh->index_count);
            mesh->mat_index = mesh_info->mat_index;
            parse_ply(&ply_header, fmem, (u32)fsize, mesh->vertices, mesh->indices);
        }
        else if(string_compare(extension_buffer, (char *)"obj"))
        {
            PreParseObjResult pre = pre_parse_obj(fmem, fsize);
            mesh->vertex_count = pre.position_count;
            mesh->index_count = pre.index_count;
            mesh->vertices = (v3 *)malloc(sizeof(v3) * mesh->vertex_count);
            mesh->indices = (u32 *)malloc(sizeof(u32) * mesh->index_count);
            mesh->mat_index = mesh_info->mat_index;
            parse_obj(&pre, fmem, (u32)fsize, mesh->vertices, 0, 0, mesh->indices);
        }
        free(fmem);

        v3 min = v3_(Flt_Max, Flt_Max, Flt_Max);
        v3 max = v3_(Flt_Min, Flt_Min, Flt_Min);

        for(u32 vertex_index = 0; vertex_index < mesh->vertex_count; ++vertex_index)
        {
            v3 *v = mesh->vertices + vertex_index;

            *v *= mesh_info->scale;
            *v = quaternion_rotation(mesh_info->quaternion, quaternion_rotation(v3_(0, 1, 0), 0.0174533f * mesh_info->axis.degree, *v));
            *v += mesh_info->translate;

            min.x = minimum(min.x, v->x);
            min.y = minimum(min.y, v->y);
            min.z = minimum(min.z, v->z);

            max.x = maximum(max.x, v->x);
            max.y = maximum(max.y, v->y);
            max.z = maximum(max.z, v->z);

            mesh->aabb_min = min;
            mesh->aabb_max = max;
        }
    }
    rs->mesh_count = mesh_count;

    size_t node_arena_size = (size_t)megabytes(node_arena_mb);
    rs->node_mem = calloc(1, node_arena_size);
    rs->bvh_node_arena = start_memory_arena(rs->node_mem, node_arena_size, false);
    MemoryArena &bvh_node_arena = rs->bvh_node_arena;

    BVHOctreeNode *top_most_node = push_struct(&bvh_node_arena, BVHOctreeNode);
    zero(top_most_node);
    top_most_node->aabb_min = v3_(Flt_Max, Flt_Max, Flt_Max);
    top_most_node->aabb_max = v3_(Flt_Min, Flt_Min, Flt_Min);

    for(u32 i = 0; i < scene.mesh_count; ++i)
        update_aabb_min_max(&top_most_node->aabb_min, &top_most_node->aabb_max, meshes + i, Shape_Type_Mesh);
    for(u32 i = 0; i < scene.cylinder_count; ++i)
        update_aabb_min_max(&top_most_node->aabb_min, &top_most_node->aabb_max, scene.cylinders + i, Shape_Type_Cylinder);
    for(u32 i = 0; i < scene.box_count; ++i)
        update_aabb_min_max(&top_most_node->aabb_min, &top_most_node->aabb_max, scene.boxes + i, Shape_Type_AAB);
    for(u32 i = 0; i < scene.sphere_count; ++i)
        update_aabb_min_max(&top_most_node->aabb_min, &top_most_node->aabb_max, scene.spheres + i, Shape_Type_Sphere);
    for(u32 i = 0; i < rs->csg_count; ++i)
    {
        CSG *csg = rs->csgs + i;
        csg->aabb_min = Flt_Max * v3_(1, 1, 1);
        csg->aabb_max = Flt_Min * v3_(1, 1, 1);
        update_aabb_min_max(&csg->aabb_min, &csg->aabb_max, &csg->sphere, Shape_Type_Sphere);
        update_aabb_min_max(&csg->aabb_min, &csg->aabb_max, &csg->aab, Shape_Type_AAB);
    }

    v3 top_most_node_center = 0.5f * (top_most_node->aabb_min + top_most_node->aabb_max);
    v3 top_most_node_half_dim = top_most_node->aabb_max - top_most_node_center;

    u32 desired_depth = 10;

    for(u32 mesh_index = 0; mesh_index < mesh_count; ++mesh_index)
    {
        Mesh *mesh = meshes + mesh_index;
        for(u32 index_index = 0; index_index < mesh->index_count; index_index += 3)
        {
            Triangle triangle = {};
            triangle.mesh = mesh;
            triangle.i_0 = mesh->indices[index_index];
            triangle.i_1 = mesh->indices[index_index + 1];
            triangle.i_2 = mesh->indices[index_index + 2];
            push_shape_inside_node(&bvh_node_arena, top_most_node, top_most_node_center, top_most_node_half_dim, 0, desired_depth, &triangle, Shape_Type_Triangle);
        }
    }
    for(u32 i = 0; i < scene.cylinder_count; ++i)
        push_shape_inside_node(&bvh_node_arena, top_most_node, top_most_node_center, top_most_node_half_dim, 0, desired_depth, scene.cylinders + i, Shape_Type_Cylinder);
    for(u32 i = 0; i < scene.box_count; ++i)
        push_shape_inside_node(&bvh_node_arena, top_most_node, top_most_node_center, top_most_node_half_dim, 0, desired_depth, scene.boxes + i, Shape_Type_AAB);
    for(u32 i = 0; i < scene.sphere_count; ++i)
        push_shape_inside_node(&bvh_node_arena, top_most_node, top_most_node_center, top_most_node_half_dim, 0, desired_depth, scene.spheres + i, Shape_Type_Sphere);
    for(u32 i = 0; i < rs->csg_count; ++i)
        push_shape_inside_node(&bvh_node_arena, top_most_node, top_most_node_center, top_most_node_half_dim, 0, desired_depth, rs->csgs + i, Shape_Type_CSG);

    size_t shape_arena_size = (size_t)megabytes(shape_arena_mb);
    rs->shape_mem = calloc(1, shape_arena_size);
    rs->shape_arena = start_memory_arena(rs->shape_mem, shape_arena_size, false);
    validate_nodes_and_reallocate_shapes(&rs->shape_arena, top_most_node, &rs->validate);
    rs->top_most_node = top_most_node;

    Camera &camera = rs->camera;
    camera.p = scene.camera_p;
    f32 rx = scene.camera_height_ratio * ((f32)scene.output_width / scene.output_height);
    camera.x_axis = rx * quaternion_rotation(scene.camera_quaternion, v3_(1, 0, 0));
    camera.y_axis = scene.camera_height_ratio * quaternion_rotation(scene.camera_quaternion, v3_(0, 1, 0));
    camera.z_axis = quaternion_rotation(scene.camera_quaternion, v3_(0, 0, 1));

    free(scene_file);
    return rs;
}

void *ref_scene_world(void *h) { return &((RefScene *)h)->world; }
void *ref_scene_camera(void *h) { return &((RefScene *)h)->camera; }
void *ref_scene_root(void *h) { return ((RefScene *)h)->top_most_node; }
void *ref_scene_meshes(void *h) { return ((RefScene *)h)->meshes; }
uint32_t ref_scene_mesh_count(void *h) { return ((RefScene *)h)->mesh_count; }

// counts[0..9] = spheres, boxes, cylinders, materials, meshes, lights,
// light-buffer bytes, node count, node arena bytes, shape arena bytes
void
ref_scene_counts(void *h, uint64_t *counts)
{
    RefScene *rs = (RefScene *)h;
    counts[0] = rs->scene.sphere_count;
    counts[1] = rs->scene.box_count;
    counts[2] = rs->scene.cylinder_count;
    counts[3] = rs->scene.mat_count;
    counts[4] = rs->scene.mesh_count;
    counts[5] = rs->scene.light_count;
    counts[6] = rs->scene.light_push_buffer.used;
    counts[7] = rs->bvh_node_arena.used / sizeof(BVHOctreeNode);
    counts[8] = rs->bvh_node_arena.used;
    counts[9] = rs->shape_arena.used;
}

// struct sizes of the data contract
void
ref_struct_sizes(uint32_t *s)
{
    s[0] = sizeof(Sphere); s[1] = sizeof(AAB); s[2] = sizeof(Cylinder);
    s[3] = sizeof(Material); s[4] = sizeof(Camera); s[5] = sizeof(Mesh);
    s[6] = sizeof(Triangle); s[7] = sizeof(World); s[8] = sizeof(BVHOctreeNode);
    s[9] = sizeof(TempMemory); s[10] = sizeof(BVHShapeHeader); s[11] = sizeof(CSG);
}

// ---------------------------------------------------------------------------
// the hot path, unmodified
// ---------------------------------------------------------------------------
uint64_t
ref_tiled_raytrace_bvh(void *h, float *output_buffer, int output_width, int output_height,
                       int tile_min_x, int tile_min_y, int tile_one_past_max_x, int tile_one_past_max_y,
                       uint32_t *series_state, uint32_t ray_per_pixel_count, float russian_roulette_value)
{
    RefScene *rs = (RefScene *)h;
    RandomSeries series = start_random_series(*series_state);
    u64 r = tiled_raytrace_bvh(&rs->world, &rs->camera, rs->top_most_node, (v3 *)output_buffer,
                               output_width, output_height, tile_min_x, tile_min_y,
                               tile_one_past_max_x, tile_one_past_max_y, &series,
                               ray_per_pixel_count, russian_roulette_value);
    *series_state = series.next_random;
    return r;
}

// Per-pixel seed mode: every pixel is rendered by the UNMODIFIED
// tiled_raytrace_bvh on a 1x1 tile whose RandomSeries starts at
// stream_seed(base_seed, y*W + x, 0).  Rows are distributed over n_threads
// std::threads.  Returns total shape-test count.
uint64_t
ref_render_pixel_seeds(void *h, float *output_buffer, int output_width, int output_height,
                       int min_x, int min_y, int one_past_max_x, int one_past_max_y,
                       uint32_t base_seed, uint32_t ray_per_pixel_count, float russian_roulette_value,
                       int n_threads)
{
    RefScene *rs = (RefScene *)h;
    if(n_threads < 1) n_threads = 1;
    std::atomic<int> next_row(min_y);
    std::atomic<unsigned long long> total(0);
    auto work = [&]()
    {
        u64 local = 0;
        for(;;)
        {
            int y = next_row.fetch_add(1);
            if(y >= one_past_max_y) break;
            for(int x = min_x; x < one_past_max_x; ++x)
            {
                RandomSeries series = start_random_series(stream_seed(base_seed, (u32)(y * output_width + x), 0));
                local += tiled_raytrace_bvh(&rs->world, &rs->camera, rs->top_most_node, (v3 *)output_buffer,
                                            output_width, output_height, x, y, x + 1, y + 1, &series,
                                            ray_per_pixel_count, russian_roulette_value);
            }
        }
        total += local;
    };
    std::vector<std::thread> threads;
    for(int i = 1; i < n_threads; ++i) threads.emplace_back(work);
    work();
    for(auto &t : threads) t.join();
    return total.load();
}

// The reference's own scheduling: 32x32 tile grid (macos_main.mm:602-605),
// each tile with its own RandomSeries seeded from a master series
// (macos_main.mm:642), tiles pulled by n_threads workers.  This is the CPU
// baseline workload.
uint64_t
ref_render_tiles(void *h, float *output_buffer, int output_width, int output_height,
                 uint32_t master_seed, uint32_t ray_per_pixel_count, float russian_roulette_value,
                 int n_threads, int max_tiles)
{
    RefScene *rs = (RefScene *)h;
    struct Tile { int x0, y0, x1, y1; RandomSeries series; };
    std::vector<Tile> tiles;
    RandomSeries series = start_random_series(master_seed);
    i32 tile_x_count = 32, tile_y_count = 32;
    i32 px = ceil_r32_i32(output_width / (f32)tile_x_count);
    i32 py = ceil_r32_i32(output_height / (f32)tile_y_count);
    for(i32 ty = 0; ty < tile_y_count; ++ty)
    {
        i32 y0 = ty * py, y1 = y0 + py;
        if(y1 > output_height) y1 = output_height;
        for(i32 tx = 0; tx < tile_x_count; ++tx)
        {
            i32 x0 = tx * px, x1 = x0 + px;
            if(x1 > output_width) x1 = output_width;
            Tile t = { x0, y0, x1, y1, start_random_series(random_u32(&series)) };
            tiles.push_back(t);
        }
    }
    int tile_count = (int)tiles.size();
    if(max_tiles > 0 && max_tiles < tile_count) tile_count = max_tiles;
    if(n_threads < 1) n_threads = 1;
    std::atomic<int> next_tile(0);
    std::atomic<unsigned long long> total(0);
    auto work = [&]()
    {
        u64 local = 0;
        for(;;)
        {
            int i = next_tile.fetch_add(1);
            if(i >= tile_count) break;
            Tile &t = tiles[i];
            if(t.x0 >= t.x1 || t.y0 >= t.y1) continue;
            local += tiled_raytrace_bvh(&rs->world, &rs->camera, rs->top_most_node, (v3 *)output_buffer,
                                        output_width, output_height, t.x0, t.y0, t.x1, t.y1, &t.series,
                                        ray_per_pixel_count, russian_roulette_value);
        }
        total += local;
    };
    std::vector<std::thread> threads;
    for(int i = 1; i < n_threads; ++i) threads.emplace_back(work);
    work();
    for(auto &t : threads) t.join();
    return total.load();
}

// raycast_top_most_node on explicit ray buffers (origins/dirs are xyz triples)
uint64_t
ref_raycast_batch(void *h, uint64_t n, const float *origins, const float *dirs,
                  float *hit_t, uint32_t *hit_mat_index, float *hit_normal, int32_t *inner_hit,
                  int n_threads)
{
    RefScene *rs = (RefScene *)h;
    if(n_threads < 1) n_threads = 1;
    std::atomic<unsigned long long> total(0);
    auto work = [&](int tid)
    {
        BVHQueue queue = {};
        queue.size = megabytes(4);
        queue.base = (u8 *)malloc(queue.size);
        u64 tests = 0;
        for(u64 i = (u64)tid; i < n; i += (u64)n_threads)
        {
            v3 o = v3_(origins[3*i], origins[3*i+1], origins[3*i+2]);
            v3 d = v3_(dirs[3*i], dirs[3*i+1], dirs[3*i+2]);
            RaycastBVHResult r = raycast_top_most_node(&queue, rs->top_most_node, &tests, o, d);
            hit_t[i] = r.hit_t;
            hit_mat_index[i] = r.hit_mat_index;
            if(hit_normal) { hit_normal[3*i] = r.hit_normal.x; hit_normal[3*i+1] = r.hit_normal.y; hit_normal[3*i+2] = r.hit_normal.z; }
            if(inner_hit) inner_hit[i] = r.inner_hit;
        }
        free(queue.base);
        total += tests;
    };
    std::vector<std::thread> threads;
    for(int i = 1; i < n_threads; ++i) threads.emplace_back(work, i);
    work(0);
    for(auto &t : threads) t.join();
    return total.load();
}

// ---------------------------------------------------------------------------
// single functions, for differential tests
// ---------------------------------------------------------------------------
// out = { hit_t, nx, ny, nz, inner_hit }
static void
pack_isect(IntersectionTestResult r, float *out)
{
    out[0] = r.hit_t; out[1] = r.hit_normal.x; out[2] = r.hit_normal.y; out[3] = r.hit_normal.z;
    out[4] = (float)r.inner_hit;
}

void ref_intersect_triangle(const float *v0, const float *v1, const float *v2, const float *o, const float *d, float *out)
{
    pack_isect(ray_intersect_with_triangle(v3_(v0[0], v0[1], v0[2]), v3_(v1[0], v1[1], v1[2]), v3_(v2[0], v2[1], v2[2]),
                                           v3_(o[0], o[1], o[2]), v3_(d[0], d[1], d[2])), out);
}
void ref_intersect_sphere(const float *c, float r, const float *o, const float *d, float *out)
{
    pack_isect(ray_intersect_with_sphere(v3_(c[0], c[1], c[2]), r, v3_(o[0], o[1], o[2]), v3_(d[0], d[1], d[2])), out);
}
void ref_intersect_aab(const float *mn, const float *mx, const float *o, const float *d, float *out)
{
    pack_isect(ray_intersect_with_aab(v3_(mn[0], mn[1], mn[2]), v3_(mx[0], mx[1], mx[2]), v3_(o[0], o[1], o[2]), v3_(d[0], d[1], d[2])), out);
}
void ref_intersect_cylinder(const float *base, const float *axis, float r, const float *o, const float *d, float *out)
{
    pack_isect(ray_intersect_with_cylinder(v3_(base[0], base[1], base[2]), v3_(axis[0], axis[1], axis[2]), r,
                                           v3_(o[0], o[1], o[2]), v3_(d[0], d[1], d[2])), out);
}
int32_t ref_in_rect(const float *p, const float *mn, const float *mx)
{
    return in_rect(v3_(p[0], p[1], p[2]), v3_(mn[0], mn[1], mn[2]), v3_(mx[0], mx[1], mx[2]));
}

// RNG (code/random.h)
uint32_t ref_xor_shift_32(uint32_t state) { xor_shift_32(&state); return state; }
float ref_random_between_0_1(uint32_t *state) { RandomSeries s = { *state }; float r = random_between_0_1(&s); *state = s.next_random; return r; }
float ref_random_between(uint32_t *state, float mn, float mx) { RandomSeries s = { *state }; float r = random_between(&s, mn, mx); *state = s.next_random; return r; }
uint32_t ref_random_between_u32(uint32_t *state, uint32_t mn, uint32_t mx) { RandomSeries s = { *state }; u32 r = random_between_u32(&s, mn, mx); *state = s.next_random; return r; }

// BSDF (code/ray.cpp:825-1161).  mat = {Kd[3], Ks[3], Kt[3], ior}
void ref_sample_brdf(uint32_t *state, const float *N, const float *wo, float roughness, const float *mat, float *wi_out, int32_t *is_transmission)
{
    RandomSeries s = { *state };
    SampleBRDFResult r = sample_brdf(&s, v3_(N[0], N[1], N[2]), v3_(wo[0], wo[1], wo[2]), roughness,
                                     v3_(mat[0], mat[1], mat[2]), v3_(mat[3], mat[4], mat[5]), v3_(mat[6], mat[7], mat[8]), mat[9]);
    *state = s.next_random;
    wi_out[0] = r.wi.x; wi_out[1] = r.wi.y; wi_out[2] = r.wi.z;
    *is_transmission = r.is_transmission;
}
float ref_pdf_brdf(const float *N, const float *wi, const float *wo, float roughness, const float *mat)
{
    return pdf_brdf(0, v3_(N[0], N[1], N[2]), v3_(wi[0], wi[1], wi[2]), v3_(wo[0], wo[1], wo[2]), roughness,
                    v3_(mat[0], mat[1], mat[2]), v3_(mat[3], mat[4], mat[5]), v3_(mat[6], mat[7], mat[8]), mat[9]);
}
void ref_eval_scattering(const float *N, const float *wi, const float *wo, const float *mat, float roughness, float distance, float *out)
{
    v3 r = eval_scattering(v3_(N[0], N[1], N[2]), v3_(wi[0], wi[1], wi[2]), v3_(wo[0], wo[1], wo[2]),
                           v3_(mat[0], mat[1], mat[2]), v3_(mat[3], mat[4], mat[5]), v3_(mat[6], mat[7], mat[8]), mat[9],
                           roughness, distance);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// light pick (code/ray.cpp:537-601): returns the advanced RNG state
uint32_t ref_sample_random_lights(void *h, uint32_t state)
{
    RefScene *rs = (RefScene *)h;
    RandomSeries s = { state };
    sample_random_lights(&rs->world.light_push_buffer, rs->world.light_count, &s);
    return s.next_random;
}

// loaders (code/parser.cpp)
// returns is_float; value bits in *bits
int32_t ref_eat_numeric(const char *text, uint32_t *bits)
{
    Tokenizer t = {};
    t.at = (u8 *)text;
    t.one_past_end = (u8 *)text + strlen(text);
    ParseNumericResult r = eat_numeric(&t);
    memcpy(bits, &r.value_f32, 4);
    return r.is_float;
}

// loads a mesh file with the reference parsers; caller frees with ref_free
int32_t ref_load_mesh(const char *path, float **vertices, uint32_t *vertex_count, uint32_t **indices, uint32_t *index_count)
{
    u64 fsize = 0;
    u8 *fmem = read_whole_file(path, &fsize);
    if(!fmem) return -1;
    char extension_buffer[16] = {};
    get_extension(extension_buffer, (char *)path);
    if(string_compare(extension_buffer, (char *)"ply"))
    {
        ParsePlyHeaderResult hd = parse_ply_header(fmem, (u32)fsize);
        *vertex_count = hd.vertex_count; *index_count = hd.index_count;
        *vertices = (float *)malloc(sizeof(v3) * hd.vertex_count);
        *indices = (u32 *)malloc(sizeof(u32) * hd.index_count);
        parse_ply(&hd, fmem, (u32)fsize, (v3 *)*vertices, *indices);
    }
    else if(string_compare(extension_buffer, (char *)"obj"))
    {
        PreParseObjResult pre = pre_parse_obj(fmem, fsize);
        *vertex_count = pre.position_count; *index_count = pre.index_count;
        *vertices = (float *)malloc(sizeof(v3) * pre.position_count);
        *indices = (u32 *)malloc(sizeof(u32) * pre.index_count);
        parse_obj(&pre, fmem, (u32)fsize, (v3 *)*vertices, 0, 0, *indices);
    }
    else { free(fmem); return -2; }
    free(fmem);
    return 0;
}
void ref_free(void *p) { free(p); }

} // extern "C"
