// host_scene.cpp -- scene assembly on the host: the drop-in for the setup half of
// the reference's main() (code/macos_main.mm:310-562) and for its octree builder
// (code/ray.cpp:1469-2045), plus the Radiance .hdr writer (macos_main.mm:242-287,
// 682-707) and the extern "C" host entry points of include/ort_b200.h.
//
// The octree is needed for one thing only: the ORDER of its leaf records is the
// tie-break rank of the hot path (bvh.h).  That order depends on every float
// comparison of the insertion, so the arithmetic below keeps the reference's
// operation order (f32, no contraction; vector helpers from core_math.h follow
// code/math.h) and its quirks: boxes start at (FLT_MAX, FLT_MIN) -- FLT_MIN being
// the smallest POSITIVE float (ray.cpp:1760-1761, macos_main.mm:383,424) --, the
// cylinder bound r*(1 - sqrt(a_i^2/|a|^2)) (ray.cpp:1694), the mesh bake order
// scale -> rotate about (0,1,0) by `degree` -> quaternion -> translate
// (macos_main.mm:397-400), insertion order triangles -> cylinders -> boxes ->
// spheres -> CSG (macos_main.mm:478-538).
#include "host_scene.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <utility>

#include <new>
#include <stdexcept>
#include "core_math.h"

namespace ort {

namespace {

inline f3 V(const ort_v3 &v) { return mk3(v.x, v.y, v.z); }
inline ort_v3 U(f3 v) { ort_v3 r; r.x = v.x; r.y = v.y; r.z = v.z; return r; }
inline f3 gather_min(f3 a, f3 b) { return mk3(ref_min(a.x, b.x), ref_min(a.y, b.y), ref_min(a.z, b.z)); }   // math.h:1085
inline f3 gather_max(f3 a, f3 b) { return mk3(ref_max(a.x, b.x), ref_max(a.y, b.y), ref_max(a.z, b.z)); }   // math.h:1097

// quaternion_rotation(v4 q, v3 v), math.h:774-795
f3 quat_rotate(ort_v4 q, f3 v)
{
    float m00 = 1.0f - 2 * q.y * q.y - 2 * q.z * q.z;
    float m01 = 2 * q.x * q.y - 2 * q.w * q.z;
    float m02 = 2 * q.x * q.z + 2 * q.w * q.y;
    float m10 = 2 * q.x * q.y + 2 * q.w * q.z;
    float m11 = 1.0f - 2 * q.x * q.x - 2 * q.z * q.z;
    float m12 = 2 * q.y * q.z - 2 * q.w * q.x;
    float m20 = 2 * q.x * q.z - 2 * q.w * q.y;
    float m21 = 2 * q.y * q.z + 2 * q.w * q.x;
    float m22 = 1 - 2 * q.x * q.x - 2 * q.y * q.y;
    return mk3(m00 * v.x + m01 * v.y + m02 * v.z,
               m10 * v.x + m11 * v.y + m12 * v.z,
               m20 * v.x + m21 * v.y + m22 * v.z);
}

// quaternion_rotation(v3 axis, r32 rad, v3 v), math.h:746-772
f3 axis_rotate(f3 axis, float rad, f3 v)
{
    ort_v4 q;
    q.w = cosf(rad / 2);
    q.x = axis.x * sinf(rad / 2);
    q.y = axis.y * sinf(rad / 2);
    q.z = axis.z * sinf(rad / 2);
    return quat_rotate(q, v);
}

struct ShapeBox { f3 center, half_dim; };

// get_shape_aabb, ray.cpp:1675-1746
ShapeBox shape_box(const void *shape, uint32_t type)
{
    ShapeBox r; r.center = mk3(0, 0, 0); r.half_dim = mk3(0, 0, 0);
    switch(type)
    {
        case ORT_SHAPE_SPHERE:
        {
            const OrtSphere *s = (const OrtSphere *)shape;
            r.center = V(s->center);
            r.half_dim = s->r * mk3(1, 1, 1);
        } break;
        case ORT_SHAPE_CYLINDER:
        {
            const OrtCylinder *c = (const OrtCylinder *)shape;
            f3 base = V(c->base), axis = V(c->axis);
            f3 other = base + axis;
            f3 q = hadamard(axis, axis) / dot(axis, axis);
            f3 e = c->r * (mk3(1, 1, 1) - mk3(sqrtf(q.x), sqrtf(q.y), sqrtf(q.z)));
            f3 mn = gather_min(base - e, other - e);
            f3 mx = gather_max(base + e, other + e);
            r.center = 0.5f * (mn + mx);
            r.half_dim = mx - r.center;
        } break;
        case ORT_SHAPE_AAB:
        {
            const OrtAAB *b = (const OrtAAB *)shape;
            r.center = 0.5f * (V(b->min) + V(b->max));
            r.half_dim = V(b->max) - r.center;
        } break;
        case ORT_SHAPE_MESH:
        {
            const OrtMesh *m = (const OrtMesh *)shape;
            r.center = 0.5f * (V(m->aabb_min) + V(m->aabb_max));
            r.half_dim = V(m->aabb_max) - r.center;
        } break;
        case ORT_SHAPE_TRIANGLE:
        {
            const OrtTriangle *t = (const OrtTriangle *)shape;
            f3 v0 = V(t->mesh->vertices[t->i_0]), v1 = V(t->mesh->vertices[t->i_1]), v2 = V(t->mesh->vertices[t->i_2]);
            f3 mn = gather_min(gather_min(v0, v1), v2);
            f3 mx = gather_max(gather_max(v0, v1), v2);
            r.center = 0.5f * (mn + mx);
            r.half_dim = mx - r.center;
        } break;
        case ORT_SHAPE_CSG:
        {
            const OrtCSG *c = (const OrtCSG *)shape;
            r.center = 0.5f * (V(c->aabb_min) + V(c->aabb_max));
            r.half_dim = V(c->aabb_max) - r.center;
        } break;
    }
    return r;
}

// update_aabb_min_max, ray.cpp:1765-1777
void grow_box(ort_v3 *dst_min, ort_v3 *dst_max, const void *shape, uint32_t type)
{
    ShapeBox b = shape_box(shape, type);
    f3 mn = b.center - b.half_dim;
    f3 mx = b.center + b.half_dim;
    *dst_min = U(gather_min(V(*dst_min), mn));
    *dst_max = U(gather_max(V(*dst_max), mx));
}

uint32_t payload_bytes(uint32_t type)
{
    switch(type)
    {
        case ORT_SHAPE_SPHERE:   return (uint32_t)sizeof(OrtSphere);
        case ORT_SHAPE_AAB:      return (uint32_t)sizeof(OrtAAB);
        case ORT_SHAPE_CYLINDER: return (uint32_t)sizeof(OrtCylinder);
        case ORT_SHAPE_TRIANGLE: return (uint32_t)sizeof(OrtTriangle);
        case ORT_SHAPE_CSG:      return (uint32_t)sizeof(OrtCSG);
        default: return 0;
    }
}

struct OctreeBuilder
{
    OrtHostScene *hs;
    uint32_t desired_depth;

    // push_shape, ray.cpp:1524-1629: header + a COPY of the shape struct.  During
    // the build a leaf's records live in a malloc'd buffer; compact() moves them
    // into the shape arena (validate_nodes_and_reallocate_shapes, ray.cpp:1960).
    void append_record(OrtTempMemory *buf, const void *shape, uint32_t type)
    {
        size_t need = 4 + payload_bytes(type);
        if(buf->used + need > buf->total_size)
        {
            size_t cap = buf->total_size ? buf->total_size * 2 : 128;
            while(cap < buf->used + need) cap *= 2;
            buf->base = realloc(buf->base, cap);
            buf->total_size = cap;
        }
        memcpy((uint8_t *)buf->base + buf->used, &type, 4);
        memcpy((uint8_t *)buf->base + buf->used + 4, shape, need - 4);
        buf->used += need;
    }

    OrtBVHOctreeNode *alloc_children()
    {
        hs->node_blocks.push_back(std::vector<OrtBVHOctreeNode>(8));
        OrtBVHOctreeNode *c = hs->node_blocks.back().data();
        for(int i = 0; i < 8; ++i)                                        // initialize_all_childs, ray.cpp:1748-1763
        {
            memset(&c[i], 0, sizeof(c[i]));
            c[i].is_leaf = 1;
            c[i].aabb_min = U(mk3(FLT_MAX, FLT_MAX, FLT_MAX));
            c[i].aabb_max = U(mk3(FLT_MIN, FLT_MIN, FLT_MIN));
        }
        hs->node_count += 8;
        return c;
    }

    // get_bvh_octree_node_child_info + find_least_significant_bit (ray.cpp:1476-1522,
    // platform.h:140-162): octant of the shape-box CENTRE; bit0 = +x, bit1 = +y, bit2 = +z
    static uint32_t pick_child(f3 node_center, f3 node_half, f3 shape_center, f3 *child_center, f3 *child_half)
    {
        f3 h = 0.5f * node_half;
        f3 c = node_center;
        uint32_t idx = 0;
        if(shape_center.x >= node_center.x) { idx |= 1u; c.x += h.x; } else c.x -= h.x;
        if(shape_center.y >= node_center.y) { idx |= 2u; c.y += h.y; } else c.y -= h.y;
        if(shape_center.z >= node_center.z) { idx |= 4u; c.z += h.z; } else c.z -= h.z;
        *child_center = c; *child_half = h;
        return idx;
    }

    // push_shape_inside_node, ray.cpp:1799-1948
    void insert(OrtBVHOctreeNode *node, f3 center, f3 half, uint32_t depth, const void *shape, uint32_t type)
    {
        grow_box(&node->aabb_min, &node->aabb_max, shape, type);
        if(depth >= desired_depth)
        {
            append_record(&node->push_buffer, shape, type);
            return;
        }
        if(!node->first_child)
        {
            if(node->push_buffer.used == 0)
            {
                append_record(&node->push_buffer, shape, type);
                return;
            }
            // occupied leaf: split, re-insert what it held (in buffer order), then the new shape
            node->first_child = alloc_children();
            OrtTempMemory old = node->push_buffer;
            memset(&node->push_buffer, 0, sizeof(node->push_buffer));
            for(size_t consumed = 0; consumed < old.used; )
            {
                uint32_t t; memcpy(&t, (uint8_t *)old.base + consumed, 4);
                const void *existing = (uint8_t *)old.base + consumed + 4;
                consumed += 4 + payload_bytes(t);
                // the payload may sit at a 4-byte boundary: copy it out before use
                uint8_t tmp[sizeof(OrtCSG)];
                memcpy(tmp, existing, payload_bytes(t));
                descend(node, center, half, depth, tmp, t);
            }
            free(old.base);
            node->is_leaf = 0;
        }
        descend(node, center, half, depth, shape, type);
    }

    void descend(OrtBVHOctreeNode *node, f3 center, f3 half, uint32_t depth, const void *shape, uint32_t type)
    {
        ShapeBox b = shape_box(shape, type);
        f3 cc, ch;
        uint32_t idx = pick_child(center, half, b.center, &cc, &ch);
        insert(node->first_child + idx, cc, ch, depth + 1, shape, type);
    }

    size_t total_record_bytes(const OrtBVHOctreeNode *node)
    {
        size_t n = node->push_buffer.used;
        if(node->first_child) for(int i = 0; i < 8; ++i) n += total_record_bytes(node->first_child + i);
        return n;
    }

    // validate_nodes_and_reallocate_shapes, ray.cpp:1960-2045 (depth-first)
    void compact(OrtBVHOctreeNode *node, size_t *cursor)
    {
        if(node->push_buffer.used)
        {
            uint8_t *dst = hs->shape_arena.data() + *cursor;
            memcpy(dst, node->push_buffer.base, node->push_buffer.used);
            free(node->push_buffer.base);
            node->push_buffer.base = dst;
            node->push_buffer.total_size = node->push_buffer.used;
            node->push_buffer.memory_arena = 0;
            *cursor += node->push_buffer.used;
        }
        if(node->first_child) for(int i = 0; i < 8; ++i) compact(node->first_child + i, cursor);
    }
};

} // namespace

int assemble_scene(const char *scn_path, const char *base_dir, int32_t width, int32_t height, int with_csg,
                   OrtHostScene *hs, std::string *err)
{
    if(width <= 0 || height <= 0) { *err = "output size must be positive"; return ORT_ERR_ARG; }
    std::vector<uint8_t> text;
    int rc = read_file(scn_path, &text, err);
    if(rc != ORT_OK) return rc;
    rc = parse_scene_text(text.data(), text.size(), base_dir, &hs->parsed, err);
    if(rc != ORT_OK) return rc;
    ParsedScene &sc = hs->parsed;
    sc.output_width = width;              // the driver overrides the .scn's `screen` (macos_main.mm:319-320)
    sc.output_height = height;

    // hard-coded, inert CSG (macos_main.mm:322-332); it occupies a rank
    const bool build_octree = (with_csg & ORT_HOST_NO_OCTREE) == 0;
    const bool bake_on_device = (with_csg & ORT_HOST_BAKE_ON_DEVICE) != 0;
    hs->has_csg = (with_csg & 1) != 0;
    memset(&hs->csg, 0, sizeof(hs->csg));
    if(hs->has_csg)
    {
        f3 c = mk3(0, 0, 0.8f);
        hs->csg.sphere.center = U(c);
        hs->csg.sphere.r = 0.35f;
        hs->csg.aab.min = U(c - mk3(0.3f, 0.3f, 0.3f));
        hs->csg.aab.max = U(c + mk3(0.3f, 0.3f, 0.3f));
        hs->csg.mat_index = 5;
    }

    // light push buffer: packed (u32 ShapeType, pointer) pairs (parser.cpp:1144-1182)
    hs->light_buffer.clear();
    for(size_t i = 0; i < sc.lights.size(); ++i)
    {
        const void *p = sc.lights[i].type == ORT_SHAPE_SPHERE ? (const void *)&sc.spheres[sc.lights[i].index]
                                                              : (const void *)&sc.cylinders[sc.lights[i].index];
        uint32_t type = sc.lights[i].type;
        size_t at = hs->light_buffer.size();
        hs->light_buffer.resize(at + 4 + sizeof(void *));
        memcpy(&hs->light_buffer[at], &type, 4);
        memcpy(&hs->light_buffer[at + 4], &p, sizeof(void *));
    }
    memset(&hs->world, 0, sizeof(hs->world));
    hs->world.ambient = sc.ambient;
    hs->world.materials = sc.materials.data();
    hs->world.mat_count = (uint32_t)sc.materials.size();
    hs->world.light_push_buffer.base = hs->light_buffer.data();
    hs->world.light_push_buffer.used = hs->light_buffer.size();
    hs->world.light_push_buffer.total_size = hs->light_buffer.size();
    hs->world.light_count = (uint32_t)sc.lights.size();

    // meshes: load, bake, bound (macos_main.mm:341-414)
    size_t nm = sc.meshes.size();
    hs->meshes.assign(nm, OrtMesh());
    hs->mesh_vertices.assign(nm, std::vector<ort_v3>());
    hs->mesh_indices.assign(nm, std::vector<uint32_t>());
    std::map<std::string, size_t> parsed_as;
    std::vector<std::pair<std::vector<ort_v3>, std::vector<uint32_t> > > pristine;      // as parsed, before the bake
    for(size_t mi = 0; mi < nm; ++mi)
    {
        const ParsedMesh &info = sc.meshes[mi];
        // the reference re-parses the file for every `mesh` line (macos_main.mm:344-381); parsing is a
        // pure function of the file, so a file seen before is copied instead (config 5 names bunny.ply 729 times)
        {
            std::map<std::string, size_t>::const_iterator seen = parsed_as.find(info.file_path);
            if(seen == parsed_as.end())
            {
                rc = load_mesh_file(info.file_path.c_str(), &hs->mesh_vertices[mi], &hs->mesh_indices[mi], err);
                if(rc != ORT_OK) return rc;
                parsed_as[info.file_path] = pristine.size();
                pristine.push_back(std::make_pair(hs->mesh_vertices[mi], hs->mesh_indices[mi]));
            }
            else
            {
                hs->mesh_vertices[mi] = pristine[seen->second].first;
                hs->mesh_indices[mi] = pristine[seen->second].second;
            }
        }
        OrtMesh &mesh = hs->meshes[mi];
        memset(&mesh, 0, sizeof(mesh));
        mesh.vertices = hs->mesh_vertices[mi].data();
        mesh.vertex_count = (uint32_t)hs->mesh_vertices[mi].size();
        mesh.indices = hs->mesh_indices[mi].data();
        mesh.index_count = (uint32_t)hs->mesh_indices[mi].size();
        mesh.mat_index = info.mat_index;
        for(size_t k = 0; k < hs->mesh_indices[mi].size(); ++k)
            if(hs->mesh_indices[mi][k] >= mesh.vertex_count) { *err = "mesh index out of range: " + info.file_path; return ORT_ERR_PARSE; }

        if(bake_on_device)
        {
            // the same loop as a CUDA kernel (ort_tools.cu: k_bake_mesh), bit-identical vertices and box
            int brc = ort_bake_mesh(0, mesh.vertex_count, mesh.vertices, mesh.vertices, info.scale, info.degree,
                                    info.quaternion, info.translate, &mesh.aabb_min, &mesh.aabb_max);
            if(brc != ORT_OK) { *err = std::string("device bake: ") + ort_last_error(); return brc; }
            continue;
        }
        f3 mn = mk3(FLT_MAX, FLT_MAX, FLT_MAX);
        f3 mx = mk3(FLT_MIN, FLT_MIN, FLT_MIN);
        for(uint32_t vi = 0; vi < mesh.vertex_count; ++vi)
        {
            f3 v = V(mesh.vertices[vi]);
            v = mk3(v.x * info.scale, v.y * info.scale, v.z * info.scale);                 // *v *= scale
            v = quat_rotate(info.quaternion, axis_rotate(mk3(0, 1, 0), 0.0174533f * info.degree, v));
            v = v + V(info.translate);
            mesh.vertices[vi] = U(v);
            mn = mk3(ref_min(mn.x, v.x), ref_min(mn.y, v.y), ref_min(mn.z, v.z));
            mx = mk3(ref_max(mx.x, v.x), ref_max(mx.y, v.y), ref_max(mx.z, v.z));
            mesh.aabb_min = U(mn);
            mesh.aabb_max = U(mx);
        }
    }

    // root node and its box over all shapes (macos_main.mm:421-472)
    hs->node_blocks.clear();
    hs->node_blocks.push_back(std::vector<OrtBVHOctreeNode>(1));
    OrtBVHOctreeNode *root = hs->node_blocks.back().data();
    memset(root, 0, sizeof(*root));
    root->aabb_min = U(mk3(FLT_MAX, FLT_MAX, FLT_MAX));
    root->aabb_max = U(mk3(FLT_MIN, FLT_MIN, FLT_MIN));
    hs->node_count = 1;
    for(size_t i = 0; i < nm; ++i) grow_box(&root->aabb_min, &root->aabb_max, &hs->meshes[i], ORT_SHAPE_MESH);
    for(size_t i = 0; i < sc.cylinders.size(); ++i) grow_box(&root->aabb_min, &root->aabb_max, &sc.cylinders[i], ORT_SHAPE_CYLINDER);
    for(size_t i = 0; i < sc.boxes.size(); ++i) grow_box(&root->aabb_min, &root->aabb_max, &sc.boxes[i], ORT_SHAPE_AAB);
    for(size_t i = 0; i < sc.spheres.size(); ++i) grow_box(&root->aabb_min, &root->aabb_max, &sc.spheres[i], ORT_SHAPE_SPHERE);
    if(hs->has_csg)
    {
        hs->csg.aabb_min = U(FLT_MAX * mk3(1, 1, 1));
        hs->csg.aabb_max = U(FLT_MIN * mk3(1, 1, 1));
        grow_box(&hs->csg.aabb_min, &hs->csg.aabb_max, &hs->csg.sphere, ORT_SHAPE_SPHERE);
        grow_box(&hs->csg.aabb_min, &hs->csg.aabb_max, &hs->csg.aab, ORT_SHAPE_AAB);
    }
    f3 root_center = 0.5f * (V(root->aabb_min) + V(root->aabb_max));
    f3 root_half = V(root->aabb_max) - root_center;
    hs->root_min = root->aabb_min; hs->root_max = root->aabb_max;
    hs->top_most_node = 0;
    if(build_octree)
    {

    // insertion (macos_main.mm:474-538), depth 10
    OctreeBuilder ob; ob.hs = hs; ob.desired_depth = 10;
    for(size_t mi = 0; mi < nm; ++mi)
    {
        OrtMesh *mesh = &hs->meshes[mi];
        for(uint32_t k = 0; k + 2 < mesh->index_count; k += 3)
        {
            OrtTriangle tri; memset(&tri, 0, sizeof(tri));
            tri.mesh = mesh;
            tri.i_0 = mesh->indices[k]; tri.i_1 = mesh->indices[k + 1]; tri.i_2 = mesh->indices[k + 2];
            ob.insert(root, root_center, root_half, 0, &tri, ORT_SHAPE_TRIANGLE);
        }
    }
    for(size_t i = 0; i < sc.cylinders.size(); ++i) ob.insert(root, root_center, root_half, 0, &sc.cylinders[i], ORT_SHAPE_CYLINDER);
    for(size_t i = 0; i < sc.boxes.size(); ++i) ob.insert(root, root_center, root_half, 0, &sc.boxes[i], ORT_SHAPE_AAB);
    for(size_t i = 0; i < sc.spheres.size(); ++i) ob.insert(root, root_center, root_half, 0, &sc.spheres[i], ORT_SHAPE_SPHERE);
    if(hs->has_csg) ob.insert(root, root_center, root_half, 0, &hs->csg, ORT_SHAPE_CSG);

    hs->shape_arena.assign(ob.total_record_bytes(root) + 16, 0);
    size_t cursor = 0;
    ob.compact(root, &cursor);
    hs->top_most_node = root;
    }

    // camera (macos_main.mm:548-556): axes pre-scaled by the image-plane extents
    hs->camera.p = sc.camera_p;
    float rx = sc.camera_height_ratio * ((float)sc.output_width / sc.output_height);
    hs->camera.x_axis = U(rx * quat_rotate(sc.camera_quaternion, mk3(1, 0, 0)));
    hs->camera.y_axis = U(sc.camera_height_ratio * quat_rotate(sc.camera_quaternion, mk3(0, 1, 0)));
    hs->camera.z_axis = U(quat_rotate(sc.camera_quaternion, mk3(0, 0, 1)));
    return ORT_OK;
}

// v3_to_rgbe, macos_main.mm:242-261
uint32_t v3_to_rgbe(ort_v3 color)
{
    uint32_t result = 0;
    float max_component = ref_max(ref_max(color.x, color.y), color.z);
    int e;
    if(max_component >= 1e-32f)
    {
        float denom = frexpf(max_component, &e) * 255.0f / max_component;
        result = (((uint32_t)roundf(color.x * denom) << 0) |
                  ((uint32_t)roundf(color.y * denom) << 8) |
                  ((uint32_t)roundf(color.z * denom) << 16) |
                  ((uint32_t)(e + 128) << 24));
    }
    return result;
}

// write_hdr_header + the output loop of main(), macos_main.mm:263-287, 682-707:
// flat RGBE, buffer rows emitted height-1 -> 0 (the top of the picture first)
int write_hdr(const char *path, const ort_v3 *pixels, int32_t width, int32_t height, std::string *err)
{
    if(width <= 0 || height <= 0 || !pixels) { *err = "bad image"; return ORT_ERR_ARG; }
    FILE *f = fopen(path, "wb");
    if(!f) { *err = std::string("cannot create ") + path; return ORT_ERR_IO; }
    fprintf(f, "#?RADIANCE\n");
    fprintf(f, "FORMAT=32-bit_rle_rgbe\n\n");
    fprintf(f, "+Y %d +X %d\n", height, width);
    std::vector<uint32_t> row((size_t)width);
    for(int32_t y = height - 1; y >= 0; --y)
    {
        const ort_v3 *src = pixels + (size_t)y * width;
        for(int32_t x = 0; x < width; ++x) row[x] = v3_to_rgbe(src[x]);
        if(fwrite(row.data(), 4, (size_t)width, f) != (size_t)width) { fclose(f); *err = "short write"; return ORT_ERR_IO; }
    }
    fclose(f);
    return ORT_OK;
}

// the same file from RGBE words already in file order (encoded on the device, kernels.cuh)
int write_hdr_rgbe(const char *path, const uint32_t *words, int32_t width, int32_t height, std::string *err)
{
    if(width <= 0 || height <= 0 || !words) { *err = "bad image"; return ORT_ERR_ARG; }
    FILE *f = fopen(path, "wb");
    if(!f) { *err = std::string("cannot create ") + path; return ORT_ERR_IO; }
    fprintf(f, "#?RADIANCE\n");
    fprintf(f, "FORMAT=32-bit_rle_rgbe\n\n");
    fprintf(f, "+Y %d +X %d\n", height, width);
    size_t n = (size_t)width * height;
    bool ok = fwrite(words, 4, n, f) == n;
    fclose(f);
    if(!ok) { *err = "short write"; return ORT_ERR_IO; }
    return ORT_OK;
}

} // namespace ort

// ---------------------------------------------------------------------------
// extern "C" host entry points (include/ort_b200.h)
// ---------------------------------------------------------------------------
extern "C" void ort_set_last_error_(const char *msg);   // ort_b200.cu

extern "C" {

int ort_host_scene_load(const char *scn_path, const char *base_dir, int32_t width, int32_t height, int with_csg, OrtHostScene **out)
{
    if(!scn_path || !base_dir || !out) { ort_set_last_error_("null argument"); return ORT_ERR_ARG; }
    *out = 0;
    OrtHostScene *hs = 0;
    // no C++ exception crosses the C ABI: a file that makes a container throw is a reported error, not std::terminate
    try
    {
        hs = new OrtHostScene();
        std::string err;
        int rc = ort::assemble_scene(scn_path, base_dir, width, height, with_csg, hs, &err);
        if(rc != ORT_OK) { ort_set_last_error_(err.c_str()); delete hs; return rc; }
    }
    catch(const std::bad_alloc &) { delete hs; ort_set_last_error_("out of host memory while loading the scene"); return ORT_ERR_LIMIT; }
    catch(const std::exception &e) { delete hs; ort_set_last_error_((std::string("scene loader: ") + e.what()).c_str()); return ORT_ERR_PARSE; }
    *out = hs;
    return ORT_OK;
}

int ort_host_scene_destroy(OrtHostScene *hs) { delete hs; return ORT_OK; }
const OrtWorld *ort_host_scene_world(const OrtHostScene *hs) { return hs ? &hs->world : 0; }
const OrtCamera *ort_host_scene_camera(const OrtHostScene *hs) { return hs ? &hs->camera : 0; }
const OrtBVHOctreeNode *ort_host_scene_root(const OrtHostScene *hs) { return hs ? hs->top_most_node : 0; }
const OrtMesh *ort_host_scene_meshes(const OrtHostScene *hs, uint32_t *mesh_count)
{
    if(mesh_count) *mesh_count = hs ? (uint32_t)hs->meshes.size() : 0;
    return hs && !hs->meshes.empty() ? hs->meshes.data() : 0;
}

int ort_host_scene_lists(const OrtHostScene *hs, OrtShapeLists *out)
{
    if(!hs || !out) { ort_set_last_error_("null argument"); return ORT_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    const ort::ParsedScene &sc = hs->parsed;
    out->meshes = hs->meshes.empty() ? 0 : hs->meshes.data(); out->mesh_count = (uint32_t)hs->meshes.size();
    out->cylinders = sc.cylinders.empty() ? 0 : sc.cylinders.data(); out->cylinder_count = (uint32_t)sc.cylinders.size();
    out->boxes = sc.boxes.empty() ? 0 : sc.boxes.data(); out->box_count = (uint32_t)sc.boxes.size();
    out->spheres = sc.spheres.empty() ? 0 : sc.spheres.data(); out->sphere_count = (uint32_t)sc.spheres.size();
    out->csg = hs->has_csg ? &hs->csg : 0;
    out->root_min = hs->root_min; out->root_max = hs->root_max;
    return ORT_OK;
}

int ort_load_mesh(const char *path, float **vertices, uint32_t *vertex_count, uint32_t **indices, uint32_t *index_count)
{
    if(!path || !vertices || !vertex_count || !indices || !index_count) { ort_set_last_error_("null argument"); return ORT_ERR_ARG; }
    std::vector<ort_v3> v; std::vector<uint32_t> i; std::string err;
    int rc = ORT_OK;
    try { rc = ort::load_mesh_file(path, &v, &i, &err); }
    catch(const std::bad_alloc &) { ort_set_last_error_("out of host memory while loading the mesh"); return ORT_ERR_LIMIT; }
    catch(const std::exception &e) { ort_set_last_error_((std::string("mesh loader: ") + e.what()).c_str()); return ORT_ERR_PARSE; }
    if(rc != ORT_OK) { ort_set_last_error_(err.c_str()); return rc; }
    *vertex_count = (uint32_t)v.size(); *index_count = (uint32_t)i.size();
    *vertices = (float *)malloc(sizeof(ort_v3) * (v.size() ? v.size() : 1));
    *indices = (uint32_t *)malloc(sizeof(uint32_t) * (i.size() ? i.size() : 1));
    if(!v.empty()) memcpy(*vertices, v.data(), sizeof(ort_v3) * v.size());
    if(!i.empty()) memcpy(*indices, i.data(), sizeof(uint32_t) * i.size());
    return ORT_OK;
}

int ort_parse_numeric(const char *text, uint32_t *value_bits)
{
    uint32_t bits = 0;
    int is_float = ort::parse_numeric_text(text ? text : "", &bits);
    if(value_bits) *value_bits = bits;
    return is_float;
}

void ort_free(void *p) { free(p); }

int ort_write_hdr(const char *path, const ort_v3 *pixels, int32_t width, int32_t height)
{
    std::string err;
    int rc = ort::write_hdr(path, pixels, width, height, &err);
    if(rc != ORT_OK) ort_set_last_error_(err.c_str());
    return rc;
}

int ort_write_hdr_rgbe(const char *path, const uint32_t *rgbe_words, int32_t width, int32_t height)
{
    std::string err;
    int rc = ort::write_hdr_rgbe(path, rgbe_words, width, height, &err);
    if(rc != ORT_OK) ort_set_last_error_(err.c_str());
    return rc;
}

uint32_t ort_v3_to_rgbe(ort_v3 color) { return ort::v3_to_rgbe(color); }

} // extern "C"
