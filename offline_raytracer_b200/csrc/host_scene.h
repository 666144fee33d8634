// host_scene.h -- host side of the drop-in: loaders + scene assembly producing
// the reference's own structs (include/ort_scene.h).  See host_loader.cpp and
// host_scene.cpp.
#pragma once

#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "ort_b200.h"

namespace ort {

struct ParsedLight { uint32_t type; uint32_t index; };     // entry of the light push buffer (push_light, parser.cpp:1144)

struct ParsedMesh                                          // MeshInfo, parser.h:177-189
{
    std::string file_path;
    ort_v3 translate;
    float scale;
    ort_v4 quaternion;
    float degree;
    uint32_t mat_index;
};

struct ParsedScene                                         // ParseSceneResult, parser.h:191-213 (without the 100-entry caps)
{
    int32_t output_width = 0, output_height = 0;
    ort_v3 camera_p = { 0, 0, 0 };
    float camera_height_ratio = 0;
    ort_v4 camera_quaternion = { 0, 0, 0, 0 };
    ort_v3 ambient = { 0, 0, 0 };
    std::vector<OrtSphere> spheres;
    std::vector<OrtAAB> boxes;
    std::vector<OrtCylinder> cylinders;
    std::vector<OrtMaterial> materials;
    std::vector<ParsedMesh> meshes;
    std::vector<ParsedLight> lights;
};

int read_file(const char *path, std::vector<uint8_t> *out, std::string *err);
int parse_numeric_text(const char *text, uint32_t *bits);
int load_ply(const uint8_t *mem, size_t size, std::vector<ort_v3> *vertices, std::vector<uint32_t> *indices, std::string *err);
int load_obj(const uint8_t *mem, size_t size, std::vector<ort_v3> *positions, std::vector<uint32_t> *indices, std::string *err);
int load_mesh_file(const char *path, std::vector<ort_v3> *vertices, std::vector<uint32_t> *indices, std::string *err);
int parse_scene_text(const uint8_t *mem, size_t size, const char *base_dir, ParsedScene *scene, std::string *err);

} // namespace ort

// The assembled scene.  Everything the reference structs point to is owned here
// and stays at a fixed address for the lifetime of the object.
struct OrtHostScene
{
    ort::ParsedScene parsed;
    OrtWorld world;
    OrtCamera camera;
    std::vector<OrtMesh> meshes;
    std::vector<std::vector<ort_v3> > mesh_vertices;
    std::vector<std::vector<uint32_t> > mesh_indices;
    std::vector<uint8_t> light_buffer;                      // packed (u32 type, pointer) pairs
    std::deque<std::vector<OrtBVHOctreeNode> > node_blocks; // octree nodes, 8 children per block
    OrtBVHOctreeNode *top_most_node;
    std::vector<uint8_t> shape_arena;                       // compacted leaf push buffers
    OrtCSG csg;
    bool has_csg;
    uint32_t node_count;
    ort_v3 root_min, root_max;                              // box of the top-most node (macos_main.mm:421-472)
};

namespace ort {
int assemble_scene(const char *scn_path, const char *base_dir, int32_t width, int32_t height, int with_csg,
                   OrtHostScene *hs, std::string *err);
uint32_t v3_to_rgbe(ort_v3 color);
int write_hdr(const char *path, const ort_v3 *pixels, int32_t width, int32_t height, std::string *err);
int write_hdr_rgbe(const char *path, const uint32_t *words, int32_t width, int32_t height, std::string *err);
}
