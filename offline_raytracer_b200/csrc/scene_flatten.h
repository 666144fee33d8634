// scene_flatten.h -- host side of ort_scene_create: walks the reference's octree,
// ranks its leaf records, builds the wide BVH and produces the arrays that are
// uploaded to HBM (layout in bvh.h).
#pragma once

#include <string>
#include <vector>

#include "bvh.h"
#include "ort_b200.h"

namespace ort {

// one leaf record of the octree (code/ray.cpp:637-774), decoded
struct HostPrim
{
    uint32_t kind;      // PRIM_*
    uint32_t rank;      // reference test order (see bvh.h)
    uint32_t mat;
    f3 a, b, c;
    float radius;       // sphere / cylinder
    float lo[3], hi[3]; // conservative, padded bounds
};

// device material, 96 B = 6 x 16 B: OrtMaterial re-packed for 128-bit loads, plus the per-material
// constants that pdf_brdf / sample_brdf / eval_scattering (ray.cpp:1007-1161, 936-1005) would
// otherwise re-derive at every bounce -- six square roots and eight divisions.  They are formed
// here with the same IEEE single-precision operations in the same order (no contraction), so the
// device reads the very bits it would have computed.
struct alignas(16) DevMaterial
{
    float diffuse[3];      int32_t is_light;
    float specular[3];     float ior;
    float transmission[3]; uint32_t lobes;    // bit 0/1/2: length_square(Kd/Ks/Kt) > 0
    float emit[3];         float pad1;
    float pd_c, ps_c, pt_c, pad2;             // |Kd|/s, |Ks|/s, |Kt|/s with s = |Kd|+|Ks|+|Kt| (ray.cpp:1013-1019)
    float ed[3];           float pad3;        // Kd / pi (ray.cpp:940)
};
static_assert(sizeof(DevMaterial) == 96, "DevMaterial is 6 x 16 B");

struct BuildOptions
{
    float traversal_cost;   // SAH cost of visiting a (wide) node relative to one primitive test
    uint32_t max_leaf;      // primitives per leaf child, <= 3
    float pad_rel;          // box padding relative to the primitive's own extent
    float pad_scene;        // box padding relative to the scene's largest |coordinate|
    bool merge_shapes;      // boxes + cylinders in the triangle tree instead of a tree of their own
    bool optimal_collapse;  // binary -> 8-wide by dynamic programming over the SAH cost instead of greedily
    float wide_node_cost;   // SAH cost of visiting a wide node, relative to one primitive test (optimal collapse)
    // optimal collapse, measured on B200 (EXTEND ms, greedy / optimal with node cost 0.6 / 1.0): C3 162.3 / 159.9 / 160.4,
    // 4.4 M-triangle grid 211.2 / 206.6 / 205.4 (node visits per ray 8.73 -> 8.35, primitive tests 2.44 -> 2.55)
    // measured on B200, C3 scene, EXTEND ms per 1080p x 128 spp: cost 0.25/0.5/1.0 = 200/202/217;
    // boxes and cylinders in the triangle tree 202 vs in a tree of their own 221
    BuildOptions() : traversal_cost(0.3f), max_leaf(3), pad_rel(1e-3f), pad_scene(4e-6f), merge_shapes(true), optimal_collapse(true), wide_node_cost(1.0f) {}
};

struct FlatScene
{
    std::vector<WideNode> nodes;
    std::vector<PrimRec> prims;
    std::vector<CylinderAux> cylinders;
    std::vector<DevMaterial> materials;
    std::vector<uint8_t> light_is_sphere;   // one entry per light-list entry (parser.cpp:1144-1182)
    OrtSceneInfo info;
    uint32_t wide_depth;
    uint32_t main_root;     // nodes [0, main_root) = sphere tree (absent if 0)
    uint32_t tri_root;      // nodes [main_root, tri_root) = box/cylinder tree (absent if equal), triangles from tri_root
    std::vector<uint32_t> rank_to_prim;   // record index of each rank (0xFFFFFFFF for dropped CSG records)
};

// decode the octree into ranked records (+ world materials / light list)
int collect_records(const OrtWorld *world, const OrtBVHOctreeNode *root,
                    std::vector<HostPrim> *prims, FlatScene *out, std::string *err);
// the same records and ranks from the shape lists alone, no octree (include/ort_b200.h, OrtShapeLists)
int collect_records_from_lists(const OrtWorld *world, const OrtShapeLists *lists,
                               std::vector<HostPrim> *prims, FlatScene *out, std::string *err);
// build the wide BVH over `prims` (consumed) into out->nodes / prims / cylinders
int build_wide_bvh(std::vector<HostPrim> &prims, const BuildOptions &opt, FlatScene *out, std::string *err);

// ---- the data-parallel builder (bvh_build.h; CUDA execution in bvh_build.cuh) ----
struct ParallelBuildInput
{
    std::vector<float> boxes;        // 6 floats per primitive: padded lo, hi -- input order
    std::vector<PrimRec> recs;       // the device records, input order
    double scene_lo[3], scene_scale[3];   // box-centre bounds -> 21-bit grid
    uint32_t max_leaf;
    float traversal_cost;
    float node_cost;                 // cost of a wide-node visit relative to a primitive test (optimal collapse)
    uint32_t sphere_depth;           // depth of the sphere tree already emitted into the FlatScene
};
int prepare_parallel_build(std::vector<HostPrim> &prims, const BuildOptions &opt, FlatScene *out, ParallelBuildInput *in, std::string *err);
int build_wide_bvh_parallel_host(std::vector<HostPrim> &prims, const BuildOptions &opt, uint32_t radius, FlatScene *out, std::string *err);
#define ORT_PLOC_RADIUS 16u
#define ORT_PLOC_TOP_CLUSTERS 65536u     // the last so many clusters (at most n / 64) go to the SAH builder (ORT_PLOC_TOP overrides; 0 = none)
uint32_t parallel_top_clusters(uint32_t n);
namespace build { struct B2; struct Dp; }
int build_top_tree(const std::vector<uint32_t> &clusters, const BuildOptions &opt, uint32_t max_leaf, float node_cost,
                   build::B2 *nodes, uint32_t *sizes, build::Dp *cost, uint32_t *next_node, uint32_t *root_out, std::string *err);
// pieces of the host path the device execution (bvh_build.cuh) reuses for the few analytic shapes
int collect_analytic(const OrtWorld *world, const OrtShapeLists *lists, f3 root_center, f3 root_half,
                     std::vector<HostPrim> *analytic, uint32_t *keys, std::string *err);
int fill_world_tables_public(const OrtWorld *world, FlatScene *out, std::string *err);
int pad_and_split_public(std::vector<HostPrim> &prims, const BuildOptions &opt, double scene_abs,
                         std::vector<HostPrim> *spheres, std::vector<HostPrim> *shapes, std::vector<HostPrim> *tris, std::string *err);
int emit_tree_public(const std::vector<HostPrim> &prims, const BuildOptions &opt, FlatScene *out, uint32_t *depth_out, std::string *err);
PrimRec make_record_public(const HostPrim &hp, FlatScene *out);

inline int flatten_scene(const OrtWorld *world, const OrtBVHOctreeNode *root, const BuildOptions &opt,
                         FlatScene *out, std::string *err)
{
    std::vector<HostPrim> prims;
    int rc = collect_records(world, root, &prims, out, err);
    if(rc != ORT_OK) return rc;
    return build_wide_bvh(prims, opt, out, err);
}

} // namespace ort
