/* include/ort_scene.h -- the scene DATA CONTRACT of the drop-in boundary.
 *
 * These PODs have, field for field, the memory layout of the reference's host
 * structs (gyuhyun-lee/offline_raytracer), so that a pointer to the reference's
 * own World / Camera / BVHOctreeNode can be handed to the C ABI in ort_b200.h
 * unchanged.  They are prefixed `Ort` so that this header can be included next
 * to the reference's ray.h without name clashes.
 *
 *   reference struct        file:line                 here           sizeof
 *   v3 / v4                 code/types.h:61-123       ort_v3/ort_v4  12 / 16
 *   TempMemory              code/platform.h:218-226   OrtTempMemory  32
 *   Sphere                  code/ray.h:4-10           OrtSphere      20
 *   AAB                     code/ray.h:12-18          OrtAAB         28
 *   Cylinder                code/ray.h:20-27          OrtCylinder    32
 *   Material                code/ray.h:30-40          OrtMaterial    60
 *   Camera                  code/ray.h:42-49          OrtCamera      48
 *   Mesh                    code/ray.h:51-65          OrtMesh        56
 *   Triangle                code/ray.h:67-74          OrtTriangle    24
 *   World                   code/ray.h:76-88          OrtWorld       80
 *   ShapeType               code/ray.h:97-106         OrtShapeType   4
 *   BVHShapeHeader          code/ray.h:108-111        (u32 type tag) 4
 *   BVHOctreeNode           code/ray.h:115-133        OrtBVHOctreeNode 80
 *   CSG                     code/ray.h:160-176        OrtCSG         76
 *   RandomSeries            code/random.h:18-21       OrtRandomSeries 4
 *
 * Sizes are those measured on the reference itself built for x86-64 Linux
 * (oracle/_ref, ref_struct_sizes) and are statically asserted below.
 */
#ifndef ORT_SCENE_H
#define ORT_SCENE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ort_v3 { float x, y, z; } ort_v3;
typedef struct ort_v4 { float x, y, z, w; } ort_v4;

typedef struct OrtTempMemory
{
    void  *memory_arena;   /* MemoryArena*, never dereferenced by the hot path */
    void  *base;
    size_t total_size;
    size_t used;
} OrtTempMemory;

typedef struct OrtSphere   { ort_v3 center; float r; uint32_t mat_index; } OrtSphere;
typedef struct OrtAAB      { ort_v3 min; ort_v3 max; uint32_t mat_index; } OrtAAB;
typedef struct OrtCylinder { ort_v3 base; ort_v3 axis; float r; uint32_t mat_index; } OrtCylinder;

typedef struct OrtMaterial
{
    ort_v3  diffuse;
    ort_v4  specular;      /* .w is parsed but unused: roughness is the literal 0.01 (ray.cpp:1194) */
    ort_v3  transmission;
    float   ior;
    ort_v3  emit_color;
    int32_t is_light;
} OrtMaterial;

typedef struct OrtCamera { ort_v3 p; ort_v3 x_axis; ort_v3 y_axis; ort_v3 z_axis; } OrtCamera;

typedef struct OrtMesh
{
    ort_v3   *vertices;    /* already scaled / rotated / translated */
    uint32_t  vertex_count;
    uint32_t *indices;
    uint32_t  index_count;
    uint32_t  mat_index;
    ort_v3    aabb_min;
    ort_v3    aabb_max;
} OrtMesh;

typedef struct OrtTriangle { uint32_t i_0, i_1, i_2; OrtMesh *mesh; } OrtTriangle;

typedef struct OrtWorld
{
    ort_v3            ambient;                      /* set, never read (macos_main.mm:335) */
    uint32_t          total_tile_to_render_count;
    volatile uint32_t rendered_tile_count;
    OrtMaterial      *materials;                    /* index 0 == "miss" */
    uint32_t          mat_count;
    OrtTempMemory     light_push_buffer;            /* packed (u32 ShapeType, void*) pairs, 12 B stride */
    uint32_t          light_count;
} OrtWorld;

typedef enum OrtShapeType
{
    ORT_SHAPE_NULL = 0,
    ORT_SHAPE_SPHERE = 1,
    ORT_SHAPE_CYLINDER = 2,
    ORT_SHAPE_AAB = 3,
    ORT_SHAPE_MESH = 4,
    ORT_SHAPE_TRIANGLE = 5,
    ORT_SHAPE_CSG = 6
} OrtShapeType;

typedef struct OrtBVHOctreeNode
{
    uint8_t                  child_bitmask;
    struct OrtBVHOctreeNode *first_child;   /* 8 children laid out contiguously */
    int32_t                  is_leaf;
    OrtTempMemory            push_buffer;   /* records: u32 ShapeType + payload struct */
    ort_v3                   aabb_min;
    ort_v3                   aabb_max;
} OrtBVHOctreeNode;

typedef struct OrtCSG
{
    OrtSphere sphere;
    OrtAAB    aab;
    ort_v3    aabb_min;
    ort_v3    aabb_max;
    uint32_t  mat_index;
} OrtCSG;

typedef struct OrtRandomSeries { uint32_t next_random; } OrtRandomSeries;

#ifdef __cplusplus
}
#define ORT_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define ORT_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

ORT_STATIC_ASSERT(sizeof(OrtTempMemory) == 32, "TempMemory layout");
ORT_STATIC_ASSERT(sizeof(OrtSphere) == 20, "Sphere layout");
ORT_STATIC_ASSERT(sizeof(OrtAAB) == 28, "AAB layout");
ORT_STATIC_ASSERT(sizeof(OrtCylinder) == 32, "Cylinder layout");
ORT_STATIC_ASSERT(sizeof(OrtMaterial) == 60, "Material layout");
ORT_STATIC_ASSERT(sizeof(OrtCamera) == 48, "Camera layout");
ORT_STATIC_ASSERT(sizeof(OrtMesh) == 56, "Mesh layout");
ORT_STATIC_ASSERT(sizeof(OrtTriangle) == 24, "Triangle layout");
ORT_STATIC_ASSERT(sizeof(OrtWorld) == 80, "World layout");
ORT_STATIC_ASSERT(sizeof(OrtBVHOctreeNode) == 80, "BVHOctreeNode layout");
ORT_STATIC_ASSERT(sizeof(OrtCSG) == 76, "CSG layout");
ORT_STATIC_ASSERT(offsetof(OrtBVHOctreeNode, first_child) == 8, "BVHOctreeNode.first_child");
ORT_STATIC_ASSERT(offsetof(OrtBVHOctreeNode, push_buffer) == 24, "BVHOctreeNode.push_buffer");
ORT_STATIC_ASSERT(offsetof(OrtBVHOctreeNode, aabb_min) == 56, "BVHOctreeNode.aabb_min");
ORT_STATIC_ASSERT(offsetof(OrtWorld, materials) == 24, "World.materials");
ORT_STATIC_ASSERT(offsetof(OrtWorld, light_push_buffer) == 40, "World.light_push_buffer");
ORT_STATIC_ASSERT(offsetof(OrtWorld, light_count) == 72, "World.light_count");

#endif /* ORT_SCENE_H */
