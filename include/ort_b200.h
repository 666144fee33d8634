/* include/ort_b200.h -- C ABI of libort_b200.so, the B200 (sm_100a) replacement
 * for the hot path of gyuhyun-lee/offline_raytracer: the per-pixel radiance loop
 * `tiled_raytrace_bvh` (code/ray.cpp:1178-1466) and everything it calls.
 *
 * Plain pointers and sizes only -- no C++/torch/CUDA types -- so the library can
 * be bound from the reference's C++ driver, from ctypes, or from any FFI.
 * Scene inputs use the reference's own struct layouts (ort_scene.h).
 *
 * Every entry point cites the reference interface it replaces.  All functions
 * return ORT_OK (0) or a negative error code and never abort; the message of
 * the last error on the calling thread is available from ort_last_error().
 * (The reference itself reports nothing: failures are assert() null-derefs,
 * code/platform.h:16-20.)
 *
 * There is NO CPU fallback: every render / raycast entry point fails with
 * ORT_ERR_CUDA when no CUDA device is usable.
 *
 * THREADS.  An OrtScene may be used from any number of host threads: every entry
 * point that touches the handle's scratch (framebuffers, counters, path pools,
 * streams) takes the handle's lock for the whole call, so concurrent calls on one
 * handle are serialised -- which is what the GPU would do with them anyway.  This
 * is what lets ort_tiled_raytrace_bvh sit inside the reference's thread callback,
 * which nine workers run at once (code/macos_main.mm:150-163, 574-598); tiles of
 * one image rendered that way are bit-identical to one whole-image call.  Handles
 * on different devices run concurrently (one handle per GPU).  ort_last_error()
 * is per thread.
 */
#ifndef ORT_B200_H
#define ORT_B200_H

#include "ort_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORT_OK              0
#define ORT_ERR_ARG        -1   /* bad argument */
#define ORT_ERR_CUDA       -2   /* CUDA runtime error / no device */
#define ORT_ERR_IO         -3   /* file could not be read / written */
#define ORT_ERR_PARSE      -4   /* malformed .scn / .ply / .obj */
#define ORT_ERR_LIMIT      -5   /* a scene limit was exceeded */

#define ORT_MISS_RANK 0xFFFFFFFFu

/* which render kernel family runs the path */
#define ORT_KERNEL_DEFAULT    0
#define ORT_KERNEL_MEGAKERNEL 1   /* persistent threads, path regeneration, state in registers */
#define ORT_KERNEL_WAVEFRONT  2   /* generate / extend / shade / accumulate kernels over path queues */

/* fixed-point scale of the multi-chunk accumulation buffer (see OrtRenderParams.chunk_spp) */
#define ORT_ACCUM_FRAC_BITS 24
#define ORT_ACCUM_SAT_BITS  28   /* a chunk sum saturates at +-2^28 before conversion */

typedef struct OrtScene OrtScene;         /* flattened, device-resident scene */
typedef struct OrtHostScene OrtHostScene; /* host scene built by this library's own loaders */

const char *ort_last_error(void);
/* 0 when a usable CUDA device exists, else ORT_ERR_CUDA. */
int ort_device_count(int *count);
/* Measures the FP32-pipe roofline denominator of `device`: TFLOP/s of a dependent-chain
 * FMUL+FADD kernel (no FMA contraction -- the instruction mix of the intersectors).
 * MEASURED_PEAKS.json carries no FP32 figure.  The third argument is unused. */
int ort_measure_fp32_peak(int device, float *tflops_non_fma, float *reserved);
/* Self-test of the device's shared-reciprocal vector division (csrc/core_math.h, div3_shared: the
 * three quotients of math.h:235 `v3 / f32` as used by `normalize`, math.h:299) against the
 * compiler's IEEE-754 division on `triples` random operand sets; *mismatches must come back 0. */
int ort_selftest_div3(int device, uint64_t triples, uint32_t seed, uint64_t *tested, uint64_t *mismatches);
/* Measures the L2 roofline denominator of `device` ("the build must measure both"): GB/s of
 * a kernel in which every block streams the same `buffer_mib` MiB buffer (0 = 32; it must stay inside the
 * 126 MB L2) with L1-bypassing 128-bit loads -- the way a wide node that misses L1 is served. */
int ort_measure_l2_bandwidth(int device, uint32_t buffer_mib, float *gb_per_s);

/* ------------------------------------------------------------------------
 * Device differential harness (tests).  Runs the DEVICE build of the intersectors
 * (csrc/core_math.h; code/ray.cpp:63-352) and of the BSDF (csrc/path.h;
 * code/ray.cpp:825-1161) -- the very functions the render kernels inline -- on
 * explicit inputs, so that the reference's golden vectors are checked on the GPU.
 *   kind 0 triangle: v0 v1 v2 o d (15 floats per case)     1 sphere: c r o d (10)
 *        2 box: min max o d (12)                           3 cylinder: base axis r o d (13)
 *   out: 5 floats per case = t (-1 = miss), unnormalised normal xyz, inner-hit flag
 *        (IntersectionTestResult, ray.cpp:54-59).
 * ort_selftest_bsdf: materials as 10 floats (Kd, Ks, Kt, ior); per case sample_brdf
 * from `state` (-> wi, is_transmission, state after), pdf_brdf and eval_scattering
 * of (N, wi, wo).  All pointers are HOST pointers.
 * ort_selftest_rng: the xorshift of code/random.h:5-117 on the device -- per seed 64 successive states, then (each
 * restarting from the seed) 8 x random_between_0_1, 8 x random_between(0, between_hi), 8 x
 * random_between_u32(0, u32_one_past_max).  ort_selftest_light_pick: n successive sample_random_lights calls
 * (code/ray.cpp:537-601; one byte per light-list entry: is it a sphere) -- the state after each.  All integer: exact.
 * ---------------------------------------------------------------------- */
int ort_selftest_rng(int device, uint32_t n, const uint32_t *seeds, float between_hi, uint32_t u32_one_past_max,
                     uint32_t *states64, float *f01_8, float *between_8, uint32_t *u32_8);
int ort_selftest_light_pick(int device, uint32_t light_count, const uint8_t *light_is_sphere, uint32_t state, uint32_t n,
                            uint32_t *states_out);
int ort_selftest_intersect(int device, uint32_t kind, uint32_t n, const float *cases, float *out);
int ort_selftest_bsdf(int device, uint32_t n, const float *mat10, const float *N, const float *wo, const float *wi,
                      const uint32_t *state, const float *dist, float roughness,
                      float *sample_wi, int32_t *is_transmission, uint32_t *state_after, float *pdf, float *eval);

/* ------------------------------------------------------------------------
 * Scene hand-off.
 * Replaces: the (World*, BVHOctreeNode* top_most_node) arguments of
 * tiled_raytrace_bvh (code/ray.cpp:1179) as assembled by main()
 * (code/macos_main.mm:334-545).  All inputs are borrowed, read-only host
 * pointers; nothing is retained after the call returns and nothing is freed.
 * The octree is walked in the reference's breadth-first order to give every
 * leaf record its tie-break RANK (the order in which raycast_bvh,
 * code/ray.cpp:624-822, would test it with no culling); the records are then
 * re-organised into a wide BVH with quantised child boxes and float4 triangles
 * and uploaded to `device`.
 * ---------------------------------------------------------------------- */
typedef struct OrtSceneInfo
{
    uint32_t triangle_count;
    uint32_t sphere_count, box_count, cylinder_count, csg_count;  /* csg records are inert (ray.cpp:718-767) and dropped */
    uint32_t record_count;          /* = number of ranks */
    uint32_t octree_node_count;     /* nodes reachable from top_most_node */
    uint32_t octree_max_depth;
    uint32_t material_count, light_count;
    uint32_t bvh_node_count;        /* wide nodes of the flattened BVH */
    uint32_t bvh_node_bytes;        /* bytes per wide node */
    uint64_t device_bytes;          /* total device memory held by the handle */
    float    root_min[3], root_max[3];
} OrtSceneInfo;

int ort_scene_create(const OrtWorld *world, const OrtBVHOctreeNode *top_most_node,
                     int device, OrtScene **scene_out);

/* The same, choosing where the acceleration structure is built.
 * ORT_BUILD_ON_DEVICE: Morton sort, PLOC clustering and the collapse to the 8-wide layout run as
 * CUDA kernels (csrc/bvh_build.h/.cuh) instead of the host's binned-SAH builder -- the replacement
 * for the minutes the reference spends in push_shape_inside_node / validate_nodes_and_reallocate_shapes
 * (code/ray.cpp:1799-2045) on large scenes.  Hits are bit-identical either way: the structure only
 * decides which records are tested.  ort_scene_create = ORT_BUILD_AUTO unless the environment says
 * ORT_BVH_BUILD=host or =device. */
#define ORT_BUILD_ON_DEVICE 1u
#define ORT_BUILD_AUTO 2u               /* on the device when the scene has ORT_BUILD_AUTO_RECORDS records or more */
#define ORT_BUILD_AUTO_RECORDS 1000000u
int ort_scene_create_ex(const OrtWorld *world, const OrtBVHOctreeNode *top_most_node,
                        int device, uint32_t flags, OrtScene **scene_out);

/* Scene hand-off WITHOUT the octree.  The reference builds its octree only to have
 * something to traverse; this library rebuilds the acceleration structure anyway and needs from the
 * octree nothing but each record's RANK -- its position in the order raycast_bvh would test records
 * with no culling (breadth-first over nodes, push-buffer order within a node), which breaks exact-t
 * ties.  That order is a pure function of the shape lists: push_shape_inside_node (ray.cpp:1799-1948)
 * sends a shape down the octant of its box centre until it is alone in a node or depth 10 is reached,
 * so a record's node is the shortest prefix of its 10-digit octant path that no other record shares,
 * and the rank order is "sort by (depth, path prefix, insertion index)".  ort_scene_create_from_lists
 * computes exactly that with two radix sorts instead of minutes of pointer-chasing inserts; the
 * resulting scene is identical to ort_scene_create on the octree the reference would have built
 * (tests/test_host_scene.py).  Lists are in the reference's insertion order (macos_main.mm:474-538):
 * triangles of mesh 0, 1, ..., then cylinders, boxes, spheres, then the inert CSG record if any. */
typedef struct OrtShapeLists
{
    const OrtMesh     *meshes;    uint32_t mesh_count;
    const OrtCylinder *cylinders; uint32_t cylinder_count;
    const OrtAAB      *boxes;     uint32_t box_count;
    const OrtSphere   *spheres;   uint32_t sphere_count;
    const OrtCSG      *csg;                 /* NULL, or the one hard-coded record (macos_main.mm:322-332) */
    ort_v3 root_min, root_max;              /* box of the top-most node (macos_main.mm:421-472) */
} OrtShapeLists;
int ort_scene_create_from_lists(const OrtWorld *world, const OrtShapeLists *lists,
                                int device, uint32_t flags, OrtScene **scene_out);

typedef struct OrtBuildStats
{
    uint32_t on_device;          /* 1 = built by the CUDA kernels */
    uint32_t ploc_iterations;    /* clustering rounds (device build) */
    uint32_t wide_depth;         /* levels of the 8-wide tree */
    float    collect_s;          /* host: decoding the octree's push buffers into ranked records */
    float    prepare_s;          /* host: padding boxes, packing records (device build) */
    float    build_s;            /* wall time of the build proper, uploads included */
    float    device_build_ms;    /* CUDA-event time of sort + clustering + collapse (device build) */
    uint32_t reserved;
} OrtBuildStats;
int ort_scene_build_stats(const OrtScene *scene, OrtBuildStats *out);
/* copies the flattened tree back (tests: the CUDA build must equal the host execution of the same
 * algorithm byte for byte); either pointer may be NULL; sizes must match ort_scene_info */
int ort_scene_download(OrtScene *scene, void *nodes, uint64_t nodes_bytes, void *prims, uint64_t prims_bytes);
int ort_scene_destroy(OrtScene *scene);
int ort_scene_info(const OrtScene *scene, OrtSceneInfo *info);
int ort_scene_device(const OrtScene *scene, int *device);

/* ------------------------------------------------------------------------
 * Render entry point, reference signature.
 * Replaces: u64 tiled_raytrace_bvh(World*, Camera*, BVHOctreeNode*, v3 *output_buffer,
 *   i32 output_width, i32 output_height, i32 tile_min_x, i32 tile_min_y,
 *   i32 tile_one_past_max_x, i32 tile_one_past_max_y, RandomSeries *series,
 *   u32 ray_per_pixel_count, f32 russian_roulette_value)   code/ray.cpp:1178-1183
 * as called from thread_callback_tiled_raytrace (code/macos_main.mm:150-163).
 *
 * Same argument meaning and output layout: output_buffer is a HOST v3 array of
 * output_width*output_height pixels, row 0 = bottom of the picture; only the
 * tile's pixels are written, each with the mean radiance over
 * ray_per_pixel_count samples.  *test_shape_count receives the number of
 * primitive (shape) tests executed -- the reference's return value, summed as at
 * ray.cpp:661-715, 1173; the tally is far smaller than the reference's for the
 * same image because the rebuilt acceleration structure culls better.  It may be
 * called from several host threads at once on disjoint tiles (see THREADS above).
 *
 * RNG: the reference shares one sequential RandomSeries between all pixels of
 * the tile, which cannot be parallelised.  Here the series' current state is
 * the BASE SEED of independent per-pixel streams, stream(x, y) starting at
 * ort_stream_seed(series->next_random, y*output_width + x, 0), each consumed in
 * the reference's draw order; *series is then advanced by one
 * xorshift step.  The reference reproduces any such seeding exactly by calling
 * the unmodified tiled_raytrace_bvh on 1x1 tiles.
 * ---------------------------------------------------------------------- */
int ort_tiled_raytrace_bvh(OrtScene *scene, const OrtCamera *camera,
                           ort_v3 *output_buffer, int32_t output_width, int32_t output_height,
                           int32_t tile_min_x, int32_t tile_min_y,
                           int32_t tile_one_past_max_x, int32_t tile_one_past_max_y,
                           OrtRandomSeries *series, uint32_t ray_per_pixel_count,
                           float russian_roulette_value, uint64_t *test_shape_count);

/* ------------------------------------------------------------------------
 * Render entry point, full form.
 * The hot-loop literals of the reference are fields with the reference's
 * values as defaults (ort_render_params_default).
 * ---------------------------------------------------------------------- */
typedef struct OrtRenderParams
{
    int32_t  output_width, output_height;
    int32_t  tile_min_x, tile_min_y, tile_one_past_max_x, tile_one_past_max_y;
    uint32_t ray_per_pixel_count;      /* total samples per pixel of the image */
    float    russian_roulette_value;   /* 0.8 in the reference driver, macos_main.mm:656 */
    uint32_t base_seed;
    /* Sample streams.  The samples of one pixel are cut into chunks of chunk_spp
     * samples; chunk c of pixel p is an independent xorshift stream seeded with
     * ort_stream_seed(base_seed, p, c) and consumed sequentially.  chunk_spp = 0
     * means one chunk per pixel (= ray_per_pixel_count): then the pixel is the
     * float sum in sample order divided by spp, exactly the reference's
     * accumulate (ray.cpp:1257,1364,1428).  With several chunks per pixel the
     * chunk sums are combined in 64-bit fixed point (ORT_ACCUM_FRAC_BITS
     * fractional bits, round-to-nearest-even, each chunk sum saturating at
     * +-2^ORT_ACCUM_SAT_BITS), which is
     * order-independent, so the image is bit-identical for any scheduling and
     * any number of GPUs. */
    uint32_t chunk_spp;
    uint32_t chunk_begin, chunk_end;   /* sub-range of chunks rendered by this call; 0,0 = all */
    uint32_t kernel;                   /* ORT_KERNEL_* */
    float    roughness;                /* 0.01f   ray.cpp:1194 */
    float    dont_get_too_close_epsilon; /* 0.0001f ray.cpp:1196 */
    float    aperture_radius;          /* 0.1f    ray.cpp:1199 */
    float    lens_z_offset;            /* 0.1f    ray.cpp:1234 */
    float    focus_target[3];          /* (0,0,0.2) ray.cpp:1198 */
} OrtRenderParams;

typedef struct OrtRenderStats
{
    uint64_t samples;          /* pixel samples traced */
    uint64_t rays;             /* extend calls: primary + bounce rays */
    uint64_t node_visits;      /* wide-node visits       (counters build only, else 0) */
    uint64_t box_tests;        /* child-box slab tests   (counters build only, else 0) */
    uint64_t shape_tests;      /* primitive tests        (counters build only, else 0) */
    float    device_ms;        /* CUDA-event time of the kernels of this call */
    uint32_t kernel_launches;  /* kernels launched by this call */
    /* per-stage times: only when the environment has ORT_WF_TIMING=1 (the event records between the kernels cost
     * 2-3 % of the frame), else 0 */
    float    extend_ms;        /* wavefront: summed CUDA-event time of the EXTEND launches */
    float    shade_ms;         /* wavefront: ... of the SHADE launches */
    float    sort_ms;          /* wavefront: ... of the key scatter launches (the scan runs inside EXTEND) */
    uint32_t extend_launches;  /* wavefront: number of EXTEND launches of this call */
} OrtRenderStats;

void ort_render_params_default(OrtRenderParams *params, int32_t width, int32_t height,
                               uint32_t ray_per_pixel_count);

/* host output (v3, row 0 = bottom); host<->device copies inside the call */
int ort_render(OrtScene *scene, const OrtCamera *camera, const OrtRenderParams *params,
               ort_v3 *output_buffer, OrtRenderStats *stats);

/* Device-resident form for the multi-GPU path: adds the chunk sums of
 * [chunk_begin, chunk_end) into `accum_device`, an int64[height*width*4]
 * fixed-point buffer on the scene's device (zero it first with
 * ort_accum_zero_device).  The work is ordered after whatever the caller queued on
 * CUDA stream `stream` (a cudaStream_t passed as void*, NULL = default stream); the
 * call returns when the chunk range is done, whichever kernel family ran it (it
 * synchronises `stream`), so the caller's next launch on any stream sees the sums.  After the sum over ranks (ncclSum on int64 is
 * exact), ort_accum_resolve_device writes float3 pixels = sum / spp. */
int ort_render_accumulate_device(OrtScene *scene, const OrtCamera *camera,
                                 const OrtRenderParams *params, void *accum_device,
                                 void *stream, OrtRenderStats *stats);
int ort_accum_zero_device(OrtScene *scene, void *accum_device, int32_t width, int32_t height, void *stream);
int ort_accum_resolve_device(OrtScene *scene, const void *accum_device, int32_t width, int32_t height,
                             uint32_t ray_per_pixel_count, void *rgb_device, void *stream);

/* ------------------------------------------------------------------------
 * Ray-cast entry point, batched.
 * Replaces: RaycastBVHResult raycast_top_most_node(BVHQueue*, BVHOctreeNode*,
 *   u64 *test_shape_count, v3 ray_origin, v3 ray_dir)   code/ray.cpp:1165-1176
 * (result struct code/ray.cpp:613-622) for n rays at once.  origins/dirs are
 * xyz triples; dirs need not be unit (t is in units of |dir|).  Outputs (any
 * may be NULL): hit_t (FLT_MAX on a miss, as ray.cpp:627), prim_rank (the
 * winning record's rank, ORT_MISS_RANK on a miss -- the one addition to the
 * reference's result), mat_index (0 on a miss), hit_normal (normalised xyz, as
 * ray.cpp:817).  Closest hit with the reference's acceptance rule
 * t >= 1e-6 && t < best (ray.cpp:653,670,686,708); exact-t ties go to the
 * lowest rank, i.e. to the record the reference would have tested first.
 * ---------------------------------------------------------------------- */
int ort_raycast_batch(OrtScene *scene, uint64_t n, const float *origins, const float *dirs,
                      float *hit_t, uint32_t *prim_rank, uint32_t *mat_index, float *hit_normal,
                      OrtRenderStats *stats);
/* same with device pointers on the scene's device; asynchronous on `stream` */
int ort_raycast_batch_device(OrtScene *scene, uint64_t n, const float *origins, const float *dirs,
                             float *hit_t, uint32_t *prim_rank, uint32_t *mat_index, float *hit_normal,
                             void *stream);
/* exhaustive closest hit over ALL records (no acceleration structure) with the
 * same intersectors: the structure-free definition of the right answer
 * (SURVEY.md 8c), used to validate the BVH at sizes the CPU oracle cannot reach. */
int ort_raycast_brute_device(OrtScene *scene, uint64_t n, const float *origins, const float *dirs,
                             float *hit_t, uint32_t *prim_rank, uint32_t *mat_index, void *stream);
/* per-ray work counters of the traversal kernel on the given rays (device
 * pointers): sums over the n rays of wide-node visits, child-box tests,
 * primitive tests. */
int ort_raycast_counters_device(OrtScene *scene, uint64_t n, const float *origins, const float *dirs,
                                uint64_t *node_visits, uint64_t *box_tests, uint64_t *shape_tests);

/* Explicit ray buffers of the traversal microbenchmark (BASELINE config 2), generated on `device`
 * into device arrays of n xyz triples; ray i draws from its own xorshift stream
 * ort_stream_seed(seed, i, 0) (code/random.h:5-38).
 *   camera rays: ray i is a sample of pixel (i % w, i / w) of the params' w x h grid from the
 *     reference camera model with its depth-of-field lens sample (code/ray.cpp:1215-1246) -- the
 *     render kernels' own generate code; n <= w*h.
 *   random rays: origin uniform in [box_min, box_max], direction uniform on the sphere. */
int ort_generate_camera_rays_device(int device, const OrtCamera *camera, const OrtRenderParams *params, uint32_t seed,
                                    uint64_t n, float *origins_device, float *dirs_device, void *stream);
int ort_generate_random_rays_device(int device, const float box_min[3], const float box_max[3], uint32_t seed,
                                    uint64_t n, float *origins_device, float *dirs_device, void *stream);

/* ------------------------------------------------------------------------
 * Multi-GPU.  The path shards with no data-path dependency: the
 * scene is replicated, GPU r renders its own range of sample chunks of the same
 * frame into its own int64 fixed-point framebuffer, and the framebuffers are
 * summed.  The sum is ONE kernel on the root GPU that reads the peers'
 * framebuffers in place over NVLink peer memory, adds them into the root's and
 * resolves in the same pass (float3 pixels and/or RGBE words) -- integer addition,
 * hence bit-identical to the 1-GPU image for any number of GPUs.
 * Replaces the reference's nine tile workers on one shared-memory machine
 * (code/macos_main.mm:165-240, 574-671).
 *
 * (a) one process per GPU: every process allocates its framebuffer with
 *     ort_accum_alloc_device, exports a CUDA IPC handle (ort_accum_ipc_export), the
 *     host-side plumbing moves the 64-byte handles to the root, which opens them once
 *     (ort_accum_ipc_open) and calls ort_accum_reduce_resolve_device after the ranks
 *     have finished their ort_render_accumulate_device (a barrier of the plumbing).
 * (b) one process, N devices: OrtMulti below.
 * ---------------------------------------------------------------------- */
#define ORT_IPC_HANDLE_BYTES 64
int ort_accum_alloc_device(OrtScene *scene, int32_t width, int32_t height, void **accum_device);   /* zeroed */
int ort_accum_free_device(OrtScene *scene, void *accum_device);
int ort_accum_ipc_export(OrtScene *scene, const void *accum_device, uint8_t handle[ORT_IPC_HANDLE_BYTES]);
int ort_accum_ipc_open(OrtScene *scene, const uint8_t handle[ORT_IPC_HANDLE_BYTES], void **peer_accum);
int ort_accum_ipc_close(OrtScene *scene, void *peer_accum);
/* accum_device += sum of the n_peers peer framebuffers (device pointers readable from the scene's
 * device: peer access or opened IPC handles); rgb_device / rgbe_device, when not NULL, receive the
 * resolved image (sum / spp as float3, row 0 = bottom; RGBE words in .hdr file order).  Asynchronous
 * on `stream`.  n_peers = 0 is a plain resolve. */
int ort_accum_reduce_resolve_device(OrtScene *scene, void *accum_device, const void *const *peer_accums, uint32_t n_peers,
                                    int32_t width, int32_t height, uint32_t ray_per_pixel_count,
                                    void *rgb_device, void *rgbe_device, void *stream);

/* (b) N devices in one process.  scenes[i] are handles of the SAME scene on distinct devices; scenes[0]'s
 * device is the root.  ort_multi_render = ort_render for a whole image with the sample chunks split over
 * the devices (set params->chunk_spp; with a single chunk only one device has work).  The image is
 * bit-identical to ort_render's on one GPU with the same params whenever the frame has more than one chunk. */
typedef struct OrtMulti OrtMulti;
int ort_multi_create(OrtScene *const *scenes, uint32_t n, OrtMulti **out);
int ort_multi_destroy(OrtMulti *multi);
int ort_multi_device_count(const OrtMulti *multi, uint32_t *n_devices, uint32_t *n_peer_direct);
int ort_multi_render(OrtMulti *multi, const OrtCamera *camera, const OrtRenderParams *params,
                     ort_v3 *output_buffer, OrtRenderStats *stats);

/* Progressive accumulation with dynamic chunk dispatch and checkpoint / resume (the
 * reference's ThreadWorkQueue, code/platform.h:307-339, hands out 32x32 tiles and cannot stop and
 * continue).  The unit of work is a chunk index: ort_progress_render lets every device of the OrtMulti
 * pull the next unrendered chunk from a shared counter until max_chunks more are done (0 = all that
 * remain); ort_progress_resolve returns the image of the chunks done so far (sum / samples done);
 * ort_progress_save / ort_progress_load write / read a checkpoint (format: csrc/ort_multi.cu) from which
 * any number of devices continues.  Chunks may be rendered in any order, by any device, in any number of
 * sessions: the final image is bit-identical to ort_render's.  One OrtProgress at a time per OrtMulti. */
typedef struct OrtProgress OrtProgress;
int ort_progress_create(OrtMulti *multi, const OrtCamera *camera, const OrtRenderParams *params, OrtProgress **out);
int ort_progress_destroy(OrtProgress *progress);
int ort_progress_render(OrtProgress *progress, uint32_t max_chunks, OrtRenderStats *stats);
int ort_progress_state(const OrtProgress *progress, uint32_t *chunks_done, uint32_t *chunks_total, uint32_t *spp_done);
int ort_progress_resolve(OrtProgress *progress, ort_v3 *output_buffer);
int ort_progress_save(OrtProgress *progress, const char *path);
int ort_progress_load(OrtMulti *multi, const char *path, OrtProgress **out);

/* ------------------------------------------------------------------------
 * Host side of the drop-in: this library's own re-implementation of the
 * reference's loaders and scene assembly, producing the reference's structs.
 * Replaces: parse_scene (code/parser.cpp:1184-1446), parse_ply_header/parse_ply
 * (:384-570), pre_parse_obj/parse_obj (:687-982) and the scene assembly of
 * main() (code/macos_main.mm:310-562: mesh bake, octree insertion with
 * push_shape_inside_node code/ray.cpp:1799-1948, compaction, camera axes).
 * with_csg != 0 inserts the reference's hard-coded (inert) CSG record
 * (macos_main.mm:322-332) so that record ranks equal the reference's.
 * ---------------------------------------------------------------------- */
int ort_host_scene_load(const char *scn_path, const char *base_dir, int32_t width, int32_t height,
                        int with_csg, OrtHostScene **out);
int ort_host_scene_destroy(OrtHostScene *hs);
const OrtWorld *ort_host_scene_world(const OrtHostScene *hs);
const OrtCamera *ort_host_scene_camera(const OrtHostScene *hs);
const OrtBVHOctreeNode *ort_host_scene_root(const OrtHostScene *hs);
const OrtMesh *ort_host_scene_meshes(const OrtHostScene *hs, uint32_t *mesh_count);
/* the shape lists of the loaded scene, for ort_scene_create_from_lists; with_csg & ORT_HOST_NO_OCTREE
 * makes ort_host_scene_load skip the octree (ort_host_scene_root then returns NULL) */
#define ORT_HOST_NO_OCTREE 2
int ort_host_scene_lists(const OrtHostScene *hs, OrtShapeLists *out);
/* Mesh bake on the device.  Replaces the per-vertex loop of main() (code/macos_main.mm:382-413):
 * v *= scale; rotate by `degree` about (0, 1, 0); rotate by `quaternion`; v += translate -- and the mesh AABB folded
 * from (FLT_MAX, FLT_MIN) as the reference does (FLT_MIN = smallest positive float: a mesh in negative space keeps
 * a max of 1.2e-38).  Host pointers; in and out may be the same array.  Baked vertices and the box are bit-identical
 * to the host bake (tests/test_gpu_build.py).  ort_host_scene_load uses it when with_csg has ORT_HOST_BAKE_ON_DEVICE. */
int ort_bake_mesh(int device, uint32_t vertex_count, const ort_v3 *vertices_in, ort_v3 *vertices_out,
                  float scale, float degree, ort_v4 quaternion, ort_v3 translate, ort_v3 *aabb_min, ort_v3 *aabb_max);
#define ORT_HOST_BAKE_ON_DEVICE 4
/* mesh loader alone; vertices (xyz floats) and indices are malloc'd, free with ort_free */
int ort_load_mesh(const char *path, float **vertices, uint32_t *vertex_count,
                  uint32_t **indices, uint32_t *index_count);
/* the reference's number tokenizer (eat_numeric, code/parser.cpp:158-250), which is
 * NOT strtof-equivalent; returns 1 for a float token, 0 for an integer token */
int ort_parse_numeric(const char *text, uint32_t *value_bits);
void ort_free(void *p);

/* Radiance .hdr writer: flat (non-RLE) RGBE, header "+Y h +X w", buffer rows
 * emitted h-1 -> 0.  Replaces v3_to_rgbe / write_hdr_header and the output loop
 * of main() (code/macos_main.mm:242-287, 682-707). */
int ort_write_hdr(const char *path, const ort_v3 *pixels, int32_t width, int32_t height);
uint32_t ort_v3_to_rgbe(ort_v3 color);

/* The same encoding on the device.  RGBE words come out in FILE order -- output
 * row r is buffer row height-1-r, as the reference's writer walks the buffer (macos_main.mm:686-705)
 * -- so the file is the header followed by the words as they are:
 *   ort_rgbe_encode_device         float3 pixels on the device -> uint32 RGBE[height*width]
 *   ort_accum_resolve_rgbe_device  the int64 fixed-point sums (e.g. straight out of the NCCL reduce)
 *                                  -> RGBE, resolve (ray.cpp:1428) and encode fused in one pass
 *   ort_render_rgbe                ort_render for a whole image, returning RGBE words: 4 bytes per
 *                                  pixel cross PCIe instead of 12
 *   ort_render_hdr                 the same, written to `path`: byte-identical to ort_render
 *                                  followed by ort_write_hdr
 *   ort_write_hdr_rgbe             header + words (host) */
int ort_rgbe_encode_device(OrtScene *scene, const void *rgb_device, int32_t width, int32_t height,
                           void *rgbe_device, void *stream);
int ort_accum_resolve_rgbe_device(OrtScene *scene, const void *accum_device, int32_t width, int32_t height,
                                  uint32_t ray_per_pixel_count, void *rgbe_device, void *stream);
int ort_render_rgbe(OrtScene *scene, const OrtCamera *camera, const OrtRenderParams *params,
                    uint32_t *rgbe_out, OrtRenderStats *stats);
int ort_render_hdr(OrtScene *scene, const OrtCamera *camera, const OrtRenderParams *params,
                   const char *path, OrtRenderStats *stats);
int ort_write_hdr_rgbe(const char *path, const uint32_t *rgbe_words, int32_t width, int32_t height);

/* ------------------------------------------------------------------------
 * Seed of the sample stream of (pixel, chunk).  Part of the boundary's
 * contract: the CPU oracle uses the same function.  Never returns 0 (0 is the
 * absorbing state of the reference's xorshift, code/random.h:5-16).
 * ---------------------------------------------------------------------- */
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline uint32_t ort_stream_seed(uint32_t base, uint32_t pixel_index, uint32_t chunk)
{
    uint32_t h = base ^ (pixel_index * 0x9E3779B1u) ^ (chunk * 0x85EBCA77u);
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    if(h == 0) h = 0x6D2B79F5u;
    return h;
}

#ifdef __cplusplus
}
#endif

#endif /* ORT_B200_H */
