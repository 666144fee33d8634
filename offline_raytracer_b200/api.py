"""ctypes binding of libort_b200.so -- the C ABI of include/ort_b200.h.

Host-side mirror of the reference's render interface (code/ray.cpp):

    reference                                        here
    ---------------------------------------------    -----------------------------------------
    main() scene assembly (macos_main.mm:310-562)    HostScene.load(scn_path, base_dir, w, h)
    World* / BVHOctreeNode* hand-off                 Scene(world_ptr, root_ptr, device)
    tiled_raytrace_bvh(...)        ray.cpp:1178      Scene.tiled_raytrace_bvh(...)
    raycast_top_most_node(...)     ray.cpp:1165      Scene.raycast_batch(origins, dirs)

There is no CPU fallback: if the CUDA library is missing or no device is usable
every compute call raises OrtError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libort_b200.so")

ORT_KERNEL_DEFAULT, ORT_KERNEL_MEGAKERNEL, ORT_KERNEL_WAVEFRONT = 0, 1, 2
MISS_RANK = 0xFFFFFFFF
ACCUM_FRAC_BITS = 24

c_f, c_u32, c_i32, c_u64, vp = C.c_float, C.c_uint32, C.c_int32, C.c_uint64, C.c_void_p


class OrtError(RuntimeError):
    pass


class V3(C.Structure):
    _fields_ = [("x", c_f), ("y", c_f), ("z", c_f)]


class Camera(C.Structure):
    """OrtCamera == reference Camera (code/ray.h:42-49)"""
    _fields_ = [("p", c_f * 3), ("x_axis", c_f * 3), ("y_axis", c_f * 3), ("z_axis", c_f * 3)]


class RenderParams(C.Structure):
    """OrtRenderParams (include/ort_b200.h)"""
    _fields_ = [
        ("output_width", c_i32), ("output_height", c_i32),
        ("tile_min_x", c_i32), ("tile_min_y", c_i32),
        ("tile_one_past_max_x", c_i32), ("tile_one_past_max_y", c_i32),
        ("ray_per_pixel_count", c_u32), ("russian_roulette_value", c_f),
        ("base_seed", c_u32), ("chunk_spp", c_u32),
        ("chunk_begin", c_u32), ("chunk_end", c_u32), ("kernel", c_u32),
        ("roughness", c_f), ("dont_get_too_close_epsilon", c_f),
        ("aperture_radius", c_f), ("lens_z_offset", c_f),
        ("focus_target", c_f * 3),
    ]


class RenderStats(C.Structure):
    _fields_ = [("samples", c_u64), ("rays", c_u64), ("node_visits", c_u64), ("box_tests", c_u64),
                ("shape_tests", c_u64), ("device_ms", c_f), ("kernel_launches", c_u32),
                ("extend_ms", c_f), ("shade_ms", c_f), ("sort_ms", c_f), ("extend_launches", c_u32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SceneInfo(C.Structure):
    _fields_ = [("triangle_count", c_u32), ("sphere_count", c_u32), ("box_count", c_u32),
                ("cylinder_count", c_u32), ("csg_count", c_u32), ("record_count", c_u32),
                ("octree_node_count", c_u32), ("octree_max_depth", c_u32),
                ("material_count", c_u32), ("light_count", c_u32),
                ("bvh_node_count", c_u32), ("bvh_node_bytes", c_u32), ("device_bytes", c_u64),
                ("root_min", c_f * 3), ("root_max", c_f * 3)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_[:13]}
        d["root_min"] = list(self.root_min); d["root_max"] = list(self.root_max)
        return d


class ShapeLists(C.Structure):
    _fields_ = [("meshes", vp), ("mesh_count", c_u32), ("cylinders", vp), ("cylinder_count", c_u32),
                ("boxes", vp), ("box_count", c_u32), ("spheres", vp), ("sphere_count", c_u32),
                ("csg", vp), ("root_min", V3), ("root_max", V3)]


ORT_HOST_NO_OCTREE = 2
ORT_HOST_BAKE_ON_DEVICE = 4


class V4(C.Structure):
    _fields_ = [("x", c_f), ("y", c_f), ("z", c_f), ("w", c_f)]


class Mesh(C.Structure):
    """OrtMesh == reference Mesh (code/ray.h:51-65)"""
    _fields_ = [("vertices", vp), ("vertex_count", c_u32), ("indices", vp), ("index_count", c_u32), ("mat_index", c_u32),
                ("aabb_min", c_f * 3), ("aabb_max", c_f * 3)]


class BuildStats(C.Structure):
    _fields_ = [("on_device", c_u32), ("ploc_iterations", c_u32), ("wide_depth", c_u32), ("collect_s", c_f),
                ("prepare_s", c_f), ("build_s", c_f), ("device_build_ms", c_f), ("reserved", c_u32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_[:7]}


ORT_BUILD_ON_DEVICE = 1
ORT_BUILD_AUTO = 2

_lib = None


def lib(path=None):
    """loads libort_b200.so (once); raises OrtError if it is not built"""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise OrtError("CUDA library %s is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                       "there is no CPU fallback" % p)
    L = C.CDLL(p)
    L.ort_last_error.restype = C.c_char_p
    L.ort_device_count.argtypes = [C.POINTER(C.c_int)]
    L.ort_scene_create.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    L.ort_scene_create_ex.argtypes = [vp, vp, C.c_int, c_u32, C.POINTER(vp)]
    L.ort_scene_create_from_lists.argtypes = [vp, C.POINTER(ShapeLists), C.c_int, c_u32, C.POINTER(vp)]
    L.ort_scene_build_stats.argtypes = [vp, C.POINTER(BuildStats)]
    L.ort_scene_download.argtypes = [vp, vp, c_u64, vp, c_u64]
    L.ort_scene_destroy.argtypes = [vp]
    L.ort_scene_info.argtypes = [vp, C.POINTER(SceneInfo)]
    L.ort_render_params_default.argtypes = [C.POINTER(RenderParams), c_i32, c_i32, c_u32]
    L.ort_render_params_default.restype = None
    L.ort_render.argtypes = [vp, vp, C.POINTER(RenderParams), vp, C.POINTER(RenderStats)]
    L.ort_tiled_raytrace_bvh.argtypes = [vp, vp, vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                                         C.POINTER(c_u32), c_u32, c_f, C.POINTER(c_u64)]
    L.ort_render_accumulate_device.argtypes = [vp, vp, C.POINTER(RenderParams), vp, vp, C.POINTER(RenderStats)]
    L.ort_accum_zero_device.argtypes = [vp, vp, c_i32, c_i32, vp]
    L.ort_accum_resolve_device.argtypes = [vp, vp, c_i32, c_i32, c_u32, vp, vp]
    L.ort_rgbe_encode_device.argtypes = [vp, vp, c_i32, c_i32, vp, vp]
    L.ort_accum_resolve_rgbe_device.argtypes = [vp, vp, c_i32, c_i32, c_u32, vp, vp]
    L.ort_render_rgbe.argtypes = [vp, vp, C.POINTER(RenderParams), vp, C.POINTER(RenderStats)]
    L.ort_render_hdr.argtypes = [vp, vp, C.POINTER(RenderParams), C.c_char_p, C.POINTER(RenderStats)]
    L.ort_write_hdr_rgbe.argtypes = [C.c_char_p, vp, c_i32, c_i32]
    L.ort_raycast_batch.argtypes = [vp, c_u64, vp, vp, vp, vp, vp, vp, C.POINTER(RenderStats)]
    L.ort_raycast_batch_device.argtypes = [vp, c_u64, vp, vp, vp, vp, vp, vp, vp]
    L.ort_raycast_brute_device.argtypes = [vp, c_u64, vp, vp, vp, vp, vp, vp]
    L.ort_raycast_counters_device.argtypes = [vp, c_u64, vp, vp, C.POINTER(c_u64), C.POINTER(c_u64), C.POINTER(c_u64)]
    L.ort_scene_device.argtypes = [vp, C.POINTER(C.c_int)]
    if hasattr(L, "ort_bake_mesh"):
        L.ort_bake_mesh.argtypes = [C.c_int, c_u32, vp, vp, c_f, c_f, V4, V3, C.POINTER(V3), C.POINTER(V3)]
    L.ort_measure_l2_bandwidth.argtypes = [C.c_int, c_u32, C.POINTER(c_f)]
    L.ort_selftest_intersect.argtypes = [C.c_int, c_u32, c_u32, vp, vp]
    if hasattr(L, "ort_selftest_rng"):
        L.ort_selftest_rng.argtypes = [C.c_int, c_u32, vp, c_f, c_u32, vp, vp, vp, vp]
        L.ort_selftest_light_pick.argtypes = [C.c_int, c_u32, vp, c_u32, c_u32, vp]
    L.ort_selftest_bsdf.argtypes = [C.c_int, c_u32, vp, vp, vp, vp, vp, vp, c_f, vp, vp, vp, vp, vp]
    L.ort_generate_camera_rays_device.argtypes = [C.c_int, vp, C.POINTER(RenderParams), c_u32, c_u64, vp, vp, vp]
    L.ort_generate_random_rays_device.argtypes = [C.c_int, C.POINTER(c_f), C.POINTER(c_f), c_u32, c_u64, vp, vp, vp]
    L.ort_accum_alloc_device.argtypes = [vp, c_i32, c_i32, C.POINTER(vp)]
    L.ort_accum_free_device.argtypes = [vp, vp]
    L.ort_accum_ipc_export.argtypes = [vp, vp, vp]
    L.ort_accum_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.ort_accum_ipc_close.argtypes = [vp, vp]
    L.ort_accum_reduce_resolve_device.argtypes = [vp, vp, C.POINTER(vp), c_u32, c_i32, c_i32, c_u32, vp, vp, vp]
    L.ort_multi_create.argtypes = [C.POINTER(vp), c_u32, C.POINTER(vp)]
    L.ort_multi_destroy.argtypes = [vp]
    L.ort_multi_device_count.argtypes = [vp, C.POINTER(c_u32), C.POINTER(c_u32)]
    L.ort_multi_render.argtypes = [vp, vp, C.POINTER(RenderParams), vp, C.POINTER(RenderStats)]
    L.ort_progress_create.argtypes = [vp, vp, C.POINTER(RenderParams), C.POINTER(vp)]
    L.ort_progress_destroy.argtypes = [vp]
    L.ort_progress_render.argtypes = [vp, c_u32, C.POINTER(RenderStats)]
    L.ort_progress_state.argtypes = [vp, C.POINTER(c_u32), C.POINTER(c_u32), C.POINTER(c_u32)]
    L.ort_progress_resolve.argtypes = [vp, vp]
    L.ort_progress_save.argtypes = [vp, C.c_char_p]
    L.ort_progress_load.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    if hasattr(L, "ort_host_scene_load"):
        L.ort_host_scene_load.argtypes = [C.c_char_p, C.c_char_p, c_i32, c_i32, C.c_int, C.POINTER(vp)]
        L.ort_host_scene_destroy.argtypes = [vp]
        for n in ("ort_host_scene_world", "ort_host_scene_camera", "ort_host_scene_root"):
            getattr(L, n).argtypes = [vp]
            getattr(L, n).restype = vp
        L.ort_host_scene_meshes.argtypes = [vp, C.POINTER(c_u32)]
        L.ort_host_scene_meshes.restype = vp
        L.ort_load_mesh.argtypes = [C.c_char_p, C.POINTER(C.POINTER(c_f)), C.POINTER(c_u32),
                                    C.POINTER(C.POINTER(c_u32)), C.POINTER(c_u32)]
        L.ort_parse_numeric.argtypes = [C.c_char_p, C.POINTER(c_u32)]
        L.ort_free.argtypes = [vp]
        L.ort_write_hdr.argtypes = [C.c_char_p, vp, c_i32, c_i32]
        L.ort_v3_to_rgbe.argtypes = [V3]
        L.ort_v3_to_rgbe.restype = c_u32
    if path is None:
        _lib = L
    return L


def _check(rc, L=None):
    if rc != 0:
        L = L or lib()
        raise OrtError("ort error %d: %s" % (rc, (L.ort_last_error() or b"").decode()))


def device_count():
    n = C.c_int(0)
    rc = lib().ort_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def measure_fp32_peak(device=0):
    """TFLOP/s of the FMUL+FADD (non-FMA) chain microbenchmark: the FP32 roofline denominator"""
    L = lib()
    L.ort_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(c_f), C.POINTER(c_f)]
    v = c_f(0)
    _check(L.ort_measure_fp32_peak(device, C.byref(v), None))
    return float(v.value)


def measure_l2_bandwidth(device=0, buffer_mib=32):
    """GB/s of the L2-resident streaming-read microbenchmark: the L2 roofline denominator"""
    v = c_f(0)
    _check(lib().ort_measure_l2_bandwidth(device, buffer_mib, C.byref(v)))
    return float(v.value)


_ISECT_KIND = {"triangle": (0, 15), "sphere": (1, 10), "aab": (2, 12), "cylinder": (3, 13)}


def selftest_intersect(kind, cases, device=0):
    """the DEVICE build of the intersector `kind` on explicit cases [n, stride] -> [n, 5] = t, normal xyz, inner"""
    k, stride = _ISECT_KIND[kind]
    a = np.ascontiguousarray(cases, np.float32).reshape(-1, stride)
    out = np.zeros((a.shape[0], 5), np.float32)
    _check(lib().ort_selftest_intersect(device, k, a.shape[0], _ptr(a), _ptr(out)))
    return out


def selftest_rng(seeds, between_hi=6.2831855, u32_one_past_max=12, device=0):
    """the DEVICE xorshift (code/random.h): (states [n, 64], f01 [n, 8], between [n, 8], u32 [n, 8]) per seed"""
    s = np.ascontiguousarray(seeds, np.uint32).reshape(-1)
    n = s.shape[0]
    st = np.zeros((n, 64), np.uint32); f01 = np.zeros((n, 8), np.float32)
    bt = np.zeros((n, 8), np.float32); u = np.zeros((n, 8), np.uint32)
    _check(lib().ort_selftest_rng(device, n, _ptr(s), c_f(between_hi), u32_one_past_max, _ptr(st), _ptr(f01), _ptr(bt), _ptr(u)))
    return st, f01, bt, u


def selftest_light_pick(light_is_sphere, state, n, device=0):
    """n successive sample_random_lights calls on the DEVICE: the RNG state after each"""
    l = np.ascontiguousarray(light_is_sphere, np.uint8).reshape(-1)
    out = np.zeros(n, np.uint32)
    _check(lib().ort_selftest_light_pick(device, l.shape[0], _ptr(l), state, n, _ptr(out)))
    return out


def world_light_is_sphere(world_ptr):
    """decodes World.light_push_buffer (packed (u32 ShapeType, pointer) pairs, parser.cpp:1144-1182): 1 per sphere entry"""
    base = C.cast(_as_ptr(world_ptr), C.POINTER(C.c_uint8))
    raw = bytes(base[40:80])
    import struct
    _arena, buf, _total, used = struct.unpack("<QQQQ", raw[:32])
    count = struct.unpack("<I", raw[32:36])[0]
    data = C.string_at(buf, used) if buf and used else b""
    out = [1 if struct.unpack_from("<I", data, 12 * k)[0] == 1 else 0 for k in range(len(data) // 12)]
    assert len(out) == count
    return np.array(out, np.uint8)


def selftest_bsdf(mat10, N, wo, wi, state, dist, roughness=0.01, device=0):
    """the DEVICE build of sample_brdf / pdf_brdf / eval_scattering (csrc/path.h) on explicit tuples"""
    f = lambda x, w: np.ascontiguousarray(x, np.float32).reshape(-1, w)
    m, N, wo, wi = f(mat10, 10), f(N, 3), f(wo, 3), f(wi, 3)
    n = m.shape[0]
    st = np.ascontiguousarray(state, np.uint32).reshape(n)
    d = np.ascontiguousarray(dist, np.float32).reshape(n)
    s_wi = np.zeros((n, 3), np.float32); is_t = np.zeros(n, np.int32); st2 = np.zeros(n, np.uint32)
    pdf = np.zeros(n, np.float32); ev = np.zeros((n, 3), np.float32)
    _check(lib().ort_selftest_bsdf(device, n, _ptr(m), _ptr(N), _ptr(wo), _ptr(wi), _ptr(st), _ptr(d), c_f(roughness),
                                   _ptr(s_wi), _ptr(is_t), _ptr(st2), _ptr(pdf), _ptr(ev)))
    return dict(sample_wi=s_wi, is_transmission=is_t, state_after=st2, pdf=pdf, eval=ev)


def generate_camera_rays_device(camera_ptr, params, seed, n, origins_ptr, dirs_ptr, device=0, stream=None):
    """BASELINE config 2, coherent buffer: n lens-sampled primaries of the params' w x h grid (device arrays)"""
    _check(lib().ort_generate_camera_rays_device(device, _as_ptr(camera_ptr), C.byref(params), seed, n,
                                                 vp(origins_ptr), vp(dirs_ptr), vp(stream) if stream else None))


def generate_random_rays_device(box_min, box_max, seed, n, origins_ptr, dirs_ptr, device=0, stream=None):
    """BASELINE config 2, incoherent buffer: origins uniform in the box, directions uniform on the sphere"""
    lo = (c_f * 3)(*[float(x) for x in box_min]); hi = (c_f * 3)(*[float(x) for x in box_max])
    _check(lib().ort_generate_random_rays_device(device, lo, hi, seed, n, vp(origins_ptr), vp(dirs_ptr),
                                                 vp(stream) if stream else None))


def selftest_div3(triples=1 << 28, seed=12345, device=0):
    """(tested, mismatches) of the shared-reciprocal division self-test; mismatches must be 0"""
    L = lib()
    L.ort_selftest_div3.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    t, m = C.c_uint64(0), C.c_uint64(0)
    _check(L.ort_selftest_div3(device, triples, seed, C.byref(t), C.byref(m)))
    return int(t.value), int(m.value)


def default_params(width, height, spp, rr=0.8, seed=1234567, chunk_spp=0, kernel=ORT_KERNEL_DEFAULT):
    p = RenderParams()
    lib().ort_render_params_default(C.byref(p), width, height, spp)
    p.russian_roulette_value = rr
    p.base_seed = seed
    p.chunk_spp = chunk_spp
    p.kernel = kernel
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


def _as_ptr(x):
    """accepts an int address, a ctypes pointer/void_p, or None"""
    if x is None:
        return None
    if isinstance(x, int):
        return vp(x)
    return C.cast(x, vp)


class HostScene:
    """Scene assembled on the host by this library's own loaders (the drop-in for
    main(), code/macos_main.mm:310-562).  Holds the reference-layout World / Camera /
    octree that Scene() consumes."""

    def __init__(self, handle, width, height):
        self.h, self.width, self.height = handle, width, height
        L = lib()
        self.world = L.ort_host_scene_world(handle)
        self.camera = L.ort_host_scene_camera(handle)
        self.root = L.ort_host_scene_root(handle)

    @classmethod
    def load(cls, scn_path, base_dir, width, height, with_csg=True, octree=True, bake_on_device=False):
        """octree=False skips the reference's octree (root is then None): for Scene.from_lists;
        bake_on_device=True runs the mesh bake (macos_main.mm:382-413) as a CUDA kernel (ort_bake_mesh)"""
        L = lib()
        if not base_dir.endswith("/"):
            base_dir += "/"
        h = vp(0)
        flags = (1 if with_csg else 0) | (0 if octree else ORT_HOST_NO_OCTREE) | (ORT_HOST_BAKE_ON_DEVICE if bake_on_device else 0)
        _check(L.ort_host_scene_load(scn_path.encode(), base_dir.encode(), width, height, flags, C.byref(h)))
        return cls(h, width, height)

    def lists(self):
        """the scene's shape lists in the reference's insertion order (OrtShapeLists)"""
        L = lib()
        L.ort_host_scene_lists.argtypes = [vp, C.POINTER(ShapeLists)]
        out = ShapeLists()
        _check(L.ort_host_scene_lists(self.h, C.byref(out)))
        return out

    def camera_array(self):
        return np.ctypeslib.as_array(C.cast(self.camera, C.POINTER(c_f)), shape=(12,)).copy()

    def meshes(self):
        """[(vertices [n, 3] float32 copy, indices copy, aabb_min, aabb_max, mat_index)] of the baked meshes"""
        n = c_u32(0)
        p = lib().ort_host_scene_meshes(self.h, C.byref(n))
        arr = C.cast(p, C.POINTER(Mesh))
        out = []
        for i in range(n.value):
            m = arr[i]
            v = np.ctypeslib.as_array(C.cast(m.vertices, C.POINTER(c_f)), shape=(m.vertex_count * 3,)).copy().reshape(-1, 3)
            idx = np.ctypeslib.as_array(C.cast(m.indices, C.POINTER(c_u32)), shape=(m.index_count,)).copy()
            out.append((v, idx, np.array(list(m.aabb_min), np.float32), np.array(list(m.aabb_max), np.float32), m.mat_index))
        return out

    def close(self):
        if self.h:
            lib().ort_host_scene_destroy(self.h)
            self.h = None


class Scene:
    """Device-resident flattened scene (OrtScene)."""

    def __init__(self, world_ptr, root_ptr, device=0, library=None, build_on_device=None):
        """build_on_device: True = CUDA builder (ort_scene_create_ex, ORT_BUILD_ON_DEVICE), False = host
        binned-SAH builder, None = ort_scene_create (host unless ORT_BVH_BUILD=device)"""
        self.L = library or lib()
        h = vp(0)
        if build_on_device is None:
            _check(self.L.ort_scene_create(_as_ptr(world_ptr), _as_ptr(root_ptr), device, C.byref(h)), self.L)
        else:
            _check(self.L.ort_scene_create_ex(_as_ptr(world_ptr), _as_ptr(root_ptr), device,
                                              ORT_BUILD_ON_DEVICE if build_on_device else 0, C.byref(h)), self.L)
        self.h = h
        self.device = device

    @classmethod
    def from_lists(cls, world_ptr, lists, device=0, build_on_device=None, library=None):
        """ort_scene_create_from_lists: no octree needed; `lists` is a ShapeLists (HostScene.lists()).
        build_on_device: True / False / None = by scene size (ORT_BUILD_AUTO)"""
        self = cls.__new__(cls)
        self.L = library or lib()
        h = vp(0)
        flags = ORT_BUILD_AUTO if build_on_device is None else (ORT_BUILD_ON_DEVICE if build_on_device else 0)
        _check(self.L.ort_scene_create_from_lists(_as_ptr(world_ptr), C.byref(lists), device, flags, C.byref(h)), self.L)
        self.h = h
        self.device = device
        return self

    def build_stats(self):
        b = BuildStats()
        _check(self.L.ort_scene_build_stats(self.h, C.byref(b)), self.L)
        return b.as_dict()

    def download(self):
        """(nodes uint8[n, 80], records uint32[m, 12]) of the flattened tree"""
        i = SceneInfo()
        _check(self.L.ort_scene_info(self.h, C.byref(i)), self.L)
        nodes = np.zeros((i.bvh_node_count, i.bvh_node_bytes), np.uint8)
        kept = i.record_count - i.csg_count
        prims = np.zeros((kept, 12), np.uint32)
        _check(self.L.ort_scene_download(self.h, _ptr(nodes), nodes.nbytes, _ptr(prims), prims.nbytes), self.L)
        return nodes, prims

    def close(self):
        if self.h:
            self.L.ort_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        i = SceneInfo()
        _check(self.L.ort_scene_info(self.h, C.byref(i)), self.L)
        return i.as_dict()

    # -- render ---------------------------------------------------------------
    def render(self, camera_ptr, params, out=None):
        """ort_render: returns (image[H,W,3] float32 with row 0 = bottom, stats dict)"""
        W, H = params.output_width, params.output_height
        if out is None:
            out = np.zeros((H, W, 3), np.float32)
        st = RenderStats()
        _check(self.L.ort_render(self.h, _as_ptr(camera_ptr), C.byref(params), _ptr(out), C.byref(st)), self.L)
        return out, st.as_dict()

    def render_rgbe(self, camera_ptr, params):
        """ort_render_rgbe: (uint32[H,W] RGBE words in .hdr file order -- row 0 = top --, stats dict)"""
        W, H = params.output_width, params.output_height
        out = np.zeros((H, W), np.uint32)
        st = RenderStats()
        _check(self.L.ort_render_rgbe(self.h, _as_ptr(camera_ptr), C.byref(params), _ptr(out), C.byref(st)), self.L)
        return out, st.as_dict()

    def render_hdr(self, camera_ptr, params, path):
        """ort_render_hdr: renders and writes the Radiance .hdr file; returns the stats dict"""
        st = RenderStats()
        _check(self.L.ort_render_hdr(self.h, _as_ptr(camera_ptr), C.byref(params), str(path).encode(), C.byref(st)), self.L)
        return st.as_dict()

    def rgbe_encode_device(self, rgb_ptr, width, height, rgbe_ptr, stream=None):
        _check(self.L.ort_rgbe_encode_device(self.h, vp(rgb_ptr), width, height, vp(rgbe_ptr),
                                             vp(stream) if stream else None), self.L)

    def accum_resolve_rgbe_device(self, accum_ptr, width, height, spp, rgbe_ptr, stream=None):
        _check(self.L.ort_accum_resolve_rgbe_device(self.h, vp(accum_ptr), width, height, spp, vp(rgbe_ptr),
                                                    vp(stream) if stream else None), self.L)

    def tiled_raytrace_bvh(self, camera_ptr, output_buffer, output_width, output_height,
                           tile_min_x, tile_min_y, tile_one_past_max_x, tile_one_past_max_y,
                           series, ray_per_pixel_count, russian_roulette_value):
        """Same argument list as the reference's tiled_raytrace_bvh (code/ray.cpp:1178-1183) minus
        the scene pointers.  `series` is a one-element uint32 array (RandomSeries.next_random),
        advanced in place.  Returns the work counter."""
        st = c_u32(int(series[0]))
        cnt = c_u64(0)
        _check(self.L.ort_tiled_raytrace_bvh(self.h, _as_ptr(camera_ptr), _ptr(output_buffer), output_width, output_height,
                                             tile_min_x, tile_min_y, tile_one_past_max_x, tile_one_past_max_y,
                                             C.byref(st), ray_per_pixel_count, c_f(russian_roulette_value), C.byref(cnt)),
               self.L)
        series[0] = st.value
        return cnt.value

    def render_accumulate_device(self, camera_ptr, params, accum_ptr, stream=None, want_stats=False):
        st = RenderStats()
        _check(self.L.ort_render_accumulate_device(self.h, _as_ptr(camera_ptr), C.byref(params), vp(accum_ptr),
                                                   vp(stream) if stream else None,
                                                   C.byref(st) if want_stats else None), self.L)
        return st.as_dict() if want_stats else None

    def accum_zero_device(self, accum_ptr, width, height, stream=None):
        _check(self.L.ort_accum_zero_device(self.h, vp(accum_ptr), width, height, vp(stream) if stream else None), self.L)

    def accum_resolve_device(self, accum_ptr, width, height, spp, rgb_ptr, stream=None):
        _check(self.L.ort_accum_resolve_device(self.h, vp(accum_ptr), width, height, spp, vp(rgb_ptr),
                                               vp(stream) if stream else None), self.L)

    # -- multi-GPU: framebuffers shared over NVLink peer memory ---------------------
    def accum_alloc_device(self, width, height):
        p = vp(0)
        _check(self.L.ort_accum_alloc_device(self.h, width, height, C.byref(p)), self.L)
        return p.value

    def accum_free_device(self, accum_ptr):
        _check(self.L.ort_accum_free_device(self.h, vp(accum_ptr)), self.L)

    def accum_ipc_export(self, accum_ptr):
        """64-byte CUDA IPC handle of a framebuffer from accum_alloc_device (bytes)"""
        h = (C.c_uint8 * 64)()
        _check(self.L.ort_accum_ipc_export(self.h, vp(accum_ptr), h), self.L)
        return bytes(h)

    def accum_ipc_open(self, handle_bytes):
        h = (C.c_uint8 * 64)(*handle_bytes)
        p = vp(0)
        _check(self.L.ort_accum_ipc_open(self.h, h, C.byref(p)), self.L)
        return p.value

    def accum_ipc_close(self, peer_ptr):
        _check(self.L.ort_accum_ipc_close(self.h, vp(peer_ptr)), self.L)

    def accum_reduce_resolve_device(self, accum_ptr, peer_ptrs, width, height, spp, rgb_ptr=0, rgbe_ptr=0, stream=None):
        """accum += sum(peers) and resolve, one kernel reading the peers' framebuffers in place"""
        n = len(peer_ptrs)
        arr = (vp * max(n, 1))(*[vp(p) for p in peer_ptrs])
        _check(self.L.ort_accum_reduce_resolve_device(self.h, vp(accum_ptr), arr, n, width, height, spp,
                                                      vp(rgb_ptr) if rgb_ptr else None, vp(rgbe_ptr) if rgbe_ptr else None,
                                                      vp(stream) if stream else None), self.L)

    # -- ray cast ---------------------------------------------------------------
    def raycast_batch(self, origins, dirs, want_normal=True):
        """ort_raycast_batch on host arrays [n,3]; returns dict(t, rank, mat, normal, device_ms)"""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        t = np.zeros(n, np.float32); rank = np.zeros(n, np.uint32); mat = np.zeros(n, np.uint32)
        nrm = np.zeros((n, 3), np.float32) if want_normal else None
        st = RenderStats()
        _check(self.L.ort_raycast_batch(self.h, n, _ptr(o), _ptr(d), _ptr(t), _ptr(rank), _ptr(mat), _ptr(nrm),
                                        C.byref(st)), self.L)
        return dict(t=t, rank=rank, mat=mat, normal=nrm, device_ms=st.device_ms)

    def raycast_batch_device(self, n, origins_ptr, dirs_ptr, t_ptr=0, rank_ptr=0, mat_ptr=0, normal_ptr=0, stream=None):
        _check(self.L.ort_raycast_batch_device(self.h, n, vp(origins_ptr), vp(dirs_ptr), vp(t_ptr) if t_ptr else None,
                                               vp(rank_ptr) if rank_ptr else None, vp(mat_ptr) if mat_ptr else None,
                                               vp(normal_ptr) if normal_ptr else None,
                                               vp(stream) if stream else None), self.L)

    def raycast_brute_device(self, n, origins_ptr, dirs_ptr, t_ptr=0, rank_ptr=0, mat_ptr=0, stream=None):
        _check(self.L.ort_raycast_brute_device(self.h, n, vp(origins_ptr), vp(dirs_ptr), vp(t_ptr) if t_ptr else None,
                                               vp(rank_ptr) if rank_ptr else None, vp(mat_ptr) if mat_ptr else None,
                                               vp(stream) if stream else None), self.L)

    def raycast_counters_device(self, n, origins_ptr, dirs_ptr):
        a, b, c = c_u64(0), c_u64(0), c_u64(0)
        _check(self.L.ort_raycast_counters_device(self.h, n, vp(origins_ptr), vp(dirs_ptr),
                                                  C.byref(a), C.byref(b), C.byref(c)), self.L)
        return dict(node_visits=a.value, box_tests=b.value, shape_tests=c.value)


class Multi:
    """OrtMulti: the same scene on several devices of this process (ort_multi_*)"""

    def __init__(self, scenes):
        self.L = scenes[0].L
        self.scenes = list(scenes)
        arr = (vp * len(scenes))(*[sc.h for sc in scenes])
        h = vp(0)
        _check(self.L.ort_multi_create(arr, len(scenes), C.byref(h)), self.L)
        self.h = h

    def device_count(self):
        n, d = c_u32(0), c_u32(0)
        _check(self.L.ort_multi_device_count(self.h, C.byref(n), C.byref(d)), self.L)
        return n.value, d.value

    def render(self, camera_ptr, params, out=None):
        W, H = params.output_width, params.output_height
        if out is None:
            out = np.zeros((H, W, 3), np.float32)
        st = RenderStats()
        _check(self.L.ort_multi_render(self.h, _as_ptr(camera_ptr), C.byref(params), _ptr(out), C.byref(st)), self.L)
        return out, st.as_dict()

    def close(self):
        if self.h:
            self.L.ort_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Progress:
    """OrtProgress: progressive accumulation, dynamic chunk dispatch, checkpoint / resume (ort_progress_*)"""

    def __init__(self, multi, camera_ptr=None, params=None, path=None):
        self.L, self.multi = multi.L, multi
        h = vp(0)
        if path is not None:
            _check(self.L.ort_progress_load(multi.h, str(path).encode(), C.byref(h)), self.L)
        else:
            _check(self.L.ort_progress_create(multi.h, _as_ptr(camera_ptr), C.byref(params), C.byref(h)), self.L)
        self.h = h

    def render(self, max_chunks=0):
        st = RenderStats()
        _check(self.L.ort_progress_render(self.h, max_chunks, C.byref(st)), self.L)
        return st.as_dict()

    def state(self):
        a, b, c = c_u32(0), c_u32(0), c_u32(0)
        _check(self.L.ort_progress_state(self.h, C.byref(a), C.byref(b), C.byref(c)), self.L)
        return dict(chunks_done=a.value, chunks_total=b.value, spp_done=c.value)

    def resolve(self, width, height):
        out = np.zeros((height, width, 3), np.float32)
        _check(self.L.ort_progress_resolve(self.h, _ptr(out)), self.L)
        return out

    def save(self, path):
        _check(self.L.ort_progress_save(self.h, str(path).encode()), self.L)

    def close(self):
        if self.h:
            self.L.ort_progress_destroy(self.h)
            self.h = None


def bake_mesh(vertices, scale, degree, quaternion_xyzw, translate, device=0):
    """ort_bake_mesh: (baked vertices [n, 3], aabb_min, aabb_max) -- the reference's mesh bake as a CUDA kernel"""
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    out = np.zeros_like(v)
    mn, mx = V3(), V3()
    q = V4(*[float(x) for x in quaternion_xyzw]); t = V3(*[float(x) for x in translate])
    _check(lib().ort_bake_mesh(device, v.shape[0], _ptr(v), _ptr(out), c_f(scale), c_f(degree), q, t, C.byref(mn), C.byref(mx)))
    return out, np.array([mn.x, mn.y, mn.z], np.float32), np.array([mx.x, mx.y, mx.z], np.float32)


def load_mesh(path):
    L = lib()
    v = C.POINTER(c_f)(); i = C.POINTER(c_u32)()
    nv = c_u32(0); ni = c_u32(0)
    _check(L.ort_load_mesh(path.encode(), C.byref(v), C.byref(nv), C.byref(i), C.byref(ni)))
    verts = np.ctypeslib.as_array(v, shape=(nv.value * 3,)).copy().reshape(-1, 3) if nv.value else np.zeros((0, 3), np.float32)
    idx = np.ctypeslib.as_array(i, shape=(ni.value,)).copy() if ni.value else np.zeros(0, np.uint32)
    L.ort_free(v); L.ort_free(i)
    return verts, idx


def parse_numeric(text):
    bits = c_u32(0)
    is_float = lib().ort_parse_numeric(text.encode(), C.byref(bits))
    return int(is_float), bits.value


def write_hdr(path, image):
    img = np.ascontiguousarray(image, np.float32)
    H, W = img.shape[:2]
    _check(lib().ort_write_hdr(path.encode(), _ptr(img), W, H))


def v3_to_rgbe(rgb):
    return int(lib().ort_v3_to_rgbe(V3(float(rgb[0]), float(rgb[1]), float(rgb[2]))))
