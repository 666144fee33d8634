"""ctypes bindings of the TEST ORACLES (oracle/liboracle.so, oracle/_ref/libref.so).

Test infrastructure only: nothing under offline_raytracer_b200/ imports this.
  * ``Ref``    -- the UNMODIFIED reference sources compiled for Linux (oracle/ref_shim.cpp)
  * ``Oracle`` -- the restated hot path (oracle/oracle.cpp)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
DATA_DIR = os.path.join(REF_DIR, "data")
SCENES_DIR = os.path.join(ROOT, "scenes")

c_f = C.c_float
c_u32 = C.c_uint32
c_u64 = C.c_uint64
c_i32 = C.c_int32
vp = C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


def f32a(x, shape=None):
    a = np.ascontiguousarray(x, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class RenderParams(C.Structure):
    """Mirror of OrtRenderParams (include/ort_b200.h)."""
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


def default_params(width, height, spp, rr=0.8, seed=1234567, chunk_spp=0, kernel=0):
    p = RenderParams()
    p.output_width, p.output_height = width, height
    p.tile_min_x, p.tile_min_y = 0, 0
    p.tile_one_past_max_x, p.tile_one_past_max_y = width, height
    p.ray_per_pixel_count = spp
    p.russian_roulette_value = rr
    p.base_seed = seed
    p.chunk_spp = chunk_spp
    p.chunk_begin = p.chunk_end = 0
    p.kernel = kernel
    p.roughness = 0.01
    p.dont_get_too_close_epsilon = 0.0001
    p.aperture_radius = 0.1
    p.lens_z_offset = 0.1
    p.focus_target[0], p.focus_target[1], p.focus_target[2] = 0.0, 0.0, 0.2
    return p


def build_oracles():
    """make -C oracle (liboracle.so always; _ref when /root/reference exists)."""
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref.so"))


def have_data():
    return os.path.exists(os.path.join(DATA_DIR, "testscene.scn"))


_ISECT_SIGS = {
    "intersect_triangle": [vp] * 6,
    "intersect_sphere": [vp, c_f, vp, vp, vp],
    "intersect_aab": [vp] * 5,
    "intersect_cylinder": [vp, vp, c_f, vp, vp, vp],
}


class _FuncsMixin:
    """single-function entry points shared by libref (ref_*) and liboracle (oracle_*)"""

    def _bind_funcs(self, lib, prefix):
        self._p = prefix
        for name, sig in _ISECT_SIGS.items():
            f = getattr(lib, prefix + name)
            f.argtypes = sig
            f.restype = None
        g = lambda n: getattr(lib, prefix + n)
        g("in_rect").argtypes = [vp, vp, vp]; g("in_rect").restype = c_i32
        g("xor_shift_32").argtypes = [c_u32]; g("xor_shift_32").restype = c_u32
        g("random_between_0_1").argtypes = [C.POINTER(c_u32)]; g("random_between_0_1").restype = c_f
        g("random_between").argtypes = [C.POINTER(c_u32), c_f, c_f]; g("random_between").restype = c_f
        g("random_between_u32").argtypes = [C.POINTER(c_u32), c_u32, c_u32]; g("random_between_u32").restype = c_u32
        g("sample_brdf").argtypes = [C.POINTER(c_u32), vp, vp, c_f, vp, vp, C.POINTER(c_i32)]; g("sample_brdf").restype = None
        g("pdf_brdf").argtypes = [vp, vp, vp, c_f, vp]; g("pdf_brdf").restype = c_f
        g("eval_scattering").argtypes = [vp, vp, vp, vp, c_f, c_f, vp]; g("eval_scattering").restype = None
        g("sample_random_lights").argtypes = [vp, c_u32]; g("sample_random_lights").restype = c_u32
        self._lib = lib

    def _f(self, name):
        return getattr(self._lib, self._p + name)

    def intersect(self, kind, *args):
        """kind in triangle/sphere/aab/cylinder; vector args as 3-float arrays, returns [t, nx, ny, nz, inner]"""
        out = np.zeros(5, np.float32)
        conv = []
        keep = []
        for a in args:
            if isinstance(a, (float, np.floating)):
                conv.append(c_f(a))
            else:
                arr = f32a(a)
                keep.append(arr)
                conv.append(_ptr(arr))
        self._f("intersect_" + kind)(*conv, _ptr(out))
        return out

    def in_rect(self, p, mn, mx):
        a, b, c = f32a(p), f32a(mn), f32a(mx)
        return int(self._f("in_rect")(_ptr(a), _ptr(b), _ptr(c)))

    def xor_shift_32(self, s):
        return int(self._f("xor_shift_32")(c_u32(s)))

    def random_between_0_1(self, s):
        st = c_u32(s)
        v = self._f("random_between_0_1")(C.byref(st))
        return np.float32(v), st.value

    def random_between(self, s, mn, mx):
        st = c_u32(s)
        v = self._f("random_between")(C.byref(st), c_f(mn), c_f(mx))
        return np.float32(v), st.value

    def random_between_u32(self, s, mn, mx):
        st = c_u32(s)
        v = self._f("random_between_u32")(C.byref(st), c_u32(mn), c_u32(mx))
        return int(v), st.value

    def sample_brdf(self, s, N, wo, roughness, mat):
        st = c_u32(s)
        N, wo, mat = f32a(N), f32a(wo), f32a(mat)
        wi = np.zeros(3, np.float32)
        is_t = c_i32(0)
        self._f("sample_brdf")(C.byref(st), _ptr(N), _ptr(wo), c_f(roughness), _ptr(mat), _ptr(wi), C.byref(is_t))
        return wi, is_t.value, st.value

    def pdf_brdf(self, N, wi, wo, roughness, mat):
        N, wi, wo, mat = f32a(N), f32a(wi), f32a(wo), f32a(mat)
        return np.float32(self._f("pdf_brdf")(_ptr(N), _ptr(wi), _ptr(wo), c_f(roughness), _ptr(mat)))

    def eval_scattering(self, N, wi, wo, mat, roughness, distance):
        N, wi, wo, mat = f32a(N), f32a(wi), f32a(wo), f32a(mat)
        out = np.zeros(3, np.float32)
        self._f("eval_scattering")(_ptr(N), _ptr(wi), _ptr(wo), _ptr(mat), c_f(roughness), c_f(distance), _ptr(out))
        return out


class Ref(_FuncsMixin):
    """the unmodified reference (oracle/_ref/libref.so)"""

    def __init__(self):
        lib = C.CDLL(os.path.join(REF_DIR, "libref.so"))
        self.lib = lib
        lib.ref_scene_load.restype = vp
        lib.ref_scene_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint]
        for n in ("ref_scene_world", "ref_scene_camera", "ref_scene_root", "ref_scene_meshes"):
            getattr(lib, n).restype = vp
            getattr(lib, n).argtypes = [vp]
        lib.ref_scene_mesh_count.restype = c_u32
        lib.ref_scene_mesh_count.argtypes = [vp]
        lib.ref_scene_counts.argtypes = [vp, vp]
        lib.ref_struct_sizes.argtypes = [vp]
        lib.ref_tiled_raytrace_bvh.restype = c_u64
        lib.ref_tiled_raytrace_bvh.argtypes = [vp, vp] + [C.c_int] * 6 + [C.POINTER(c_u32), c_u32, c_f]
        lib.ref_render_pixel_seeds.restype = c_u64
        lib.ref_render_pixel_seeds.argtypes = [vp, vp] + [C.c_int] * 6 + [c_u32, c_u32, c_f, C.c_int]
        lib.ref_render_tiles.restype = c_u64
        lib.ref_render_tiles.argtypes = [vp, vp, C.c_int, C.c_int, c_u32, c_u32, c_f, C.c_int, C.c_int]
        lib.ref_raycast_batch.restype = c_u64
        lib.ref_raycast_batch.argtypes = [vp, c_u64, vp, vp, vp, vp, vp, vp, C.c_int]
        lib.ref_eat_numeric.restype = c_i32
        lib.ref_eat_numeric.argtypes = [C.c_char_p, C.POINTER(c_u32)]
        lib.ref_load_mesh.restype = c_i32
        lib.ref_load_mesh.argtypes = [C.c_char_p, C.POINTER(C.POINTER(c_f)), C.POINTER(c_u32),
                                      C.POINTER(C.POINTER(c_u32)), C.POINTER(c_u32)]
        lib.ref_free.argtypes = [vp]
        self._bind_funcs(lib, "ref_")

    def struct_sizes(self):
        s = np.zeros(12, np.uint32)
        self.lib.ref_struct_sizes(_ptr(s))
        return s

    def scene_load(self, scn_path, base_dir, width, height, with_csg=1, node_mb=0, shape_mb=0):
        if not base_dir.endswith("/"):
            base_dir += "/"
        h = self.lib.ref_scene_load(scn_path.encode(), base_dir.encode(), width, height, with_csg, node_mb, shape_mb)
        if not h:
            raise RuntimeError("ref_scene_load failed: " + scn_path)
        return RefScene(self, h, width, height)

    def eat_numeric(self, text):
        bits = c_u32(0)
        is_float = self.lib.ref_eat_numeric(text.encode(), C.byref(bits))
        return int(is_float), bits.value

    def load_mesh(self, path):
        v = C.POINTER(c_f)(); i = C.POINTER(c_u32)()
        nv = c_u32(0); ni = c_u32(0)
        rc = self.lib.ref_load_mesh(path.encode(), C.byref(v), C.byref(nv), C.byref(i), C.byref(ni))
        if rc != 0:
            raise RuntimeError("ref_load_mesh rc=%d" % rc)
        verts = np.ctypeslib.as_array(v, shape=(nv.value * 3,)).copy().reshape(-1, 3)
        idx = np.ctypeslib.as_array(i, shape=(ni.value,)).copy()
        self.lib.ref_free(v); self.lib.ref_free(i)
        return verts, idx


class RefScene:
    def __init__(self, ref, h, width, height):
        self.ref, self.h, self.width, self.height = ref, h, width, height
        lib = ref.lib
        self.world = lib.ref_scene_world(h)
        self.camera = lib.ref_scene_camera(h)
        self.root = lib.ref_scene_root(h)

    def counts(self):
        c = np.zeros(10, np.uint64)
        self.ref.lib.ref_scene_counts(self.h, _ptr(c))
        keys = ["spheres", "boxes", "cylinders", "materials", "meshes", "lights", "light_bytes",
                "nodes", "node_bytes", "shape_bytes"]
        return dict(zip(keys, [int(x) for x in c]))

    def camera_array(self):
        return np.ctypeslib.as_array(C.cast(self.camera, C.POINTER(c_f)), shape=(12,)).copy()

    def render_tile(self, rect, seed, spp, rr=0.8, out=None):
        W, H = self.width, self.height
        if out is None:
            out = np.zeros((H, W, 3), np.float32)
        st = c_u32(seed)
        n = self.ref.lib.ref_tiled_raytrace_bvh(self.h, _ptr(out), W, H, rect[0], rect[1], rect[2], rect[3],
                                                C.byref(st), spp, c_f(rr))
        return out, int(n), st.value

    def render_pixel_seeds(self, base_seed, spp, rr=0.8, rect=None, threads=8):
        W, H = self.width, self.height
        rect = rect or (0, 0, W, H)
        out = np.zeros((H, W, 3), np.float32)
        n = self.ref.lib.ref_render_pixel_seeds(self.h, _ptr(out), W, H, rect[0], rect[1], rect[2], rect[3],
                                                base_seed, spp, c_f(rr), threads)
        return out, int(n)

    def render_tiles(self, master_seed, spp, rr=0.8, threads=8, max_tiles=0):
        W, H = self.width, self.height
        out = np.zeros((H, W, 3), np.float32)
        n = self.ref.lib.ref_render_tiles(self.h, _ptr(out), W, H, master_seed, spp, c_f(rr), threads, max_tiles)
        return out, int(n)

    def raycast(self, origins, dirs, threads=8):
        o, d = f32a(origins, (-1, 3)), f32a(dirs, (-1, 3))
        n = o.shape[0]
        t = np.zeros(n, np.float32); mat = np.zeros(n, np.uint32)
        nrm = np.zeros((n, 3), np.float32); inner = np.zeros(n, np.int32)
        tests = self.ref.lib.ref_raycast_batch(self.h, n, _ptr(o), _ptr(d), _ptr(t), _ptr(mat), _ptr(nrm), _ptr(inner), threads)
        return dict(t=t, mat=mat, normal=nrm, inner=inner, tests=int(tests))

    def sample_random_lights(self, state):
        return int(self.ref.lib.ref_sample_random_lights(self.h, c_u32(state)))


class Oracle(_FuncsMixin):
    """the restated hot path (oracle/liboracle.so)"""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracles()
        lib = C.CDLL(path)
        self.lib = lib
        lib.oracle_scene_create.restype = vp
        lib.oracle_scene_create.argtypes = [vp, vp]
        lib.oracle_scene_destroy.argtypes = [vp]
        lib.oracle_scene_info.argtypes = [vp, vp]
        lib.oracle_scene_records.argtypes = [vp, vp, vp]
        lib.oracle_raycast_batch.argtypes = [vp, c_u64, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
        lib.oracle_render.argtypes = [vp, vp, C.POINTER(RenderParams), vp, vp, C.c_int, C.c_int]
        lib.oracle_render_accum.argtypes = [vp, vp, C.POINTER(RenderParams), vp, C.c_int]
        lib.oracle_accum_resolve.argtypes = [vp, C.c_int, C.c_int, c_u32, vp]
        lib.oracle_to_fixed.argtypes = [c_f]
        lib.oracle_to_fixed.restype = C.c_int64
        lib.oracle_make_rays.argtypes = [C.c_int, vp, C.POINTER(RenderParams), vp, vp, c_u32, c_u64, vp, vp]
        self._bind_funcs(lib, "oracle_")

    def scene(self, world_ptr, root_ptr):
        return OracleScene(self, world_ptr, root_ptr)

    def make_camera_rays(self, camera_ptr, params, seed, n):
        """BASELINE config 2, coherent buffer: n lens-sampled primaries of the params' grid (xorshift streams of `seed`)"""
        o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32)
        self.lib.oracle_make_rays(0, camera_ptr, C.byref(params), None, None, seed, n, _ptr(o), _ptr(d))
        return o, d

    def make_random_rays(self, box_min, box_max, seed, n):
        """BASELINE config 2, incoherent buffer: origins uniform in the box, directions uniform on the sphere"""
        o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32)
        lo, hi = f32a(box_min), f32a(box_max)
        self.lib.oracle_make_rays(1, None, None, _ptr(lo), _ptr(hi), seed, n, _ptr(o), _ptr(d))
        return o, d


class OracleScene:
    def __init__(self, oracle, world_ptr, root_ptr):
        self.o = oracle
        self.h = oracle.lib.oracle_scene_create(world_ptr, root_ptr)

    def close(self):
        if self.h:
            self.o.lib.oracle_scene_destroy(self.h)
            self.h = None

    def info(self):
        a = np.zeros(4, np.uint32)
        self.o.lib.oracle_scene_info(self.h, _ptr(a))
        return dict(records=int(a[0]), nodes=int(a[1]), max_depth=int(a[2]), lights=int(a[3]))

    def records(self):
        n = self.info()["records"]
        tm = np.zeros((n, 2), np.uint32); g = np.zeros((n, 9), np.float32)
        self.o.lib.oracle_scene_records(self.h, _ptr(tm), _ptr(g))
        return tm, g

    def raycast(self, origins, dirs, mode=0, threads=8):
        o, d = f32a(origins, (-1, 3)), f32a(dirs, (-1, 3))
        n = o.shape[0]
        t = np.zeros(n, np.float32); rank = np.zeros(n, np.uint32); mat = np.zeros(n, np.uint32)
        nrm = np.zeros((n, 3), np.float32); inner = np.zeros(n, np.int32)
        cnt = np.zeros(3, np.uint64)
        self.o.lib.oracle_raycast_batch(self.h, n, _ptr(o), _ptr(d), mode, _ptr(t), _ptr(rank), _ptr(mat),
                                        _ptr(nrm), _ptr(inner), _ptr(cnt), threads)
        return dict(t=t, rank=rank, mat=mat, normal=nrm, inner=inner,
                    shape_tests=int(cnt[0]), box_tests=int(cnt[1]), node_visits=int(cnt[2]))

    def render(self, camera_ptr, params, threads=8, brute=0):
        W, H = params.output_width, params.output_height
        out = np.zeros((H, W, 3), np.float32)
        cnt = np.zeros(4, np.uint64)
        self.o.lib.oracle_render(self.h, camera_ptr, C.byref(params), _ptr(out), _ptr(cnt), threads, brute)
        return out, dict(rays=int(cnt[0]), shape_tests=int(cnt[1]), box_tests=int(cnt[2]), node_visits=int(cnt[3]))

    def sample_random_lights(self, state):
        return int(self.o.lib.oracle_sample_random_lights(self.h, c_u32(state)))

    def render_accum(self, camera_ptr, params, accum, threads=8):
        """adds the fixed-point chunk sums of params' chunk range into accum (int64 [H, W, 4])"""
        assert accum.dtype == np.int64 and accum.flags["C_CONTIGUOUS"]
        self.o.lib.oracle_render_accum(self.h, camera_ptr, C.byref(params), _ptr(accum), threads)

    def accum_resolve(self, accum, spp):
        H, W = accum.shape[:2]
        out = np.zeros((H, W, 3), np.float32)
        self.o.lib.oracle_accum_resolve(_ptr(accum), W, H, spp, _ptr(out))
        return out


def stream_seed(base, pixel_index, chunk):
    """ort_stream_seed (include/ort_b200.h)"""
    M = 0xFFFFFFFF
    h = (base ^ ((pixel_index * 0x9E3779B1) & M) ^ ((chunk * 0x85EBCA77) & M)) & M
    h ^= h >> 16; h = (h * 0x85EBCA6B) & M
    h ^= h >> 13; h = (h * 0xC2B2AE35) & M
    h ^= h >> 16
    return h or 0x6D2B79F5


def xorshift_np(state):
    """vectorised reference xorshift (code/random.h:5-16) on uint32 arrays"""
    x = state.astype(np.uint32).copy()
    x ^= (x << np.uint32(13))
    x ^= (x >> np.uint32(17))
    x ^= (x >> np.uint32(5))
    return x


def make_primary_rays(camera12, width, height, n_per_pixel=1, seed=1, focus=(0.0, 0.0, 0.2),
                      aperture=0.1, lens_z=0.1):
    """explicit primary-ray buffers from the reference camera model (code/ray.cpp:1198-1237),
    computed in float32 with numpy (the rays are INPUTS shared by all sides, so their own
    rounding does not matter)."""
    cam = f32a(camera12, (4, 3))
    p, X, Y, Z = cam
    f = np.float32
    ys, xs = np.mgrid[0:height, 0:width]
    xs = np.repeat(xs.reshape(-1), n_per_pixel); ys = np.repeat(ys.reshape(-1), n_per_pixel)
    px = (f(2.0) * xs.astype(np.float32) / f(width)) - f(1.0)
    py = (f(2.0) * ys.astype(np.float32) / f(height)) - f(1.0)
    c2p = px[:, None] * X + py[:, None] * Y - Z
    c2p /= np.linalg.norm(c2p, axis=1, keepdims=True).astype(np.float32)
    focal = np.float32(np.linalg.norm(p - f32a(focus)))
    fp = p + focal * c2p
    st = np.arange(1, xs.size + 1, dtype=np.uint32) * np.uint32(2654435761) + np.uint32(seed)
    st[st == 0] = 1
    st = xorshift_np(xorshift_np(st))
    rad = (st.astype(np.float32) / f(4294967295.0)) * f(2 * np.pi)
    lens = p + (f(aperture) * np.cos(rad))[:, None] * X + (f(aperture) * np.sin(rad))[:, None] * Y - f(lens_z) * Z
    d = fp - lens
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    return np.ascontiguousarray(lens, np.float32), np.ascontiguousarray(d, np.float32)


def make_incoherent_rays(n, box_min, box_max, seed=2, inflate=1.0):
    """origins uniform in an (inflated) box, directions uniform on the sphere"""
    rng = np.random.default_rng(seed)
    mn, mx = np.asarray(box_min, np.float64), np.asarray(box_max, np.float64)
    c, hd = 0.5 * (mn + mx), 0.5 * (mx - mn) * inflate
    o = (c + (rng.random((n, 3)) * 2 - 1) * hd).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(o), np.ascontiguousarray(d.astype(np.float32))
