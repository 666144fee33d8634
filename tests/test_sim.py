"""CPU checks of the PRODUCT's flattener and shared host/device headers.

tests/sim/libort_sim.so is a host build (g++ -ffp-contract=off) of csrc/bvh.h + csrc/path.h +
csrc/scene_flatten.cpp -- the same source the CUDA kernels are compiled from -- so the wide-BVH
layout, the traversal logic and the path logic can be compared with the oracle on a box without
a GPU.  It is test infrastructure: the product library never contains or calls it.
On the CPU both sides use glibc's libm, so even whole images must agree bit for bit, except where
the reference's own octree culling is not conservative ("A vs B").
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import _ptr

vp = C.c_void_p


class Sim:
    def __init__(self):
        L = C.CDLL(os.path.join(ol.ROOT, "tests", "sim", "libort_sim.so"))
        L.sim_scene_create.restype = vp
        L.sim_scene_create.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float]
        L.sim_scene_destroy.argtypes = [vp]
        L.sim_raycast_batch.argtypes = [vp, C.c_uint64] + [vp] * 7 + [C.c_int]
        L.sim_render.argtypes = [vp, vp, C.POINTER(ol.RenderParams), vp, vp, C.c_int]
        self.L = L

    def scene(self, world, root, traversal_cost=-1.0, pad_rel=-1.0, pad_scene=-1.0):
        h = self.L.sim_scene_create(world, root, traversal_cost, pad_rel, pad_scene)
        assert h
        return h

    def raycast(self, h, O, D):
        O = ol.f32a(O, (-1, 3)); D = ol.f32a(D, (-1, 3))
        n = len(O)
        t = np.zeros(n, np.float32); rank = np.zeros(n, np.uint32); mat = np.zeros(n, np.uint32)
        nrm = np.zeros((n, 3), np.float32); cnt = np.zeros(3, np.uint64)
        self.L.sim_raycast_batch(h, n, _ptr(O), _ptr(D), _ptr(t), _ptr(rank), _ptr(mat), _ptr(nrm), _ptr(cnt), 8)
        return dict(t=t, rank=rank, mat=mat, normal=nrm, node_visits=int(cnt[0]), box_tests=int(cnt[1]), shape_tests=int(cnt[2]))

    def render(self, h, camera, params):
        img = np.zeros((params.output_height, params.output_width, 3), np.float32)
        cnt = np.zeros(1, np.uint64)
        self.L.sim_render(h, camera, C.byref(params), _ptr(img), _ptr(cnt), 8)
        return img, int(cnt[0])


@pytest.fixture(scope="module")
def sim(built):
    return Sim()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_wide_bvh_traversal_equals_reference_traversal(sim, testscene_host, testscene_oracle):
    """hit t / rank / material / normal bit-identical to the octree BFS of the reference"""
    hs, osc = testscene_host, testscene_oracle
    h = sim.scene(hs.world, hs.root)
    o1, d1 = ol.make_primary_rays(hs.camera_array(), 480, 270)
    o2, d2 = ol.make_incoherent_rays(200000, [-2.9, -2.9, 0.0], [14.9, 14.9, 8.8])
    O = np.concatenate([o1, o2]); D = np.concatenate([d1, d2])
    a = osc.raycast(O, D, mode=0)
    b = sim.raycast(h, O, D)
    assert np.array_equal(a["rank"], b["rank"])
    assert np.array_equal(bits(a["t"]), bits(b["t"]))
    assert np.array_equal(a["mat"], b["mat"])
    assert np.array_equal(bits(a["normal"]), bits(b["normal"]))
    # and it does far less work: that is the point of re-building the structure
    n = len(O)
    assert b["shape_tests"] / n < 0.1 * a["shape_tests"] / n


def test_tangent_sphere_rays(sim, testscene_host, testscene_oracle):
    """the reference reports a ray TANGENT to a sphere at t = -b/(2a), half way to the sphere
    (ray.cpp:174-183); the unclipped sphere tree must reproduce that"""
    hs, osc = testscene_host, testscene_oracle
    h = sim.scene(hs.world, hs.root)
    c = np.array([0.0, 0.0, 0.8], np.float32); r = 0.4      # the glass sphere of testscene.scn
    rng = np.random.default_rng(11)
    n = 20000
    o = np.tile(np.array([5.5, 2.9, 2.9], np.float32), (n, 1)) + rng.normal(scale=0.05, size=(n, 3)).astype(np.float32)
    # aim at points at distance ~r from the centre, perpendicular to the view direction
    to_c = c - o
    perp = np.cross(to_c, rng.normal(size=(n, 3)))
    perp /= np.linalg.norm(perp, axis=1, keepdims=True)
    L = np.linalg.norm(to_c, axis=1, keepdims=True)
    rho = r * (1 + rng.uniform(-1e-4, 1e-4, (n, 1)))          # closest approach of the ray line to the centre
    d = (c + perp * (L * np.tan(np.arcsin(rho / L))) - o)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    a = osc.raycast(o, d, mode=1)          # brute force: structure-free answer
    b = sim.raycast(h, o, d)
    assert np.array_equal(a["rank"], b["rank"]) and np.array_equal(bits(a["t"]), bits(b["t"]))
    tm, _ = osc.records()
    sphere_hits = tm[a["rank"][a["rank"] != 0xFFFFFFFF], 0] == 1
    assert sphere_hits.sum() > 100          # the case is exercised
    # some of them are the half-way hits: t much smaller than the distance to the sphere
    dist = np.linalg.norm(c - o, axis=1)
    half_way = (a["t"] < 0.6 * dist) & (a["rank"] != 0xFFFFFFFF)
    assert half_way.sum() > 0


def test_degenerate_directions(sim, testscene_host, testscene_oracle):
    """axis-parallel rays (zero direction components) and non-unit directions"""
    hs, osc = testscene_host, testscene_oracle
    h = sim.scene(hs.world, hs.root)
    rng = np.random.default_rng(5)
    o = rng.uniform([-2.5, -2.5, 0.2], [14, 14, 8.5], (6000, 3)).astype(np.float32)
    d = np.zeros((6000, 3), np.float32)
    d[np.arange(6000), rng.integers(0, 3, 6000)] = rng.choice([-1.0, 1.0], 6000)
    d[3000:] += (rng.integers(0, 2, (3000, 3)) * rng.normal(size=(3000, 3))).astype(np.float32)
    d[4500:] *= rng.uniform(0.6, 3.0, (1500, 1)).astype(np.float32)      # t is in units of |d|
    a = osc.raycast(o, d, mode=1)
    b = sim.raycast(h, o, d)
    assert np.array_equal(a["rank"], b["rank"]) and np.array_equal(bits(a["t"]), bits(b["t"]))


def test_path_logic_equals_oracle_bitwise(sim, oracle, ort):
    """csrc/path.h on the host == oracle, whole image, brute-force traversal on both sides"""
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, 48, 27)
    osc = oracle.scene(hs.world, hs.root)
    h = sim.scene(hs.world, hs.root)
    for chunk in (0, 3):
        P = ol.default_params(48, 27, 8, chunk_spp=chunk)
        img_o, _ = osc.render(hs.camera, P, brute=1)
        img_s, rays = sim.render(h, hs.camera, P)
        assert np.array_equal(bits(img_o), bits(img_s)), chunk
        assert rays > 48 * 27 * 8


def test_image_equals_reference_traversal_except_octree_leaks(sim, testscene_host, testscene_oracle):
    """against the reference's OWN traversal the image is bit-identical except for the few
    pixels where the octree culling drops a hit (box entry t < 1e-6 from outside, ray.cpp:800)"""
    hs, osc = testscene_host, testscene_oracle
    h = sim.scene(hs.world, hs.root)
    P = ol.default_params(480, 270, 2)
    P.tile_min_x, P.tile_one_past_max_x, P.tile_min_y, P.tile_one_past_max_y = 120, 360, 60, 200
    img_o, cnt = osc.render(hs.camera, P)
    img_s, rays = sim.render(h, hs.camera, P)
    differ = (bits(img_o) != bits(img_s)).any(axis=2)
    assert differ.sum() <= 5, differ.sum()
    assert abs(rays - cnt["rays"]) <= 100


def test_chunked_accumulation_is_order_independent(sim, oracle, ort):
    """chunk sub-ranges rendered separately and summed in fixed point == one call"""
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, 32, 18)
    osc = oracle.scene(hs.world, hs.root)
    P = ol.default_params(32, 18, 12, chunk_spp=4)
    whole, _ = osc.render(hs.camera, P)
    parts = []
    for b, e in ((0, 1), (1, 3)):
        Q = ol.default_params(32, 18, 12, chunk_spp=4)
        Q.chunk_begin, Q.chunk_end = b, e
        img, _ = osc.render(hs.camera, Q)
        parts.append(img)
    # each partial image is (fixed-point sum of its chunks)/spp rounded to f32; the sums themselves add
    # exactly, so the only slack is the f32 rounding of the three quotients
    total = parts[0].astype(np.float64) + parts[1].astype(np.float64)
    assert np.allclose(total, whole.astype(np.float64), rtol=3e-7, atol=1e-9)
    assert np.abs(whole).sum() > 0


def test_padding_is_what_makes_culling_safe(sim, testscene_host, testscene_oracle):
    """with the box padding switched off the quantised slab test may cull true hits; with the
    default padding it never does on this buffer -- guards the BuildOptions defaults"""
    hs, osc = testscene_host, testscene_oracle
    o, d = ol.make_incoherent_rays(150000, [-2.9, -2.9, 0.0], [14.9, 14.9, 8.8], seed=21)
    a = osc.raycast(o, d, mode=0)
    h = sim.scene(hs.world, hs.root)
    b = sim.raycast(h, o, d)
    assert np.array_equal(a["rank"], b["rank"])


@pytest.mark.parametrize("top", ["default", "0"])
def test_parallel_builder_gives_the_same_hits_with_comparable_work(sim, testscene_host, testscene_oracle, top, monkeypatch):
    """the data-parallel builder (csrc/bvh_build.h: Morton order, PLOC clustering,
    level-by-level collapse to 8-wide), executed on the host, yields a tree that returns the same
    (t, rank, material, normal) for every ray as the binned-SAH tree, at a comparable traversal cost"""
    if top != "default":
        monkeypatch.setenv("ORT_PLOC_TOP", top)          # pure clustering, no SAH-built top
    hs, osc = testscene_host, testscene_oracle
    L = sim.L
    L.sim_scene_create_parallel.restype = vp
    L.sim_scene_create_parallel.argtypes = [vp, vp, C.c_uint32]
    h_sah = sim.scene(hs.world, hs.root)
    h_par = L.sim_scene_create_parallel(hs.world, hs.root, 0)
    assert h_par
    o1, d1 = ol.make_primary_rays(hs.camera_array(), 320, 180)
    o2, d2 = ol.make_incoherent_rays(100000, [-2.9, -2.9, 0.0], [14.9, 14.9, 8.8])
    O = np.concatenate([o1, o2]); D = np.concatenate([d1, d2])
    a = sim.raycast(h_sah, O, D)
    b = sim.raycast(h_par, O, D)
    for k in ("rank", "mat"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(bits(a["t"]), bits(b["t"])) and np.array_equal(bits(a["normal"]), bits(b["normal"]))
    ref = osc.raycast(O, D, mode=0)
    assert np.array_equal(ref["rank"], b["rank"]) and np.array_equal(bits(ref["t"]), bits(b["t"]))
    # tree quality: within 1.5x of the SAH tree's node visits and primitive tests on these rays
    print("SAH  nodes/ray %.2f tests/ray %.2f | PLOC nodes/ray %.2f tests/ray %.2f" % (
        a["node_visits"] / len(O), a["shape_tests"] / len(O), b["node_visits"] / len(O), b["shape_tests"] / len(O)))
    assert b["node_visits"] < 1.5 * a["node_visits"] and b["shape_tests"] < 1.5 * a["shape_tests"]
    # every record appears exactly once
    p = vp()
    L.sim_scene_prims.restype = C.c_uint64; L.sim_scene_prims.argtypes = [vp, C.POINTER(vp)]
    nb = L.sim_scene_prims(h_par, C.byref(p))
    recs = np.frombuffer((C.c_char * nb).from_address(p.value), np.uint32).reshape(-1, 12)
    na = L.sim_scene_prims(h_sah, C.byref(p))
    recs_sah = np.frombuffer((C.c_char * na).from_address(p.value), np.uint32).reshape(-1, 12)
    assert nb == na and np.array_equal(np.sort(recs[:, 3]), np.sort(recs_sah[:, 3]))
    L.sim_scene_destroy(h_par); L.sim_scene_destroy(h_sah)


@pytest.mark.parametrize("scene", ["testscene", "box_spheres", "c3_bunny_box", "c4_dwarf_hdr"])
def test_ranks_from_shape_lists_equal_ranks_from_the_octree(sim, ort, scene):
    """the tie-break ranks -- breadth-first octree order, push-buffer order within a node --
    computed from the shape lists alone by two radix sorts (collect_records_from_lists) equal those read
    off the octree that push_shape_inside_node builds (ray.cpp:1799-1948), record for record; so do the
    octree statistics and, therefore, the whole flattened scene."""
    base = ol.DATA_DIR
    path = os.path.join(ol.DATA_DIR, "testscene.scn") if scene == "testscene" else os.path.join(ol.SCENES_DIR, scene + ".scn")
    if scene == "box_spheres":
        base = ol.SCENES_DIR
    hs = ort.HostScene.load(path, base, 64, 36)
    L = sim.L
    L.sim_scene_create_from_lists.restype = vp
    L.sim_scene_create_from_lists.argtypes = [vp, vp]
    L.sim_scene_nodes.restype = C.c_uint64; L.sim_scene_nodes.argtypes = [vp, C.POINTER(vp)]
    L.sim_scene_prims.restype = C.c_uint64; L.sim_scene_prims.argtypes = [vp, C.POINTER(vp)]
    L.sim_scene_info.argtypes = [vp, vp, vp]
    lists = hs.lists()
    a = sim.scene(hs.world, hs.root)
    b = L.sim_scene_create_from_lists(hs.world, C.byref(lists))
    assert b

    def dump(h):
        p = vp()
        nb = L.sim_scene_nodes(h, C.byref(p))
        nodes = np.frombuffer((C.c_char * nb).from_address(p.value), np.uint8).copy()
        nb = L.sim_scene_prims(h, C.byref(p))
        prims = np.frombuffer((C.c_char * nb).from_address(p.value), np.uint32).reshape(-1, 12).copy()
        info = ort.api.SceneInfo(); depth = C.c_uint32(0)
        L.sim_scene_info(h, C.byref(info), C.byref(depth))
        return nodes, prims, info.as_dict()

    na, pa, ia = dump(a)
    nb_, pb, ib = dump(b)
    assert ia == ib, (ia, ib)
    assert np.array_equal(pa, pb) and np.array_equal(na, nb_)
    # without the octree at all
    hs2 = ort.HostScene.load(path, base, 64, 36, octree=False)
    assert not hs2.root
    c = L.sim_scene_create_from_lists(hs2.world, C.byref(hs2.lists()))
    nc, pc, ic = dump(c)
    assert ic == ia and np.array_equal(pc, pa)
    for h in (a, b, c):
        L.sim_scene_destroy(h)
