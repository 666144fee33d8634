"""the acceleration structure built by CUDA kernels (csrc/bvh_build.h/.cuh: Morton
sort, PLOC clustering, level-by-level collapse to the 8-wide layout).

 * the device-built tree equals the host execution of the same per-element functions byte for byte
   (tests/sim, no GPU involved) -- the builder is deterministic;
 * a scene built on the device returns bit-identical hits and images to the host binned-SAH build and
   to the oracle: the structure only decides which records are tested.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import _ptr

pytestmark = pytest.mark.gpu
vp = C.c_void_p


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _sim_parallel_tree(world, root):
    L = C.CDLL(os.path.join(ol.ROOT, "tests", "sim", "libort_sim.so"))
    L.sim_scene_create_parallel.restype = vp
    L.sim_scene_create_parallel.argtypes = [vp, vp, C.c_uint32]
    L.sim_scene_nodes.restype = C.c_uint64; L.sim_scene_nodes.argtypes = [vp, C.POINTER(vp)]
    L.sim_scene_prims.restype = C.c_uint64; L.sim_scene_prims.argtypes = [vp, C.POINTER(vp)]
    L.sim_scene_destroy.argtypes = [vp]
    h = L.sim_scene_create_parallel(world, root, 0)
    assert h
    p = vp()
    nb = L.sim_scene_nodes(h, C.byref(p))
    nodes = np.frombuffer((C.c_char * nb).from_address(p.value), np.uint8).reshape(-1, 80).copy()
    nb = L.sim_scene_prims(h, C.byref(p))
    prims = np.frombuffer((C.c_char * nb).from_address(p.value), np.uint32).reshape(-1, 12).copy()
    L.sim_scene_destroy(h)
    return nodes, prims


@pytest.mark.parametrize("scene", ["testscene", "box_spheres", "c2_bunny_only"])
def test_device_build_equals_host_execution_byte_for_byte(ort, built, testscene_host, scene):
    if scene == "testscene":
        hs = testscene_host
    else:
        base = ol.DATA_DIR if scene == "c2_bunny_only" else ol.SCENES_DIR
        hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, scene + ".scn"), base, 64, 36)
    sc = ort.Scene(hs.world, hs.root, 0, build_on_device=True)
    st = sc.build_stats()
    assert st["on_device"] == 1 and st["wide_depth"] >= 1
    nodes, prims = sc.download()
    want_nodes, want_prims = _sim_parallel_tree(hs.world, hs.root)
    assert nodes.shape == want_nodes.shape and prims.shape == want_prims.shape
    assert np.array_equal(prims, want_prims)
    assert np.array_equal(nodes, want_nodes)
    sc.close()


def test_device_built_scene_gives_identical_hits_and_images(ort, testscene_host, testscene_oracle):
    hs = testscene_host
    a = ort.Scene(hs.world, hs.root, 0, build_on_device=False)
    b = ort.Scene(hs.world, hs.root, 0, build_on_device=True)
    assert b.build_stats()["ploc_iterations"] > 5
    o1, d1 = ol.make_primary_rays(hs.camera_array(), 640, 360)
    o2, d2 = ol.make_incoherent_rays(400000, [-2.9, -2.9, 0.0], [14.9, 14.9, 8.8])
    O = np.concatenate([o1, o2]); D = np.concatenate([d1, d2])
    ra, rb = a.raycast_batch(O, D), b.raycast_batch(O, D)
    for k in ("rank", "mat"):
        assert np.array_equal(ra[k], rb[k]), k
    assert np.array_equal(bits(ra["t"]), bits(rb["t"])) and np.array_equal(bits(ra["normal"]), bits(rb["normal"]))
    ref = testscene_oracle.raycast(O[:200000], D[:200000], mode=0)
    assert np.array_equal(ref["rank"], rb["rank"][:200000]) and np.array_equal(bits(ref["t"]), bits(rb["t"][:200000]))
    P = ort.default_params(160, 90, 6, chunk_spp=2, kernel=ort.ORT_KERNEL_WAVEFRONT)
    ia, _ = a.render(hs.camera, P)
    ib, _ = b.render(hs.camera, P)
    assert np.array_equal(bits(ia), bits(ib))
    a.close(); b.close()


@pytest.mark.parametrize("scene", ["testscene", "box_spheres", "c4_dwarf_hdr"])
def test_scene_from_shape_lists_on_the_device(ort, scene):
    """no octree anywhere: lists -> ranks by sorting -> records -> tree, all on the device; same ranks,
    same octree statistics, same hits, same image as the octree hand-off with the host builder
    (box_spheres has no mesh at all, c4 an .obj mesh)"""
    base = ol.SCENES_DIR if scene == "box_spheres" else ol.DATA_DIR
    path = os.path.join(ol.DATA_DIR, "testscene.scn") if scene == "testscene" else os.path.join(ol.SCENES_DIR, scene + ".scn")
    Wd, Hd = 160, 90
    hs = ort.HostScene.load(path, base, Wd, Hd)
    a = ort.Scene(hs.world, hs.root, 0, build_on_device=False)
    hs2 = ort.HostScene.load(path, base, Wd, Hd, octree=False)
    assert not hs2.root
    b = ort.Scene.from_lists(hs2.world, hs2.lists(), 0, build_on_device=True)
    ia, ib = a.info(), b.info()
    for k in ("triangle_count", "sphere_count", "box_count", "cylinder_count", "csg_count", "record_count",
              "octree_node_count", "octree_max_depth", "material_count", "light_count", "root_min", "root_max"):
        assert ia[k] == ib[k], k
    # every record carries the rank the octree would have given it: same records in rank order
    _, pa = a.download()
    _, pb = b.download()
    pa = pa[np.argsort(pa[:, 3])]; pb = pb[np.argsort(pb[:, 3])]
    # (a cylinder's aux index, bits 8.. of the kind word, depends on the order of emission)
    pa[:, 11] &= 0xFF; pb[:, 11] &= 0xFF
    assert np.array_equal(pa, pb)
    assert b.build_stats()["on_device"] == 1
    lo, hi = np.array(ia["root_min"]) + 0.05, np.array(ia["root_max"]) - 0.05
    o, d = ol.make_incoherent_rays(200000, list(lo), list(hi))
    ra, rb = a.raycast_batch(o, d), b.raycast_batch(o, d)
    assert np.array_equal(ra["rank"], rb["rank"]) and np.array_equal(bits(ra["t"]), bits(rb["t"]))
    assert np.array_equal(ra["mat"], rb["mat"]) and np.array_equal(bits(ra["normal"]), bits(rb["normal"]))
    P = ort.default_params(Wd, Hd, 4, chunk_spp=2, kernel=ort.ORT_KERNEL_WAVEFRONT)
    ia_, _ = a.render(hs.camera, P)
    ib_, _ = b.render(hs2.camera, P)
    assert np.array_equal(bits(ia_), bits(ib_))
    a.close(); b.close()
