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


def test_mesh_bake_on_the_device_is_bit_identical(ort, data_dir):
    """the mesh bake (scale -> rotate about (0,1,0) -> quaternion -> translate + the AABB fold that
    starts at (FLT_MAX, FLT_MIN), code/macos_main.mm:382-413) as a CUDA kernel: vertices and boxes of every mesh of
    the BASELINE scenes equal the host bake bit for bit (PLY + OBJ, with and without the extra z rotation)"""
    for name in ("c3_bunny_box", "c4_dwarf_hdr"):
        scn = os.path.join(ol.SCENES_DIR, name + ".scn")
        a = ort.HostScene.load(scn, data_dir, 64, 36)
        b = ort.HostScene.load(scn, data_dir, 64, 36, bake_on_device=True)
        ma, mb = a.meshes(), b.meshes()
        assert len(ma) == len(mb) >= 1
        for (va, ia, mna, mxa, mata), (vb, ib, mnb, mxb, matb) in zip(ma, mb):
            assert np.array_equal(va.view(np.uint32), vb.view(np.uint32))
            assert np.array_equal(ia, ib) and mata == matb
            assert np.array_equal(mna.view(np.uint32), mnb.view(np.uint32)) and np.array_equal(mxa.view(np.uint32), mxb.view(np.uint32))
        # and the scenes built from them are the same scene
        sa, sb = ort.Scene(a.world, a.root, 0), ort.Scene(b.world, b.root, 0)
        na, pa = sa.download(); nb, pb = sb.download()
        assert np.array_equal(na, nb) and np.array_equal(pa, pb)
        sa.close(); sb.close(); a.close(); b.close()


def test_mesh_bake_box_quirks_on_the_device(ort):
    """the fold starts at FLT_MIN = smallest POSITIVE float, so a mesh in negative space keeps max = 1.18e-38
    (ray.cpp:1761, macos_main.mm:383); among equal values the later vertex wins ((a < b) ? a : b, types.h:50-51),
    which decides the sign of a zero bound; an empty mesh returns the start values"""
    flt_min = np.float32(1.17549435e-38)
    v = np.array([[-1.0, -2.0, -3.0], [-0.5, -0.25, -4.0], [-2.0, -1.0, -0.125]], np.float32)
    out, mn, mx = ort.bake_mesh(v, 1.0, 0.0, (0, 0, 0, 1), (0, 0, 0))
    assert np.array_equal(out, v)
    assert np.array_equal(mn, np.array([-2.0, -2.0, -4.0], np.float32)) and np.all(mx == flt_min)
    # a baked coordinate of -0 needs every term of the rotation row and the translation to be a negative zero:
    # x = 1*(-0) + 0*(-1) + 0*(-2) + (-0) = -0, while x = 1*(+0) + ... = +0
    z = np.array([[0.0, -1.0, -2.0], [-0.0, -1.0, -2.0], [3.0, -1.0, -2.0]], np.float32)
    out, mn, _ = ort.bake_mesh(z, 1.0, 0.0, (0, 0, 0, 1), (-0.0, 0, 0))
    assert list(out[:, 0].view(np.uint32)) == [0x00000000, 0x80000000, 0x40400000]
    assert mn.view(np.uint32)[0] == 0x80000000          # -0 came later than +0
    _, mn, _ = ort.bake_mesh(z[[1, 0, 2]], 1.0, 0.0, (0, 0, 0, 1), (-0.0, 0, 0))
    assert mn.view(np.uint32)[0] == 0x00000000          # +0 came later
    _, mn, mx = ort.bake_mesh(np.zeros((0, 3), np.float32), 2.0, 30.0, (0, 0, 0.707107, 0.707106), (1, 2, 3))
    assert np.all(mn == np.finfo(np.float32).max) and np.all(mx == flt_min)


def test_non_finite_geometry_is_rejected(ort, tmp_path):
    """NaN / Inf vertices cannot be placed in any tree: ort_scene_create* reports ORT_ERR_ARG on every path
    (octree hand-off, shape lists on the host, shape lists on the device) instead of building a wrong tree"""
    for bad in ("nan", "inf"):
        ply = tmp_path / ("bad_%s.ply" % bad)
        ply.write_text("ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n"
                       "element face 2\nproperty list uchar int vertex_indices\nend_header\n"
                       "0.0 0.0 0.0\n1.0 0.0 0.0\n0.0 1.0 0.0\n1.0 1.0 0.5\n3 0 1 2\n3 1 3 2\n")
        scn = tmp_path / ("bad_%s.scn" % bad)
        scn.write_text("camera 0.0 0.0 5.0 b 0.2 q 1.0 0.0 0.0 0.0\nbrdf 0.5 0.5 0.5 0.0 0.0 0.0 10\n"
                       "mesh bad_%s.ply 0.0 0.0 0.0 1.0 q 1.0 0.0 0.0 0.0\n" % bad)
        hs = ort.HostScene.load(str(scn), str(tmp_path), 32, 18, octree=False)
        # poison one baked vertex in place (the structs are the data contract: host pointers)
        import ctypes as C2
        n = C2.c_uint32(0)
        m = C2.cast(ort.lib().ort_host_scene_meshes(hs.h, C2.byref(n)), C2.POINTER(ort.api.Mesh))[0]
        C2.cast(m.vertices, C2.POINTER(C2.c_float))[4] = float(bad)
        for on_device in (False, True):
            with pytest.raises(ort.OrtError, match="non-finite"):
                ort.Scene.from_lists(hs.world, hs.lists(), 0, build_on_device=on_device)
        hs.close()
