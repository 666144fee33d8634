"""GPU parity on the BASELINE configurations beyond config 1 (BASELINE.json `configs`):

  C2  bunny.ply alone: explicit ray buffers generated ONCE by the oracle (reference xorshift, seeds 1 and 2:
      lens-sampled primaries of the reference camera, >= 50 % of them on the bunny; incoherent rays in the
      bunny's box inflated x2), 1 M rays each, CUDA path vs the oracle's restated raycast_bvh AND vs the
      unmodified reference's raycast_top_most_node: primitive rank equal on 100 % of the rays, t bit-equal
      (the spec's 1e-5 relative is asserted as well).
  C3  bunny in the closed box with shaped lights (the benchmarked scene), C4 dwarf.obj (OBJ loader, emitters
      > 1): explicit rays rank-exact + a reduced-size image against the oracle on identical per-pixel seeds
      (same criteria as tests/test_gpu_render.py: >= 99.5 % of pixels within 1e-4, RMSE <= 0.25 x seed-to-seed
      noise, mean luminance within 0.5 %).
  C5-lite  64 baked bunnies = 4.44 M triangles (exceeds L2; device-built BVH): a stated SUBSAMPLE of rays --
      the reference's depth-10 octree makes one CPU ray cost ~2 ms here -- rank-exact vs the oracle, the
      exhaustive GPU kernel on 20 k more, and a 24 x 14 image.
  +   the rays on which the CUDA path and the LIVE reference differ are isolated and counted: every one must be
      a hit the reference's culling lost ("definition B"), i.e. equal to the exhaustive answer.
"""
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

NCPU = os.cpu_count() or 8


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def lum(img):
    return float((img * np.array([0.2126, 0.7152, 0.0722], np.float32)).sum(-1).mean())


def rmse(a, b):
    return np.sqrt(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean((0, 1)))


def scene_path(name):
    return os.path.join(ol.SCENES_DIR, name + ".scn")


def assert_hits_equal(g, r, what, normal=True):
    assert np.array_equal(g["rank"], r["rank"]), "%s: %d ranks differ" % (what, (g["rank"] != r["rank"]).sum())
    assert np.array_equal(bits(g["t"]), bits(r["t"])), what
    hit = r["rank"] != 0xFFFFFFFF
    if hit.any():
        assert (np.abs(g["t"][hit] - r["t"][hit]) <= 1e-5 * r["t"][hit]).all()       # the tolerance the spec states
    assert np.array_equal(g["mat"], r["mat"]), what
    if normal:
        assert np.array_equal(bits(g["normal"]), bits(r["normal"])), what


def assert_hits_equal_or_reference_leak(g, r, osc, O, D, what, max_frac):
    """`r` is the reference's own traversal ("A").  Where its child-box cull is not conservative (ray.cpp:792-800:
    strict `<` of a box ENTRY t against the best hit so far -- a flat box around an axis-parallel quad enters at
    the very t of its triangle, so a hit one ulp nearer can be dropped) the CUDA path returns the exhaustive
    argmin over all records ("B").  Every differing ray must be exactly that: equal to the CPU
    brute force over all records, and NEARER than what the reference kept.  Returns the number of such rays."""
    diff = np.nonzero((bits(g["t"]) != bits(r["t"])) | (g["rank"] != r["rank"]))[0]
    assert len(diff) <= max_frac * len(O), (what, len(diff))
    if len(diff):
        b = osc.raycast(O[diff], D[diff], mode=1, threads=NCPU)
        assert np.array_equal(g["rank"][diff], b["rank"]) and np.array_equal(bits(g["t"][diff]), bits(b["t"])), what
        assert (g["t"][diff] < r["t"][diff]).all(), what
    same = np.ones(len(O), bool); same[diff] = False
    assert np.array_equal(g["mat"][same], r["mat"][same]) and np.array_equal(bits(g["normal"][same]), bits(r["normal"][same])), what
    return len(diff)


# ----------------------------------------------------------------------------------------------- C2
C2_N = 1_000_000


@pytest.fixture(scope="module")
def c2(ort, oracle, data_dir):
    W, H = 1920, 1080
    hs = ort.HostScene.load(scene_path("c2_bunny_only"), data_dir, W, H)
    sc = ort.Scene(hs.world, hs.root, 0)
    info = sc.info()
    # (i) coherent: the first C2_N cells of a 16:9 grid of 1333 x 750 lens-sampled primaries, seed 1
    grid = ol.default_params(1333, 750, 1)
    o1, d1 = oracle.make_camera_rays(hs.camera, grid, 1, C2_N)
    # (ii) incoherent: the bunny's box inflated x2 about its centre, seed 2
    lo, hi = np.array(info["root_min"]), np.array(info["root_max"])
    c, half = 0.5 * (lo + hi), (hi - lo)
    o2, d2 = oracle.make_random_rays(c - half, c + half, 2, C2_N)
    yield dict(hs=hs, sc=sc, coherent=(o1, d1), incoherent=(o2, d2))
    sc.close()


def test_c2_ray_buffers_vs_oracle(c2, oracle):
    osc = oracle.scene(c2["hs"].world, c2["hs"].root)
    for name in ("coherent", "incoherent"):
        O, D = c2[name]
        r = osc.raycast(O, D, mode=0, threads=NCPU)
        g = c2["sc"].raycast_batch(O, D)
        assert_hits_equal(g, r, "C2 " + name)
        frac = (r["rank"] != 0xFFFFFFFF).mean()
        if name == "coherent":
            assert frac >= 0.5, frac           # the camera is aimed so that most primaries hit the bunny
        else:
            assert 0.02 < frac < 0.9, frac
    osc.close()


def test_c2_ray_buffers_vs_live_reference(c2, ort, ref, data_dir):
    """the same buffers through the UNMODIFIED reference: its own loader, octree and raycast_top_most_node"""
    rs = ref.scene_load(scene_path("c2_bunny_only"), data_dir, 1920, 1080)
    sc = ort.Scene(rs.world, rs.root, 0)           # the drop-in case: the reference's own World / octree pointers
    for name in ("coherent", "incoherent"):
        O, D = c2[name]
        n = 500_000                                 # stated subsample of each 1 M buffer for the second CPU pass
        a = rs.raycast(O[:n], D[:n], threads=NCPU)
        g = sc.raycast_batch(O[:n], D[:n])
        assert np.array_equal(bits(g["t"]), bits(a["t"])), name
        assert np.array_equal(g["mat"], a["mat"]), name
        assert np.array_equal(bits(g["normal"]), bits(a["normal"])), name
        g2 = c2["sc"].raycast_batch(O[:n], D[:n])   # and the product's own loader gives the same scene
        assert np.array_equal(g["rank"], g2["rank"]) and np.array_equal(bits(g["t"]), bits(g2["t"]))
    sc.close()


def test_c2_device_generated_buffers(c2, ort, oracle):
    """bench.py generates its C2 buffers on the device (ort_generate_*_rays_device): same streams, same camera
    model -- they agree with the oracle's buffers up to libm (sinf / cosf), and the hits on them are checked
    against the oracle like any other explicit buffer"""
    import torch
    n = 200_000
    dev = torch.device("cuda:0")
    o = torch.empty((n, 3), device=dev); d = torch.empty((n, 3), device=dev)
    grid = ort.default_params(1333, 750, 1)
    ort.generate_camera_rays_device(c2["hs"].camera, grid, 1, n, o.data_ptr(), d.data_ptr())
    torch.cuda.synchronize()
    O, D = c2["coherent"]
    assert np.abs(o.cpu().numpy() - O[:n]).max() <= 2e-6 and np.abs(d.cpu().numpy() - D[:n]).max() <= 2e-6
    info = c2["sc"].info()
    lo, hi = np.array(info["root_min"]), np.array(info["root_max"])
    c, half = 0.5 * (lo + hi), (hi - lo)
    ort.generate_random_rays_device(c - half, c + half, 2, n, o.data_ptr(), d.data_ptr())
    torch.cuda.synchronize()
    O2, D2 = c2["incoherent"]
    assert np.array_equal(bits(o.cpu().numpy()), bits(O2[:n]))                   # origins: RNG + IEEE only
    assert np.abs(d.cpu().numpy() - D2[:n]).max() <= 2e-6
    osc = oracle.scene(c2["hs"].world, c2["hs"].root)
    oh, dh = o.cpu().numpy()[:50_000], d.cpu().numpy()[:50_000]
    assert_hits_equal(c2["sc"].raycast_batch(oh, dh), osc.raycast(oh, dh, mode=0, threads=NCPU), "device-generated")
    osc.close()


# ----------------------------------------------------------------------------------------------- C3 / C4
@pytest.mark.parametrize("name,w,h,spp", [("c3_bunny_box", 240, 135, 8), ("c4_dwarf_hdr", 256, 144, 8)])
def test_c3_c4_rays_and_image_vs_oracle(ort, oracle, data_dir, name, w, h, spp):
    hs = ort.HostScene.load(scene_path(name), data_dir, w, h)
    osc = oracle.scene(hs.world, hs.root)
    sc = ort.Scene(hs.world, hs.root, 0)
    info = sc.info()
    # explicit rays: primaries of the scene's camera + incoherent rays inside the room
    o1, d1 = oracle.make_camera_rays(hs.camera, ol.default_params(640, 360, 1), 1, 640 * 360)
    lo, hi = np.array(info["root_min"]), np.array(info["root_max"])
    o2, d2 = oracle.make_random_rays(lo + 0.2, hi - 0.2, 2, 300_000)
    # primaries: identical to the reference's traversal on every ray (the spec's "primary hit IDs bit-exact")
    assert_hits_equal(sc.raycast_batch(o1, d1), osc.raycast(o1, d1, mode=0, threads=NCPU), name + " primary")
    # incoherent rays reach the flat emissive quads of letterX.ply from all sides: there the reference's cull
    # drops a hit one ulp nearer on ~1e-4 of the rays (measured on the host build: 30 of 300 000 on C3, 0 on C4)
    leaks = assert_hits_equal_or_reference_leak(sc.raycast_batch(o2, d2), osc.raycast(o2, d2, mode=0, threads=NCPU),
                                                osc, o2, d2, name + " incoherent", 3e-4)
    if name == "c4_dwarf_hdr":
        assert leaks == 0
    # reduced-size image, identical per-pixel seeds (single chunk: float sums in sample order, ray.cpp:1428)
    img_a, cnt = osc.render(hs.camera, ol.default_params(w, h, spp, seed=1234567), threads=NCPU)
    img_b, _ = osc.render(hs.camera, ol.default_params(w, h, spp, seed=7654321), threads=NCPU)
    for kernel in (ort.ORT_KERNEL_WAVEFRONT, ort.ORT_KERNEL_MEGAKERNEL):
        g, st = sc.render(hs.camera, ort.default_params(w, h, spp, seed=1234567, kernel=kernel))
        assert st["samples"] == w * h * spp
        assert abs(st["rays"] - cnt["rays"]) <= 5e-4 * cnt["rays"]
        close = (np.abs(g - img_a) <= 1e-4 * np.maximum(np.abs(img_a), 1e-3)).all(axis=2).mean()
        assert close >= 0.995, (name, kernel, close)
        noise = rmse(img_a, img_b)
        assert np.all(rmse(g, img_a) <= 0.25 * noise), (rmse(g, img_a), noise)
        assert abs(lum(g) - lum(img_a)) <= 0.005 * lum(img_a)
        assert not np.isnan(g).any()
    # converged-image criterion on different seeds: RMSE <= 1.25 x noise, luminance within 2 %
    g2, _ = sc.render(hs.camera, ort.default_params(w, h, spp, seed=7654321))
    assert np.all(rmse(g2, img_a) <= 1.25 * rmse(img_a, img_b)) and abs(lum(g2) - lum(img_a)) <= 0.02 * lum(img_a)
    if name == "c4_dwarf_hdr":
        assert img_a.max() > 1.0 and g.max() > 1.0          # "HDR": unclamped radiance from emitters > 1
    sc.close(); osc.close()


def test_c4_rays_vs_live_reference(ort, ref, data_dir):
    """the OBJ path (dwarf.obj, v/vt/vn corners) through the unmodified reference's parser and octree"""
    rs = ref.scene_load(scene_path("c4_dwarf_hdr"), data_dir, 384, 216)
    sc = ort.Scene(rs.world, rs.root, 0)
    o1, d1 = ol.make_primary_rays(rs.camera_array(), 384, 216, seed=3)
    o2, d2 = ol.make_incoherent_rays(200_000, [-4.8, -4.8, 0.05], [8.8, 8.8, 6.9], seed=4)
    O, D = np.concatenate([o1, o2]), np.concatenate([d1, d2])
    a = rs.raycast(O, D, threads=NCPU)
    g = sc.raycast_batch(O, D)
    assert np.array_equal(bits(g["t"]), bits(a["t"])) and np.array_equal(g["mat"], a["mat"])
    assert np.array_equal(bits(g["normal"]), bits(a["normal"]))
    sc.close()


# ----------------------------------------------------------------------------------------------- C5-lite
def test_c5_lite_subsampled_rays_and_image_vs_oracle(ort, oracle, data_dir):
    import torch
    w, h = 24, 14
    hs = ort.HostScene.load(scene_path("c5_bunny_grid_64"), data_dir, w, h)
    osc = oracle.scene(hs.world, hs.root)
    sc = ort.Scene(hs.world, hs.root, 0)                  # 4.44 M records: ORT_BUILD_AUTO builds on the device
    assert sc.build_stats()["on_device"] == 1
    info = sc.info()
    assert info["triangle_count"] == 64 * 69451
    # stated subsample: 3000 primaries (stride over a 640 x 360 grid) + 3000 incoherent rays
    o1, d1 = oracle.make_camera_rays(hs.camera, ol.default_params(640, 360, 1), 1, 640 * 360)
    sel = np.arange(0, 640 * 360, 640 * 360 // 3000)[:3000]
    lo, hi = np.array(info["root_min"]), np.array(info["root_max"])
    o2, d2 = oracle.make_random_rays(lo + 0.2, hi - 0.2, 2, 3000)
    O, D = np.concatenate([o1[sel], o2]), np.concatenate([d1[sel], d2])
    r = osc.raycast(O, D, mode=0, threads=NCPU)
    assert_hits_equal(sc.raycast_batch(O, D), r, "C5-lite subsample")
    assert (r["rank"] != 0xFFFFFFFF).all()                # closed room
    # the host-built tree of the same scene answers identically (structure-free result)
    sc_host = ort.Scene(hs.world, hs.root, 0, build_on_device=False)
    g_host = sc_host.raycast_batch(O, D)
    assert np.array_equal(g_host["rank"], r["rank"]) and np.array_equal(bits(g_host["t"]), bits(r["t"]))
    sc_host.close()
    # 20 k more rays against the exhaustive GPU kernel (all 4.44 M records per ray)
    n = 20_000
    o3, d3 = oracle.make_random_rays(lo + 0.2, hi - 0.2, 5, n)
    dev = torch.device("cuda:0")
    to, td = torch.from_numpy(o3).to(dev), torch.from_numpy(d3).to(dev)
    ta, tb = torch.empty(n, device=dev), torch.empty(n, device=dev)
    ra, rb = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev)
    sc.raycast_batch_device(n, to.data_ptr(), td.data_ptr(), ta.data_ptr(), ra.data_ptr())
    sc.raycast_brute_device(n, to.data_ptr(), td.data_ptr(), tb.data_ptr(), rb.data_ptr())
    torch.cuda.synchronize()
    assert bool((ra == rb).all()) and bool((ta.view(torch.int32) == tb.view(torch.int32)).all())
    # a 24 x 14 x 2 spp image on identical seeds
    img, _ = osc.render(hs.camera, ol.default_params(w, h, 2), threads=NCPU)
    g, st = sc.render(hs.camera, ort.default_params(w, h, 2))
    close = (np.abs(g - img) <= 1e-4 * np.maximum(np.abs(img), 1e-3)).all(axis=2).mean()
    assert close >= 0.98, close                           # 336 pixels: one diverged stream is 0.3 %
    sc.close(); osc.close()


# ----------------------------------------------------------------------------------------------- definition B
def test_divergences_from_the_live_reference_are_only_its_culling_leaks(ort, ref, oracle, data_dir):
    """Bounce-like rays (origins 1e-4 in front of surfaces, ray.cpp:1262) are where the reference's child-box
    cull is not conservative (ray.cpp:792-800).  Count the rays on which the CUDA path differs from the LIVE
    reference; each must be a nearer hit that equals the exhaustive answer (the CUDA path never loses a hit the
    reference finds), and they must be rare."""
    rs = ref.scene_load(os.path.join(data_dir, "testscene.scn"), data_dir, 480, 270)
    sc = ort.Scene(rs.world, rs.root, 0)
    hs = ort.HostScene.load(os.path.join(data_dir, "testscene.scn"), data_dir, 480, 270)
    osc = oracle.scene(hs.world, hs.root)
    o1, d1 = ol.make_primary_rays(rs.camera_array(), 960, 540, n_per_pixel=3, seed=21)
    g1 = sc.raycast_batch(o1, d1)
    hit = g1["mat"] != 0
    rng = np.random.default_rng(22)
    O = (o1 + (g1["t"] - np.float32(1e-4))[:, None] * d1)[hit].astype(np.float32)
    D = rng.normal(size=O.shape)
    D = (D / np.linalg.norm(D, axis=1, keepdims=True)).astype(np.float32)
    a = rs.raycast(O, D, threads=NCPU)
    g = sc.raycast_batch(O, D)
    diff = np.nonzero(bits(g["t"]) != bits(a["t"]))[0]
    n = len(O)
    assert n > 1_400_000
    assert len(diff) <= 5e-5 * n, (len(diff), n)                     # measured on the host build: 21 of 1 555 200
    if len(diff):
        b = osc.raycast(O[diff], D[diff], mode=1)                    # exhaustive over all records, on the CPU
        assert np.array_equal(bits(g["t"][diff]), bits(b["t"])) and np.array_equal(g["rank"][diff], b["rank"])
        assert (g["t"][diff] < a["t"][diff]).all()                   # the reference lost a NEARER hit, never the reverse
    same = np.ones(n, bool); same[diff] = False
    assert np.array_equal(g["mat"][same], a["mat"][same]) and np.array_equal(bits(g["normal"][same]), bits(a["normal"][same]))
    sc.close(); osc.close()


# ----------------------------------------------------------------------------------------------- C5, full size
def test_c5_full_size_rays_vs_oracle_and_exhaustive_kernel(ort, oracle, data_dir):
    """BASELINE config 5 at FULL size: 729 baked bunnies = 50 629 779 triangles, 3.2 GB of nodes and records built on
    the device from the shape lists.  The reference itself cannot load this scene (100-mesh limit, fixed arenas), but its
    algorithm can be run on it: the product's host code builds the reference's depth-10 octree (bit-identical to the
    reference's on every scene both can load, tests/test_host_scene.py) and the oracle -- pinned to the reference --
    traverses it.  A CPU ray costs ~10 ms on that octree, so the subsample is stated: 320 primaries + 320 incoherent rays
    rank- and t-exact against the oracle; 20 000 more against the exhaustive GPU kernel (all 50.6 M records per ray)."""
    import torch
    scn = scene_path("c5_bunny_grid_729")
    hs = ort.HostScene.load(scn, data_dir, 3840, 2160)                    # with the octree: ~20 s, the oracle needs it
    sc = ort.Scene.from_lists(hs.world, hs.lists(), 0)                    # records, ranks and BVH on the device
    info = sc.info()
    assert info["triangle_count"] == 729 * 69451 and sc.build_stats()["on_device"] == 1
    osc = oracle.scene(hs.world, hs.root)
    o1, d1 = oracle.make_camera_rays(hs.camera, ol.default_params(640, 360, 1), 1, 640 * 360)
    sel = np.arange(0, 640 * 360, 640 * 360 // 320)[:320]
    lo, hi = np.array(info["root_min"]), np.array(info["root_max"])
    o2, d2 = oracle.make_random_rays(lo + 0.2, hi - 0.2, 2, 320)
    O, D = np.concatenate([o1[sel], o2]), np.concatenate([d1[sel], d2])
    r = osc.raycast(O, D, mode=0, threads=NCPU)
    assert_hits_equal(sc.raycast_batch(O, D), r, "C5 full subsample")
    assert (r["rank"] != 0xFFFFFFFF).all()
    osc.close()
    n = 20_000
    o3, d3 = oracle.make_random_rays(lo + 0.2, hi - 0.2, 5, n)
    dev = torch.device("cuda:0")
    to, td = torch.from_numpy(o3).to(dev), torch.from_numpy(d3).to(dev)
    ta, tb = torch.empty(n, device=dev), torch.empty(n, device=dev)
    ra, rb = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev)
    sc.raycast_batch_device(n, to.data_ptr(), td.data_ptr(), ta.data_ptr(), ra.data_ptr())
    sc.raycast_brute_device(n, to.data_ptr(), td.data_ptr(), tb.data_ptr(), rb.data_ptr())
    torch.cuda.synchronize()
    assert bool((ra == rb).all()) and bool((ta.view(torch.int32) == tb.view(torch.int32)).all())
    sc.close(); hs.close()
