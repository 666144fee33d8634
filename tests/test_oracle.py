"""Pins the restated CPU oracle (oracle/oracle.cpp) to the reference.

Two sources of truth:
  * tests/golden/reference_vectors.npz -- outputs of the UNMODIFIED reference
    (oracle/_ref/libref.so) recorded by tests/golden/make_golden.py; always available;
  * the reference itself, live, when oracle/_ref was built on this box (`ref` fixture).
Everything on this path is IEEE f32 with the same libm, so the bar is BIT-EXACT.
"""
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_bits(a, b):
    """bit-equal, except that any NaN equals any NaN"""
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


# ---------------------------------------------------------------- RNG (a11)
def test_rng_known_answer_from_survey(oracle):
    # seed 12345 -> states and floats measured on the reference
    x = 12345
    states = []
    for _ in range(4):
        x = oracle.xor_shift_32(x)
        states.append(x)
    assert states == [104278947, 3831047122, 3324124125, 2171811514]
    f, _ = oracle.random_between_0_1(12345)
    assert abs(float(f) - 0.0242793337) < 1e-9
    assert oracle.xor_shift_32(0) == 0          # 0 is absorbing


def test_rng_golden(oracle):
    for i, s in enumerate(GOLD["rng_seeds"]):
        x = int(s)
        for k in range(64):
            x = oracle.xor_shift_32(x)
            assert x == int(GOLD["rng_states"][i, k])
        st = int(s)
        for k in range(8):
            f, st = oracle.random_between_0_1(st)
            assert bits(f) == bits(GOLD["rng_f01"][i, k])
        st = int(s)
        for k in range(8):
            f, st = oracle.random_between(st, 0.0, 6.2831855)
            assert bits(f) == bits(GOLD["rng_between"][i, k])
        st = int(s)
        for k in range(8):
            u, st = oracle.random_between_u32(st, 0, 12)
            assert u == int(GOLD["rng_u32_12"][i, k])


def test_stream_seed_matches_header(oracle):
    # ort_stream_seed is compiled into the oracle through include/ort_b200.h; the python
    # restatement in oracle_lib must agree (never 0)
    for base, pix, chunk in [(1234567, 0, 0), (1234567, 129599, 3), (0, 0, 0), (0xFFFFFFFF, 77, 9)]:
        assert ol.stream_seed(base, pix, chunk) != 0


# ---------------------------------------------------------------- intersectors (a3-a6)
@pytest.mark.parametrize("kind", ["tri", "sph", "box", "cyl"])
def test_intersectors_golden(oracle, kind):
    n = len(GOLD[kind + "_out"])
    hits = 0
    for i in range(n):
        if kind == "tri":
            out = oracle.intersect("triangle", GOLD["tri_v0"][i], GOLD["tri_v1"][i], GOLD["tri_v2"][i], GOLD["tri_o"][i], GOLD["tri_d"][i])
        elif kind == "sph":
            out = oracle.intersect("sphere", GOLD["sph_c"][i], float(GOLD["sph_r"][i]), GOLD["sph_o"][i], GOLD["sph_d"][i])
        elif kind == "box":
            out = oracle.intersect("aab", GOLD["box_min"][i], GOLD["box_max"][i], GOLD["box_o"][i], GOLD["box_d"][i])
            assert oracle.in_rect(GOLD["box_o"][i], GOLD["box_min"][i], GOLD["box_max"][i]) == int(GOLD["box_in_rect"][i])
        else:
            out = oracle.intersect("cylinder", GOLD["cyl_base"][i], GOLD["cyl_axis"][i], float(GOLD["cyl_r"][i]), GOLD["cyl_o"][i], GOLD["cyl_d"][i])
        assert same_bits(out, GOLD[kind + "_out"][i]), (kind, i, out, GOLD[kind + "_out"][i])
        hits += out[0] > 0
    assert hits > n // 10          # the vectors do exercise the hit branches


# ---------------------------------------------------------------- BSDF (a7-a9)
def test_bsdf_golden(oracle):
    g = GOLD
    for i in range(len(g["bsdf_pdf"])):
        wi, is_t, st = oracle.sample_brdf(int(g["bsdf_state"][i]), g["bsdf_N"][i], g["bsdf_wo"][i], 0.01, g["bsdf_mat"][i])
        assert same_bits(wi, g["bsdf_sample_wi"][i]) and is_t == int(g["bsdf_sample_is_t"][i]) and st == int(g["bsdf_sample_state"][i]), i
        p = oracle.pdf_brdf(g["bsdf_N"][i], g["bsdf_wi"][i], g["bsdf_wo"][i], 0.01, g["bsdf_mat"][i])
        assert same_bits(p, g["bsdf_pdf"][i]), i
        e = oracle.eval_scattering(g["bsdf_N"][i], g["bsdf_wi"][i], g["bsdf_wo"][i], g["bsdf_mat"][i], 0.01, float(g["bsdf_dist"][i]))
        assert same_bits(e, g["bsdf_eval"][i]), i
    assert int(g["bsdf_sample_is_t"].sum()) > 50      # refraction branch exercised


# ---------------------------------------------------------------- scene level (a1, a2, a10), golden
@pytest.fixture(scope="module")
def box_scene(ort, oracle):
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, 64, 36)
    return hs, oracle.scene(hs.world, hs.root)


def test_scene_counts_and_camera_golden(box_scene):
    hs, osc = box_scene
    assert same_bits(hs.camera_array(), GOLD["scene_camera"])
    info = osc.info()
    # spheres, boxes, cylinders (+1 inert CSG record) ; lights
    assert info["records"] == int(GOLD["scene_counts"][0] + GOLD["scene_counts"][1] + GOLD["scene_counts"][2]) + 1
    assert info["lights"] == int(GOLD["scene_counts"][5])


def test_raycast_golden(box_scene):
    hs, osc = box_scene
    r = osc.raycast(GOLD["scene_ray_o"], GOLD["scene_ray_d"], mode=0)
    assert same_bits(r["t"], GOLD["scene_ray_t"])
    assert np.array_equal(r["mat"], GOLD["scene_ray_mat"])
    assert same_bits(r["normal"], GOLD["scene_ray_normal"])
    assert (r["mat"] == 0).sum() == 0          # closed scene: no ray escapes


def test_image_golden(box_scene):
    hs, osc = box_scene
    img, cnt = osc.render(hs.camera, ol.default_params(64, 36, 8))
    assert same_bits(img, GOLD["scene_image_64x36x8"])
    assert cnt["rays"] > 64 * 36 * 8


def test_tile_with_shared_series_is_not_the_boundary_mode(box_scene, oracle):
    """The reference's own tile call shares ONE series between all pixels (ray.cpp:1178);
    the golden tile image pins that the reference was run that way, and shows that per-pixel
    seeding (the boundary's mode) is a different -- equally distributed -- stream."""
    hs, osc = box_scene
    tile = GOLD["scene_tile_image"]
    assert np.abs(tile[4:12, 8:24]).sum() > 0 and np.abs(tile[:4]).sum() == 0


def test_light_pick_rng_golden(box_scene):
    hs, osc = box_scene
    st = 99
    for k in range(32):
        st = osc.sample_random_lights(st)
        assert st == int(GOLD["scene_light_states"][k])


def test_bfs_equals_brute_force_on_golden_rays(box_scene):
    """the reference traversal == argmin over all records of (t, rank)"""
    hs, osc = box_scene
    a = osc.raycast(GOLD["scene_ray_o"], GOLD["scene_ray_d"], mode=0)
    b = osc.raycast(GOLD["scene_ray_o"], GOLD["scene_ray_d"], mode=1)
    assert same_bits(a["t"], b["t"]) and np.array_equal(a["rank"], b["rank"]) and np.array_equal(a["mat"], b["mat"])


# ---------------------------------------------------------------- live reference
def test_struct_sizes_match_reference(ref):
    s = ref.struct_sizes()
    assert list(s) == [20, 28, 32, 60, 48, 56, 24, 80, 80, 32, 4, 76]


def test_testscene_known_answers(ref, data_dir):
    # SURVEY.md 8c known-answer vectors
    rs = ref.scene_load(os.path.join(data_dir, "testscene.scn"), data_dir, 480, 270)
    c = rs.counts()
    assert (c["spheres"], c["boxes"], c["cylinders"], c["materials"], c["meshes"], c["lights"], c["light_bytes"], c["nodes"]) == \
           (7, 9, 11, 19, 2, 12, 144, 52257)


def test_oracle_equals_reference_on_testscene(ref, oracle, data_dir):
    W, H = 96, 54
    rs = ref.scene_load(os.path.join(data_dir, "testscene.scn"), data_dir, W, H)
    osc = oracle.scene(rs.world, rs.root)
    assert osc.info()["records"] == 138930
    o1, d1 = ol.make_primary_rays(rs.camera_array(), 320, 180)
    o2, d2 = ol.make_incoherent_rays(60000, [-2.9, -2.9, 0.0], [14.9, 14.9, 8.8])
    O = np.concatenate([o1, o2]); D = np.concatenate([d1, d2])
    a = rs.raycast(O, D)
    b = osc.raycast(O, D, mode=0)
    assert same_bits(a["t"], b["t"]) and np.array_equal(a["mat"], b["mat"])
    assert same_bits(a["normal"], b["normal"]) and np.array_equal(a["inner"], b["inner"])
    assert a["tests"] == b["shape_tests"]
    # whole image, per-pixel seeds, through the UNMODIFIED tiled_raytrace_bvh on 1x1 tiles
    img_ref, tests = rs.render_pixel_seeds(1234567, 4)
    img_o, cnt = osc.render(rs.camera, ol.default_params(W, H, 4))
    assert same_bits(img_ref, img_o)
    assert tests == cnt["shape_tests"]
    # the reference's own tile call (one shared series) == oracle on that tile? No: the boundary's
    # mode differs by design; but a 1x1 tile is both at once
    t_img, _, _ = rs.render_tile((10, 7, 11, 8), ol.stream_seed(1234567, 7 * W + 10, 0), 4)
    assert same_bits(t_img[7, 10], img_o[7, 10])


def test_golden_showcase_mean(ref, oracle, data_dir):
    """the only golden output of the reference: showcase/2.hdr (testscene, 1280x720x2048spp)"""
    path = os.path.join(ol.REF_DIR, "showcase", "2.hdr")
    if not os.path.exists(path):
        pytest.skip("showcase not staged")
    raw = open(path, "rb").read()
    hdr_end = raw.index(b"+X 1280\n") + len(b"+X 1280\n")
    px = np.frombuffer(raw[hdr_end:], np.uint8).reshape(720, 1280, 4).astype(np.float32)
    scale = np.where(px[..., 3] > 0, np.exp2(px[..., 3] - 136.0), 0.0)
    rgb = px[..., :3] * scale[..., None]
    gold_mean = rgb.mean((0, 1))
    rs = ref.scene_load(os.path.join(data_dir, "testscene.scn"), data_dir, 240, 135)
    osc = oracle.scene(rs.world, rs.root)
    img, _ = osc.render(rs.camera, ol.default_params(240, 135, 16))
    mean = img.mean((0, 1))
    assert np.all(np.abs(mean - gold_mean) / gold_mean < 0.03), (mean, gold_mean)
