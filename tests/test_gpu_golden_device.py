"""The reference's golden vectors run ON THE DEVICE (tests/golden/reference_vectors.npz, recorded from the
unmodified reference by tests/golden/make_golden.py): the device build of the four intersectors
(csrc/core_math.h; code/ray.cpp:63-352) and of the BSDF (csrc/path.h; code/ray.cpp:825-1161), through
the C ABI's differential harness ort_selftest_intersect / ort_selftest_bsdf -- the very functions the
render kernels inline, compiled by nvcc for sm_100a with -fmad=false.

Bars:
  * intersectors -- 4 x 3000 cases incl. tangent / inside / axis-parallel / vertex- and edge-aimed rays:
    hit distance t, the unnormalised normal and the inner-hit flag are BIT-EXACT (the arithmetic is
    + - * / sqrt only, IEEE on both sides).
  * BSDF -- 1500 tuples: the RNG state after sample_brdf and the lobe decision are exact; values that pass
    through libm differ by what CUDA's sinf / cosf / atan2f / logf / expf and the device's powf -> product
    substitution (path.h:137-145) differ from glibc's: bounds stated at each assert.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def check_isect(dev, ref, name):
    # golden rows: t, n.x, n.y, n.z, inner (tests/oracle_lib.py: intersect)
    hit = ref[:, 0] >= 0
    assert np.array_equal(bits(dev[:, 0]), bits(ref[:, 0])), "%s: t differs on %d cases" % (name, (bits(dev[:, 0]) != bits(ref[:, 0])).sum())
    assert np.array_equal(bits(dev[hit, 1:4]), bits(ref[hit, 1:4])), name + ": normal"
    assert np.array_equal(dev[:, 4], ref[:, 4]), name + ": inner flag"
    return int(hit.sum())


def test_triangle_golden_on_device(ort, gold):
    cases = np.concatenate([gold["tri_v0"], gold["tri_v1"], gold["tri_v2"], gold["tri_o"], gold["tri_d"]], axis=1)
    n_hit = check_isect(ort.selftest_intersect("triangle", cases), gold["tri_out"], "triangle")
    assert n_hit > 500          # the aimed cases do hit


def test_sphere_golden_on_device(ort, gold):
    cases = np.concatenate([gold["sph_c"], gold["sph_r"][:, None], gold["sph_o"], gold["sph_d"]], axis=1)
    ref = gold["sph_out"]
    n_hit = check_isect(ort.selftest_intersect("sphere", cases), ref, "sphere")
    assert n_hit > 500 and (ref[:, 4] == 1).sum() > 100        # inner hits are exercised


def test_box_golden_on_device(ort, gold):
    cases = np.concatenate([gold["box_min"], gold["box_max"], gold["box_o"], gold["box_d"]], axis=1)
    assert check_isect(ort.selftest_intersect("aab", cases), gold["box_out"], "box") > 500


def test_cylinder_golden_on_device(ort, gold):
    cases = np.concatenate([gold["cyl_base"], gold["cyl_axis"], gold["cyl_r"][:, None], gold["cyl_o"], gold["cyl_d"]], axis=1)
    assert check_isect(ort.selftest_intersect("cylinder", cases), gold["cyl_out"], "cylinder") > 300


def test_intersectors_random_cases_vs_oracle_on_device(ort, oracle):
    """fresh seeded cases beyond the golden file, oracle (pinned to the reference) as the checker"""
    rng = np.random.default_rng(77)
    n = 4000
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    v0 = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    v1 = (v0 + rng.normal(scale=0.5, size=(n, 3))).astype(np.float32)
    v2 = (v0 + rng.normal(scale=0.5, size=(n, 3))).astype(np.float32)
    w = rng.dirichlet((1, 1, 1), n).astype(np.float32)
    d = (w[:, :1] * v0 + w[:, 1:2] * v1 + w[:, 2:] * v2 - o)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[::7] *= np.float32(2.5)                                   # non-unit directions: t in units of |d|
    dev = ort.selftest_intersect("triangle", np.concatenate([v0, v1, v2, o, d], axis=1))
    ref = np.stack([oracle.intersect("triangle", v0[i], v1[i], v2[i], o[i], d[i]) for i in range(n)])
    check_isect(dev, ref, "triangle/random")
    assert ort.selftest_intersect("triangle", np.zeros((0, 15), np.float32)).shape == (0, 5)     # empty batch


def test_rng_golden_on_device(ort, gold):
    """a11: xor_shift_32 (third shift is a RIGHT shift), random_between_0_1 = x / 2^32 in [0, 1], random_between (two
    steps), random_between_u32 -- states and float bits equal to the reference's, on the device"""
    st, f01, bt, u = ort.selftest_rng(gold["rng_seeds"], 6.2831855, 12)
    assert np.array_equal(st, gold["rng_states"])
    assert np.array_equal(bits(f01), bits(gold["rng_f01"]))
    assert np.array_equal(bits(bt), bits(gold["rng_between"]))
    assert np.array_equal(u, gold["rng_u32_12"])
    # the survey's known answer and the absorbing state
    st, f01, _, _ = ort.selftest_rng([12345, 0])
    assert list(st[0, :4]) == [104278947, 3831047122, 3324124125, 2171811514]
    assert abs(float(f01[0, 0]) - 0.0242793337) < 1e-9 and not st[1].any()


def test_light_pick_rng_golden_on_device(ort, gold):
    """a10: sample_random_lights' only surviving effect is on the RNG -- one step to pick the entry, four more when the
    entry is a sphere (ray.cpp:537-601); the device consumes the reference's stream exactly"""
    import oracle_lib as ol
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, 64, 36)
    lights = ort.world_light_is_sphere(hs.world)
    assert len(lights) == int(gold["scene_counts"][5]) and lights.any() and not lights.all()
    out = ort.selftest_light_pick(lights, 99, 32)
    assert np.array_equal(out, gold["scene_light_states"])
    # no light at all: one step, no pick (the reference would divide by zero, random.h:80)
    st, _, _, _ = ort.selftest_rng([99])
    assert ort.selftest_light_pick(np.zeros(0, np.uint8), 99, 1)[0] == st[0, 0]
    hs.close()


def rel_err(a, b, floor):
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(np.abs(b.astype(np.float64)), floor)


def test_bsdf_golden_on_device(ort, gold):
    g = ort.selftest_bsdf(gold["bsdf_mat"], gold["bsdf_N"], gold["bsdf_wo"], gold["bsdf_wi"], gold["bsdf_state"], gold["bsdf_dist"])
    n = len(gold["bsdf_state"])
    # integer side: exactly three xorshift steps, always (ray.cpp:1106-1108)
    assert np.array_equal(g["state_after"], gold["bsdf_sample_state"])
    # the lobe decision compares `choice` with sums of per-material constants formed by the same IEEE
    # operations, and the total-internal-reflection test is IEEE-only as well: exact
    assert np.array_equal(g["is_transmission"], gold["bsdf_sample_is_t"])
    # Stated bound for the libm side (CUDA's sinf / cosf / atan2f / logf / expf vs glibc's, and the device's
    # powf(x, 5) / powf(x, 4) / powf(e, y) -> products / expf substitution, path.h:137-145).  Measured on B200
    # over the 1500 tuples: sampled direction max 1.2e-7 absolute (1 ulp of a unit-vector component), pdf max
    # 3.9e-7 relative, f max 5.2e-7 relative (3-4 ulp).  Asserted: 4 ulp of 1.0 absolute for the direction,
    # 16 ulp relative (floor 1e-6 absolute) for pdf and f.
    err = np.abs(g["sample_wi"] - gold["bsdf_sample_wi"]).max(axis=1)
    assert err.max() <= 4 * 1.1920929e-07, err.max()
    assert np.allclose(np.linalg.norm(g["sample_wi"], axis=1), 1.0, atol=1e-6)
    e_pdf = rel_err(g["pdf"], gold["bsdf_pdf"], 1e-6)
    assert e_pdf.max() <= 16 * 1.1920929e-07, e_pdf.max()
    e_ev = rel_err(g["eval"], gold["bsdf_eval"], 1e-6).max(axis=1)
    assert e_ev.max() <= 16 * 1.1920929e-07, e_ev.max()
    # most tuples do not touch libm-sensitive terms at all: bit-equal
    assert (bits(g["pdf"]) == bits(gold["bsdf_pdf"])).mean() >= 0.8
    # zero stays zero: a lobe the reference does not evaluate is not evaluated here either
    z = gold["bsdf_pdf"] == 0
    assert np.all(g["pdf"][z] == 0)
    assert n == 1500
