"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref.so,
built from /root/reference by oracle/Makefile).  Run here, in the build container:

    python tests/golden/make_golden.py

The vectors pin the restated oracle (oracle/oracle.cpp) and the product's host code
on boxes where /root/reference does not exist.  Everything is seeded; the scene used
is the repo's own scenes/box_spheres.scn (no reference data needed to replay).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Ref, SCENES_DIR, make_incoherent_rays, make_primary_rays  # noqa: E402


def unit(v):
    return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(np.float32)


def main():
    ref = Ref()
    rng = np.random.default_rng(20261018)
    out = {}

    # ---- RNG (code/random.h) ----
    seeds = np.array([12345, 1, 0xdeadbeef, 0x6D2B79F5, 777], np.uint32)
    states = np.zeros((len(seeds), 64), np.uint32)
    for i, s in enumerate(seeds):
        x = int(s)
        for k in range(64):
            x = ref.xor_shift_32(x)
            states[i, k] = x
    out["rng_seeds"], out["rng_states"] = seeds, states
    f01 = np.zeros((len(seeds), 8), np.float32); fbt = np.zeros((len(seeds), 8), np.float32)
    u32 = np.zeros((len(seeds), 8), np.uint32)
    for i, s in enumerate(seeds):
        st = int(s)
        for k in range(8):
            f01[i, k], st = ref.random_between_0_1(st)
        st = int(s)
        for k in range(8):
            fbt[i, k], st = ref.random_between(st, 0.0, 6.2831855)
        st = int(s)
        for k in range(8):
            u32[i, k], st = ref.random_between_u32(st, 0, 12)
    out["rng_f01"], out["rng_between"], out["rng_u32_12"] = f01, fbt, u32

    # ---- intersectors (code/ray.cpp:54-352): random + adversarial inputs ----
    n = 3000
    o = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    d = unit(rng.normal(size=(n, 3)))
    v0 = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    v1 = (v0 + rng.normal(scale=0.7, size=(n, 3))).astype(np.float32)
    v2 = (v0 + rng.normal(scale=0.7, size=(n, 3))).astype(np.float32)
    # aim 2/3 of the rays at the triangle (interior, edges, vertices)
    bary = rng.dirichlet((1, 1, 1), n).astype(np.float32)
    bary[n // 3: n // 2] = np.round(bary[n // 3: n // 2])            # vertices
    bary[n // 2: 2 * n // 3, 2] = 0                                   # edges
    bary[n // 2: 2 * n // 3] /= bary[n // 2: 2 * n // 3].sum(1, keepdims=True)
    target = bary[:, :1] * v0 + bary[:, 1:2] * v1 + bary[:, 2:] * v2
    d[: 2 * n // 3] = unit(target - o)[: 2 * n // 3]
    d[-50:, 0] = 0.0                                                  # axis-parallel components
    d[-25:, 1] = 0.0
    d[-50:] = unit(d[-50:] + 1e-30)
    tri = np.stack([ref.intersect("triangle", v0[i], v1[i], v2[i], o[i], d[i]) for i in range(n)])
    out.update(tri_v0=v0, tri_v1=v1, tri_v2=v2, tri_o=o, tri_d=d, tri_out=tri)

    c = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    r = rng.uniform(0.05, 1.5, n).astype(np.float32)
    ds = d.copy()
    # half aimed near the sphere, including tangent rays (|root| < 1e-5, ray.cpp:174)
    off = unit(rng.normal(size=(n, 3))) * (r * rng.choice([0.0, 0.5, 0.999999, 1.0, 1.000001, 1.2], n))[:, None]
    ds[: n // 2] = unit(c + off - o)[: n // 2]
    os_ = o.copy()
    os_[n // 2: n // 2 + 300] = (c + unit(rng.normal(size=(n, 3))) * r[:, None] * 0.5)[n // 2: n // 2 + 300]   # inside
    sph = np.stack([ref.intersect("sphere", c[i], float(r[i]), os_[i], ds[i]) for i in range(n)])
    out.update(sph_c=c, sph_r=r, sph_o=os_, sph_d=ds, sph_out=sph)

    mn = rng.uniform(-2, 1, (n, 3)).astype(np.float32)
    mx = (mn + rng.uniform(0.01, 2, (n, 3))).astype(np.float32)
    db = d.copy()
    tgt = mn + (mx - mn) * rng.choice([0.0, 0.5, 1.0, rng.random()], (n, 3)).astype(np.float32)
    db[: 2 * n // 3] = unit(tgt - o)[: 2 * n // 3]
    ob = o.copy()
    ob[-300:] = (mn + (mx - mn) * rng.random((n, 3)))[-300:].astype(np.float32)                               # inside
    db[-100:, 2] = 0.0
    box = np.stack([ref.intersect("aab", mn[i], mx[i], ob[i], db[i]) for i in range(n)])
    inr = np.array([ref.in_rect(ob[i], mn[i], mx[i]) for i in range(n)], np.int32)
    out.update(box_min=mn, box_max=mx, box_o=ob, box_d=db, box_out=box, box_in_rect=inr)

    base = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    axis = rng.normal(scale=1.0, size=(n, 3)).astype(np.float32)
    axis[:200] = [0, 0, 1.5]; axis[200:300] = [0, 0, -1.5]; axis[300:400] = [0, 2.0, 0]; axis[400:500] = [2.4, 0, 0]
    rc = rng.uniform(0.05, 0.6, n).astype(np.float32)
    dc = d.copy()
    tgt = base + axis * rng.random((n, 1)).astype(np.float32) + unit(rng.normal(size=(n, 3))) * (rc * rng.choice([0.0, 0.7, 1.0], n))[:, None]
    dc[: 2 * n // 3] = unit(tgt - o)[: 2 * n // 3]
    cyl = np.stack([ref.intersect("cylinder", base[i], axis[i], float(rc[i]), o[i], dc[i]) for i in range(n)])
    out.update(cyl_base=base, cyl_axis=axis, cyl_r=rc, cyl_o=o, cyl_d=dc, cyl_out=cyl)

    # ---- BSDF (code/ray.cpp:825-1161) ----
    m = 1500
    mats = np.zeros((m, 10), np.float32)
    kinds = rng.integers(0, 5, m)
    for i in range(m):
        k = kinds[i]
        if k == 0: mats[i] = [0.6, 0.6, 0.6, 0, 0, 0, 0, 0, 0, 1.0]                    # diffuse
        elif k == 1: mats[i] = [0.2, 0.2, 0.2, 1, 1, 1, 0, 0, 0, 1.0]                  # glossy
        elif k == 2: mats[i] = [0, 0, 0, 0, 0, 0, 1, 1, 1, 1.4]                        # glass
        elif k == 3: mats[i] = [0, 0, 0, 0.2, 0, 0, 1, 0, 0, 1.2]                      # red glass
        else: mats[i] = np.concatenate([rng.random(9), [rng.uniform(1.0, 1.8)]])
    N = unit(rng.normal(size=(m, 3)))
    N[:60] = [0, 0, 1]; N[60:120] = [0, 0, -1]
    wo = unit(rng.normal(size=(m, 3)))
    wi = unit(rng.normal(size=(m, 3)))
    wi[: m // 3] = unit(2 * (wo * N).sum(1, keepdims=True) * N - wo + rng.normal(scale=0.01, size=(m, 3)))[: m // 3]   # near mirror
    st = rng.integers(1, 2 ** 32 - 1, m, dtype=np.uint64).astype(np.uint32)
    s_wi = np.zeros((m, 3), np.float32); s_t = np.zeros(m, np.int32); s_st = np.zeros(m, np.uint32)
    pdf = np.zeros(m, np.float32); ev = np.zeros((m, 3), np.float32)
    dist = rng.uniform(0.01, 5, m).astype(np.float32)
    for i in range(m):
        s_wi[i], s_t[i], s_st[i] = ref.sample_brdf(int(st[i]), N[i], wo[i], 0.01, mats[i])
        pdf[i] = ref.pdf_brdf(N[i], wi[i], wo[i], 0.01, mats[i])
        ev[i] = ref.eval_scattering(N[i], wi[i], wo[i], mats[i], 0.01, float(dist[i]))
    out.update(bsdf_mat=mats, bsdf_N=N, bsdf_wo=wo, bsdf_wi=wi, bsdf_state=st, bsdf_dist=dist,
               bsdf_sample_wi=s_wi, bsdf_sample_is_t=s_t, bsdf_sample_state=s_st, bsdf_pdf=pdf, bsdf_eval=ev)

    # ---- number tokens (eat_numeric, code/parser.cpp:158-250) ----
    toks = ["0.12794", "1.443985", "5.553228", "0.707107", "0.707106", "10", "200", "0.0174533", "2.942755",
            "-5.4335527188698052e-09", "1e+2", "3.5e-3", "0.000001", "123456789", "0.", ".5", "18.000000", "8.900000"]
    toks += ["%.*f" % (int(rng.integers(1, 10)), rng.uniform(0, 1000)) for _ in range(400)]
    toks += ["%.6e" % rng.uniform(0, 1) for _ in range(100)]
    is_f = np.zeros(len(toks), np.int32); bits = np.zeros(len(toks), np.uint32)
    for i, t in enumerate(toks):
        is_f[i], bits[i] = ref.eat_numeric(t.lstrip("-"))
    out.update(num_tokens=np.array([t.lstrip("-") for t in toks]), num_is_float=is_f, num_bits=bits)

    # ---- scene-level: scenes/box_spheres.scn through the reference ----
    W, H = 64, 36
    rs = ref.scene_load(os.path.join(SCENES_DIR, "box_spheres.scn"), SCENES_DIR, W, H)
    cnt = rs.counts()
    out["scene_counts"] = np.array([cnt[k] for k in ("spheres", "boxes", "cylinders", "materials", "meshes", "lights", "light_bytes", "nodes")], np.int64)
    out["scene_camera"] = rs.camera_array()
    o1, d1 = make_primary_rays(rs.camera_array(), 96, 54)
    o2, d2 = make_incoherent_rays(15000, [-2.9, -1.9, -2.9], [2.9, 1.9, 5.9], seed=5)
    O = np.concatenate([o1, o2]); D = np.concatenate([d1, d2])
    rc_ = rs.raycast(O, D)
    out.update(scene_ray_o=O, scene_ray_d=D, scene_ray_t=rc_["t"], scene_ray_mat=rc_["mat"], scene_ray_normal=rc_["normal"])
    img, tests = rs.render_pixel_seeds(1234567, 8)
    out["scene_image_64x36x8"] = img
    tile = np.zeros((H, W, 3), np.float32)
    tile, _, st_after = rs.render_tile((8, 4, 24, 12), 4242, 4, out=tile)
    out["scene_tile_image"] = tile
    out["scene_tile_state_after"] = np.array([st_after], np.uint32)
    st = 99
    lights = []
    for k in range(32):
        st = rs.sample_random_lights(st)
        lights.append(st)
    out["scene_light_states"] = np.array(lights, np.uint32)

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
