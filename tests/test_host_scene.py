"""Host side of the drop-in (csrc/host_loader.cpp, host_scene.cpp) and the C ABI surface.
No GPU needed: nothing here launches a kernel."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as ol

ROOT = ol.ROOT
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_abi_exports_every_declared_symbol(ort):
    """every ort_* function declared in include/ort_b200.h is exported by the shared library"""
    hdr = open(os.path.join(ROOT, "include", "ort_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ort_[a-z0-9_]+)\s*\(", hdr)) - {"ort_stream_seed"}
    assert len(declared) >= 25
    L = ort.lib()
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing


def test_struct_layouts_match_header(ort):
    assert C.sizeof(ort.Camera) == 48
    assert C.sizeof(ort.RenderParams) == 80
    assert C.sizeof(ort.RenderStats) == 64


def test_number_tokens_golden(ort):
    """eat_numeric is not strtof: 0.12794 -> 0.127940014, 1.443985 -> 1.4439851"""
    for tok, is_f, b in zip(GOLD["num_tokens"], GOLD["num_is_float"], GOLD["num_bits"]):
        assert ort.parse_numeric(str(tok)) == (int(is_f), int(b)), tok
    is_f, b = ort.parse_numeric("0.12794")
    assert is_f == 1 and abs(np.array([b], np.uint32).view(np.float32)[0] - 0.127940014) < 1e-9


def test_number_tokens_vs_reference(ort, ref):
    rng = np.random.default_rng(3)
    for _ in range(3000):
        k = rng.integers(0, 4)
        if k == 0: s = "%d" % rng.integers(0, 10 ** 9)
        elif k == 1: s = "%.*f" % (int(rng.integers(0, 12)), rng.uniform(0, 10 ** rng.integers(0, 6)))
        elif k == 2: s = "%.7e" % rng.uniform(0, 1)
        else: s = "0.%0*d" % (int(rng.integers(1, 14)), rng.integers(0, 9999))
        assert ort.parse_numeric(s) == ref.eat_numeric(s), s


@pytest.mark.parametrize("name,nv,ni", [("bunny.ply", 35947, 208353), ("dwarf.obj", 979, 5688),
                                        ("letterX.ply", 8, 12), ("letterY.ply", 12, 18)])
def test_mesh_loaders_vs_reference(ort, ref, data_dir, name, nv, ni):
    """loader facts of SURVEY.md 8c + bit-equality with the reference parsers"""
    v, i = ort.load_mesh(os.path.join(data_dir, name))
    assert v.shape == (nv, 3) and i.shape == (ni,)
    rv, ri = ref.load_mesh(os.path.join(data_dir, name))
    assert np.array_equal(bits(v), bits(rv)) and np.array_equal(i, ri)
    if name == "letterX.ply":
        assert list(i[:6]) == [3, 2, 1, 3, 1, 0]      # quad -> fan
    if name == "dwarf.obj":
        assert i.max() == 978


def test_scene_assembly_equals_reference(ort, ref, oracle, data_dir):
    """same records in the same rank order (=> same octree), same camera, same lights"""
    W, H = 480, 270
    hs = ort.HostScene.load(os.path.join(data_dir, "testscene.scn"), data_dir, W, H)
    rs = ref.scene_load(os.path.join(data_dir, "testscene.scn"), data_dir, W, H)
    a, b = oracle.scene(rs.world, rs.root), oracle.scene(hs.world, hs.root)
    assert a.info() == b.info()
    assert a.info() == dict(records=138930, nodes=25578, max_depth=10, lights=12)
    ta, ga = a.records(); tb, gb = b.records()
    assert np.array_equal(ta, tb) and np.array_equal(bits(ga), bits(gb))
    assert np.array_equal(bits(rs.camera_array()), bits(hs.camera_array()))
    st = 4711
    for _ in range(64):
        assert a.sample_random_lights(st) == b.sample_random_lights(st)
        st = a.sample_random_lights(st)


def test_scene_assembly_own_scene_vs_reference(ort, ref, oracle):
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, 64, 36)
    rs = ref.scene_load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, 64, 36)
    a, b = oracle.scene(rs.world, rs.root), oracle.scene(hs.world, hs.root)
    ta, ga = a.records(); tb, gb = b.records()
    assert np.array_equal(ta, tb) and np.array_equal(bits(ga), bits(gb))


def test_scene_without_csg_drops_one_rank(ort, oracle):
    a = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, 64, 36, with_csg=True)
    b = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, 64, 36, with_csg=False)
    assert oracle.scene(a.world, a.root).info()["records"] == oracle.scene(b.world, b.root).info()["records"] + 1


def test_loader_errors_are_reported_not_fatal(ort, tmp_path):
    with pytest.raises(ort.OrtError, match="cannot open"):
        ort.HostScene.load(str(tmp_path / "missing.scn"), str(tmp_path), 32, 32)
    bad = tmp_path / "bad.scn"
    bad.write_text("camera 1 2 3 b 0.2 q 1.0 0.0 0.0 0.0\n")     # integers where floats are required
    with pytest.raises(ort.OrtError, match="expected a float"):
        ort.HostScene.load(str(bad), str(tmp_path), 32, 32)
    nomesh = tmp_path / "nomesh.scn"
    nomesh.write_text("brdf 0.5 0.5 0.5 0.0 0.0 0.0 10\nmesh nothere.ply 0.0 0.0 0.0 1.0 q 1.0 0.0 0.0 0.0\n")
    with pytest.raises(ort.OrtError, match="cannot open"):
        ort.HostScene.load(str(nomesh), str(tmp_path) + "/", 32, 32)
    with pytest.raises(ort.OrtError):
        ort.load_mesh(str(tmp_path / "x.stl"))
    # hostile counts are parse errors, not allocations: a face arity / vertex count the file cannot hold,
    # and a face line that runs out of indices
    hdr = ("ply\nformat ascii 1.0\nelement vertex %s\nproperty float x\nproperty float y\nproperty float z\n"
           "element face 1\nproperty list uchar int vertex_indices\nend_header\n")
    f = tmp_path / "arity.ply"
    f.write_text(hdr % "3" + "0 0 0\n1 0 0\n0 1 0\n2000000000 0 1 2\n")
    with pytest.raises(ort.OrtError, match="arity"):
        ort.load_mesh(str(f))
    f = tmp_path / "count.ply"
    f.write_text(hdr % "2000000000" + "0 0 0\n")
    with pytest.raises(ort.OrtError, match="vertex count"):
        ort.load_mesh(str(f))
    f = tmp_path / "short.ply"
    f.write_text(hdr % "3" + "0 0 0\n1 0 0\n0 1 0\n4 0 1 2 1.5\n")
    with pytest.raises(ort.OrtError, match="malformed face"):
        ort.load_mesh(str(f))


def test_ragged_and_empty_meshes(ort, tmp_path):
    p = tmp_path / "tri.ply"
    p.write_text("ply\nformat ascii 1.0\nelement vertex 5\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face 2\nproperty list uchar int vertex_indices\nend_header\n"
                 "0 0 0\n1 0 0\n1 1 0\n0 1 0\n0.5 2 0\n3 0 1 2\n5 0 1 2 3 4\n")
    v, i = ort.load_mesh(str(p))
    assert v.shape == (5, 3) and list(i) == [0, 1, 2, 0, 1, 2, 0, 2, 3, 0, 3, 4]
    q = tmp_path / "quad.obj"
    q.write_text("# c\nv 0 0 0\nv 1.5 0 0\nv 1 1 -2\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\ns off\nf 1//1 3//1 4//1\n")
    v, i = ort.load_mesh(str(q))
    assert v.shape == (4, 3) and list(i) == [0, 1, 2, 0, 2, 3, 0, 2, 3] and v[2, 2] == -2.0
    e = tmp_path / "empty.ply"
    e.write_text("ply\nformat ascii 1.0\nelement vertex 0\nproperty float x\nproperty float y\nproperty float z\n"
                 "element face 0\nproperty list uchar int vertex_indices\nend_header\n")
    v, i = ort.load_mesh(str(e))
    assert v.shape == (0, 3) and i.shape == (0,)


def test_hdr_writer_layout(ort, tmp_path):
    """flat RGBE, '+Y h +X w' header, rows emitted top of the picture first (macos_main.mm:682-707)"""
    img = np.zeros((3, 2, 3), np.float32)
    img[0, 0] = [1.0, 0.5, 0.25]      # buffer row 0 = bottom of the picture
    img[2, 1] = [4.0, 4.0, 4.0]
    path = str(tmp_path / "o.hdr")
    ort.write_hdr(path, img)
    raw = open(path, "rb").read()
    head = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n+Y 3 +X 2\n"
    assert raw.startswith(head) and len(raw) == len(head) + 3 * 2 * 4
    px = np.frombuffer(raw[len(head):], np.uint8).reshape(3, 2, 4)
    assert list(px[0, 1]) == [128, 128, 128, 131]       # 4.0 = 0.5 * 2^3 -> mantissa 128, exponent 128+3
    assert list(px[2, 0]) == [128, 64, 32, 129]         # 1.0 = 0.5 * 2^1
    assert list(px[1, 0]) == [0, 0, 0, 0]
    assert ort.v3_to_rgbe([1.0, 0.5, 0.25]) == 128 | (64 << 8) | (32 << 16) | (129 << 24)


def test_hdr_writer_reproduces_showcase_bytes(ort, tmp_path):
    """decode showcase/2.hdr and re-encode it: the reference's own file must round-trip byte for byte"""
    src = os.path.join(ol.REF_DIR, "showcase", "2.hdr")
    if not os.path.exists(src):
        pytest.skip("showcase not staged")
    raw = open(src, "rb").read()
    head = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n+Y 720 +X 1280\n"
    assert raw.startswith(head)
    px = np.frombuffer(raw[len(head):], np.uint8).reshape(720, 1280, 4).astype(np.float32)
    # inverse of v3_to_rgbe (macos_main.mm:242-261): mantissa = round(c * 255 / 2^e), exponent byte = e + 128
    scale = np.where(px[..., 3] > 0, np.exp2(px[..., 3].astype(np.float64) - 128.0) / 255.0, 0.0)
    rgb = (px[..., :3].astype(np.float64) * scale[..., None]).astype(np.float32)[::-1]     # file is top-first; buffer row 0 = bottom
    out = str(tmp_path / "rt.hdr")
    ort.write_hdr(out, np.ascontiguousarray(rgb))
    again = open(out, "rb").read()
    assert again[:len(head)] == head
    a = np.frombuffer(again[len(head):], np.uint8).reshape(720, 1280, 4)
    b = np.frombuffer(raw[len(head):], np.uint8).reshape(720, 1280, 4)
    # a pixel whose largest mantissa is < 128 (or 255 = 1.0 * 2^e) is not in canonical form and re-normalises
    canonical = (b[..., :3].max(-1) >= 128) & (b[..., :3].max(-1) <= 254)
    assert canonical.mean() > 0.99
    assert np.array_equal(a[canonical], b[canonical])


def test_abi_headers_compile_as_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: include/ort_b200.h (and ort_scene.h) must be valid C99 on their own"""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "ort_b200.h"\n'
                   'int main(void) { OrtShapeLists l; OrtBuildStats b; OrtRenderParams p; OrtRenderStats s;\n'
                   '  (void)l; (void)b; (void)p; (void)s; return ort_stream_seed(1u, 2u, 3u) == 0u; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I" + os.path.join(ol.ROOT, "include"), str(src)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
