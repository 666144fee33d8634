"""A/B timing of kernel variants (development).  python tools/variant_bench.py [lib.so ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import offline_raytracer_b200 as ort  # noqa: E402

DATA = os.path.join(ROOT, "oracle", "_ref", "data")
SCENES = [("c3", "c3_bunny_box", 1920, 1080, 64), ("c4", "c4_dwarf_hdr", 3840, 2160, 16), ("c5lite", "c5_bunny_grid_64", 1920, 1080, 32)]
if os.environ.get("VB_LONG"):          # steady state dominates: what the benchmark's 6.8 s frames look like
    SCENES = [("c3", "c3_bunny_box", 1920, 1080, 256), ("c4", "c4_dwarf_hdr", 3840, 2160, 128), ("c5lite", "c5_bunny_grid_64", 1920, 1080, 128)]
if os.environ.get("VB_SCENES"):
    SCENES = [s for s in SCENES if s[0] in os.environ["VB_SCENES"].split(",")]
libs = sys.argv[1:] or [ort.LIB_PATH]
hosts = {}
for tag, name, w, h, spp in SCENES:
    hosts[tag] = ort.HostScene.load(os.path.join(ROOT, "scenes", name + ".scn"), DATA, w, h)
ref_img = {}
for path in libs:
    L = ort.lib(os.path.abspath(path))
    row = {"lib": os.path.basename(path)}
    for tag, name, w, h, spp in SCENES:
        hs = hosts[tag]
        sc = ort.Scene(hs.world, hs.root, 0, library=L)
        P = ort.default_params(w, h, spp, chunk_spp=16, kernel=2)
        os.environ["ORT_WF_POOLS"] = "1"; os.environ["ORT_WF_TIMING"] = "1"
        sc.render(hs.camera, P)
        img, st = sc.render(hs.camera, P)
        del os.environ["ORT_WF_POOLS"]; del os.environ["ORT_WF_TIMING"]
        best = 1e30
        for _ in range(2):
            img, st2 = sc.render(hs.camera, P)
            best = min(best, st2["device_ms"])
        import numpy as np
        chk = int(img.view(np.uint32).astype(np.uint64).sum())
        same = ref_img.setdefault(tag, chk) == chk
        row[tag] = {"ext": round(st["extend_ms"], 1), "sort": round(st["sort_ms"], 1), "shade": round(st["shade_ms"], 1),
                    "1pool": round(st["device_ms"], 1), "2pool": round(best, 1), "Ms/s": round(st2["samples"] / best / 1e3), "same_img": same}
        sc.close()
    print(json.dumps(row), flush=True)
