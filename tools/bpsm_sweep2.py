import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import offline_raytracer_b200 as ort
data = os.path.join(ROOT, "oracle", "_ref", "data")
for name, w, h, spp in (("c3_bunny_box", 1920, 1080, 64), ("c4_dwarf_hdr", 3840, 2160, 16)):
    hs = ort.HostScene.load(os.path.join(ROOT, "scenes", name + ".scn"), data, w, h)
    sc = ort.Scene(hs.world, hs.root, 0)
    P = ort.default_params(w, h, spp, chunk_spp=16, kernel=2)
    for slots in (6, 9, 12):
        row = {"scene": name, "slots_Mi": slots}
        for pools in (2, 3, 4):
            for bpsm in (8, 4, 3, 2):
                os.environ["ORT_WF_POOLS"] = str(pools); os.environ["ORT_WF_SLOTS"] = str(slots << 20)
                if bpsm != 8: os.environ["ORT_WF_EXTEND_BPSM"] = str(bpsm)
                else: os.environ.pop("ORT_WF_EXTEND_BPSM", None)
                sc.render(hs.camera, P)
                row["p%d_b%d" % (pools, bpsm)] = round(min(sc.render(hs.camera, P)[1]["device_ms"] for _ in range(2)), 1)
        print(json.dumps(row), flush=True)
    sc.close()
