"""Pool-size / pool-count sweep of the wavefront on one scene (development; env knobs of DESIGN.md 8)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import offline_raytracer_b200 as ort  # noqa: E402

scene, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
hs = ort.HostScene.load(os.path.join(ROOT, "scenes", scene + ".scn"), os.path.join(ROOT, "oracle", "_ref", "data"), w, h)
sc = ort.Scene(hs.world, hs.root, 0)
P = ort.default_params(w, h, spp, chunk_spp=16, kernel=2)
POOLS = [int(v) for v in os.environ.get("SWEEP_POOLS", "2,3").split(",")]
SLOTS = [float(v) for v in os.environ.get("SWEEP_SLOTS_MI", "4,6,8,12,16,24").split(",")]
for pools in POOLS:
    for slots in SLOTS:
        os.environ["ORT_WF_POOLS"] = str(pools); os.environ["ORT_WF_SLOTS"] = str(int(slots * (1 << 20)) & ~1023)
        sc.render(hs.camera, P)
        best = min(sc.render(hs.camera, P)[1]["device_ms"] for _ in range(2))
        print(json.dumps({"scene": scene, "pools": pools, "slots_Mi": slots, "ms": round(best, 1), "Msamples/s": round(w * h * spp / best / 1e3)}), flush=True)
