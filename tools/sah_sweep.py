"""SAH parameter sweep (development): wide-node cost and traversal cost of the builders vs EXTEND time."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import offline_raytracer_b200 as ort
data = os.path.join(ROOT, "oracle", "_ref", "data")
jobs = [("c3_bunny_box", 1920, 1080, 64), ("c4_dwarf_hdr", 3840, 2160, 16), ("c5_bunny_grid_64", 1920, 1080, 32)]
hosts = {n: ort.HostScene.load(os.path.join(ROOT, "scenes", n + ".scn"), data, w, h) for n, w, h, _ in jobs}
for node_cost in ("0.5", "0.75", "1.0", "1.5", "2.5"):
    for trav in ("0.2", "0.3", "0.5"):
        os.environ["ORT_BVH_NODE_COST"] = node_cost; os.environ["ORT_BVH_TRAVERSAL_COST"] = trav
        row = {"node_cost": node_cost, "traversal_cost": trav}
        for name, w, h, spp in jobs:
            hs = hosts[name]
            sc = ort.Scene(hs.world, hs.root, 0)
            P = ort.default_params(w, h, spp, chunk_spp=16, kernel=2)
            os.environ["ORT_WF_POOLS"] = "1"; os.environ["ORT_WF_TIMING"] = "1"
            sc.render(hs.camera, P)
            _, st = sc.render(hs.camera, P)
            del os.environ["ORT_WF_POOLS"]; del os.environ["ORT_WF_TIMING"]
            best = min(sc.render(hs.camera, P)[1]["device_ms"] for _ in range(2))
            row[name[:6]] = {"nodes": sc.info()["bvh_node_count"], "ext": round(st["extend_ms"], 1), "2pool": round(best, 1)}
            sc.close()
        print(json.dumps(row), flush=True)
