"""Small fixed render used as the target of `ncu --set full` (one k_render_mega launch)."""
import argparse
import os
os.environ.setdefault("ORT_WF_TIMING", "1")
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import offline_raytracer_b200 as ort  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default=os.path.join(ROOT, "scenes", "c3_bunny_box.scn"))
ap.add_argument("--base", default=os.path.join(ROOT, "oracle", "_ref", "data"))
ap.add_argument("--width", type=int, default=960)
ap.add_argument("--height", type=int, default=540)
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--chunk", type=int, default=16)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--lib", default="")
ap.add_argument("--build", default="", help="host | device (default: ort_scene_create's choice)")
ap.add_argument("--lists", action="store_true", help="hand the scene over as shape lists, without building the octree")
a = ap.parse_args()
import time
t0 = time.time()
hs = ort.HostScene.load(a.scene, a.base, a.width, a.height, octree=not a.lists)
t1 = time.time()
if a.lists:
    sc = ort.Scene.from_lists(hs.world, hs.lists(), 0, build_on_device=a.build == "device")
else:
    sc = ort.Scene(hs.world, hs.root, 0, library=ort.lib(os.path.abspath(a.lib)) if a.lib else None,
                   build_on_device={"": None, "host": False, "device": True}[a.build])
t2 = time.time()
inf, bs = sc.info(), sc.build_stats()
print("load %.2f s | scene_create %.2f s: collect %.2f prepare %.2f build %.2f (device kernels %.1f ms, %d PLOC rounds) | "
      "%d triangles, %d wide nodes, depth %d, %.1f MB on device" % (
          t1 - t0, t2 - t1, bs["collect_s"], bs["prepare_s"], bs["build_s"], bs["device_build_ms"], bs["ploc_iterations"],
          inf["triangle_count"], inf["bvh_node_count"], bs["wide_depth"], inf["device_bytes"] / 1e6))
P = ort.default_params(a.width, a.height, a.spp, chunk_spp=a.chunk, kernel=a.kernel)
for _ in range(a.reps):
    img, st = sc.render(hs.camera, P)
print("ms %.3f  Msamples/s %.1f  Mrays/s %.1f  rays/sample %.3f | extend %.2f sort %.2f shade %.2f ms, %d launches" % (
    st["device_ms"], st["samples"] / st["device_ms"] / 1e3, st["rays"] / st["device_ms"] / 1e3, st["rays"] / st["samples"],
    st["extend_ms"], st["sort_ms"], st["shade_ms"], st["kernel_launches"]))
import json
print("STATS " + json.dumps({"rays": st["rays"], "samples": st["samples"], "extend_launches": st.get("extend_launches", 0),
                            "extend_ms": st["extend_ms"], "shade_ms": st["shade_ms"], "sort_ms": st["sort_ms"], "device_ms": st["device_ms"]}))
