"""Small fixed render used as the target of `ncu --set full` (one k_render_mega launch)."""
import argparse
import os
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
a = ap.parse_args()
hs = ort.HostScene.load(a.scene, a.base, a.width, a.height)
sc = ort.Scene(hs.world, hs.root, 0, library=ort.lib(os.path.abspath(a.lib)) if a.lib else None)
P = ort.default_params(a.width, a.height, a.spp, chunk_spp=a.chunk, kernel=a.kernel)
for _ in range(a.reps):
    img, st = sc.render(hs.camera, P)
print("ms %.3f  Msamples/s %.1f  Mrays/s %.1f  rays/sample %.3f | extend %.2f sort %.2f shade %.2f ms, %d launches" % (
    st["device_ms"], st["samples"] / st["device_ms"] / 1e3, st["rays"] / st["device_ms"] / 1e3, st["rays"] / st["samples"],
    st["extend_ms"], st["sort_ms"], st["shade_ms"], st["kernel_launches"]))
