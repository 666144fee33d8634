"""Writes a synthetic many-triangle scene (BASELINE config 5 family): an n x n grid of baked
bunny.ply instances in a closed room with emissive spheres/cylinders.  The reference has no
instancing -- every `mesh` line re-parses and copies the file (code/macos_main.mm:344-414) -- so
n*n <= 99 keeps the file loadable by the reference too (parser.h:207).

    python tools/make_scene_grid.py 8 scenes/c5_bunny_grid_64.scn
"""
import sys

n = int(sys.argv[1]); out = sys.argv[2]
pitch = 1.1
half = 0.5 * pitch * n
lines = ["screen 3840 2160",
         "camera 5.553228 2.942755 2.900874 b 0.2 q 0.416981 0.279589 0.480987 0.718247",
         "ambient 0.125 0.125 0.125", ""]
lo, hi, top = -half - 2.0, half + 7.0, 7.0
size = hi - lo
lines += ["brdf 0.7 0.7 0.7 0.0 0.0 0.0 10 0.0 0.0 0.0 1.0",
          "box %.4f %.4f -0.1 %.4f %.4f 0.1" % (lo, lo, size, size),
          "box %.4f %.4f %.4f %.4f %.4f 0.1" % (lo, lo, top, size, size),
          "brdf 0.75 0.3 0.25 0.0 0.0 0.0 10 0.0 0.0 0.0 1.0",
          "box %.4f %.4f -0.1 0.1 %.4f %.4f" % (lo - 0.1, lo, size, top + 0.2),
          "brdf 0.25 0.7 0.3 0.0 0.0 0.0 10 0.0 0.0 0.0 1.0",
          "box %.4f %.4f -0.1 0.1 %.4f %.4f" % (hi, lo, size, top + 0.2),
          "brdf 0.3 0.35 0.8 0.0 0.0 0.0 10 0.0 0.0 0.0 1.0",
          "box %.4f %.4f -0.1 %.4f 0.1 %.4f" % (lo, lo - 0.1, size, top + 0.2),
          "brdf 0.8 0.75 0.3 0.0 0.0 0.0 10 0.0 0.0 0.0 1.0",
          "box %.4f %.4f -0.1 %.4f 0.1 %.4f" % (lo, hi, size, top + 0.2), ""]
mats = ["brdf 0.55 0.45 0.35 0.3 0.3 0.3 10 0.0 0.0 0.0 1.0",
        "brdf 0.2 0.5 0.7 0.1 0.1 0.1 10 0.0 0.0 0.0 1.0",
        "brdf 0.0 0.0 0.0 1.0 0.9 0.7 20 0.0 0.0 0.0 1.0",
        "brdf 0.0 0.0 0.0 0.1 0.1 0.1 10 0.9 1.0 0.9 1.4"]
k = 0
for j in range(n):
    for i in range(n):
        x = (i - (n - 1) / 2.0) * pitch - 0.0475
        y = (j - (n - 1) / 2.0) * pitch + 0.0425
        lines.append(mats[k % len(mats)])
        lines.append("mesh bunny.ply %.4f %.4f -0.165 5.0 z %d q 0 0 0.707107 0.707106" % (x, y, -90 + 37 * k))
        k += 1
lines += ["", "light 9 8 7", "sphere 0.5 0.5 %.4f 0.9" % (top - 1.2),
          "light 5 5 6", "sphere %.4f %.4f 3.2 0.5" % (lo + 1.5, hi - 1.5),
          "light 3 6 3", "cylinder %.4f %.4f 0.0 0.0 0.0 4.5 0.15" % (hi - 0.8, lo + 0.5), ""]
open(out, "w").write("\n".join(lines))
print("wrote", out, "with", n * n, "bunnies =", n * n * 69451, "triangles")
