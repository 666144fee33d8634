"""Small run of every kernel, meant to be executed under `compute-sanitizer --tool memcheck`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import offline_raytracer_b200 as ort  # noqa: E402
from oracle_lib import make_incoherent_rays, make_primary_rays  # noqa: E402

scn = os.path.join(ROOT, "oracle", "_ref", "data", "testscene.scn")
base = os.path.join(ROOT, "oracle", "_ref", "data")
if not os.path.exists(scn):
    scn, base = os.path.join(ROOT, "scenes", "box_spheres.scn"), os.path.join(ROOT, "scenes")
W, H = 96, 54
hs = ort.HostScene.load(scn, base, W, H)
sc = ort.Scene(hs.world, hs.root, 0)
info = sc.info()
o, d = make_primary_rays(hs.camera_array(), W, H)
o2, d2 = make_incoherent_rays(3000, info["root_min"], info["root_max"])
g = sc.raycast_batch(np.concatenate([o, o2]), np.concatenate([d, d2]))
print("raycast ok", int((g["mat"] != 0).sum()))
for kernel in (ort.ORT_KERNEL_MEGAKERNEL, ort.ORT_KERNEL_WAVEFRONT):
    for chunk in (0, 2):
        for env in ({}, {"ORT_WF_POOLS": "1"}):
            os.environ.pop("ORT_WF_POOLS", None)
            os.environ.update(env)
            img, st = sc.render(hs.camera, ort.default_params(W, H, 4, chunk_spp=chunk, kernel=kernel))
            print("render ok kernel", kernel, "chunk", chunk, env, "mean", img.mean((0, 1)), "launches", st["kernel_launches"])
os.environ.pop("ORT_WF_POOLS", None)
import torch  # noqa: E402
n = 2000
to = torch.from_numpy(o2[:n]).cuda(); td = torch.from_numpy(d2[:n]).cuda()
t = torch.empty(n, device="cuda"); r = torch.empty(n, dtype=torch.int32, device="cuda")
sc.raycast_brute_device(n, to.data_ptr(), td.data_ptr(), t.data_ptr(), r.data_ptr())
torch.cuda.synchronize()
print("brute ok", sc.raycast_counters_device(n, to.data_ptr(), td.data_ptr()))
sc.close()
print("done")
