"""BASELINE config 2: traversal-only microbenchmark.  N coherent primary rays (reference camera
model with lens samples) + N incoherent rays (origins uniform in the scene box inflated x2,
directions uniform on the sphere) against a scene, timed on the device with CUDA events;
optional hit-ID check against the exhaustive GPU kernel on a subsample.

    python tools/bench_raycast.py [--scene scenes/c2_bunny_only.scn] [--n 50000000] [--check 1000000]
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import offline_raytracer_b200 as ort  # noqa: E402


def primary_rays(cam12, n, dev, seed=1):
    cam = torch.tensor(cam12, dtype=torch.float32, device=dev).view(4, 3)
    p, X, Y, Z = cam[0], cam[1], cam[2], cam[3]
    h = int(math.sqrt(n * 9 / 16)); w = (n + h - 1) // h
    idx = torch.arange(n, device=dev)
    x = (idx % w).float(); y = (idx // w).float()
    px = 2.0 * x / w - 1.0; py = 2.0 * y / h - 1.0
    c2p = px[:, None] * X + py[:, None] * Y - Z
    c2p = c2p / c2p.norm(dim=1, keepdim=True)
    focal = (p - torch.tensor([0.0, 0.0, 0.2], device=dev)).norm()
    fp = p + focal * c2p
    g = torch.Generator(device=dev); g.manual_seed(seed)
    rad = torch.rand(n, generator=g, device=dev) * (2 * math.pi)
    lens = p + (0.1 * torch.cos(rad))[:, None] * X + (0.1 * torch.sin(rad))[:, None] * Y - 0.1 * Z
    d = fp - lens
    d = d / d.norm(dim=1, keepdim=True)
    return lens.contiguous(), d.contiguous()


def incoherent_rays(lo, hi, n, dev, seed=2, inflate=2.0):
    lo = torch.tensor(lo, device=dev); hi = torch.tensor(hi, device=dev)
    c, hd = 0.5 * (lo + hi), 0.5 * (hi - lo) * inflate
    g = torch.Generator(device=dev); g.manual_seed(seed)
    o = c + (torch.rand((n, 3), generator=g, device=dev) * 2 - 1) * hd
    d = torch.randn((n, 3), generator=g, device=dev)
    d = d / d.norm(dim=1, keepdim=True)
    return o.contiguous(), d.contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default=os.path.join(ROOT, "scenes", "c2_bunny_only.scn"))
    ap.add_argument("--base", default=os.path.join(ROOT, "oracle", "_ref", "data"))
    ap.add_argument("--n", type=int, default=50_000_000)
    ap.add_argument("--check", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--inflate", type=float, default=2.0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    hs = ort.HostScene.load(a.scene, a.base, 1920, 1080)
    sc = ort.Scene(hs.world, hs.root, 0)          # ORT_BVH_BUILD=device selects the CUDA builder
    info = sc.info()
    st = torch.cuda.current_stream().cuda_stream
    out = {"scene": os.path.basename(a.scene), "triangles": info["triangle_count"], "bvh_nodes": info["bvh_node_count"], "n": a.n}
    t = torch.empty(a.n, device=dev); r = torch.empty(a.n, dtype=torch.int32, device=dev)
    for name, (o, d) in (("coherent", primary_rays(hs.camera_array(), a.n, dev)),
                         ("incoherent", incoherent_rays(info["root_min"], info["root_max"], a.n, dev, inflate=a.inflate))):
        best = 1e30
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sc.raycast_batch_device(a.n, o.data_ptr(), d.data_ptr(), t.data_ptr(), r.data_ptr(), stream=st)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        hit = float((r != -1).float().mean().item())
        res = {"ms": best, "mrays_s": a.n / best / 1e3, "hit_fraction": hit}
        m = min(a.check, a.n)
        if m:
            c = sc.raycast_counters_device(m, o.data_ptr(), d.data_ptr())
            res["per_ray"] = {k: v / m for k, v in c.items()}
            tb = torch.empty(m, device=dev); rb = torch.empty(m, dtype=torch.int32, device=dev)
            sc.raycast_brute_device(m, o.data_ptr(), d.data_ptr(), tb.data_ptr(), rb.data_ptr(), stream=st)
            torch.cuda.synchronize()
            res["checked"] = m
            res["rank_mismatches_vs_brute"] = int((rb != r[:m]).sum().item())
            res["t_mismatches_vs_brute"] = int((tb.view(torch.int32) != t[:m].view(torch.int32)).sum().item())
        out[name] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
