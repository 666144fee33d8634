"""Quick end-to-end check on a GPU box: parity of the CUDA path against the CPU oracle
and first timings.  Development tool (uses the test oracles); not part of the product.

    python tools/gpu_check.py [--big]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import offline_raytracer_b200 as ort  # noqa: E402
from oracle_lib import (DATA_DIR, Oracle, default_params as oracle_params, make_incoherent_rays,  # noqa: E402
                        make_primary_rays)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--scene", default=os.path.join(DATA_DIR, "testscene.scn"))
    ap.add_argument("--base", default=DATA_DIR)
    args = ap.parse_args()
    out = {}
    W, H = 480, 270
    t0 = time.time()
    hs = ort.HostScene.load(args.scene, args.base, W, H)
    out["host_load_s"] = time.time() - t0
    t0 = time.time()
    sc = ort.Scene(hs.world, hs.root, 0)
    out["scene_create_s"] = time.time() - t0
    out["info"] = sc.info()
    print(json.dumps(out), flush=True)

    orc = Oracle()
    osc = orc.scene(hs.world, hs.root)
    cam = hs.camera_array()
    o1, d1 = make_primary_rays(cam, 960, 540)
    info = out["info"]
    o2, d2 = make_incoherent_rays(500000, info["root_min"], info["root_max"])
    O = np.concatenate([o1, o2]); D = np.concatenate([d1, d2])
    t0 = time.time()
    ref = osc.raycast(O, D, 0, threads=os.cpu_count())
    t_cpu = time.time() - t0
    g = sc.raycast_batch(O, D)
    res = {
        "rays": int(len(O)), "cpu_s": t_cpu, "gpu_ms": float(g["device_ms"]),
        "t_bit_equal": int((g["t"].view(np.uint32) == ref["t"].view(np.uint32)).sum()),
        "rank_equal": int((g["rank"] == ref["rank"]).sum()),
        "mat_equal": int((g["mat"] == ref["mat"]).sum()),
        "normal_bit_equal": int((g["normal"].view(np.uint32) == ref["normal"].view(np.uint32)).all(axis=1).sum()),
    }
    print(json.dumps({"raycast_vs_oracle": res}), flush=True)

    # GPU BVH vs GPU brute force on device buffers
    n = 2_000_000 if args.big else 300_000
    o3, d3 = make_incoherent_rays(n, info["root_min"], info["root_max"], seed=7)
    dev = torch.device("cuda:0")
    to, td = torch.from_numpy(o3).to(dev), torch.from_numpy(d3).to(dev)
    t_a = torch.empty(n, dtype=torch.float32, device=dev); r_a = torch.empty(n, dtype=torch.int32, device=dev)
    t_b = torch.empty(n, dtype=torch.float32, device=dev); r_b = torch.empty(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    stream = torch.cuda.current_stream().cuda_stream
    e0.record()
    sc.raycast_batch_device(n, to.data_ptr(), td.data_ptr(), t_a.data_ptr(), r_a.data_ptr(), stream=stream)
    e1.record()
    sc.raycast_brute_device(n, to.data_ptr(), td.data_ptr(), t_b.data_ptr(), r_b.data_ptr(), stream=stream)
    e2.record()
    torch.cuda.synchronize()
    cnt = sc.raycast_counters_device(n, to.data_ptr(), td.data_ptr())
    print(json.dumps({"bvh_vs_brute": {
        "rays": n, "bvh_ms": e0.elapsed_time(e1), "brute_ms": e1.elapsed_time(e2),
        "mrays_s": n / e0.elapsed_time(e1) / 1e3,
        "rank_equal": int((r_a == r_b).sum().item()), "t_equal": int((t_a.view(torch.int32) == t_b.view(torch.int32)).sum().item()),
        "per_ray": {k: v / n for k, v in cnt.items()}}}), flush=True)

    # image vs oracle, same seeds
    P = ort.default_params(W, H, 16)
    img, st = sc.render(hs.camera, P)
    t0 = time.time()
    oimg, ocnt = osc.render(hs.camera, oracle_params(W, H, 16), threads=os.cpu_count())
    t_cpu = time.time() - t0
    diff = np.abs(img - oimg)
    rel_close = (diff <= 1e-4 * np.maximum(np.abs(oimg), 1e-3)).all(axis=2).mean()
    print(json.dumps({"image_vs_oracle": {
        "gpu_ms": st["device_ms"], "gpu_msamples_s": W * H * 16 / st["device_ms"] / 1e3, "cpu_s": t_cpu,
        "cpu_msamples_s": W * H * 16 / t_cpu / 1e6, "gpu_rays": st["rays"], "cpu_rays": ocnt["rays"],
        "gpu_mean": img.mean((0, 1)).tolist(), "cpu_mean": oimg.mean((0, 1)).tolist(),
        "pixels_within_1e-4": float(rel_close),
        "rmse": np.sqrt(((img - oimg) ** 2).mean((0, 1))).tolist()}}), flush=True)

    # timings at larger sizes
    for (w, h, spp, chunk) in ([(1920, 1080, 64, 0), (1920, 1080, 64, 16), (1920, 1080, 256, 16)] if args.big
                               else [(960, 540, 16, 0), (960, 540, 16, 4)]):
        hs2 = ort.HostScene.load(args.scene, args.base, w, h)
        P = ort.default_params(w, h, spp, chunk_spp=chunk)
        sc.render(hs2.camera, P)
        img, st = sc.render(hs2.camera, P)
        print(json.dumps({"render": {"w": w, "h": h, "spp": spp, "chunk_spp": chunk, "ms": st["device_ms"],
                                     "msamples_s": w * h * spp / st["device_ms"] / 1e3,
                                     "mrays_s": st["rays"] / st["device_ms"] / 1e3,
                                     "rays_per_sample": st["rays"] / max(1, st["samples"]),
                                     "mean": img.mean((0, 1)).tolist()}}), flush=True)


if __name__ == "__main__":
    main()
