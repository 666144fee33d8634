"""Development diagnostics on a GPU box (uses the test oracles; not part of the product):
error distribution of the device BSDF against the reference's golden vectors, the L2 / FP32 roofline
denominators for several buffer sizes, and quick stage timings of the BASELINE scenes.

    python tools/gpu_diag.py [--scenes]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import offline_raytracer_b200 as ort  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", action="store_true")
a = ap.parse_args()

gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
g = ort.selftest_bsdf(gold["bsdf_mat"], gold["bsdf_N"], gold["bsdf_wo"], gold["bsdf_wi"], gold["bsdf_state"], gold["bsdf_dist"])


def rel(x, y, floor=1e-6):
    return np.abs(x.astype(np.float64) - y.astype(np.float64)) / np.maximum(np.abs(y.astype(np.float64)), floor)


q = [50, 90, 99, 99.9, 100]
out = {"state_equal": bool(np.array_equal(g["state_after"], gold["bsdf_sample_state"])),
       "lobe_flips": int((g["is_transmission"] != gold["bsdf_sample_is_t"]).sum())}
same = g["is_transmission"] == gold["bsdf_sample_is_t"]
out["wi_abs_err_q"] = np.percentile(np.abs(g["sample_wi"][same] - gold["bsdf_sample_wi"][same]).max(axis=1), q).tolist()
out["pdf_rel_err_q"] = np.percentile(rel(g["pdf"], gold["bsdf_pdf"]), q).tolist()
out["eval_rel_err_q"] = np.percentile(rel(g["eval"], gold["bsdf_eval"]).max(axis=1), q).tolist()
out["pdf_zero_kept"] = bool(np.all(g["pdf"][gold["bsdf_pdf"] == 0] == 0))
worst = int(np.argmax(rel(g["pdf"], gold["bsdf_pdf"])))
out["pdf_worst"] = {"i": worst, "dev": float(g["pdf"][worst]), "ref": float(gold["bsdf_pdf"][worst]), "mat": gold["bsdf_mat"][worst].tolist()}
print(json.dumps({"bsdf_device_vs_reference": out}), flush=True)

print(json.dumps({"fp32_tflops": ort.measure_fp32_peak(0),
                  "l2_gbs": {str(m): ort.measure_l2_bandwidth(0, m) for m in (8, 16, 32, 48, 64, 96)}}), flush=True)

if a.scenes:
    data = os.path.join(ROOT, "oracle", "_ref", "data")
    for name, w, h, spp in (("c3_bunny_box", 1920, 1080, 64), ("c4_dwarf_hdr", 3840, 2160, 32), ("c5_bunny_grid_64", 1920, 1080, 32)):
        hs = ort.HostScene.load(os.path.join(ROOT, "scenes", name + ".scn"), data, w, h)
        sc = ort.Scene(hs.world, hs.root, 0)
        P = ort.default_params(w, h, spp, chunk_spp=16)
        os.environ["ORT_WF_POOLS"] = "1"; os.environ["ORT_WF_TIMING"] = "1"
        for _ in range(2):
            img, st = sc.render(hs.camera, P)
        del os.environ["ORT_WF_POOLS"]; del os.environ["ORT_WF_TIMING"]
        img, st2 = sc.render(hs.camera, P)
        print(json.dumps({name: {"one_pool": {k: st[k] for k in ("device_ms", "extend_ms", "sort_ms", "shade_ms", "rays", "samples", "shape_tests", "kernel_launches")},
                                 "two_pools_ms": st2["device_ms"], "msamples_s": st2["samples"] / st2["device_ms"] / 1e3}}), flush=True)
        sc.close(); hs.close()
