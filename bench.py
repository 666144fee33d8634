#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

Workload (config.workload): BASELINE config 3 -- bunny.ply in a closed box with shaped area
lights (scenes/c3_bunny_box.scn), 1920x1080, 256 samples per pixel PER GPU, Russian roulette
0.8, the reference's depth-of-field camera.  A "step" is one full frame: generate -> extend ->
shade -> accumulate of every sample.  With N GPUs each rank renders its own 256-spp range of
sample chunks of the same image (weak scaling: N x the samples) into a 64-bit fixed-point
framebuffer; one NCCL reduce(sum, int64) to rank 0 combines them exactly.

  value  Msamples/s, whole job, scene + framebuffers resident in HBM, CUDA-event timed,
         max over ranks.
  e2e    the same metric through the C ABI with HOST buffers (ort_render: params up, image down).
  roofline / roofline_fp32   algorithmic bytes and flops per ray (48 B per triangle
         test + 80 B per wide node; 51 flop per triangle test + 25 per box test), counted by the
         counters build of the same kernel, divided by the measured launch time.
  cpu_baseline   the reference's own tiled_raytrace_bvh (oracle/_ref, built from the reference's
         sources) on the box's host cores, bounded sample, reported only -- never a fallback.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
SPP_PER_GPU = 256
CHUNK_SPP = 16
RR = 0.8
SEED = 1234567
SCENE = os.path.join(ROOT, "scenes", "c3_bunny_box.scn")
DATA_DIR = os.path.join(ROOT, "oracle", "_ref", "data")      # reference meshes staged by build()
WORKLOAD = "C3 bunny.ply in closed box + shaped area lights, 1920x1080, 256 spp per GPU, rr 0.8"
SCENE_NAME, SCALING = "scenes/c3_bunny_box.scn", "weak"
# Supplementary runs only (the driver's contract is the default above): ORT_BENCH_CONFIG=c4 renders
# BASELINE config 4 -- dwarf.obj in a closed room with emitters > 1, 3840x2160, 1024 spp IN TOTAL,
# i.e. strong scaling: every rank takes 1024 / world samples per pixel
if os.environ.get("ORT_BENCH_CONFIG", "") == "c4":
    WIDTH, HEIGHT = 3840, 2160
    SPP_PER_GPU = max(CHUNK_SPP, 1024 // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
    SCENE = os.path.join(ROOT, "scenes", "c4_dwarf_hdr.scn")
    SCENE_NAME, SCALING = "scenes/c4_dwarf_hdr.scn", "strong"
    WORKLOAD = "C4 dwarf.obj in closed room, emitters > 1, 3840x2160, 1024 spp in total (%d per GPU), rr 0.8" % SPP_PER_GPU
# ORT_BENCH_CONFIG=c5: BASELINE config 5 -- 729 baked bunnies = 50.6 M triangles in a closed room
# (tools/make_scene_grid.py 27 <file>, path in ORT_BENCH_C5_SCENE), 3840x2160, 512 spp in total; the scene is
# handed over as shape lists and ranks, records and BVH are built on the device
FROM_LISTS = False
if os.environ.get("ORT_BENCH_CONFIG", "") == "c5":
    WIDTH, HEIGHT = 3840, 2160
    SPP_PER_GPU = max(CHUNK_SPP, 512 // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
    SCENE = os.environ.get("ORT_BENCH_C5_SCENE", "/tmp/c5_729.scn")
    SCENE_NAME, SCALING, FROM_LISTS = "tools/make_scene_grid.py 27", "strong", True
    WORKLOAD = "C5 729 bunnies = 50.6 M triangles in closed room, 3840x2160, 512 spp in total (%d per GPU), rr 0.8" % SPP_PER_GPU
L2_FLUSH_BYTES = 256 << 20


def emit(d):
    print(json.dumps(d), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """samples nvidia-smi during the timed region (B200_PROFILING.md: the clocks line)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference arm
def reference_child(kind, steps, warmup, sample_spp, threads):
    """runs in a subprocess: the reference's tiled_raytrace_bvh over its own 32x32 tile grid"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    if kind == "reference":
        ref = ol.Ref()
        rs = ref.scene_load(SCENE, DATA_DIR, WIDTH, HEIGHT, node_mb=1024, shape_mb=64)
        run = lambda seed: rs.render_tiles(seed, sample_spp, rr=RR, threads=threads)
    else:
        import offline_raytracer_b200 as ort
        hs = ort.HostScene.load(SCENE, DATA_DIR, WIDTH, HEIGHT)
        osc = ol.Oracle().scene(hs.world, hs.root)
        P = ol.default_params(WIDTH, HEIGHT, sample_spp, rr=RR, seed=SEED)
        run = lambda seed: osc.render(hs.camera, P, threads=threads)
    for i in range(warmup):
        run(SEED + i)
    t0 = time.perf_counter()
    for i in range(steps):
        run(SEED + 100 + i)
    dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "samples": WIDTH * HEIGHT * sample_spp * steps}), flush=True)


def run_reference_subprocess(steps, warmup, sample_spp, timeout=900):
    threads = os.cpu_count() or 1
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    for kind in (["reference"] if have_ref else []) + ["port"]:
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--child", kind,
               "--steps", str(steps), "--warmup", str(warmup), "--sample-spp", str(sample_spp)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
            if r.returncode == 0:
                out = json.loads(r.stdout.strip().splitlines()[-1])
                return kind, threads, out
        except Exception:
            pass
    return None, threads, None


def cpu_baseline_object(kind, threads, out, sample_spp, steps):
    ms = out["samples"] / out["seconds"] / 1e6
    return {"value": ms, "unit": "Msamples/s", "cores": threads, "kind": kind,
            "sample": "%dx%d at %d spp of %d (1/%d of a step) x %d, reference 32x32 tile scheduler, all host cores"
                      % (WIDTH, HEIGHT, sample_spp, SPP_PER_GPU, SPP_PER_GPU // sample_spp, steps)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.child:
        reference_child(args.child, args.steps, args.warmup, args.sample_spp, os.cpu_count() or 1)
        return
    kind, threads, out = run_reference_subprocess(args.steps, args.warmup, args.sample_spp)
    if out is None:
        emit({"impl": "reference", "unavailable": "neither oracle/_ref nor the oracle port could run"})
        return
    cb = cpu_baseline_object(kind, threads, out, args.sample_spp, args.steps)
    v = cb["value"]
    emit({"impl": "reference", "metric": "path_samples_per_second", "value": v, "unit": "Msamples/s",
          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": out["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": WORKLOAD, "sample": cb["sample"]},
          "cpu_baseline": cb,
          "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


# ----------------------------------------------------------------------------- GPU arm
def main_gpu(args):
    import torch
    import torch.distributed as dist
    import numpy as np
    import offline_raytracer_b200 as ort

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available() or ort.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL may print a version banner on stdout; the contract is ONE JSON line there
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    if not os.path.exists(os.path.join(DATA_DIR, "bunny.ply")):
        raise SystemExit("bench.py: %s/bunny.ply is not staged; run __graft_entry__.build() where /root/reference exists" % DATA_DIR)
    hs = ort.HostScene.load(SCENE, DATA_DIR, WIDTH, HEIGHT, octree=not FROM_LISTS)
    scene = ort.Scene.from_lists(hs.world, hs.lists(), local) if FROM_LISTS else ort.Scene(hs.world, hs.root, local)
    info = scene.info()

    spp_total = SPP_PER_GPU * world
    chunks_per_rank = SPP_PER_GPU // CHUNK_SPP
    P = ort.default_params(WIDTH, HEIGHT, spp_total, rr=RR, seed=SEED, chunk_spp=CHUNK_SPP)
    P.chunk_begin, P.chunk_end = rank * chunks_per_rank, (rank + 1) * chunks_per_rank
    accum = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.int64, device=dev)
    rgb = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    launches = [0]

    def step():
        flush.zero_()                                             # evict L2 between steps
        scene.accum_zero_device(accum.data_ptr(), WIDTH, HEIGHT, stream=stream)
        scene.render_accumulate_device(hs.camera, P, accum.data_ptr(), stream=stream)
        launches[0] += 1
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)       # exact: int64 fixed point
        if rank == 0:
            scene.accum_resolve_device(accum.data_ptr(), WIDTH, HEIGHT, spp_total, rgb.data_ptr(), stream=stream)
            launches[0] += 1

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(local) if rank == 0 else None
    launches[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    ms = float(ms.item())
    samples_per_step = WIDTH * HEIGHT * spp_total
    value = samples_per_step * args.steps / (ms * 1e-3) / 1e6

    # kernel-only time and ray counts of one step (this rank), for the roofline
    # (one pool for this pass: with two overlapping pools the per-stage event times overlap too)
    os.environ["ORT_WF_POOLS"] = "1"
    st = scene.render_accumulate_device(hs.camera, P, accum.data_ptr(), stream=stream, want_stats=True)
    del os.environ["ORT_WF_POOLS"]
    kernel_ms, rays, samples = st["device_ms"], st["rays"], st["samples"]
    # every kernel this library launches in one step (wavefront: extend/scan/scatter/shade per
    # iteration) + the resolve on rank 0
    gpu_launches = (st["kernel_launches"] + (1 if rank == 0 else 0)) * args.steps

    # ---- e2e: through the C ABI with host buffers ----
    e2e_ms = None
    host_img = np.zeros((HEIGHT, WIDTH, 3), np.float32)
    pinned = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.float32).pin_memory() if world > 1 else None
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if world == 1:
            scene.render(hs.camera, P, out=host_img)              # params up, kernels, image down
        else:
            step()
            if rank == 0:
                pinned.copy_(rgb, non_blocking=False)
            sync()
    sync()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = samples_per_step * args.steps / e2e_s / 1e6
    import ctypes
    h2d = ctypes.sizeof(ort.RenderParams) + ctypes.sizeof(ort.Camera)
    d2h = WIDTH * HEIGHT * 3 * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-ray work from the counters build of the same kernels (outside any timed region) ----
    per_ray = None
    cpath = os.path.join(os.path.dirname(ort.LIB_PATH), "libort_b200_counters.so")
    if os.path.exists(cpath):
        try:
            CL = ort.lib(cpath)
            cs = (ort.Scene.from_lists(hs.world, hs.lists(), local, library=CL) if FROM_LISTS
                  else ort.Scene(hs.world, hs.root, local, library=CL))
            Pc = ort.default_params(WIDTH, HEIGHT, CHUNK_SPP, rr=RR, seed=SEED, chunk_spp=CHUNK_SPP)
            _, cst = cs.render(hs.camera, Pc)
            per_ray = {"node_visits": cst["node_visits"] / cst["rays"], "box_tests": cst["box_tests"] / cst["rays"],
                       "shape_tests": cst["shape_tests"] / cst["rays"], "rays_per_sample": cst["rays"] / cst["samples"]}
            cs.close()
        except Exception as e:          # the counters library is optional evidence, not the product
            per_ray = {"error": str(e)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp32_peak = ort.measure_fp32_peak(local)
    rays_per_s = rays / (kernel_ms * 1e-3)
    # dominant kernel: EXTEND (k_wf_extend); its launches are timed live with CUDA events on the
    # launching stream inside the library (OrtRenderStats.extend_ms = sum over the step's launches)
    extend_ms = st["extend_ms"] if st["extend_ms"] > 0 else kernel_ms
    n_extend = max(1, (st["kernel_launches"] - 2) // 4)
    extend_rays_per_s = rays / (extend_ms * 1e-3)
    roofline = roofline_fp32 = None
    if per_ray and "error" not in per_ray:
        bytes_per_ray = 48.0 * per_ray["shape_tests"] + info["bvh_node_bytes"] * per_ray["node_visits"]
        flops_per_ray = 51.0 * per_ray["shape_tests"] + 25.0 * per_ray["box_tests"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_latest.json")
        if os.path.exists(tpath):
            try:
                # ncu measured DRAM bytes of ONE launch over `slots_per_launch` rays (profiles/README.md);
                # per launch here = that per-ray figure x the rays an average launch of this run extends
                tj = json.load(open(tpath))
                traffic = tj["dram_bytes_per_launch"] / tj["slots_per_launch"] * (rays / n_extend)
            except Exception:
                traffic = None
        a = bytes_per_ray * extend_rays_per_s / 1e9
        roofline = {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak, "traffic": traffic,
                    "peak_source": peak_src, "kernel": "k_wf_extend", "launch_ms": extend_ms / n_extend,
                    "launches_per_step": n_extend, "share_of_step": extend_ms / kernel_ms,
                    "algorithmic_bytes_per_ray": bytes_per_ray, "algorithmic_bytes_per_launch": bytes_per_ray * rays / n_extend,
                    "note": "scene (%.1f MB) is L2/L1-resident by design; the binding limit is the FP32/issue pipe, see roofline_fp32"
                            % (info["device_bytes"] / 1e6)}
        f = flops_per_ray * extend_rays_per_s / 1e12
        roofline_fp32 = {"bound": "fp32", "achieved": f, "peak": fp32_peak, "unit": "TFLOP/s", "frac": f / fp32_peak,
                         "peak_source": "measured in this run: FMUL+FADD chain kernel (no FMA), ort_measure_fp32_peak",
                         "algorithmic_flops_per_ray": flops_per_ray}

    # (config 5 exceeds the reference's own limits -- 99 meshes, fixed arenas -- so it has no CPU arm)
    kind, threads, out = run_reference_subprocess(1, 0, args.sample_spp) if world == 1 and not FROM_LISTS else (None, 0, None)
    cpu_baseline = cpu_baseline_object(kind, threads, out, args.sample_spp, 1) if out else None

    emit({"metric": "path_samples_per_second", "value": value, "unit": "Msamples/s", "n_gpus": world,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
          "scaling": SCALING, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": WORKLOAD, "scene": SCENE_NAME, "width": WIDTH, "height": HEIGHT,
                     "spp_per_gpu": SPP_PER_GPU, "spp_total": spp_total, "chunk_spp": CHUNK_SPP,
                     "triangles": info["triangle_count"], "records": info["record_count"],
                     "bvh_nodes": info["bvh_node_count"], "scene_device_bytes": info["device_bytes"],
                     "parallelism": "sample-chunk split x%d, ncclReduce(int64 sum)" % world if world > 1 else "1 GPU",
                     "cache": "L2 flushed between steps (256 MiB memset); the 5 MB scene is re-fetched from HBM each step"},
          "mrays_per_s": rays_per_s * world / 1e6, "rays_per_sample": rays / max(1, samples),
          "kernel_ms_per_step": kernel_ms, "stage_ms_per_step": {"extend": st["extend_ms"], "sort": st["sort_ms"], "shade": st["shade_ms"]},
          "per_ray_work": per_ray,
          "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "api": "ort_render (host v3 buffer)" if world == 1 else "ort_render_accumulate_device + ncclReduce + D2H"},
          "gpu_launches": gpu_launches, "clocks": clocks,
          "roofline": roofline, "roofline_fp32": roofline_fp32, "cpu_baseline": cpu_baseline})
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--child", default="")
    ap.add_argument("--sample-spp", type=int, default=2, dest="sample_spp",
                    help="spp of the bounded CPU sample (of the 256 of a step)")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
