#!/usr/bin/env python
"""bench.py -- benchmark of the B200 hot path on the BASELINE.json configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

HEADLINE (config.workload): BASELINE config 4 -- dwarf.obj in a closed room lit by emitters > 1 ("HDR"),
the reference's depth-of-field camera, 3840x2160 at 1024 samples per pixel IN TOTAL, Russian roulette 0.8.
A "step" is one full frame: generate -> extend -> shade -> accumulate of all 8.49 G samples.  With N GPUs
the 64 sample chunks of the SAME frame are split over the ranks (STRONG scaling: total work fixed), every
rank accumulates into its own int64 fixed-point framebuffer, and rank 0 sums the peers' framebuffers in
place over NVLink peer memory (CUDA IPC handles, one kernel: csrc/ort_multi.cu) and resolves.  Integer sums:
the image -- and `config.image_checksum` -- is identical for every N.

  value   Msamples/s, whole job, scene + framebuffers resident in HBM, CUDA-event timed, max over ranks.
  e2e     the same metric through the C ABI with HOST buffers (N = 1: ort_render -- params up, image down;
          N > 1: per-rank ort_render_accumulate_device + peer-memory sum + device-to-host copy of the image).
  roofline   dominant kernel k_wf_extend against max(FP32 pipe, L2 bandwidth) for scenes that live in L2
          (configs 1-4) and against HBM for config 5; both denominators MEASURED in this run
          (ort_measure_fp32_peak, ort_measure_l2_bandwidth), HBM from MEASURED_PEAKS.json.  Algorithmic work per
          ray: 48 B per triangle test + 80 B per wide node, 51 flop per triangle test + 25 per box test,
          counted by the counters build of the same kernel.
  cpu_baseline   the reference's own tiled_raytrace_bvh (oracle/_ref, built from the reference's sources) on
          the box's host cores, bounded sample, reported only -- never a fallback.
  sub     the other BASELINE configurations as sub-records, each with its own roofline / e2e:
          N = 1: c1 (testscene 480x270x16), c2 (100 M explicit rays vs bunny.ply, hit-ID check against the
          reference on a 2 x 1 M-ray subsample), c3 (bunny in closed box, 1920x1080x256), c5 (729 bunnies =
          50.6 M triangles, 3840x2160x512);  N > 1: c5 (strong scaling).   ORT_BENCH_SUB=c1,c3 / =none selects.

ORT_BENCH_CONFIG=c3|c4|c5 picks another headline for development runs (the driver's contract is the default).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DATA_DIR = os.path.join(ROOT, "oracle", "_ref", "data")      # reference meshes staged by build()
RR = 0.8
SEED = 1234567
CHUNK_SPP = 16
L2_FLUSH_BYTES = 256 << 20

CONFIGS = {
    "c1": dict(scene=os.path.join(DATA_DIR, "testscene.scn"), scene_name="data/testscene.scn (as shipped)", w=480, h=270, spp=16, chunk=16,
               lists=False, workload="C1 data/testscene.scn as shipped, 480x270 at 16 spp, rr 0.8"),
    "c3": dict(scene=os.path.join(ROOT, "scenes", "c3_bunny_box.scn"), scene_name="scenes/c3_bunny_box.scn", w=1920, h=1080, spp=256, chunk=CHUNK_SPP,
               lists=False, workload="C3 bunny.ply in closed box + shaped area lights, 1920x1080 at 256 spp, rr 0.8"),
    "c4": dict(scene=os.path.join(ROOT, "scenes", "c4_dwarf_hdr.scn"), scene_name="scenes/c4_dwarf_hdr.scn", w=3840, h=2160, spp=1024, chunk=CHUNK_SPP,
               lists=False, workload="C4 dwarf.obj in closed room, emitters > 1 (HDR), depth of field, 3840x2160 at 1024 spp in total, rr 0.8"),
    "c5": dict(scene=os.path.join(ROOT, "scenes", "c5_bunny_grid_729.scn"), scene_name="scenes/c5_bunny_grid_729.scn (tools/make_scene_grid.py 27)",
               w=3840, h=2160, spp=512, chunk=CHUNK_SPP, lists=True,
               workload="C5 729 baked bunnies = 50.6 M triangles in closed room (BVH exceeds L2), 3840x2160 at 512 spp in total, rr 0.8"),
}
HEADLINE = os.environ.get("ORT_BENCH_CONFIG", "c4")


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


# ----------------------------------------------------------------------------- reference arm (CPU)
def reference_child(args):
    """runs in a subprocess: the reference's own code on all host threads.
    render: tiled_raytrace_bvh over its 32x32 tile grid (code/macos_main.mm:602-671)
    raycast: raycast_top_most_node on explicit ray buffers read from a .npz (C2 hit-ID check)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as ol
    cfg = CONFIGS.get(args.config) or {}
    threads = os.cpu_count() or 1
    if args.child_mode == "raycast":
        z = np.load(args.rays)
        scene = os.path.join(ROOT, "scenes", "c2_bunny_only.scn")
        if args.child == "reference":
            rs = ol.Ref().scene_load(scene, DATA_DIR, 1920, 1080)
            cast = lambda o, d: rs.raycast(o, d, threads=threads)
        else:
            import offline_raytracer_b200 as ort
            hs = ort.HostScene.load(scene, DATA_DIR, 1920, 1080)
            osc = ol.Oracle().scene(hs.world, hs.root)
            cast = lambda o, d: osc.raycast(o, d, mode=0, threads=threads)
        out = {}
        secs = 0.0
        for name in ("coherent", "incoherent"):
            t0 = time.perf_counter()
            r = cast(z[name + "_o"], z[name + "_d"])
            secs += time.perf_counter() - t0
            out[name + "_t"] = r["t"]; out[name + "_mat"] = r["mat"]
        np.savez(args.rays_out, **out)
        print(json.dumps({"seconds": secs, "rays": int(len(z["coherent_o"]) + len(z["incoherent_o"]))}), flush=True)
        return
    W, H = cfg["w"], cfg["h"]
    if args.child == "reference":
        rs = ol.Ref().scene_load(cfg["scene"], DATA_DIR, W, H, node_mb=1024, shape_mb=64)
        run = lambda seed: rs.render_tiles(seed, args.sample_spp, rr=RR, threads=threads)
    else:
        import offline_raytracer_b200 as ort
        hs = ort.HostScene.load(cfg["scene"], DATA_DIR, W, H)
        osc = ol.Oracle().scene(hs.world, hs.root)
        P = ol.default_params(W, H, args.sample_spp, rr=RR, seed=SEED)
        run = lambda seed: osc.render(hs.camera, P, threads=threads)
    for i in range(args.warmup):
        run(SEED + i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        run(SEED + 100 + i)
    dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "samples": W * H * args.sample_spp * args.steps}), flush=True)


def run_reference_subprocess(config, steps, warmup, sample_spp, timeout=1500, extra=()):
    threads = os.cpu_count() or 1
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    for kind in (["reference"] if have_ref else []) + ["port"]:
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--child", kind, "--config", config,
               "--steps", str(steps), "--warmup", str(warmup), "--sample-spp", str(sample_spp)] + list(extra)
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
            if r.returncode == 0:
                return kind, threads, json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            pass
    return None, threads, None


def cpu_baseline_object(config, kind, threads, out, sample_spp, steps):
    cfg = CONFIGS[config]
    return {"value": out["samples"] / out["seconds"] / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
            "sample": "%dx%d at %d spp of %d (1/%d of a step) x %d, reference 32x32 tile scheduler, all host cores"
                      % (cfg["w"], cfg["h"], sample_spp, cfg["spp"], max(1, cfg["spp"] // sample_spp), steps)}


def main_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.child:
        reference_child(args)
        return
    config = HEADLINE
    spp = args.sample_spp or (2 if CONFIGS[config]["w"] <= 1920 else 1)
    kind, threads, out = run_reference_subprocess(config, args.steps, args.warmup, spp)
    if out is None:
        emit({"impl": "reference", "unavailable": "neither oracle/_ref nor the oracle port could run"})
        return
    cb = cpu_baseline_object(config, kind, threads, out, spp, args.steps)
    v = cb["value"]
    emit({"impl": "reference", "metric": "path_samples_per_second", "value": v, "unit": "Msamples/s",
          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": out["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": CONFIGS[config]["workload"], "sample": cb["sample"]},
          "cpu_baseline": cb,
          "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


# ----------------------------------------------------------------------------- GPU arm
class Env:
    """rank plumbing: torch.distributed (NCCL) for barriers and the exchange of IPC handles"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            sys.stdout.flush()                       # NCCL may print a banner on stdout; the contract is ONE JSON line there
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                warm = torch.zeros(1, device=self.dev)
                dist.all_reduce(warm)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_bytes(self, b):
        """every rank's 64-byte blob, on every rank"""
        torch = self.torch
        mine = torch.tensor(list(b), dtype=torch.uint8, device=self.dev)
        if self.world == 1:
            return [bytes(b)]
        out = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(out, mine)
        return [bytes(t.cpu().tolist()) for t in out]


class Peaks:
    """roofline denominators: FP32 and L2 measured here, once per process; HBM from MEASURED_PEAKS.json"""

    def __init__(self, ort, local):
        self.fp32 = ort.measure_fp32_peak(local)
        self.l2 = ort.measure_l2_bandwidth(local, 32)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.hbm = peaks.get("hbm_gbs", 6650.0)
        self.hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        self.traffic = {}
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
        except Exception:
            pass


def per_ray_counters(ort, cfg, hs, local):
    """per-ray work from the counters build of the same kernels (outside any timed region)"""
    cpath = os.path.join(os.path.dirname(ort.LIB_PATH), "libort_b200_counters.so")
    if not os.path.exists(cpath):
        return {"error": "counters library not built"}
    try:
        CL = ort.lib(cpath)
        cs = (ort.Scene.from_lists(hs.world, hs.lists(), local, library=CL) if cfg["lists"] else ort.Scene(hs.world, hs.root, local, library=CL))
        spp = min(cfg["spp"], CHUNK_SPP)
        if cfg["w"] >= 3840:
            spp = min(spp, 4)
        Pc = ort.default_params(cfg["w"], cfg["h"], spp, rr=RR, seed=SEED, chunk_spp=cfg["chunk"] and spp)
        _, cst = cs.render(hs.camera, Pc)
        cs.close()
        return {"node_visits": cst["node_visits"] / cst["rays"], "box_tests": cst["box_tests"] / cst["rays"],
                "shape_tests": cst["shape_tests"] / cst["rays"], "rays_per_sample": cst["rays"] / cst["samples"]}
    except Exception as e:          # the counters library is optional evidence, not the product
        return {"error": str(e)}


def roofline_objects(key, cfg, info, st, per_ray, peaks):
    """extend is bound by max(FP32 pipe, L2) while the scene lives in L2 and by HBM when it does not.
    All three readings are reported; `roofline` is the one that binds this configuration."""
    if not per_ray or "error" in per_ray:
        return None, {}
    rays, kernel_ms = st["rays"], st["device_ms"]
    wavefront = st["extend_ms"] > 0             # the stats pass asks for per-stage events (ORT_WF_TIMING)
    extend_ms = st["extend_ms"] if wavefront else kernel_ms
    n_launch = max(1, st.get("extend_launches", 0)) if wavefront else 1
    rps = rays / (extend_ms * 1e-3)
    bytes_per_ray = 48.0 * per_ray["shape_tests"] + info["bvh_node_bytes"] * per_ray["node_visits"]
    flops_per_ray = 51.0 * per_ray["shape_tests"] + 25.0 * per_ray["box_tests"]
    gbs, tfs = bytes_per_ray * rps / 1e9, flops_per_ray * rps / 1e12
    common = {"kernel": "k_wf_extend" if wavefront else "k_render_mega (whole loop in one launch)", "launch_ms": extend_ms / n_launch,
              "launches_per_step": n_launch, "share_of_step": extend_ms / kernel_ms,
              "algorithmic_bytes_per_ray": bytes_per_ray, "algorithmic_flops_per_ray": flops_per_ray,
              "algorithmic_bytes_per_launch": bytes_per_ray * rays / n_launch, "rays_per_launch": rays / n_launch}
    tr = peaks.traffic.get(key)
    traffic = None
    if tr and wavefront and "dram_bytes_per_launch" in tr:
        # ncu dram__bytes.sum (read + write) of ONE steady-state k_wf_extend launch of this very configuration over a
        # full 6 Mi-slot pool -- the pool this run's stats pass uses (profiles/README.md); per launch, like `achieved`
        traffic = tr["dram_bytes_per_launch"]
    r_fp32 = dict(common, bound="fp32", achieved=tfs, peak=peaks.fp32, unit="TFLOP/s", frac=tfs / peaks.fp32, traffic=traffic,
                  peak_source="measured in this run: FMUL+FADD chain kernel (no FMA, the intersectors' mix), ort_measure_fp32_peak")
    r_l2 = dict(common, bound="l2", achieved=gbs, peak=peaks.l2, unit="GB/s", frac=gbs / peaks.l2, traffic=traffic,
                peak_source="measured in this run: L2-resident 32 MiB streaming read, ld.global.cg, ort_measure_l2_bandwidth")
    r_hbm = dict(common, bound="hbm", achieved=gbs, peak=peaks.hbm, unit="GB/s", frac=gbs / peaks.hbm, traffic=traffic, peak_source=peaks.hbm_src)
    l2_resident = info["device_bytes"] < 100e6
    if l2_resident:
        main = r_l2 if r_l2["frac"] >= r_fp32["frac"] else r_fp32
        main = dict(main, note="scene (%.1f MB) lives in L2: bound = max(FP32 pipe, L2 bandwidth), SURVEY.md 8d; the other readings are in roofline_all"
                               % (info["device_bytes"] / 1e6))
    else:
        main = dict(r_hbm, note="scene (%.2f GB) exceeds the 126 MB L2: bound = HBM, SURVEY.md 8d" % (info["device_bytes"] / 1e9))
    return main, {"fp32": r_fp32, "l2": r_l2, "hbm": r_hbm}


def cache_note(info):
    mb = info["device_bytes"] / 1e6
    if mb < 100:
        return "L2 flushed between steps (256 MiB memset inside the timed bracket, see l2_flush_ms_per_step); the %.1f MB scene is re-fetched from HBM once per step and then lives in L2" % mb
    return "inputs exceed L2: the %.2f GB scene streams from HBM throughout (also flushed between steps)" % (mb / 1e3)


class FrameJob:
    """one BASELINE render configuration on this rank's GPU: scene, framebuffers, the step"""

    def __init__(self, env, ort, key):
        torch = env.torch
        self.env, self.ort, self.key, self.cfg = env, ort, key, CONFIGS[key]
        cfg = self.cfg
        if not os.path.exists(os.path.join(DATA_DIR, "bunny.ply")):
            raise SystemExit("bench.py: %s/bunny.ply is not staged; run __graft_entry__.build() where /root/reference exists" % DATA_DIR)
        t0 = time.perf_counter()
        self.hs = ort.HostScene.load(cfg["scene"], DATA_DIR, cfg["w"], cfg["h"], octree=not cfg["lists"])
        t1 = time.perf_counter()
        self.scene = (ort.Scene.from_lists(self.hs.world, self.hs.lists(), env.local) if cfg["lists"]
                      else ort.Scene(self.hs.world, self.hs.root, env.local))
        self.load_s, self.create_s = t1 - t0, time.perf_counter() - t1
        self.info = self.scene.info()
        W, H = cfg["w"], cfg["h"]
        self.chunked = cfg["chunk"] > 0
        self.P = ort.default_params(W, H, cfg["spp"], rr=RR, seed=SEED, chunk_spp=cfg["chunk"])
        self.n_chunks = (cfg["spp"] + cfg["chunk"] - 1) // cfg["chunk"] if self.chunked else 1
        from offline_raytracer_b200.dist import shard_chunks
        self.P.chunk_begin, self.P.chunk_end = shard_chunks(self.n_chunks, env.world, env.rank) if self.chunked else (0, 0)
        self.my_chunks = (self.P.chunk_end - self.P.chunk_begin) if self.chunked else (1 if env.rank == 0 else 0)
        self.rgb = torch.empty((H, W, 3), dtype=torch.float32, device=env.dev)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.side = torch.cuda.Stream(device=env.dev)
        self.launches = 0
        self.reduce_kind = "1 GPU: resolve only"
        # framebuffers: two per rank, alternating between steps, so that ONE barrier per step suffices -- the root
        # reads the buffers of step k while the peers already render step k+1 into the other ones
        self.acc, self.peers = [], [[], []]
        if self.chunked:
            self.acc = [self.scene.accum_alloc_device(W, H) for _ in range(2 if env.world > 1 else 1)]
            if env.world > 1:
                self._open_peers()
        self.k = 0

    def _open_peers(self):
        env = self.env
        ok = 1.0
        try:
            if os.environ.get("ORT_BENCH_REDUCE", "") == "nccl":
                raise RuntimeError("ORT_BENCH_REDUCE=nccl")          # evidence runs: force the library collective
            for b in range(2):
                handles = env.gather_bytes(self.scene.accum_ipc_export(self.acc[b]))
                if env.rank == 0:
                    self.peers[b] = [self.scene.accum_ipc_open(h) for r, h in enumerate(handles) if r != 0]
            self.reduce_kind = "peer-memory sum over NVLink (CUDA IPC), fused with the resolve: k_accum_reduce_resolve"
        except Exception as e:          # noqa: BLE001
            ok = 0.0
            self.reduce_error = str(e)
        if env.sum_over_ranks(ok) < env.world:
            # the framebuffers of other processes cannot be mapped on this box: sum them with ncclReduce instead
            self.peers = None
            W, H = self.cfg["w"], self.cfg["h"]
            self.nccl_acc = [env.torch.zeros((H, W, 4), dtype=env.torch.int64, device=env.dev)]
            self.reduce_kind = "ncclReduce(int64 sum) -- CUDA IPC unavailable: %s" % getattr(self, "reduce_error", "on another rank")

    def step(self, host_out=None):
        env, sc, cfg = self.env, self.scene, self.cfg
        W, H = cfg["w"], cfg["h"]
        if not self.chunked:
            # single chunk per pixel (config 1): float sums in sample order, exactly ray.cpp:1428
            img = host_out if host_out is not None else self._host_img()
            _, st = sc.render(self.hs.camera, self.P, out=img)
            self.launches += st["kernel_launches"]
            return
        b = self.k % len(self.acc)
        self.k += 1
        if self.peers is None:                      # ncclReduce arm
            acc_t = self.nccl_acc[0]
            acc_ptr = acc_t.data_ptr()
        else:
            acc_ptr = self.acc[b]
        sc.accum_zero_device(acc_ptr, W, H, stream=self.stream)
        if self.my_chunks:
            sc.render_accumulate_device(self.hs.camera, self.P, acc_ptr, stream=self.stream)     # returns when done
        if env.world > 1:
            if self.peers is None:
                env.dist.reduce(acc_t, dst=0, op=env.dist.ReduceOp.SUM)
                if env.rank == 0:
                    sc.accum_resolve_device(acc_ptr, W, H, cfg["spp"], self.rgb.data_ptr(), stream=self.stream)
            else:
                env.dist.barrier()                  # every rank's sums are complete (their calls are synchronous)
                if env.rank == 0:
                    self.side.synchronize()         # the previous step's sum (it ran beside this step's render)
                    sc.accum_reduce_resolve_device(acc_ptr, self.peers[b], W, H, cfg["spp"], self.rgb.data_ptr(),
                                                   stream=self.side.cuda_stream)
        else:
            sc.accum_reduce_resolve_device(acc_ptr, [], W, H, cfg["spp"], self.rgb.data_ptr(), stream=self.stream)
        self.launches += 1

    def finish(self):
        self.side.synchronize()
        self.env.torch.cuda.synchronize()

    def _host_img(self):
        import numpy as np
        if not hasattr(self, "_img"):
            self._img = np.zeros((self.cfg["h"], self.cfg["w"], 3), np.float32)
        return self._img

    def stats_pass(self):
        """kernel-only time, stage times and ray counts of one step of THIS rank's share (one pool: with two
        overlapping pools the per-stage event times overlap too)"""
        os.environ["ORT_WF_POOLS"] = "1"
        os.environ["ORT_WF_TIMING"] = "1"           # per-stage events: opt-in, they cost 2-3 % (the timed region runs without)
        try:
            if self.chunked:
                if not self.my_chunks:
                    return None
                acc_ptr = self.acc[0] if self.peers is not None else self.nccl_acc[0].data_ptr()
                return self.scene.render_accumulate_device(self.hs.camera, self.P, acc_ptr, stream=self.stream, want_stats=True)
            _, st = self.scene.render(self.hs.camera, self.P, out=self._host_img())
            return st
        finally:
            del os.environ["ORT_WF_POOLS"]
            del os.environ["ORT_WF_TIMING"]

    def checksum(self):
        """64-bit sum of the resolved float image's bit patterns on rank 0: identical for every N"""
        torch = self.env.torch
        return int(self.rgb.view(torch.int32).to(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF

    def close(self):
        if self.peers:
            for lst in self.peers:
                for p in lst:
                    self.scene.accum_ipc_close(p)
        self.env.barrier()
        for a in self.acc:
            self.scene.accum_free_device(a)
        self.scene.close()
        self.hs.close()


def run_frame_config(env, ort, peaks, key, steps, warmup, e2e_steps, with_cpu=True, sampler_cb=None):
    """value / e2e / roofline / cpu_baseline of one render configuration (headline or sub-record)"""
    import numpy as np
    torch = env.torch
    job = FrameJob(env, ort, key)
    cfg, W, H = job.cfg, job.cfg["w"], job.cfg["h"]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=env.dev)
    samples_per_step = W * H * cfg["spp"]

    for _ in range(warmup):
        flush.zero_()
        job.step()
    job.finish(); env.barrier()
    # what the L2 flush itself costs (it sits inside the timed bracket; reported so that it can be told apart)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(); flush.zero_(); f1.record(); torch.cuda.synchronize()
    flush_ms = f0.elapsed_time(f1)
    sampler = sampler_cb() if sampler_cb else None
    job.launches = 0
    # timed region: EXACTLY `steps` steps between two device events, barrier + synchronize on both sides
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        flush.zero_()                                            # evict L2 between steps
        job.step()
    job.finish()
    e1.record()
    torch.cuda.synchronize(); env.barrier()
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    value = samples_per_step * steps / (ms * 1e-3) / 1e6
    checksum = job.checksum() if (env.rank == 0 and job.chunked) else None

    st = job.stats_pass()
    launches_per_step = ((st["kernel_launches"] if st else 0) + 1)
    rays_total = env.sum_over_ranks(st["rays"] if st else 0)
    kernel_ms_max = env.max_over_ranks(st["device_ms"] if st else 0.0)

    # ---- e2e: through the C ABI with HOST buffers ----
    host_img = np.zeros((H, W, 3), np.float32)
    pinned = torch.empty((H, W, 3), dtype=torch.float32).pin_memory() if env.world > 1 else None
    Pw = ort.default_params(W, H, cfg["spp"], rr=RR, seed=SEED, chunk_spp=cfg["chunk"])
    if env.world == 1 and ms / steps < 1000.0:
        job.scene.render(job.hs.camera, Pw, out=host_img)        # short steps: the entry point's first call allocates its buffers
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if env.world == 1:
            job.scene.render(job.hs.camera, Pw, out=host_img)    # params up, kernels, image down
        else:
            job.step()
            if env.rank == 0:
                job.side.synchronize()
                pinned.copy_(job.rgb, non_blocking=False)
            env.barrier()
    env.barrier()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    e2e_value = samples_per_step * e2e_steps / e2e_s / 1e6
    h2d = ctypes.sizeof(ort.RenderParams) + ctypes.sizeof(ort.Camera)
    d2h = W * H * 3 * 4

    rec = None
    if env.rank == 0:
        per_ray = per_ray_counters(ort, cfg, job.hs, env.local)
        roof, roof_all = roofline_objects(key, cfg, job.info, st, per_ray, peaks) if st else (None, {})
        cpu = None
        if with_cpu and env.world == 1 and not cfg["lists"]:
            spp = 2 if W <= 1920 else 1
            if key == "c1":
                spp = cfg["spp"]
            kind, threads, out = run_reference_subprocess(key, 1, 0, spp)
            cpu = cpu_baseline_object(key, kind, threads, out, spp, 1) if out else None
        elif cfg["lists"]:
            cpu = {"value": None, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": "none: 729 meshes / 50.6 M triangles exceed the reference's own limits (100 meshes, fixed arenas: parser.h:195-208); "
                             "its CPU path is timed on configs 1-4"}
        rec = {"metric": "path_samples_per_second", "value": value, "unit": "Msamples/s", "n_gpus": env.world,
               "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": cfg["workload"], "scene": cfg["scene_name"], "width": W, "height": H,
                          "spp_total": cfg["spp"], "chunk_spp": cfg["chunk"], "chunks_total": job.n_chunks,
                          "chunks_per_rank": [job.n_chunks // env.world + (1 if r < job.n_chunks % env.world else 0) for r in range(env.world)],
                          "triangles": job.info["triangle_count"], "records": job.info["record_count"],
                          "bvh_nodes": job.info["bvh_node_count"], "scene_device_bytes": job.info["device_bytes"],
                          "scene_load_s": job.load_s, "scene_create_s": job.create_s,
                          "parallelism": ("sample-chunk split x%d of one frame, %s" % (env.world, job.reduce_kind)) if env.world > 1 else "1 GPU",
                          "cache": cache_note(job.info), "l2_flush_ms_per_step": flush_ms, "image_checksum": checksum},
               "mrays_per_s": rays_total / (kernel_ms_max * 1e-3) / 1e6 if kernel_ms_max else None,
               "rays_per_sample": rays_total / samples_per_step,
               "kernel_ms_per_step": kernel_ms_max,
               "stage_ms_per_step": {"extend": st["extend_ms"], "sort": st["sort_ms"], "shade": st["shade_ms"]} if st else None,
               "per_ray_work": per_ray,
               "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                       "api": "ort_render (host v3 buffer)" if env.world == 1 else "ort_render_accumulate_device + peer-memory sum + D2H"},
               "gpu_launches": launches_per_step * steps, "clocks": clocks,
               "roofline": roof, "roofline_all": roof_all, "cpu_baseline": cpu}
    job.close()
    del flush
    torch.cuda.empty_cache()
    return rec


def run_c2(env, ort, peaks):
    """BASELINE config 2: 50 M coherent + 50 M incoherent explicit rays against bunny.ply; Mrays/s on the device;
    hit-ID check of a 2 x 1 M-ray subsample against the reference's raycast_top_most_node on the host cores"""
    import numpy as np
    torch = env.torch
    n = int(os.environ.get("ORT_BENCH_C2_RAYS", "50000000"))
    check = min(n, 1_000_000)
    W, H = 1920, 1080
    scene = os.path.join(ROOT, "scenes", "c2_bunny_only.scn")
    hs = ort.HostScene.load(scene, DATA_DIR, W, H)
    sc = ort.Scene(hs.world, hs.root, env.local)
    info = sc.info()
    gw = 9428
    gh = (n + gw - 1) // gw                                   # 16:9 grid of >= n cells (9428 x 5304 for 50 M)
    grid = ort.default_params(gw, gh, 1)
    o = torch.empty((n, 3), device=env.dev); d = torch.empty((n, 3), device=env.dev)
    t = torch.empty(n, device=env.dev); r = torch.empty(n, dtype=torch.int32, device=env.dev); m = torch.empty(n, dtype=torch.int32, device=env.dev)
    lo, hi = np.array(info["root_min"]), np.array(info["root_max"])
    c, ext = 0.5 * (lo + hi), (hi - lo)
    st = torch.cuda.current_stream().cuda_stream
    out = {"workload": "C2 traversal only: %d coherent primaries (reference camera + lens samples, xorshift seed 1) + %d incoherent rays "
                       "(origins in the bunny's box x2, uniform directions, xorshift seed 2) vs bunny.ply" % (n, n),
           "scene": "scenes/c2_bunny_only.scn", "triangles": info["triangle_count"], "bvh_nodes": info["bvh_node_count"],
           "scene_device_bytes": info["device_bytes"], "metric": "traversal_rays_per_second", "unit": "Mrays/s"}
    sub = {}
    total_ms, launches = 0.0, 0
    for name in ("coherent", "incoherent"):
        if name == "coherent":
            ort.generate_camera_rays_device(hs.camera, grid, 1, n, o.data_ptr(), d.data_ptr(), device=env.local, stream=st)
        else:
            ort.generate_random_rays_device(c - ext, c + ext, 2, n, o.data_ptr(), d.data_ptr(), device=env.local, stream=st)
        best = 1e30
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sc.raycast_batch_device(n, o.data_ptr(), d.data_ptr(), t.data_ptr(), r.data_ptr(), m.data_ptr(), stream=st)
            e1.record(); torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
            launches += 1
        total_ms += best
        cnt = sc.raycast_counters_device(check, o.data_ptr(), d.data_ptr())
        bytes_per_ray = (48.0 * cnt["shape_tests"] + 80.0 * cnt["node_visits"]) / check
        flops_per_ray = (51.0 * cnt["shape_tests"] + 25.0 * cnt["box_tests"]) / check
        rps = n / (best * 1e-3)
        f_fp32, f_l2 = flops_per_ray * rps / 1e12 / peaks.fp32, bytes_per_ray * rps / 1e9 / peaks.l2
        sub[name] = {"ms": best, "mrays_per_s": rps / 1e6, "hit_fraction": float((r != -1).float().mean().item()),
                     "per_ray_work": {k: v / check for k, v in cnt.items()},
                     "roofline": {"bound": "l2" if f_l2 >= f_fp32 else "fp32", "frac": max(f_l2, f_fp32), "frac_fp32": f_fp32, "frac_l2": f_l2,
                                  "achieved_tflops": flops_per_ray * rps / 1e12, "achieved_gbs": bytes_per_ray * rps / 1e9,
                                  "peak_tflops": peaks.fp32, "peak_l2_gbs": peaks.l2, "kernel": "k_raycast", "launch_ms": best, "traffic": None}}
        # e2e: the host-buffer entry point on the check subsample (origins + directions up, t / rank / material down)
        oh, dh = o[:check].cpu().numpy(), d[:check].cpu().numpy()
        t0 = time.perf_counter()
        g = sc.raycast_batch(oh, dh, want_normal=False)
        sub[name]["e2e"] = {"value": check / (time.perf_counter() - t0) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": check * 24,
                            "d2h_bytes_per_step": check * 12, "api": "ort_raycast_batch (host arrays), %d rays" % check}
        sub[name]["_check"] = (oh, dh, g)
    out["value"] = 2 * n / (total_ms * 1e-3) / 1e6
    out["gpu_launches"] = launches
    # ---- hit-ID check against the reference on the host cores (the CPU leg) ----
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        rays, res = os.path.join(tmp, "rays.npz"), os.path.join(tmp, "hits.npz")
        np.savez(rays, coherent_o=sub["coherent"]["_check"][0], coherent_d=sub["coherent"]["_check"][1],
                 incoherent_o=sub["incoherent"]["_check"][0], incoherent_d=sub["incoherent"]["_check"][1])
        kind, threads, cout = run_reference_subprocess("c2", 1, 0, 1, extra=["--child-mode", "raycast", "--rays", rays, "--rays-out", res])
        if cout:
            z = np.load(res)
            mism = {}
            for name in ("coherent", "incoherent"):
                g = sub[name]["_check"][2]
                mism[name] = {"t_bits": int((g["t"].view(np.uint32) != z[name + "_t"].view(np.uint32)).sum()),
                              "material": int((g["mat"] != z[name + "_mat"]).sum())}
            out["hit_check"] = {"against": "the unmodified reference's raycast_top_most_node (code/ray.cpp:1165-1176)" if kind == "reference" else "oracle port",
                                "rays_checked": 2 * check, "mismatches": mism,
                                "note": "the reference returns no primitive ID (ray.cpp:613-622): hit distance bits + material identify the hit; "
                                        "rank equality vs the oracle's restated traversal is asserted in tests/test_gpu_configs.py"}
            out["cpu_baseline"] = {"value": cout["rays"] / cout["seconds"] / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind,
                                   "sample": "the first %d rays of each buffer" % check}
    for name in sub:
        del sub[name]["_check"]
    out.update(sub)
    sc.close(); hs.close()
    return out


def main_gpu(args):
    import torch
    import offline_raytracer_b200 as ort

    if not torch.cuda.is_available() or ort.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    env = Env(args)
    peaks = Peaks(ort, env.local)
    e2e_steps = max(1, min(args.steps, 3))

    rec = run_frame_config(env, ort, peaks, HEADLINE, args.steps, args.warmup, e2e_steps,
                           sampler_cb=(lambda: ClockSampler(env.local)) if env.rank == 0 else None)

    want = os.environ.get("ORT_BENCH_SUB")
    if want is None:
        subs = [k for k in (("c1", "c2", "c3", "c5") if env.world == 1 else ("c5",)) if k != HEADLINE]
    else:
        subs = [k for k in want.split(",") if k in ("c1", "c2", "c3", "c4", "c5") and k != HEADLINE]
    sub = {}
    for key in subs:
        try:
            if key == "c2":
                if env.world == 1:
                    sub[key] = run_c2(env, ort, peaks)
                continue
            if env.world > 1 and not CONFIGS[key]["chunk"]:
                continue
            k_steps, k_warm = (2, 1) if key == "c5" else ((3, 3) if key != "c1" else (5, 3))
            sub[key] = run_frame_config(env, ort, peaks, key, k_steps, k_warm, 1)
        except Exception as e:          # a sub-record must never cost the headline
            sub[key] = {"error": "%s: %s" % (type(e).__name__, e)}
    if env.rank == 0:
        rec["sub"] = sub
        rec["peaks_measured"] = {"fp32_tflops_non_fma": peaks.fp32, "l2_read_gbs": peaks.l2, "hbm_gbs": peaks.hbm, "hbm_source": peaks.hbm_src}
        emit(rec)
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--child", default="")
    ap.add_argument("--child-mode", default="render", dest="child_mode")
    ap.add_argument("--config", default=HEADLINE)
    ap.add_argument("--rays", default="")
    ap.add_argument("--rays-out", default="", dest="rays_out")
    ap.add_argument("--sample-spp", type=int, default=0, dest="sample_spp",
                    help="spp of the bounded CPU sample of a step (default: 1 at 4K, 2 at 1080p)")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
