"""Digest of an `ncu --page raw --csv` export: the metrics this project argues with, per kernel launch.
    python tools/ncu_digest.py gpurun_out/r2b/c4_full_raw.csv [...]"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "sm__inst_executed.sum": "warp_inst",
    "smsp__thread_inst_executed.sum": "thread_inst",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_elapsed",
    "sm__inst_executed.avg.per_cycle_active": "ipc_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "l1tex__t_bytes.sum": "l1_bytes",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fmaheavy_pct",
    "smsp__issue_active.avg.pct": "issue_active_pct",
    "launch__registers_per_thread": "regs", "launch__occupancy_limit_registers": "occ_limit_regs",
    "sm__icc_request_hit_rate.pct": "icc_hit_pct",
    "launch__grid_size": "grid",
}
STALL = "smsp__average_warps_issue_stalled_"


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return x


def digest(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        rec = {"kernel": d["Kernel Name"].split("(")[0]}
        for k, short in KEYS.items():
            if k in d:
                v = num(d[k])
                if short == "time_us" and isinstance(v, float):
                    v = v / 1e3 if u[k].startswith("ns") else (v * 1e3 if u[k].startswith("ms") else v)
                if short in ("dram_read", "dram_write", "l2_bytes", "l1_bytes") and isinstance(v, float):
                    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[k], 1)
                    v = v * m
                rec[short] = v
        stalls = {k[len(STALL):].replace("_per_warp_active.pct", "").replace(".pct", "").replace("_per_warp_active.ratio", ""): num(v)
                  for k, v in d.items() if k.startswith(STALL) and k.endswith("per_warp_active.pct")}
        rec["top_stalls"] = dict(sorted(((k, v) for k, v in stalls.items() if isinstance(v, float)), key=lambda kv: -kv[1])[:5])
        out.append(rec)
    return out


if __name__ == "__main__":
    for p in sys.argv[1:]:
        for rec in digest(p):
            print(p, json.dumps(rec))
