#!/bin/bash
# ncu evidence of one round (run on the GPU box through gpurun): launch list of the bench command + one
# `--set full` capture of k_wf_extend / k_wf_shade per configuration.  Usage: tools/ncu_round.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export ORT_WF_POOLS=1
# 1. the bench command itself must exit 0 without ncu first
ORT_BENCH_SUB=none python bench.py --steps 1 --warmup 1 > $OUT/bench_plain.json 2> $OUT/bench_plain.err || { echo "plain bench failed"; exit 1; }
# 2. launch list of the same command (first 3000 launches: the warm-up step's fill and steady state)
ORT_BENCH_SUB=none timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/bench_launches.csv \
    python bench.py --steps 1 --warmup 1 > $OUT/bench_ncu.json 2> $OUT/bench_ncu.err
python - <<PY
import csv, collections, sys
rows = [r for r in csv.reader(open("$OUT/bench_launches.csv")) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0]
    try: v = float(r[-1].replace(",", ""))
    except ValueError: continue
    unit = r[-2]
    v = v / 1e3 if unit.startswith("ns") else (v * 1e3 if unit.startswith("ms") else v)   # -> us
    tot[name][0] += 1; tot[name][1] += v
s = sum(v[1] for v in tot.values())
with open("$OUT/bench_launches_summary.csv", "w") as f:
    f.write("kernel,launches,total_us,avg_us,share\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%.1f,%.2f,%.4f\n" % (k, v[0], v[1], v[1] / v[0], v[1] / s))
print(open("$OUT/bench_launches_summary.csv").read())
PY
gzip -f $OUT/bench_launches.csv
# 3. --set full of one steady-state extend + shade launch per configuration
cap() {  # name, skip, args...
    local name=$1 skip=$2; shift 2
    python tools/profile_render.py "$@" > $OUT/${name}_plain.log 2>&1 || { echo "$name plain failed"; return; }
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wf_(shade|extend)" -s $skip -c 2 -f -o $OUT/${name}_full \
        python tools/profile_render.py "$@" > $OUT/${name}_ncu.log 2>&1
    ncu -i $OUT/${name}_full.ncu-rep --page raw --csv > $OUT/${name}_full_raw.csv 2>/dev/null
    ncu -i $OUT/${name}_full.ncu-rep --page source --csv --kernel-name regex:k_wf_extend > $OUT/${name}_extend_source.csv 2>/dev/null
    ncu -i $OUT/${name}_full.ncu-rep --page source --csv --kernel-name regex:k_wf_shade > $OUT/${name}_shade_source.csv 2>/dev/null
    gzip -f $OUT/${name}_extend_source.csv $OUT/${name}_shade_source.csv
    ls -la $OUT/${name}_full.ncu-rep
}
cap c3 120 --scene scenes/c3_bunny_box.scn --width 1920 --height 1080 --spp 64 --kernel 2 --reps 1
cap c4 120 --scene scenes/c4_dwarf_hdr.scn --width 3840 --height 2160 --spp 32 --kernel 2 --reps 1
cap c5lite 120 --scene scenes/c5_bunny_grid_64.scn --width 1920 --height 1080 --spp 32 --kernel 2 --reps 1
cap c5 120 --scene scenes/c5_bunny_grid_729.scn --lists --width 3840 --height 2160 --spp 16 --kernel 2 --reps 1
rm -f $OUT/c3_full.ncu-rep $OUT/c5lite_full.ncu-rep $OUT/c5_full.ncu-rep      # keep the csv pages; one .ncu-rep (c4) is enough to bring back
# 4. DRAM traffic of the captured EXTEND launch per configuration (bench.py: roofline.traffic)
python - <<PY
import json, sys
sys.path.insert(0, "tools")
from ncu_digest import digest
out = {}
for tag in ("c3", "c4", "c5lite", "c5"):
    try:
        recs = digest("$OUT/%s_full_raw.csv" % tag)
        stats = [json.loads(l[6:]) for l in open("$OUT/%s_plain.log" % tag) if l.startswith("STATS ")][-1]
        for r in recs:
            if "extend" in r["kernel"]:
                out[tag] = {"dram_bytes_per_launch": r["dram_read"] + r["dram_write"], "launch_us": r["time_us"],
                            "rays_per_launch": stats["rays"] / max(1, stats["extend_launches"]),
                            "l2_hit_pct": r.get("l2_hit_pct"), "l1_hit_pct": r.get("l1_hit_pct"), "ipc_active": r.get("ipc_active"),
                            "lanes_per_inst": r.get("lanes_per_inst"),
                            "source": "ncu --set full, one steady-state k_wf_extend launch, one pool of 6 Mi slots ($TAG)"}
    except Exception as e:
        out[tag] = {"error": str(e)}
json.dump(out, open("$OUT/traffic.json", "w"), indent=1)
print(json.dumps(out))
PY
