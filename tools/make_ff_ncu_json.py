"""gpurun_out/ffncu_<workload>.csv (tools/ff_ncu_capture.sh) -> profiles/ff_ncu.json, keyed by workload (bench.py reads it)."""
import csv
import glob
import json
import os
import re

out = {}
for path in sorted(glob.glob("gpurun_out/ffncu_*.csv")):
    name = re.match(r"ffncu_(.*)\.csv", os.path.basename(path)).group(1)
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    if len(rows) < 2:
        continue
    h = rows[0]
    mi, vi = h.index("Metric Name"), h.index("Metric Value")
    m = {r[mi]: r[vi].replace(",", "") for r in rows[1:]}
    f = lambda k: float(m[k]) if k in m and m[k] not in ("", "n/a") else None
    out[name] = {
        "kernel": "k_ff_tiles<0>", "workload": name,
        "source": "ncu --metrics ... --clock-control none -k regex:k_ff_tiles -c 1 python tools/ff_build_only.py %s (round 2, face-grid kernel; tools/ff_ncu_capture.sh)" % name,
        "warp_instructions": f("smsp__inst_executed.sum"),
        "threads_active_per_warp_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_slots_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_lsu_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "l1tex_throughput_pct": f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        "lts_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1_hit_rate_pct": f("l1tex__t_sector_hit_rate.pct"), "l2_hit_rate_pct": f("lts__t_sector_hit_rate.pct"),
        "dram_bytes": (f("dram__bytes_read.sum") or 0) + (f("dram__bytes_write.sum") or 0),
        "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": f("launch__registers_per_thread"), "grid": f("launch__grid_size"),
        "gpu_time_under_ncu": m.get("gpu__time_duration.sum"),
        "bound": "instruction issue / fixed-latency dependencies (DRAM far below 1 % of peak)",
    }
json.dump(out, open("profiles/ff_ncu.json", "w"), indent=1)
print(json.dumps({k: (v["warp_instructions"], v["threads_active_per_warp_instruction"], v["issue_slots_active_pct"]) for k, v in out.items()}, indent=1))
