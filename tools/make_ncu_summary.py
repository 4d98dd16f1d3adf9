"""Turn ncu reports / launch lists under gpurun_out/ into the markdown summary committed under profiles/.

    python tools/make_ncu_summary.py OUT.md --title T --rep NAME=path.ncu-rep[:launch_index] ... --launches path.csv [--note TEXT]
"""
import argparse
import collections
import csv
import io
import re
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h0]
    data = [r for r in rows[h0 + 1:] if len(r) == len(hdr)]
    reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    si = hdr.index("Warp Stall Sampling (All Samples)")
    iv = lambda x: int(x) if x.isdigit() else 0
    tot = sum(iv(r[si]) for r in data) or 1
    agg = {hdr[i][6:]: sum(iv(r[i]) for r in data) for i in reasons}
    return tot, len(data), sorted(((v * 100.0 / tot, k) for k, v in agg.items() if v * 100.0 / tot >= 1), reverse=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--launches")
    ap.add_argument("--note", action="append", default=[])
    a = ap.parse_args()
    md = [f"# {a.title}", ""] + [n + "\n" for n in a.note]
    for spec in a.rep:
        name, path = spec.split("=", 1)
        idx = 0
        if re.search(r":\d+$", path):
            path, idx = path.rsplit(":", 1)
            idx = int(idx)
        rows = raw(path)
        h, units, vals = rows[0], rows[1], rows[2 + idx]
        kn = vals[h.index("Kernel Name")]
        md += [f"## {name}", "", f"kernel: `{kn}`  (report `{path}`)", "", "| metric | value | unit |", "|---|---|---|"]
        for m in METRICS:
            if m in h:
                md.append(f"| {m} | {vals[h.index(m)]} | {units[h.index(m)]} |")
        tot, ninst, st = stalls(path)
        md += ["", f"warp-stall samples {tot} over {ninst} SASS instructions: " + ", ".join(f"{k} {p:.1f} %" for p, k in st), ""]
    if a.launches:
        rows = [r for r in csv.reader(open(a.launches)) if len(r) > 10 and r[0].isdigit()]
        agg = collections.OrderedDict()
        for r in rows:
            k = re.sub(r"\(.*", "", r[4])
            t = agg.setdefault(k, [0, 0.0])
            t[0] += 1
            t[1] += float(r[-1]) / 1e6
        total = sum(v[1] for v in agg.values())
        md += [f"## Launch list (`{a.launches}`, `--metrics gpu__time_duration.sum --clock-control none`)", "",
               "| kernel | launches | total ms | avg us | share |", "|---|---|---|---|---|"]
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| {k} | {n} | {ms:.3f} | {ms * 1e3 / n:.1f} | {100 * ms / total:.2f} % |")
        g = {k: v for k, v in agg.items() if "gather" in k or "split" in k or "wait" in k}
        gt = sum(v[1] for v in g.values())
        if gt:
            md += ["", "Within the gather passes: " + ", ".join(f"`{k}` {100 * v[1] / gt:.1f} %" for k, v in g.items()) + "."]
    open(a.out, "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
