"""Aggregate `ncu --page source --print-source cuda,sass --csv` by CUDA source line: instructions executed and stall samples."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])  # inst, thread inst, samples, text
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[0] in ("Function Name",):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    ie, te, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    def iv(x):
        try:
            return int(x)
        except ValueError:
            return 0
    if r[2] in ("", "-"):  # CUDA line row (no SASS address): carries the per-line totals
        key = (cur_file, line)
        agg[key][0] += iv(r[ie]); agg[key][1] += iv(r[te]); agg[key][2] += iv(r[ss]); agg[key][3] = r[1].strip()[:90]
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[2] for v in agg.values()) or 1
print(f"total warp-instructions {tot_i:.3e}, samples {tot_s}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{100*v[0]/tot_i:5.1f}% inst {100*v[2]/tot_s:5.1f}% stall  util {v[1]/max(1,v[0]):4.1f}  {f}:{l:4d}  {v[3]}")
