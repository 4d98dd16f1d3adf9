"""Attribute the warp instructions of `ncu --page source --print-source cuda,sass --csv` to code regions of formfactor.cu: SASS
instructions are taken in address order and an instruction whose innermost source line lies in a helper header is charged
to the last formfactor.cu line seen before it (inlined helpers follow their call site).  usage: ncu_regions.py file.csv lo:hi:name ..."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
regions = []
for a in sys.argv[2:]:
    lo, hi, name = a.split(":")
    regions.append((int(lo), int(hi), name))
cur_file, hdr, cur_line = None, None, None
ins = []  # (addr, file, line, inst, thread_inst, samples)
func = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, te, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] not in ("", "-") and r[2] in ("", "-"):
        try:
            cur_line = int(r[0])
        except ValueError:
            pass
        continue
    if r[2].startswith("0x"):
        def iv(x):
            try:
                return int(x)
            except ValueError:
                return 0
        ins.append((int(r[2], 16), cur_file, cur_line, iv(r[ie]), iv(r[te]), iv(r[ss]), func))
ins.sort()
# an instruction inlined from a header appears once per file section: keep one row per address, preferring formfactor.cu's
uniq = {}
for rec in ins:
    a = rec[0]
    if a not in uniq or (rec[1] == "formfactor.cu" and uniq[a][1] != "formfactor.cu"):
        uniq[a] = rec
ins = [uniq[a] for a in sorted(uniq)]
tot = sum(i[3] for i in ins) or 1
tots = sum(i[5] for i in ins) or 1
agg = {}
last = None
for addr, f, line, ie_, te_, ss_, fn in ins:
    if f == "formfactor.cu":
        last = line
    key = "?"
    if last is not None:
        for lo, hi, name in regions:
            if lo <= last <= hi:
                key = name
                break
        else:
            key = f"ff:{last // 20 * 20}"
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += ie_; a[1] += te_; a[2] += ss_
print(f"total warp-instructions {tot:.3e}  samples {tots}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] * 1000 > tot:
        print(f"{100*v[0]/tot:5.1f}% inst  {100*v[2]/tots:5.1f}% samples  lanes {v[1]/max(1,v[0]):4.1f}  {k}")
