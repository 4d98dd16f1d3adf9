"""Print the instructions with the most warp-stall samples from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for bi, h0 in enumerate(hi):
    hdr = rows[h0]
    end = hi[bi + 1] - 1 if bi + 1 < len(hi) else len(rows)
    data = [r for r in rows[h0 + 1:end] if len(r) == len(hdr)]
    si, src = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
    reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]

    def iv(x):
        try:
            return int(x)
        except ValueError:
            return 0
    tot = sum(iv(r[si]) for r in data) or 1
    print("== block", bi, rows[h0 - 1][:2], "samples", tot, "instructions", len(data))
    agg = {hdr[i]: sum(iv(r[i]) for r in data) for i in reasons}
    print("   by reason:", ", ".join(f"{k[6:]}={100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 / tot >= 1))
    for r in sorted(data, key=lambda r: -iv(r[si]))[:n]:
        why = " ".join(f"{hdr[i][6:]}={r[i]}" for i in reasons if iv(r[i]) * 10 > iv(r[si]) and iv(r[i]))
        print(f"{iv(r[si]):7d} {100*iv(r[si])/tot:5.1f}%  {r[src][:100]:100s} | {why}")
    break
