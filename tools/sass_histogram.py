"""SASS mnemonic histogram per kernel of libdaisy_b200.so (cuobjdump -sass): the instructions that prove what each kernel is
made of (UTCHMMA / LDTM / STTM = tcgen05 + TMEM, UTMALDG = tensor-map TMA, UBLKCP = bulk TMA, ...).

    python tools/sass_histogram.py > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "daisyriot_b200", "libdaisy_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "REDUX", "FFMA", "FMNMX3", "DFMA", "MUFU", "BAR", "MEMBAR", "ATOMS", "ATOMG", "RED")
print("# SASS mnemonic histograms, `cuobjdump -sass %s` (sm_100a)\n" % os.path.basename(so))
print("Instruction counts are static (per kernel image).  Tensor-core / TMEM / TMA mnemonics first, then the top 24 of the rest.\n")
for k, h in hist.items():
    if sum(h.values()) < 40:
        continue
    print("## `%s`  (%d instructions)\n" % (k, sum(h.values())))
    keys = [m for m in KEY if h.get(m)]
    print("key: " + (", ".join("%s %d" % (m, h[m]) for m in keys) if keys else "-") + "\n")
    print("top: " + ", ".join("%s %d" % (m, c) for m, c in h.most_common(24)) + "\n")
