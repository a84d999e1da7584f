"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    tot[name] += v; cnt[name] += 1
allms = sum(tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("        ms launches  share  kernel")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{tot[k]:10.2f} {cnt[k]:8d} {100*tot[k]/allms:5.1f}%  {k[:120]}")
print(f"{allms:10.2f} total")
