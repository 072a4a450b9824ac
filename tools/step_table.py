"""Per-kernel device time of ONE eager training step from an `ncu --metrics gpu__time_duration.sum --csv` launch list:
the launches between two consecutive `adam_prox_kernel`s.  usage: python tools/step_table.py launches.csv"""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    rows.append((r[ki], us))
marks = [i for i, (k, _) in enumerate(rows) if "adam_prox_kernel" in k]
assert len(marks) >= 2, "need two optimizer steps in the capture"
a, b = marks[-2], marks[-1]
step = rows[a + 1: b + 1]
agg = OrderedDict()
for k, us in step:
    k = re.sub(r"\(.*", "", k).replace("fedb200::", "").replace("void ", "").strip()
    k = k[:96]
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + us)
total = sum(t for _, t in agg.values())
print("| kernel | launches | sum µs | share |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (k, c, t, 100 * t / total))
print("| **total** | %d | %.1f | |" % (sum(c for c, _ in agg.values()), total))
