"""Per-kernel share of a step from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = defaultdict(list)
for r in rows[h + 1:]:
    if len(r) > mv:
        d[r[kn].split("(")[0].replace("void ", "")].append(float(r[mv].replace(",", "")))
tot = sum(sum(v) for v in d.values())
print(f"# {sys.argv[1]}: launches profiled = {sum(len(v) for v in d.values())}; times are cold-cache, serialised (compare shares)")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:44s} n={len(v):3d} avg={sum(v) / len(v) / 1000:9.2f} us share={sum(v) / tot * 100:5.1f}%")
