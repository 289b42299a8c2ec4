"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, total = collections.OrderedDict(), 0.0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1000.0 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
    name = re.sub(r"\(.*", "", r[kn]).replace("pcseg::", "").replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    total += v
print(f"total {total:.1f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / total:5.1f}%  n={n:3d}  avg {t / n:8.1f} us  {k[:100]}")
