"""Per-kernel counts of the SASS mnemonics that show the Blackwell-native paths (tcgen05 MMA, TMEM loads, TMA) in the built
library:  python tools/sass_summary.py > profiles/r01_sass_mnemonics.txt   (needs cuobjdump, no GPU)"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "point-cloud-cnn-segmentation_b200", "lib", "libpcseg_b200.so")
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "ATOMG", "RED", "REDG", "ATOM"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("pcseg::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["total"] += 1
        if op in KEYS:
            counts[cur][op] += 1
cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "total"]
print(f"{'kernel':58s}" + "".join(f"{c:>9s}" for c in cols))
for k, c in counts.items():
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        print(f"{k[:58]:58s}" + "".join(f"{c[x]:9d}" for x in cols))
others = [k for k, c in counts.items() if not (c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"])]
print(f"\n{len(others)} CUDA-core kernels (k_*: ingest, BN, head, pooling, Adam, ragged packing ...): no tensor-core / TMA instructions;"
      f" legacy HMMA anywhere in the library: {sum(c['HMMA'] for c in counts.values())}")
print("\nUTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA),")
print("UTCBAR = tcgen05.commit -> mbarrier, HMMA = legacy mma.sync (must be 0).")
