"""Hot SASS instructions of one kernel of an ncu report (warp-stall samples per instruction).
    python tools/ncu_hot_sass.py report.ncu-rep <invocation-nr (1-based)> [top]"""
import csv
import io
import subprocess
import sys

rep, inv = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{inv}"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:100])
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) > ix["# Samples"]]
S = ix["# Samples"]


def n(r):
    try:
        return int(r[S])
    except ValueError:
        return 0


tot = sum(n(r) for r in data)
print("total samples", tot, "instructions", len(data))
hot = sorted(range(len(data)), key=lambda i: -n(data[i]))[:top]
for i in sorted(hot):
    print(f"{i:5d} {n(data[i]):7d} {100.0 * n(data[i]) / tot:5.1f}%  {data[i][ix['Source']].strip()[:100]}")
