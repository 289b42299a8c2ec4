"""Per-kernel summary of an ncu --set full report: duration, tensor pipe %, DRAM bytes / throughput, registers."""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = {
    "gpu__time_duration.sum": "dur_us",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit",
    "launch__grid_size": "grid",
}
idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
kn = idx["Kernel Name"]
cols = [h for h in hdr if h in want]
print("kernel".ljust(58), " ".join(want[c].rjust(11) for c in cols))


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[kn]).replace("pcseg::", "").replace("void ", "")
    vals = []
    for c in cols:
        v = fnum(r[idx[c]])
        u = units[idx[c]]
        if want[c] == "dur_us":
            v = v / 1000.0 if u == "ns" else (v * 1000.0 if u == "ms" else v)
        if want[c] in ("dram_rd", "dram_wr"):
            mult = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            v = v * mult   # MB
        vals.append(v)
    print(name[:58].ljust(58), " ".join(f"{v:11.1f}" for v in vals))
