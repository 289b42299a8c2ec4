"""Append the '# traffic tag=<t> MB=<x>' lines bench.py parses (roofline.traffic) to an ncu_summary.py table.
    python tools/ncu_traffic_lines.py <train summary txt> [<eval summary txt>] >> profiles/rNN_ncu_full_gemm_cfg2_train.txt
tag 5 = global_feat forward (gemm_kernel<256, 6, 0>), 21 = its data gradient (the longest gemm_kernel<256, 8, 0>), 53 = the Gram
matrix of its input (the longest gemm_kernel<256, 4, 1>), 69 = the inference forward (gemm_kernel<256, 1, 0>)."""
import sys


def rows(path):
    out = []
    for ln in open(path):
        if ln.startswith("gemm_kernel"):
            name = ln[:58].strip()
            vals = ln[58:].split()
            out.append((name, float(vals[0]), float(vals[1]), float(vals[3])))      # dram_rd MB, dram_wr MB, dur us
    return out


def pick(rs, name):
    c = [r for r in rs if r[0].replace(", 0, 0>", ">") == name]       # (the kernel has five template arguments since round 2)
    return max(c, key=lambda r: r[3]) if c else None


train = rows(sys.argv[1])
for tag, name in ((5, "gemm_kernel<256, 6, 0>"), (21, "gemm_kernel<256, 8, 0>"), (53, "gemm_kernel<256, 4, 1>")):
    r = pick(train, name)
    if r:
        print(f"# traffic tag={tag} MB={r[1] + r[2]:.1f} ({name}: dram read {r[1]:.1f} + write {r[2]:.1f} MB, {r[3]:.1f} us under ncu)")
if len(sys.argv) > 2:
    r = pick(rows(sys.argv[2]), "gemm_kernel<256, 1, 0>")
    if r:
        print(f"# traffic tag=69 MB={r[1] + r[2]:.1f} (gemm_kernel<256, 1, 0>: dram read {r[1]:.1f} + write {r[2]:.1f} MB, {r[3]:.1f} us under ncu)")
