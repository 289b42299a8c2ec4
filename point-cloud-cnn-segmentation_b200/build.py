"""In-tree build of the CUDA library (sm_100a only).  `python -m pcseg_b200.build` or
`__graft_entry__.build()`.  The .so is git-ignored but travels to the GPU box with the tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "pcseg_api.cu")
DEPS = [SRC] + [os.path.join(HERE, "csrc", f) for f in ("gemm.cuh", "head_chain.cuh", "pointwise.cuh", "ptx.cuh", "bn.cuh", "peer_allreduce.cuh")] + [
    os.path.join(os.path.dirname(HERE), "include", "pcseg_b200.h")]
LIB_DIR = os.path.join(HERE, "lib")
# A/B experiments: PCSEG_LIB_SUFFIX=_x builds / loads lib/libpcseg_b200_x.so, PCSEG_NVCC_EXTRA adds compile flags (-D...)
LIB = os.path.join(LIB_DIR, "libpcseg_b200%s.so" % os.environ.get("PCSEG_LIB_SUFFIX", ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("PCSEG_NVCC_EXTRA", "").split() + ["-o", LIB, SRC]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpcseg_b200.so")
    with open(os.path.join(LIB_DIR, "ptxas%s.log" % os.environ.get("PCSEG_LIB_SUFFIX", "")), "w") as f:
        f.write(res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
