"""Builds libxq_b200.so (hand-written sm_100a kernels + the C ABI of include/xq.h) in-tree with nvcc.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.environ.get("XQ_LIB_PATH") or os.path.join(PKG, "libxq_b200.so")      # XQ_LIB_PATH: a profiling build (-DXQ_TIMELINE) next to the product library
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    """Compile every .cu under csrc/ into libxq_b200.so for sm_100a.  Returns the library path."""
    if not force and not stale():
        return LIB
    cmd = [NVCC] + FLAGS + os.environ.get("XQ_NVCC_EXTRA", "").split() + ["-o", LIB] + sources() + ["-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libxq_b200.so")
    with open(os.path.join(PKG, "build_ptxas.log"), "w") as f:      # git-ignored
        f.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_native(force="-f" in sys.argv, verbose=True))
