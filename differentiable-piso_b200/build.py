"""Build libdpiso.so (hand-written CUDA kernels for sm_100a behind the C ABI of include/dpiso.h).

    python differentiable-piso_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library is written next to the Python package
(differentiable-piso_b200/diffpiso_b200/libdpiso.so) so that it travels with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "diffpiso_b200", "libdpiso.so")
SOURCES = ["piso_ops.cu", "pressure_cg.cu", "bicgstab.cu", "bicgstab_band.cu", "bicgstab_tile.cu", "bicgstab_f64.cu", "piso_adjoint.cu", "tables.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--threads", "4"]


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(HERE, "..", "include", "dpiso.h"))
    return files


def build(force=False, verbose=False, timing=False):
    """timing=True: diagnostics build libdpiso_timing.so (-DDPISO_CG_TIMING: %clock stamps of the pressure-CG phases,
    scripts/cg_timing.py); never loaded by the package unless DPISO_LIBRARY points at it."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    out = OUT[:-3] + "_timing.so" if timing else OUT
    extra = ["-DDPISO_CG_TIMING"] if timing else []
    suffix = ".timing.o" if timing else ".o"
    if not force and os.path.exists(out):
        t = os.path.getmtime(out)
        if all(os.path.getmtime(f) <= t for f in _deps()):
            return out
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(CSRC, os.path.basename(s)[:-3] + suffix)
        cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((subprocess.Popen(cmd), cmd))
        objs.append(o)
    for p, cmd in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = ["nvcc", "-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    for o in objs:
        os.remove(o)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, timing="--timing" in sys.argv))
