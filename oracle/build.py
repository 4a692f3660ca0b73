"""Build recipe for the oracle (test infrastructure; see oracle/piso_oracle.c header).

  python oracle/build.py            -> oracle/_build/libpiso_oracle.so   (gcc, CPU restatement)
  python oracle/build.py --ref      -> oracle/_ref/libdiffpiso_ref.so    (nvcc, the reference's own
                                       CUDA sources compiled where they lie under /root/reference)

The reference sources are never copied: nvcc reads them in place.  Only
central_difference_csr_op.cu.cc, laplace_op.cu.cc and pressure_solve_op.cu.cc build against
CUDA 12.9; multi_bicgstab_ilu_linear_solve_op.cu.cc needs cusparse csrsv2/CsrmvEx/csr2csc which
were removed from cusparse.h, so it is not part of oracle/_ref.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/CUDAsrc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_oracle(verbose=False):
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(HERE, "piso_oracle.c")
    src2 = os.path.join(HERE, "piso_oracle_adjoint.c")
    srcs = [s for s in (src, src2) if os.path.exists(s)]
    out = os.path.join(out_dir, "libpiso_oracle.so")
    if _stale(out, srcs):
        cmd = ["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-o", out] + srcs + ["-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return out


def build_ref(verbose=False):
    """Compile the reference's own kernels (in place) + our extern-C shim into oracle/_ref."""
    out_dir = os.path.join(HERE, "_ref")
    out = os.path.join(out_dir, "libdiffpiso_ref.so")
    if not os.path.isdir(REF_SRC):
        return out if os.path.exists(out) else None      # GPU box: use the prebuilt file
    os.makedirs(out_dir, exist_ok=True)
    shim = os.path.join(HERE, "ref_shim.cu")
    ref_files = [os.path.join(REF_SRC, f) for f in
                 ("central_difference_csr_op.cu.cc", "laplace_op.cu.cc", "pressure_solve_op.cu.cc")]
    if _stale(out, [shim] + ref_files):
        objs = []
        for f in ref_files + [shim]:
            o = os.path.join(out_dir, os.path.basename(f).split(".")[0] + ".o")
            cmd = ["nvcc", "-x", "cu", "-std=c++17", "-O3", "-w"] + ARCH + \
                  ["-Xcompiler", "-fPIC", "-I", REF_SRC, "-c", f, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            objs.append(o)
        cmd = ["nvcc", "-shared"] + ARCH + ["-o", out] + objs + ["-lcublas", "-lcurand"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        for o in objs:
            os.remove(o)
    return out


if __name__ == "__main__":
    print(build_oracle(verbose=True))
    if "--ref" in sys.argv:
        print(build_ref(verbose=True))
