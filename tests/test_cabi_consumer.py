"""A C++ program written against include/dpiso.h alone (tests/cabi_smoke.cpp): the C ABI is self-sufficient -- structure
tables, workspaces and both solvers without any Python-side construction (what an op shim of the
reference, CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cc:50-58 / pressure_solve_op.cc:48-84, would call)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "differentiable-piso_b200", "diffpiso_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build():
    from diffpiso_b200 import _native  # noqa: F401  (builds libdpiso.so when missing)
    out_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "cabi_smoke")
    src = os.path.join(ROOT, "tests", "cabi_smoke.cpp")
    deps = [src, os.path.join(ROOT, "include", "dpiso.h"), os.path.join(LIBDIR, "libdpiso.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(CUDA, "include"), src, "-L", LIBDIR, "-ldpiso",
                               "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-Wl,-rpath," + LIBDIR, "-o", exe])
    return exe


def test_header_is_plain_c_and_consumer_links():
    """dpiso.h compiles as C99 (no C++-isms, no torch / CUDA types) and the C++ consumer links against libdpiso.so."""
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", "-I", os.path.join(ROOT, "include"), "-"],
                   input='#include "dpiso.h"\nint main(void) { return dpiso_version() == 0; }\n', text=True, check=True)
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_cabi_consumer_runs_both_solvers_without_python():
    exe = _build()
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + ":" + os.path.join(CUDA, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    print(r.stdout)
    print(r.stderr)
    assert r.returncode == 0 and "CABI_SMOKE_OK" in r.stdout
