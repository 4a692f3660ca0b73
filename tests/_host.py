"""Test helper: builds tests/host_shim.cpp (the kernels' per-row code compiled for the host) and wraps it."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_lib = None


def lib():
    global _lib
    if _lib is None:
        out_dir = os.path.join(ROOT, "oracle", "_build")
        os.makedirs(out_dir, exist_ok=True)
        out = os.path.join(out_dir, "libhost_shim.so")
        src = os.path.join(HERE, "host_shim.cpp")
        deps = [src] + [os.path.join(ROOT, "differentiable-piso_b200", "csrc", f) for f in
                        ("rows.cuh", "structure.cuh", "common.cuh")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src])
        _lib = C.CDLL(out)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def csr_structure(ny, nx, per_x, per_y, n_u, n_v, nnz):
    rp = np.zeros(n_u + n_v + 2, np.int32)
    ci = np.zeros(nnz, np.int32)
    lib().hs_csr_structure(ny, nx, int(per_x), int(per_y), _p(rp), _p(ci))
    return rp, ci


def assemble(ny, nx, per_x, per_y, dy, dx, beta, vel, dirichlet, active, noslip, visc, nnz, areas=None):
    from oracle import oracle as O
    area_x, area_y = O.cell_areas(dy, dx) if areas is None else areas
    vel = np.ascontiguousarray(vel, np.float32)
    visc = np.ascontiguousarray(np.atleast_1d(visc), np.float32).ravel()
    values = np.zeros(nnz, np.float32)
    a_diag = np.zeros(vel.size, np.float32)
    f = lib().hs_assemble
    f.argtypes = [C.c_int] * 4 + [C.c_float] * 5 + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 2
    f(ny, nx, int(per_x), int(per_y), dy, dx, area_x, area_y, beta, _p(vel), _p(np.ascontiguousarray(dirichlet, np.uint8)),
      _p(np.ascontiguousarray(active, np.float32)), _p(np.ascontiguousarray(noslip, np.uint8)), _p(visc),
      int(visc.size > 1), _p(values), _p(a_diag))
    return values, a_diag


def fv_gradient(ny, nx, dy, dx, pbc, access, p):
    g = np.zeros(ny * (nx + 1) + (ny + 1) * nx, np.float32)
    f = lib().hs_fv_gradient
    f.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 4
    f(ny, nx, dy, dx, _p(np.ascontiguousarray(pbc, np.int32)), _p(np.ascontiguousarray(access, np.float32)),
      _p(np.ascontiguousarray(p, np.float32)), _p(g))
    return g


def fv_divergence(ny, nx, dy, dx, vel, a_diag=None, beta=0.0):
    div = np.zeros(ny * nx, np.float32)
    f = lib().hs_fv_divergence
    f.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    f(ny, nx, dy, dx, _p(np.ascontiguousarray(vel, np.float32)),
      None if a_diag is None else _p(np.ascontiguousarray(a_diag, np.float32)), beta, _p(div))
    return div


def laplace(ny, nx, active, fluid, k_faces, dtype=np.float64):
    lap = np.zeros(5 * ny * nx, dtype)
    f = lib().hs_laplace_f64 if dtype == np.float64 else lib().hs_laplace_f32
    f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4
    f(ny, nx, _p(np.ascontiguousarray(active, np.float32)), _p(np.ascontiguousarray(fluid, np.float32)),
      _p(np.ascontiguousarray(k_faces, np.float32)), _p(lap))
    return lap
