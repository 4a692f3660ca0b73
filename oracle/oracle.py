"""ctypes front-end of the CPU oracle (oracle/piso_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of piso_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
All arrays are numpy; every function handles ONE sample (the reference's batch size).
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import build as _build  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_oracle())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def sizes(ny, nx, per_x, per_y):
    n = np.zeros(2, np.int32)
    nnz = np.zeros(2, np.int32)
    rc = lib().orc_sizes(ny, nx, int(per_x), int(per_y), _p(n, C.c_int), _p(nnz, C.c_int))
    if rc:
        raise ValueError("grid too small for the 5-point pattern (need ny,nx >= 3)")
    return int(n[0]), int(n[1]), int(nnz[0]), int(nnz[1])


def csr_structure(ny, nx, per_x, per_y):
    n_u, n_v, z_u, z_v = sizes(ny, nx, per_x, per_y)
    row_ptr = np.zeros(n_u + n_v + 2, np.int32)
    col_ind = np.zeros(z_u + z_v, np.int32)
    lib().orc_csr_structure(ny, nx, int(per_x), int(per_y), _p(row_ptr, C.c_int), _p(col_ind, C.c_int))
    return row_ptr, col_ind


def pad_velocity(ny, nx, per_x, per_y, u, v):
    u, v = _f32(u), _f32(v)
    assert u.shape == (ny, nx + 1) and v.shape == (ny + 1, nx)
    up = np.zeros((ny + 2, nx + 3), np.float32)
    vp = np.zeros((ny + 3, nx + 2), np.float32)
    lib().orc_pad_velocity(ny, nx, int(per_x), int(per_y), _p(u, C.c_float), _p(v, C.c_float),
                           _p(up, C.c_float), _p(vp, C.c_float))
    return up, vp


def cell_areas(dy, dx):
    """cell_area op input (diffpiso/piso_tf.py:97): prod(dx) / dx[::-1].astype(float32) -> fp32"""
    prod = float(dy) * float(dx)
    return float(np.float32(prod / float(np.float32(dx)))), float(np.float32(prod / float(np.float32(dy))))


def assemble(ny, nx, per_x, per_y, dy, dx, beta, up, vp, dirichlet, active, noslip, visc, row_ptr, areas=None):
    n_u, n_v, z_u, z_v = sizes(ny, nx, per_x, per_y)
    up, vp, active, visc = _f32(up), _f32(vp), _f32(active).ravel(), _f32(np.atleast_1d(visc)).ravel()
    dirichlet, noslip, row_ptr = _u8(dirichlet).ravel(), _u8(noslip).ravel(), _i32(row_ptr)
    assert dirichlet.size == n_u + n_v and active.size == (ny + 2) * (nx + 2) == noslip.size
    assert visc.size in (1, n_u + n_v)
    values = np.zeros(z_u + z_v, np.float32)
    a_diag = np.zeros(n_u + n_v, np.float32)
    f = lib().orc_assemble
    f.argtypes = [C.c_int] * 4 + [C.c_float] * 5 + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 3
    area_x, area_y = cell_areas(dy, dx) if areas is None else areas
    rc = f(ny, nx, int(per_x), int(per_y), dy, dx, area_x, area_y, beta, up.ctypes.data, vp.ctypes.data, dirichlet.ctypes.data,
           active.ctypes.data, noslip.ctypes.data, visc.ctypes.data, int(visc.size > 1), row_ptr.ctypes.data,
           values.ctypes.data, a_diag.ctypes.data)
    assert rc == 0
    return values, a_diag


def fv_gradient(ny, nx, dy, dx, pbc, access, p):
    pbc, access, p = _i32(pbc), _f32(access).ravel(), _f32(p).ravel()
    g = np.zeros(ny * (nx + 1) + (ny + 1) * nx, np.float32)
    f = lib().orc_fv_gradient
    f.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 4
    f(ny, nx, dy, dx, pbc.ctypes.data, access.ctypes.data, p.ctypes.data, g.ctypes.data)
    return g


def fv_divergence(ny, nx, dy, dx, vel):
    vel = _f32(vel).ravel()
    div = np.zeros(ny * nx, np.float32)
    f = lib().orc_fv_divergence
    f.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 2
    f(ny, nx, dy, dx, vel.ctypes.data, div.ctypes.data)
    return div


def spmv(rp, ci, val, x):
    rp, ci, val, x = _i32(rp), _i32(ci), _f32(val), _f32(x)
    n = rp.size - 1
    y = np.zeros(n, np.float32)
    f = lib().orc_spmv_f32
    f.argtypes = [C.c_int] + [C.c_void_p] * 5
    f(n, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, x.ctypes.data, y.ctypes.data)
    return y


def csr_transpose(rp, ci, val):
    rp, ci, val = _i32(rp), _i32(ci), _f32(val)
    n = rp.size - 1
    trp, tci, tval = np.zeros_like(rp), np.zeros_like(ci), np.zeros_like(val)
    f = lib().orc_csr_transpose_f32
    f.argtypes = [C.c_int] + [C.c_void_p] * 6
    f(n, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, trp.ctypes.data, tci.ctypes.data, tval.ctypes.data)
    return trp, tci, tval


def ilu0(rp, ci, val):
    rp, ci, val = _i32(rp), _i32(ci), _f32(val)
    n = rp.size - 1
    lu = np.zeros_like(val)
    f = lib().orc_ilu0_f32
    f.argtypes = [C.c_int] + [C.c_void_p] * 4
    zp = f(n, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, lu.ctypes.data)
    return lu, zp


def lu_solve(rp, ci, lu, b):
    rp, ci, lu, b = _i32(rp), _i32(ci), _f32(lu), _f32(b)
    n = rp.size - 1
    y, x = np.zeros(n, np.float32), np.zeros(n, np.float32)
    f = lib().orc_lu_solve_f32
    f.argtypes = [C.c_int] + [C.c_void_p] * 6
    f(n, rp.ctypes.data, ci.ctypes.data, lu.ctypes.data, b.ctypes.data, y.ctypes.data, x.ctypes.data)
    return x


def bicgstab_ilu(rp, ci, val, rhs, x0, tol, max_it, transpose=False, fp64=False):
    """Returns x, dict(iterations, restarts, warn, exit_kind, residual).  fp64: the cast_to_double=True variant (fp32 in,
    fp64 solve, fp32 out)."""
    rp, ci, val, rhs, x0 = _i32(rp), _i32(ci), _f32(val), _f32(rhs), _f32(x0)
    n = rp.size - 1
    x = np.zeros(n, np.float32)
    stats = np.zeros(4, np.int32)
    res = C.c_float(0)
    f = lib().orc_bicgstab_ilu_f64 if fp64 else lib().orc_bicgstab_ilu_f32
    f.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_float, C.c_int, C.c_int] + [C.c_void_p] * 2 + [C.POINTER(C.c_float)]
    f(n, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, rhs.ctypes.data, x0.ctypes.data, tol, max_it,
      int(transpose), x.ctypes.data, stats.ctypes.data, C.byref(res))
    return x, dict(iterations=int(stats[0]), restarts=int(stats[1]), warn=int(stats[2]),
                   exit_kind=int(stats[3]), residual=float(res.value))


def laplace(ny, nx, active, fluid, k_faces, dtype=np.float64):
    active, fluid, k_faces = _f32(active).ravel(), _f32(fluid).ravel(), _f32(k_faces).ravel()
    assert k_faces.size == (ny + 1) * nx + ny * (nx + 1)
    lap = np.zeros(5 * ny * nx, dtype)
    f = lib().orc_laplace_f64 if dtype == np.float64 else lib().orc_laplace_f32
    f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4
    f(ny, nx, active.ctypes.data, fluid.ctypes.data, k_faces.ctypes.data, lap.ctypes.data)
    return lap


def pressure_cg(ny, nx, per_x, per_y, lap, div, accuracy, max_it, residual_reset, rank_deficient):
    dtype = lap.dtype
    lap = np.ascontiguousarray(lap)
    div = np.ascontiguousarray(div, dtype=dtype).ravel()
    x = np.zeros(ny * nx, dtype)
    it = C.c_int(0)
    f = lib().orc_pressure_cg_f64 if dtype == np.float64 else lib().orc_pressure_cg_f32
    f.argtypes = [C.c_int] * 4 + [C.c_void_p] * 2 + [C.c_float] + [C.c_int] * 3 + [C.c_void_p, C.POINTER(C.c_int)]
    f(ny, nx, int(per_x), int(per_y), lap.ctypes.data, div.ctypes.data, accuracy, max_it, residual_reset,
      int(rank_deficient), x.ctypes.data, C.byref(it))
    return x, it.value


def h_apply(rp, ci, val, a_diag, beta, d):
    rp, ci, val, a_diag, d = _i32(rp), _i32(ci), _f32(val), _f32(a_diag), _f32(d)
    n = rp.size - 1
    h = np.zeros(n, np.float32)
    f = lib().orc_h_apply
    f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_float] + [C.c_void_p] * 2
    f(n, rp.ctypes.data, ci.ctypes.data, val.ctypes.data, a_diag.ctypes.data, beta, d.ctypes.data, h.ctypes.data)
    return h


def step_constants(dy, dx, dt):
    """fp32 graph constants formed in fp64 on the Python side of the reference: beta = prod(dx)/dt (piso_tf.py:26),
    prod(dx), dx_factor = prod(dx)/dx[0]**2 (piso_tf.py:53; dx[0] = dy)."""
    prod = float(dy) * float(dx)
    return dict(beta=float(np.float32(prod / float(dt))), prod=float(np.float32(prod)),
                dx_factor=float(np.float32(prod / (float(dy) ** 2))))


class _Extra(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("values", "a_diag", "rhs", "u_star", "div1", "p1", "u_s2", "h", "div2", "p2", "lap64", "lap32")]


def piso_step(setup, vel, pres, forcing=None, dirichlet_values=None, full_output=False):
    """One forward PISO step of ONE sample.

    setup: dict with ny, nx, per_y, per_x, dy, dx, dt, pbc, pbc_inc, dirichlet (flat [u,v] uint8),
    dirichlet_values (flat [u,v]), active, access, noslip ((ny+2)(nx+2)), visc (scalar or flat field),
    bicg_tol, bicg_max_it, cg_tol, cg_max_it, cg_reset, rank_deficient, cg_fp64; optional vel_pad_periodic = (y, x):
    False pads the velocity by replication on a periodic axis (steps >= 2 of the reference's run_piso_steps).
    vel flat [u,v] float32, pres (ny*nx).  Returns vel_next, pres_next, stats[, extras].
    """
    s = setup
    ny, nx = s["ny"], s["nx"]
    n_u, n_v, z_u, z_v = sizes(ny, nx, s["per_x"], s["per_y"])
    nf, nc, nz = n_u + n_v, ny * nx, z_u + z_v
    visc = _f32(np.atleast_1d(s["visc"])).ravel()
    ip = np.array([ny, nx, int(s["per_y"]), int(s["per_x"])] + list(s["pbc"]) + list(s["pbc_inc"]) +
                  [int(visc.size > 1), s["bicg_max_it"], s["cg_max_it"], s["cg_reset"], int(s["rank_deficient"]),
                   int(s.get("cg_fp64", True))] + [int(k) for k in s.get("vel_pad_periodic", (True, True))], np.int32)
    c = step_constants(s["dy"], s["dx"], s["dt"])
    ax, ay = cell_areas(s["dy"], s["dx"])
    fp = np.array([s["dy"], s["dx"], s["dt"], s["bicg_tol"], s["cg_tol"], c["beta"], c["prod"], c["dx_factor"], ax, ay],
                  np.float32)
    vel, pres = _f32(vel).ravel(), _f32(pres).ravel()
    assert vel.size == nf and pres.size == nc
    dmask = _u8(s["dirichlet"]).ravel()
    dvals = _f32(s["dirichlet_values"] if dirichlet_values is None else dirichlet_values).ravel()
    active, access, noslip = _f32(s["active"]).ravel(), _f32(s["access"]).ravel(), _u8(s["noslip"]).ravel()
    assert dmask.size == nf == dvals.size and active.size == (ny + 2) * (nx + 2) == access.size == noslip.size
    frc = None if forcing is None else _f32(forcing).ravel()
    vel_out, pres_out = np.zeros(nf, np.float32), np.zeros(nc, np.float32)
    stats = np.zeros(12, np.int32)
    ex = _Extra()
    extras = {}
    if full_output:
        shapes = dict(values=nz, a_diag=nf, rhs=nf, u_star=nf, div1=nc, p1=nc, u_s2=nf, h=nf, div2=nc, p2=nc)
        for k, n in shapes.items():
            extras[k] = np.zeros(n, np.float32)
            setattr(ex, k, extras[k].ctypes.data)
        if s.get("cg_fp64", True):
            extras["lap"] = np.zeros(5 * nc, np.float64)
            ex.lap64 = extras["lap"].ctypes.data
        else:
            extras["lap"] = np.zeros(5 * nc, np.float32)
            ex.lap32 = extras["lap"].ctypes.data
    f = lib().orc_piso_step
    f.argtypes = [C.c_void_p] * 14 + [C.POINTER(_Extra)]
    rc = f(ip.ctypes.data, fp.ctypes.data, vel.ctypes.data, pres.ctypes.data, dmask.ctypes.data, dvals.ctypes.data,
           active.ctypes.data, access.ctypes.data, noslip.ctypes.data, visc.ctypes.data,
           None if frc is None else frc.ctypes.data, vel_out.ctypes.data, pres_out.ctypes.data, stats.ctypes.data,
           C.byref(ex))
    assert rc == 0
    st = dict(bicg_u=stats[0:4].tolist(), bicg_v=stats[4:8].tolist(), cg1=int(stats[8]), cg2=int(stats[9]))
    if full_output:
        return vel_out, pres_out, st, extras
    return vel_out, pres_out, st
