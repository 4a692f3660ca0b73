"""CPU restatement of the BACKWARD pass of one PISO step (TEST INFRASTRUCTURE -- see oracle/piso_oracle.c header).

What TF-1.14 autodiff assembles from the reference's gradient registrations:
  * assembly op: no gradient (diffpiso/piso_tf.py:125-126) -> M, A, 1/(beta-A), the pressure matrix are constants;
  * predictor solve: same op on the incoming gradient with `not transpose`, same initial-guess tensor, times (1-warn)
    (diffpiso/linear_solver.py:169-173);
  * pressure solve: same op on the incoming gradient, zero initial guess (diffpiso/piso_cuda_pressure_solver.py:97-107);
  * finite_volume_divergence: registered gradient (diffpiso/piso_helpers.py:291-305), periodic branch = quirk Q19;
  * finite_volume_gradient_tensor: autodiff of pad/slice/multiply; periodic axes use circular_padded_gradient's
    registered gradient (diffpiso/piso_helpers.py:226-233), quirk Q20;
  * explicit_H_csr: autodiff of gather / segment_sum = transposed product (diffpiso/piso_helpers.py:214-223).
Parity status: pinned against the reference's own Python -- every `grad` closure registered with tf.custom_gradient
executed from source on PhiFlow's torch backend (tests/golden/reference_runner.py, tests/golden/ref_python/step_*.npz,
tests/test_cpu_reference_python.py): op order, adjoint right-hand sides / initial guess, and the gradients w.r.t.
velocity, pressure, forcing and Dirichlet values to solver tolerance.  The non-periodic branches are also checked to be
exact transposes of the forward operators by dot-product tests in tests/.
All arrays float32 numpy, one sample.
"""
import numpy as np
import scipy.sparse as sp

from . import oracle as O

f32 = np.float32


def fv_gradient_adj(ny, nx, dy, dx, pbc, access, gs):
    """g_p = G^T gs (gs on faces, flat [u, v])."""
    prod = f32(np.float64(f32(dy)) * np.float64(f32(dx)))
    n_u = ny * (nx + 1)
    acc = np.asarray(access, f32).reshape(ny + 2, nx + 2)
    gu = np.asarray(gs[:n_u], f32).reshape(ny, nx + 1)
    gv = np.asarray(gs[n_u:], f32).reshape(ny + 1, nx)
    mk_u = np.minimum(acc[1:-1, :-1], acc[1:-1, 1:])          # [ny, nx+1]
    mk_v = np.minimum(acc[:-1, 1:-1], acc[1:, 1:-1])          # [ny+1, nx]
    su = ((gu * mk_u) / f32(dx)) * prod
    sv = ((gv * mk_v) / f32(dy)) * prod
    wl = np.ones(nx, f32); wr = np.ones(nx, f32); wb = np.ones(ny, f32); wt = np.ones(ny, f32)
    if pbc[2] == 0: wl[0] = 0
    if pbc[3] == 0: wr[-1] = 0
    if pbc[0] == 0: wb[0] = 0
    if pbc[1] == 0: wt[-1] = 0
    gy = wb[:, None] * sv[:-1, :] - wt[:, None] * sv[1:, :]
    gx = wl[None, :] * su[:, :-1] - wr[None, :] * su[:, 1:]
    return (gy + gx).astype(f32).ravel()


def fv_divergence_adj(ny, nx, per_x, per_y, dy, dx, gc):
    """g_vel = registered gradient of finite_volume_divergence (flat [u, v])."""
    prod = f32(np.float64(f32(dy)) * np.float64(f32(dx)))
    g = np.asarray(gc, f32).reshape(ny, nx)
    # x component (u faces, axis 1)
    if per_x:
        hi = np.concatenate([g, g[:, 0:1]], axis=1)
        lo = np.concatenate([g[:, -2:-1], g], axis=1)
    else:
        z = np.zeros((ny, 1), f32)
        hi = np.concatenate([g, z], axis=1)
        lo = np.concatenate([z, g], axis=1)
    gu = (-hi * prod) / f32(dx) + (lo * prod) / f32(dx)
    if per_y:
        hi = np.concatenate([g, g[0:1, :]], axis=0)
        lo = np.concatenate([g[-2:-1, :], g], axis=0)
    else:
        z = np.zeros((1, nx), f32)
        hi = np.concatenate([g, z], axis=0)
        lo = np.concatenate([z, g], axis=0)
    gv = (-hi * prod) / f32(dy) + (lo * prod) / f32(dy)
    return np.concatenate([gu.ravel(), gv.ravel()]).astype(f32)


def h_apply_adj(setup, values, a_diag, beta, gh):
    ny, nx = setup["ny"], setup["nx"]
    n_u, n_v, z_u, z_v = O.sizes(ny, nx, setup["per_x"], setup["per_y"])
    rp, ci = O.csr_structure(ny, nx, setup["per_x"], setup["per_y"])
    out = np.zeros(n_u + n_v, f32)
    for (r0, r1, z0, z1, rpc) in ((0, n_u, 0, z_u, rp[:n_u + 1]), (n_u, n_u + n_v, z_u, z_u + z_v, rp[n_u + 1:])):
        m = sp.csr_matrix((values[z0:z1].astype(np.float64), ci[z0:z1], rpc), shape=(r1 - r0, r1 - r0))
        t = (m.T @ gh[r0:r1].astype(np.float64)).astype(f32)
        out[r0:r1] = t - (a_diag[r0:r1] - f32(beta)) * gh[r0:r1]
    return out


def _cg(setup, lap, rhs):
    T = np.float64 if setup.get("cg_fp64", True) else np.float32
    x, it = O.pressure_cg(setup["ny"], setup["nx"], setup["per_x"], setup["per_y"], lap, rhs.astype(T), setup["cg_tol"],
                          setup["cg_max_it"], setup["cg_reset"], setup["rank_deficient"])
    return x.astype(f32), it


def piso_step_adjoint(setup, vel, pres, g_vel, g_pres, forcing=None):
    """Gradients of <g_vel, u_next> + <g_pres, p_next> w.r.t. (vel, pres, forcing, dirichlet_values)."""
    s = setup
    ny, nx = s["ny"], s["nx"]
    n_u, n_v, z_u, z_v = O.sizes(ny, nx, s["per_x"], s["per_y"])
    c = O.step_constants(s["dy"], s["dx"], s["dt"])
    beta = f32(c["beta"])
    prod = f32(np.float64(f32(s["dy"])) * np.float64(f32(s["dx"])))
    vel_next, pres_next, st, ex = O.piso_step(s, vel, pres, forcing=forcing, full_output=True)
    values, a_diag, lap = ex["values"], ex["a_diag"], ex["lap"]
    kt = beta - a_diag                                    # (beta - A)
    g_vel, g_pres = np.asarray(g_vel, f32), np.asarray(g_pres, f32)
    # p_next = p + p1 + p2 ; u_next = u** + (h - G(p2)/prod)/(beta-A)
    t = -((g_vel / kt) / prod)
    p2_bar = g_pres + fv_gradient_adj(ny, nx, s["dy"], s["dx"], s["pbc_inc"], s["access"], t)
    d2_bar, it_a = _cg(s, lap, p2_bar)
    # the registered gradient of finite_volume_divergence follows the velocity grid's extrapolation (piso_helpers.py:291-305)
    vpy, vpx = s.get("vel_pad_periodic", (True, True))
    dpx, dpy = bool(s["per_x"] and vpx), bool(s["per_y"] and vpy)
    h_bar = (g_vel + fv_divergence_adj(ny, nx, dpx, dpy, s["dy"], s["dx"], d2_bar)) / kt
    delta_bar = h_apply_adj(s, values, a_diag, beta, h_bar)
    us2_bar = g_vel + delta_bar
    t = -((us2_bar / kt) / prod)
    p1_bar = g_pres + fv_gradient_adj(ny, nx, s["dy"], s["dx"], s["pbc_inc"], s["access"], t)
    d1_bar, it_b = _cg(s, lap, p1_bar)
    ustar_bar = (us2_bar - delta_bar) + fv_divergence_adj(ny, nx, dpx, dpy, s["dy"], s["dx"], d1_bar)
    rp, ci = O.csr_structure(ny, nx, s["per_x"], s["per_y"])
    neg = -values
    xu, su = O.bicgstab_ilu(rp[:n_u + 1], ci[:z_u], neg[:z_u], ustar_bar[:n_u], vel[:n_u], s["bicg_tol"], s["bicg_max_it"], True)
    xv, sv = O.bicgstab_ilu(rp[n_u + 1:], ci[z_u:], neg[z_u:], ustar_bar[n_u:], vel[n_u:], s["bicg_tol"], s["bicg_max_it"], True)
    warn = max(su["warn"], sv["warn"])
    rhs_bar = np.concatenate([xu, xv]).astype(f32) * f32(1 - warn)
    m = s["dirichlet"].astype(bool)
    gfree = np.where(m, f32(0), rhs_bar).astype(f32)
    out = dict(g_vel=(gfree * beta).astype(f32), g_forcing=(gfree * prod).astype(f32),
               g_dvals=np.where(m, -rhs_bar, f32(0)).astype(f32),
               g_pres=(g_pres + fv_gradient_adj(ny, nx, s["dy"], s["dx"], s["pbc"], s["access"], -gfree)).astype(f32),
               vel_next=vel_next, pres_next=pres_next,
               stats=dict(forward=st, cg_adj=[it_a, it_b], bicg_adj=[su, sv]))
    return out
