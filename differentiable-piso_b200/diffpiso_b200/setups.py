"""Case setups (masks, boundary values, solver settings) for the PISO step, as plain numpy.

These restate the reference's setup code so that the same inputs can be fed to the CUDA path, the
CPU oracle and (on the GPU box) the reference's own kernels:

* lid-driven cavity            -- lid_driven_cavity_2d.py:10-47
* spatial mixing layer         -- diffpiso/combined_training_integrated.py:481-539,
                                  diffpiso/piso_helpers.py:73-133 (compute_mixingLayer_masks)
* temporal mixing layer masks  -- diffpiso/piso_helpers.py:136-166 (temporal_mixing_layer_masks)
* fully periodic box           -- no reference script; synthesised as SURVEY.md §8(d) C2 describes

Array conventions are the reference's: staggered tensors [1, ny+1, nx+1, 2] with channel 0 = v
(y-velocity), channel 1 = u; padded-centred masks [1, ny+2, nx+2, 1].
"""
import math

import numpy as np

REPLICATE, ZERO, PERIODIC = 0, 1, 2   # pressure ghost-cell rule per side ('boundary', 'constant', 'periodic')


def stack_staggered(v_comp, u_comp):
    """PhiFlow stack_staggered_components (phi/physics/field/staggered_grid.py:42-46) for 2-D:
    v [B,ny+1,nx], u [B,ny,nx+1] -> [B,ny+1,nx+1,2] (zero padded)."""
    b, ny1, nx = v_comp.shape
    out = np.zeros((b, ny1, nx + 1, 2), dtype=np.result_type(v_comp, u_comp))
    out[:, :, :nx, 0] = v_comp
    out[:, :ny1 - 1, :, 1] = u_comp
    return out


def unstack_staggered(t):
    """-> (v [B,ny+1,nx], u [B,ny,nx+1])   (staggered_grid.py:33-39)"""
    return t[:, :, :-1, 0], t[:, :-1, :, 1]


def flatten_staggered(t, coord_flip=True):
    """diffpiso/piso_helpers.py:175-185 for one sample: [u rows.., v rows..] when coord_flip."""
    v, u = unstack_staggered(np.asarray(t))
    parts = [u.reshape(u.shape[0], -1), v.reshape(v.shape[0], -1)]
    if not coord_flip:
        parts = parts[::-1]
    return np.concatenate(parts, axis=1)


def stagger_flat(flat, ny, nx, coord_flip=True):
    """Inverse of flatten_staggered (piso_helpers.py:188-206)."""
    flat = np.asarray(flat)
    if flat.ndim == 1:
        flat = flat[None]
    n_u, n_v = ny * (nx + 1), (ny + 1) * nx
    if coord_flip:
        u, v = flat[:, :n_u], flat[:, n_u:]
    else:
        v, u = flat[:, :n_v], flat[:, n_v:]
    return stack_staggered(v.reshape(-1, ny + 1, nx), u.reshape(-1, ny, nx + 1))


def spacing(length, n):
    """dx of a PhiFlow Domain: the box size is stored in fp32, dx = size / resolution is formed in fp64
    (PhiFlow/phi/physics/field/grid.py:87-89); every reference script gets its spacings this way."""
    return float(np.float32(length)) / n


def _base(ny, nx, dy, dx, dt, per_y, per_x):
    return dict(ny=ny, nx=nx, dy=float(dy), dx=float(dx), dt=float(dt), per_y=bool(per_y), per_x=bool(per_x))


def _finish(s, dirichlet_mask, dirichlet_values, active, access, noslip):
    """Adds the flat views the native code consumes next to the reference-shaped arrays."""
    s["dirichlet_mask"] = dirichlet_mask.astype(bool)
    s["dirichlet_values_staggered"] = dirichlet_values.astype(np.float32)
    s["active_mask"] = active.astype(np.float32)
    s["accessible_mask"] = access.astype(np.float32)
    s["no_slip_mask"] = noslip.astype(bool).ravel()
    s["dirichlet"] = flatten_staggered(dirichlet_mask)[0].astype(np.uint8)
    s["dirichlet_values"] = flatten_staggered(dirichlet_values)[0].astype(np.float32)
    s["active"] = s["active_mask"].ravel()
    s["access"] = s["accessible_mask"].ravel()
    s["noslip"] = s["no_slip_mask"].astype(np.uint8)
    return s


def rank_deficient_from_masks(access, active):
    """diffpiso/piso_cuda_pressure_solver.py:84-87"""
    prod = access * active + (1 - access) * (1 - active)
    v = np.prod(prod[0, 0, 1:-1, 0]) * np.prod(prod[0, -1, 1:-1, 0]) * \
        np.prod(prod[0, 1:-1, 0, 0]) * np.prod(prod[0, 1:-1, -1, 0])
    return bool(v)


def lid_driven_cavity(n=32, re=100.0, dt=0.01, bicg_tol=1e-8, bicg_max_it=100, cg_tol=1e-8, cg_max_it=1000,
                      cg_reset=10):
    """lid_driven_cavity_2d.py:10-47 (Domain([N+1, N]); the top cell row is a ghost lid row)."""
    ny, nx = n + 1, n
    s = _base(ny, nx, spacing(1.0 + 1.0 / n, ny), spacing(1.0, nx), dt, False, False)   # box[0:1+1/N, 0:1]
    dm_v = np.zeros((1, ny + 1, nx), np.float32)
    dm_v[:, 0] = 1
    dm_v[:, -2:] = 1
    dm_u = np.zeros((1, ny, nx + 1), np.float32)
    dm_u[..., 0] = 1
    dm_u[..., -1] = 1
    dm_u[:, -1] = 1
    dv_v = np.zeros_like(dm_v)
    dv_u = np.zeros_like(dm_u)
    dv_u[:, -1] = 1
    access = np.pad(np.ones((1, ny, nx, 1), np.float32), ((0, 0), (1, 1), (1, 1), (0, 0)))
    access[0, -2] = 0
    active = access.copy()
    noslip = np.zeros((1, ny + 2, nx + 2, 1), bool)
    noslip[0, 0] = 1
    noslip[0, -2:] = 1
    noslip[0, :, 0] = 1
    noslip[0, :, -1] = 1
    s.update(pbc=[REPLICATE] * 4, pbc_inc=[REPLICATE] * 4, visc=np.float32(1.0 / re), rank_deficient=True,
             bicg_tol=bicg_tol, bicg_max_it=bicg_max_it, cg_tol=cg_tol, cg_max_it=cg_max_it, cg_reset=cg_reset,
             cg_fp64=True, name="lid_driven_cavity_%d" % n)
    return _finish(s, stack_staggered(dm_v, dm_u), stack_staggered(dv_v, dv_u), active, access, noslip)


def obstacle_channel(ny=16, nx=24, block=(6, 10, 8, 13), ly=None, lx=None, visc=1e-2, dt=0.02, solver_precision=1e-8,
                     cg_reset=1000):
    """Closed box with a solid rectangular obstacle (cells [y0, y1) x [x0, x1)) -- no reference script builds one, but
    the kernels' mask logic (solid neighbours in the Laplace rows laplace_op.cu.cc:125-174, the `tBB` / no-slip terms of
    the momentum rows central_difference_csr_op.cu.cc:252-288) is only exercised by interior solid cells.  Masks follow
    the conventions of lid_driven_cavity_2d.py:19-43: active = accessible = 1 in fluid cells, Dirichlet (value 0) on
    every face that touches a solid cell or the wall, no-slip flag on solid cells."""
    ly = float(ny) / 8 if ly is None else ly
    lx = float(nx) / 8 if lx is None else lx
    s = _base(ny, nx, spacing(ly, ny), spacing(lx, nx), dt, False, False)
    y0, y1, x0, x1 = block
    fluid = np.ones((ny, nx), np.float32)
    fluid[y0:y1, x0:x1] = 0
    padded = np.pad(fluid, ((1, 1), (1, 1)))                         # solid ring = walls
    # u face (y, x) sits between padded cells (y+1, x) and (y+1, x+1); v face (y, x) between (y, x+1) and (y+1, x+1)
    dm_u = 1.0 - np.minimum(padded[1:-1, :-1], padded[1:-1, 1:])
    dm_v = 1.0 - np.minimum(padded[:-1, 1:-1], padded[1:, 1:-1])
    mask = padded[None, :, :, None].astype(np.float32)
    noslip = (padded == 0)[None, :, :, None]
    s.update(pbc=[ZERO] * 4, pbc_inc=[ZERO] * 4, visc=np.float32(visc), rank_deficient=rank_deficient_from_masks(mask, mask),
             bicg_tol=solver_precision, bicg_max_it=1000, cg_tol=solver_precision, cg_max_it=5000, cg_reset=cg_reset,
             cg_fp64=True, name="obstacle_channel_%dx%d" % (ny, nx))
    return _finish(s, stack_staggered(dm_v[None].astype(np.float32), dm_u[None].astype(np.float32)),
                   np.zeros((1, ny + 1, nx + 1, 2), np.float32), mask, mask.copy(), noslip)


def periodic_box(ny=128, nx=128, length=2 * math.pi, visc=1e-3, dt=None, cfl=0.5, umax=1.0, bicg_tol=1e-8,
                 bicg_max_it=10000, cg_tol=1e-8, cg_max_it=10000, cg_reset=1000, cg_fp64=True):
    """Fully periodic box (C2 / C5 of BASELINE.json).  All masks 1, no Dirichlet faces."""
    dy, dx = spacing(length, ny), spacing(length, nx)
    if dt is None:
        dt = cfl * min(dy, dx) / umax
    s = _base(ny, nx, dy, dx, dt, True, True)
    access = np.ones((1, ny + 2, nx + 2, 1), np.float32)
    dm = np.zeros((1, ny + 1, nx + 1, 2), np.float32)
    s.update(pbc=[PERIODIC] * 4, pbc_inc=[PERIODIC] * 4, visc=np.float32(visc),
             rank_deficient=rank_deficient_from_masks(access, access),
             bicg_tol=bicg_tol, bicg_max_it=bicg_max_it, cg_tol=cg_tol, cg_max_it=cg_max_it, cg_reset=cg_reset,
             cg_fp64=cg_fp64, name="periodic_%dx%d" % (ny, nx))
    return _finish(s, dm, dm.copy(), access, access.copy(), np.zeros((1, ny + 2, nx + 2, 1), bool))


def temporal_mixing_layer(ny=128, nx=256, ly=None, lx=None, visc=2e-3, dt=0.05, bicg_tol=1e-6, bicg_max_it=10000,
                          cg_tol=1e-6, cg_max_it=10000, cg_reset=1000):
    """Periodic in x, walls (v Dirichlet 0) in y -- masks of piso_helpers.py:136-166."""
    ly = float(ny) if ly is None else ly
    lx = float(nx) if lx is None else lx
    s = _base(ny, nx, spacing(ly, ny), spacing(lx, nx), dt, False, True)
    dm_v = np.zeros((1, ny + 1, nx), np.float32)
    dm_v[:, 0] = 1
    dm_v[:, -1] = 1
    dm_u = np.zeros((1, ny, nx + 1), np.float32)
    access = np.concatenate([np.zeros((1, nx + 2)), np.ones((ny, nx + 2)), np.zeros((1, nx + 2))], 0)
    access = access[None, :, :, None].astype(np.float32)
    s.update(pbc=[ZERO, ZERO, PERIODIC, PERIODIC], pbc_inc=[ZERO, ZERO, PERIODIC, PERIODIC], visc=np.float32(visc),
             rank_deficient=rank_deficient_from_masks(access, access),
             bicg_tol=bicg_tol, bicg_max_it=bicg_max_it, cg_tol=cg_tol, cg_max_it=cg_max_it, cg_reset=cg_reset,
             cg_fp64=True, name="temporal_mixing_layer_%dx%d" % (ny, nx))
    return _finish(s, stack_staggered(dm_v, dm_u), np.zeros((1, ny + 1, nx + 1, 2), np.float32), access,
                   access.copy(), np.zeros((1, ny + 2, nx + 2, 1), bool))


def spatial_mixing_layer(ny=128, nx=512, box=(64.0, 256.0), visc=0.002, sponge_ratio=0.875, relative_sponge_max=20.0,
                         average_velocity=1.0, velocity_difference=1.0, sharpness=2.0, dt=0.05,
                         solver_precision=1e-8, perturbation=None):
    """combined_training_integrated.py:481-539 + piso_helpers.py:73-133.

    Boundaries ((OPEN, OPEN), (OPEN, CLOSED)): v Dirichlet 0 on the y walls, u Dirichlet tanh inflow on
    column 0, free outflow.  Viscosity is a per-face field with a linear sponge ramp.
    """
    ly, lx = box
    s = _base(ny, nx, spacing(ly, ny), spacing(lx, nx), dt, False, False)
    inlet = velocity_difference / 2 * np.tanh(sharpness * (np.linspace(0, ly, ny + 2) - ly / 2)) + average_velocity
    if perturbation is not None:
        inlet = inlet + perturbation
    dm_v = np.zeros((1, ny + 1, nx), np.float32)
    dm_v[:, 0] = 1
    dm_v[:, -1] = 1
    dm_u = np.zeros((1, ny, nx + 1), np.float32)
    dm_u[..., 0] = 1
    dv_v = np.zeros_like(dm_v)
    dv_u = np.zeros_like(dm_u)
    dv_u[0, :, 0] = inlet[1:-1]
    access = np.ones((1, ny + 2, nx + 2, 1), np.float32)
    access[0, :, 0] = 0
    access[0, 0] = 0
    access[0, -1] = 0
    active = np.pad(np.ones((1, ny, nx, 1), np.float32), ((0, 0), (1, 1), (1, 1), (0, 0)))
    # viscosity field: centred array with sponge ramp, sampled at faces (":526-531")
    sponge_start = int(nx * sponge_ratio)
    vc = np.ones((ny, nx)) * visc
    vc[:, sponge_start:] += np.linspace(0, visc * relative_sponge_max, nx - sponge_start)[None, :]
    visc_u = _centered_at_u_faces(vc)
    visc_v = _centered_at_v_faces(vc)
    visc_flat = np.concatenate([visc_u.ravel(), visc_v.ravel()]).astype(np.float32)
    # pressure extrapolation: OPEN -> 'boundary' (replicate); CLOSED -> 'constant' (zero)  (material.py:85-92)
    s.update(pbc=[REPLICATE, REPLICATE, REPLICATE, ZERO], pbc_inc=[REPLICATE] * 4, visc=visc_flat,
             rank_deficient=rank_deficient_from_masks(access, active),
             bicg_tol=solver_precision, bicg_max_it=10000, cg_tol=solver_precision, cg_max_it=10000, cg_reset=1000,
             cg_fp64=True, inlet_profile=inlet.astype(np.float32), name="spatial_mixing_layer_%dx%d" % (ny, nx))
    return _finish(s, stack_staggered(dm_v, dm_u), stack_staggered(dv_v, dv_u), active, access,
                   np.zeros((1, ny + 2, nx + 2, 1), bool))


def _centered_at_u_faces(c):
    """Linear resampling of a centred field to u faces with edge replication (CenteredGrid.at, 'boundary')."""
    p = np.pad(c, ((0, 0), (1, 1)), mode="edge")
    return 0.5 * (p[:, :-1] + p[:, 1:])


def _centered_at_v_faces(c):
    p = np.pad(c, ((1, 1), (0, 0)), mode="edge")
    return 0.5 * (p[:-1, :] + p[1:, :])


# ------------------------------------------------------------------------------------------------
# synthetic initial conditions
# ------------------------------------------------------------------------------------------------

def solenoidal_field(ny, nx, length=2 * math.pi, seed=1234, modes=4, umax=1.0):
    """Divergence-free periodic velocity from a random stream function sampled at cell corners.

    psi = sum_{m,n=1..modes} a_mn sin(m x + phi_mn) sin(n y + theta_mn), a_mn ~ N(0,1)/(m^2+n^2).
    u = d(psi)/dy on u faces, v = -d(psi)/dx on v faces (discrete differences of corner values, so the
    finite-volume divergence vanishes to rounding).  Returns flat [u, v] float32 with the duplicated
    periodic faces filled consistently.
    """
    rng = np.random.RandomState(seed)
    dy, dx = length / ny, length / nx
    yc = np.arange(ny + 1) * dy
    xc = np.arange(nx + 1) * dx
    X, Y = np.meshgrid(xc, yc)
    psi = np.zeros((ny + 1, nx + 1))
    for m in range(1, modes + 1):
        for n in range(1, modes + 1):
            a = rng.randn() / (m * m + n * n)
            ph, th = rng.uniform(0, 2 * math.pi, 2)
            psi += a * np.sin(m * X + ph) * np.sin(n * Y + th)
    u = (psi[1:, :] - psi[:-1, :]) / dy          # [ny, nx+1]
    v = -(psi[:, 1:] - psi[:, :-1]) / dx         # [ny+1, nx]
    scale = umax / max(np.abs(u).max(), np.abs(v).max())
    return np.concatenate([(u * scale).ravel(), (v * scale).ravel()]).astype(np.float32)


def taylor_green(ny, nx, length=2 * math.pi, t=0.0, visc=0.1):
    """Analytic Taylor-Green vortex on the staggered grid (known-answer test for the periodic path)."""
    dy, dx = length / ny, length / nx
    decay = math.exp(-2.0 * visc * t)
    xu, yu = np.meshgrid(np.arange(nx + 1) * dx, (np.arange(ny) + 0.5) * dy)
    xv, yv = np.meshgrid((np.arange(nx) + 0.5) * dx, np.arange(ny + 1) * dy)
    u = np.cos(xu) * np.sin(yu) * decay
    v = -np.sin(xv) * np.cos(yv) * decay
    return np.concatenate([u.ravel(), v.ravel()]).astype(np.float32)


def wall_bounded_field(ny, nx, ly, lx, seed=1234, modes=3, umax=1.0):
    """Divergence-free field for a channel that is periodic in x with impermeable walls in y: discrete curl of a
    stream function that vanishes on both walls (v = 0 on the wall faces, zero net flux).  Flat [u, v] float32."""
    rng = np.random.RandomState(seed)
    dy, dx = ly / ny, lx / nx
    X, Y = np.meshgrid(np.arange(nx + 1) * dx, np.arange(ny + 1) * dy)
    psi = np.zeros((ny + 1, nx + 1))
    for m in range(1, modes + 1):
        for n in range(1, modes + 1):
            a = rng.randn() / (m * m + n * n)
            ph = rng.uniform(0, 2 * math.pi)
            psi += a * np.sin(2 * math.pi * m * X / lx + ph) * np.sin(math.pi * n * Y / ly)
    psi[0, :] = 0.0
    psi[-1, :] = 0.0
    psi[:, -1] = psi[:, 0]
    u = (psi[1:, :] - psi[:-1, :]) / dy
    v = -(psi[:, 1:] - psi[:, :-1]) / dx
    scale = umax / max(np.abs(u).max(), np.abs(v).max())
    return np.concatenate([(u * scale).ravel(), (v * scale).ravel()]).astype(np.float32)
