"""The PISO step with the reference's call surface (diffpiso/piso_tf.py:11-81, 85-137, 165-182).

`piso_step(velocity, pressure, pressure_inc1, pressure_inc2, dt, simulation_physics, dirichlet_values, ...)` returns
`(StaggeredGrid velocity, CenteredGrid pressure, warn)` (or the 17 intermediates with `full_output`).  Forward and
backward of one step are ONE `torch.autograd.Function` that enqueues the fused native kernels; the two solver plug-ins
of `SimulationParameters` (`linear_solver.solve`, `pressure_solver.solve`) are called exactly where the reference calls
them -- in forward and, for the adjoint, in backward (transposed predictor solve, two more pressure solves).

Backward follows the reference's gradient registrations: the advection matrices, 1/(beta-A), the
pressure matrix, masks and initial guesses are constants; gradients flow to velocity, pressure, the forcing term and the
Dirichlet values.
"""
import numpy as np
import torch

from . import ops
from .grids import (CenteredGrid, StaggeredGrid, as_tensor, extrapolation_codes, flatten_staggered_data,
                    stagger_flattened_data)
from .pressure_solver import ScalingFromDiagonal, _periodic_flags


class SimulationParameters(object):
    """diffpiso/piso_tf.py:165-182.  Masks as in the reference: dirichlet_mask / dirichlet_values in staggered shape
    [1, ny+1, nx+1, 2]; active_mask / accessible_mask [1, ny+2, nx+2, 1]; no_slip_mask indexable as the flattened padded
    centred grid; bool_periodic = (periodic_y, periodic_x)."""

    def __init__(self, dirichlet_mask, dirichlet_values, active_mask, accessible_mask, bool_periodic=None,
                 no_slip_mask=None, viscosity=0., linear_solver=None, pressure_solver=None, stream_groups=None):
        # stream_groups (not in the reference): how many sample groups of a batch `piso_step` runs on concurrent CUDA
        # streams; None / 1 = the whole batch on the caller's stream, "auto" = the rule of `_stream_groups`
        self.stream_groups = stream_groups
        self.pressure_solver = pressure_solver
        self.linear_solver = linear_solver
        self.dirichlet_mask = dirichlet_mask
        self.dirichlet_values = dirichlet_values
        self.active_mask = active_mask
        self.accessible_mask = accessible_mask
        self.no_slip_mask = no_slip_mask
        self.bool_periodic = bool_periodic
        self.viscosity = viscosity


def pressure_extrapolation(boundaries):
    """diffpiso/piso_tf.py:140-162: nested tuple of Material-like objects -> accessible_extrapolation_mode strings."""
    if isinstance(boundaries, tuple):
        return tuple(pressure_extrapolation(b) for b in boundaries)
    return boundaries.accessible_extrapolation_mode


class _Ctx(object):
    """Host-side constants of one step."""
    __slots__ = ("g", "m", "dy", "dx", "areas", "beta", "prod", "dx_factor", "pbc", "pbc_inc", "sim", "unrolling_step",
                 "vel_periodic")


def _linear_solve(c, values, rhs, x0, transpose, unrolling_step, pivots_out=None, pivots_in=None):
    """(-M) x = rhs through the linear_solver plug-in (piso_tf.py:42-43) -> (x, warn, stats | None).  `values` is the
    assembled M; the native solver negates on the fly, a foreign plug-in receives the materialised -M like the
    reference's."""
    ls = c.sim.linear_solver
    if getattr(ls, "_dpiso_native", False):
        return ls.solve_native(c.g, values, rhs, x0, transpose, negate=True, pivots_out=pivots_out, pivots_in=pivots_in,
                               adjoint=transpose)
    shape = (rhs.shape[0], c.g.ny + 1, c.g.nx + 1, 2)
    rp, ci = c.g.csr_structure()
    x, warn = ls.solve(torch.neg(values), rp, ci, rhs, shape, x0, offset=1, transpose=transpose,
                       unrolling_step=unrolling_step)
    return x, warn, None


def _pressure_solve(c, a_diag, div, unrolling_step, scaling=None):
    ps = c.sim.pressure_solver
    b = div.shape[0]
    div4 = div.reshape(b, c.g.ny, c.g.nx, 1)
    if scaling is not None:
        pass
    elif getattr(ps, "_dpiso_native", False):
        scaling = ScalingFromDiagonal(a_diag, c.beta, c.dx_factor)
    else:   # foreign plug-in: materialise 1/(beta - A) * dx_factor as a staggered tensor (piso_tf.py:53-54)
        scaling = stagger_flattened_data((1.0 / (c.beta - a_diag)) * c.dx_factor, (b, c.g.ny + 1, c.g.nx + 1, 2), True)
    p, its, lap = ps.solve(scaling, div4, None, False, c.sim, unrolling_step=unrolling_step)
    return p.reshape(b, c.g.nc), its, lap


class _PisoStepFn(torch.autograd.Function):
    """One PISO step on flat tensors: (vel [B,nf], pres [B,nc], dvals [1|B,nf], forcing [B,nf]|None, visc) ->
    (vel_next, pres_next, p1, p2, warn, extras...).  With the native solver plug-ins, forward and backward enqueue native
    kernels only (no torch arithmetic in between): the whole pass is CUDA-graph capturable."""

    @staticmethod
    def forward(ctx, vel, pres, dvals, forcing, visc, c):
        g, m = c.g, c.m
        vel, pres = vel.contiguous(), pres.contiguous()
        native_ls = getattr(c.sim.linear_solver, "_dpiso_native", False)
        native_ps = getattr(c.sim.pressure_solver, "_dpiso_native", False)
        needs_bwd = any(ctx.needs_input_grad[:4])
        # advection matrices (piso_tf.py:29-33)
        values, a_diag = ops.assemble(g, vel, m["dirichlet"], m["active"], m["noslip"], visc, c.dy, c.dx, c.beta, c.areas,
                                      vel_periodic=c.vel_periodic)
        # predictor (piso_tf.py:36-47); the forward ILU(0) pivots are kept for the adjoint solve (factor reuse)
        rhs = ops.predictor_rhs(g, vel, pres, m["access"], m["dirichlet"], dvals, forcing, c.dy, c.dx, c.beta, c.pbc)
        pivots = None
        if native_ls and needs_bwd and c.sim.linear_solver.reuse_factors and ops.factor_reuse_supported(g):
            pivots = torch.empty_like(vel)
        u_star, warn, _ = _linear_solve(c, values, rhs, vel, False, c.unrolling_step, pivots_out=pivots)
        u_star = u_star.contiguous()
        # corrector 1 (piso_tf.py:51-58)
        div1 = ops.fv_divergence(g, u_star, c.dy, c.dx)
        scaling = ScalingFromDiagonal(a_diag, c.beta, c.dx_factor) if native_ps else None   # both solves share the matrix
        p1, its1, lap1 = _pressure_solve(c, a_diag, div1, c.unrolling_step, scaling)
        u_s2 = ops.corrector1(g, u_star, p1, a_diag, m["access"], c.dy, c.dx, c.beta, c.pbc_inc)
        # corrector 2 (piso_tf.py:61-75)
        h = ops.h_apply(g, values, a_diag, u_star, u_s2, c.beta)
        div2 = ops.fv_divergence(g, h, c.dy, c.dx, a_diag=a_diag, beta=c.beta)
        p2, its2, lap2 = _pressure_solve(c, a_diag, div2, 1000 + c.unrolling_step, scaling)
        vel_next, pres_next = ops.corrector2(g, u_s2, h, p2, a_diag, pres, p1, m["access"], c.dy, c.dx, c.beta, c.pbc_inc)
        ctx.c = c
        ctx.set_materialize_grads(False)        # no zero-filled gradients for the 15 non-differentiable extras
        ctx.has_forcing = forcing is not None
        ctx.dvals_batched = dvals.shape[0] == vel.shape[0]
        ctx.has_pivots = pivots is not None
        ctx.lap_dtype = lap1.dtype if native_ps else None
        saved = [values, a_diag, vel]
        if pivots is not None:
            saved.append(pivots)
        if native_ps:
            saved.append(lap1)                  # the adjoint's two pressure solves use the same matrix
        ctx.save_for_backward(*saved)
        extras = (p1, p2, warn, values, a_diag, rhs, u_star, u_s2, h, div1, div2, lap1, lap2, its1, its2)
        ctx.mark_non_differentiable(*extras)
        return (vel_next, pres_next) + extras

    @staticmethod
    def backward(ctx, g_vel, g_pres, *unused):
        c = ctx.c
        g, m = c.g, c.m
        saved = list(ctx.saved_tensors)
        values, a_diag, vel = saved[:3]
        pivots = saved[3] if ctx.has_pivots else None
        b = vel.shape[0]
        g_vel = torch.zeros_like(vel) if g_vel is None else g_vel.contiguous()
        g_pres = torch.zeros((b, g.nc), dtype=torch.float32, device=vel.device) if g_pres is None else g_pres.contiguous()
        native_ps = ctx.lap_dtype is not None
        scaling = None
        if native_ps:
            scaling = ScalingFromDiagonal(a_diag, c.beta, c.dx_factor)
            scaling._lap[ctx.lap_dtype == torch.float64] = saved[-1].reshape(b, g.nc, 5)
        # p_next = p + p1 + p2 ; u_next = u** + (h - G(p2)/prod)/(beta-A)
        p2_bar = ops.fv_gradient_adj(g, g_vel, m["access"], c.dy, c.dx, c.pbc_inc, a_diag=a_diag, beta=c.beta,
                                     divisor=c.prod, negate=True, base=g_pres)
        d2_bar, _, _ = _pressure_solve(c, a_diag, p2_bar, 1100 + c.unrolling_step, scaling)
        # h_bar = (g_vel + D^T d2_bar) / (beta - A)
        h_bar = ops.fv_divergence_adj(g, d2_bar, c.dy, c.dx, base=g_vel, a_diag=a_diag, beta=c.beta, vel_periodic=c.vel_periodic)
        # delta_bar = H^T h_bar ; u**_bar = g_vel + delta_bar (same pass)
        delta_bar, us2_bar = ops.h_apply_adj(g, values, a_diag, h_bar, c.beta, base=g_vel)
        # u** = u* - G(p1)/(beta-A)/prod
        p1_bar = ops.fv_gradient_adj(g, us2_bar, m["access"], c.dy, c.dx, c.pbc_inc, a_diag=a_diag, beta=c.beta,
                                     divisor=c.prod, negate=True, base=g_pres)
        d1_bar, _, _ = _pressure_solve(c, a_diag, p1_bar, 100 + c.unrolling_step, scaling)
        # u*_bar = (us2_bar - delta_bar) + D^T d1_bar
        ustar_bar = ops.fv_divergence_adj(g, d1_bar, c.dy, c.dx, base=us2_bar, base_sub=delta_bar, vel_periodic=c.vel_periodic)
        # predictor: transposed solve, same initial-guess tensor as forward, times (1 - warn) (linear_solver.py:169-173)
        rhs_bar, warn_b, stats_b = _linear_solve(c, values, ustar_bar, vel, True, 100 + c.unrolling_step, pivots_in=pivots)
        if stats_b is None:                     # foreign plug-in: batch-wide warning scalar, as the reference applies it
            rhs_bar = (rhs_bar * (1.0 - warn_b)).contiguous()
        need_dv = ctx.needs_input_grad[2]
        gvel_in, gforce, gdvals, gfree = ops.predictor_rhs_adj(g, rhs_bar, m["dirichlet"], c.dy, c.dx, c.beta,
                                                               ctx.has_forcing and ctx.needs_input_grad[3], need_dv,
                                                               solve_stats=stats_b)
        gpres_in = ops.fv_gradient_adj(g, gfree, m["access"], c.dy, c.dx, c.pbc, negate=True, base=g_pres)
        if need_dv and not ctx.dvals_batched:
            gdvals = gdvals.sum(0, keepdim=True)
        return gvel_in, gpres_in, gdvals, gforce, None, None


# ---- sample groups on concurrent streams -------------------------------------------------------------------------
# The samples of a batch never exchange data inside a step (the batch is the partition), and on the small
# grids both solvers are latency-bound persistent kernels whose launches end in a tail of a few slow samples (the
# pressure CG runs one cluster per sample until ITS residual test passes; the predictor runs one CTA per system).
# Running sample groups on separate streams lets the solver launches of one group fill the SMs the tail of another
# group's launch leaves idle.  Results are bit-identical to the single-stream step (same kernels, same per-sample
# arithmetic).  Measured on B200, periodic 128^2 x 64, forward + adjoint, groups forked and joined inside every call:
# 10.09 ms (1 group), 9.64 (2), 9.56 (4); as independent pipelines (sharding.SampleGroups): 9.07 (2), 8.47 (4),
# profiles/r02_stream_groups.md.
_GROUP_STREAMS = {}


def _group_streams(device, n):
    key = (device.index if device.index is not None else torch.cuda.current_device(), n)
    st = _GROUP_STREAMS.get(key)
    if st is None:
        st = _GROUP_STREAMS[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
    return st


def _stream_groups(sim, c, b):
    """Number of sample groups `piso_step` itself forks for a batch of b samples: `SimulationParameters.stream_groups`
    (or DPISO_STREAM_GROUPS), default 1.  Opt-in because inside ONE call all groups pass through the same phase together
    (measured gain 4-5 %: 10.09 -> 9.56 ms with 4 groups); the larger gain needs pipelines that are independent across
    forward, adjoint and steps -- `sharding.SampleGroups`, 20 %.  "auto" = 4 groups from 32 samples, 2 from 16, where both
    solver plug-ins are native and the pressure CG keeps its state on chip (larger grids already fill the GPU with one
    cooperative launch), never while a CUDA graph is being captured."""
    import os
    want = getattr(sim, "stream_groups", None)
    if want is None:
        want = os.environ.get("DPISO_STREAM_GROUPS") or 1
    if want != "auto":
        return max(1, min(int(want), b))
    if not (getattr(sim.linear_solver, "_dpiso_native", False) and getattr(sim.pressure_solver, "_dpiso_native", False)):
        return 1
    if torch.cuda.is_current_stream_capturing():
        return 1
    if not ops.pressure_cg_on_chip(c.g, b):
        return 1
    return 4 if b >= 32 else (2 if b >= 16 else 1)


def _apply_grouped(vel, pres, dvals, forcing, visc, c, groups):
    """`_PisoStepFn` on `groups` contiguous sample blocks, each on its own stream; -> list of per-group output tuples.
    The caller's stream waits for every group before it touches the results; gradients flow back through the same
    streams (autograd runs a node's backward on the stream of its forward)."""
    dev = vel.device
    b = vel.shape[0]
    main = torch.cuda.current_stream(dev)
    streams = _group_streams(dev, groups)

    def parts(t):
        if t is None:
            return [None] * groups
        if t.shape[0] == b and b > 1:
            return list(torch.tensor_split(t, groups))
        return [t] * groups
    vs, prs, dvs, fs, vis = parts(vel.contiguous()), parts(pres.contiguous()), parts(dvals), parts(forcing), parts(visc)
    ready = torch.cuda.Event()
    ready.record(main)
    outs = []
    for i, st in enumerate(streams):
        st.wait_event(ready)
        with torch.cuda.stream(st):
            out = _PisoStepFn.apply(vs[i], prs[i], dvs[i], fs[i], vis[i], c)
        done = torch.cuda.Event()
        done.record(st)
        main.wait_event(done)
        outs.append(out)
    # No record_stream on the group outputs: the caller's stream consumes them (concatenation) before it records the next
    # step's `ready` event, which every group stream waits for before it can reuse a block of its own pool; recording
    # them would park every freed block behind an event query and drive a run-ahead host into cudaMalloc (measured:
    # 36 ms instead of 8.5 ms per step).
    return outs


def _flat_faces(x, b, g, name):
    """staggered tensor / StaggeredGrid / flat -> [1|B, nf] float32 on the right device"""
    if isinstance(x, StaggeredGrid):
        return x.flat
    t = as_tensor(x)
    if t.dim() == 4:
        return flatten_staggered_data(t, coord_flip=True).contiguous()
    if t.dim() == 1:
        t = t[None]
    if t.dim() != 2 or t.shape[1] != g.nf:
        raise ValueError("%s must be a staggered tensor [B, ny+1, nx+1, 2] or flat [B, n_u+n_v]" % name)
    return t


def make_step_context(velocity, pressure, pressure_inc, dt, simulation_physics, unrolling_step=0):
    sim = simulation_physics
    ny, nx = velocity.resolution
    per_y, per_x = _periodic_flags(sim)
    device = velocity.flat.device
    if device.type != "cuda":
        raise ops.N.DpisoError("piso_step needs CUDA tensors: the PISO path has no CPU implementation")
    c = _Ctx()
    c.g = ops.Geometry.get(ny, nx, per_y, per_x, device)
    c.m = ops.to_device_masks(sim, c.g)
    # fp32 spacings for the kernels (grid_spacing is an fp32 op input, piso_tf.py:96); the graph constants are formed
    # in fp64 on the host and rounded once to fp32, as the TF graph does with the reference's numpy expressions
    dy64, dx64 = float(velocity.dx[0]), float(velocity.dx[1])
    c.dy, c.dx = float(np.float32(dy64)), float(np.float32(dx64))
    prod = dy64 * dx64
    c.areas = ops.cell_areas(dy64, dx64)                               # piso_tf.py:97
    c.prod = float(np.float32(float(np.float32(dy64)) * float(np.float32(dx64))))   # as the kernels form it
    c.beta = float(np.float32(prod / float(dt)))                       # piso_tf.py:26
    c.dx_factor = float(np.float32(prod / (dy64 * dy64)))              # piso_tf.py:53 (dx[0] = dy)
    c.pbc = extrapolation_codes(pressure.extrapolation)
    c.pbc_inc = extrapolation_codes(pressure_inc.extrapolation)
    c.sim = sim
    c.unrolling_step = unrolling_step
    # how custom_padded pads the velocity and which registered gradient finite_volume_divergence uses: the velocity
    # grid's own extrapolation in the reference; here the simulation's flags unless the grid overrides them (Q21)
    c.vel_periodic = getattr(velocity, "pad_periodic", None)
    return c


def piso_step(velocity, pressure, pressure_inc1, pressure_inc2, dt, simulation_physics, dirichlet_values,
              viscosity_field=None, forcing_term=None, unrolling_step=0, warn=None, full_output=False, **kwargs):
    """diffpiso/piso_tf.py:11-81."""
    sim = simulation_physics
    if sim.linear_solver is None or sim.pressure_solver is None:
        raise ValueError("SimulationParameters needs linear_solver and pressure_solver")
    c = make_step_context(velocity, pressure, pressure_inc1, dt, sim, unrolling_step)
    g = c.g
    vel = velocity.flat
    b = vel.shape[0]
    pres = as_tensor(pressure.data).reshape(b, g.nc)
    dvals = _flat_faces(dirichlet_values, b, g, "dirichlet_values").to(vel.device)
    forcing = None if forcing_term is None else _flat_faces(forcing_term, b, g, "forcing_term").to(vel.device)
    if viscosity_field is None:                                              # piso_tf.py:21-24
        key = (float(sim.viscosity), str(vel.device))
        cached = getattr(sim, "_visc_tensor", None)
        if cached is None or cached[0] != key:                               # uploaded once: no host->device copy per step
            cached = (key, torch.full((1,), key[0], dtype=torch.float32, device=vel.device))
            sim._visc_tensor = cached
        visc = cached[1]
    else:
        visc = as_tensor(viscosity_field).to(vel.device)
        if visc.dim() == 4:
            visc = flatten_staggered_data(visc, coord_flip=True)
    groups = _stream_groups(sim, c, b)
    if groups > 1:
        outs = _apply_grouped(vel, pres, dvals, forcing, visc, c, groups)
        n_cat = len(outs[0]) if full_output else 2                           # the intermediates only when asked for
        out = [torch.cat([o[k] for o in outs]) if k != 4 else None for k in range(n_cat)] + [None] * (len(outs[0]) - n_cat)
        out[4] = torch.stack([o[4] for o in outs]).amax(0)                   # warn: any sample of any group
    else:
        out = _PisoStepFn.apply(vel, pres, dvals, forcing, visc, c)
    vel_next, pres_next, p1, p2, warn_new, values, a_diag, rhs, u_star, u_s2, h, div1, div2, lap1, lap2, its1, its2 = out
    if warn is not None:
        warn_new = torch.maximum(warn_new, as_tensor(warn).to(vel.device).reshape(-1)[:1].to(torch.float32))
    velocity_s3 = velocity.copied_with(flat=vel_next)
    pressure_new = pressure.copied_with(pres_next.reshape(b, g.ny, g.nx, 1))
    if not full_output:
        return velocity_s3, pressure_new, warn_new
    shape = velocity.staggered_shape
    rp, ci = g.csr_structure()
    inc1 = pressure_inc1.copied_with(p1.reshape(b, g.ny, g.nx, 1))
    inc2 = pressure_inc2.copied_with(p2.reshape(b, g.ny, g.nx, 1))
    return (velocity_s3, pressure_new, inc1, inc2, values, ci, rp,
            stagger_flattened_data(u_star, shape, True), stagger_flattened_data(u_s2, shape, True), a_diag, rhs,
            stagger_flattened_data(u_star, shape, True), velocity_s3.staggered_tensor(),
            div1.reshape(b, g.ny, g.nx, 1), lap1, lap2, warn_new)


def advection_matrix_cuda(velocity, simulation_physics, viscosity, beta, unrolling_step=0):
    """diffpiso/piso_tf.py:85-137 -> (matrix_values [B,nnz], row_pointers, column_indices, A staggered, matrix_nnz,
    A_flat).  Takes the SimulationParameters instead of the reference's loose mask arguments."""
    ny, nx = velocity.resolution
    per_y, per_x = _periodic_flags(simulation_physics)
    g = ops.Geometry.get(ny, nx, per_y, per_x, velocity.flat.device)
    m = ops.to_device_masks(simulation_physics, g)
    visc = as_tensor(viscosity).to(velocity.flat.device)
    values, a_flat = ops.assemble(g, velocity.flat, m["dirichlet"], m["active"], m["noslip"], visc,
                                  float(np.float32(velocity.dx[0])), float(np.float32(velocity.dx[1])),
                                  float(np.float32(beta)), ops.cell_areas(velocity.dx[0], velocity.dx[1]))
    rp, ci = g.csr_structure()
    return values, rp, ci, stagger_flattened_data(a_flat, velocity.staggered_shape, True), \
        np.array([g.nnz_u, g.nnz_v]), a_flat
