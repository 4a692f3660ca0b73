"""Stand-alone, differentiable versions of the pointwise operators of diffpiso/piso_helpers.py, with the reference's
names and argument meaning.  `piso_step` fuses them into its own forward/backward; these wrappers expose the same
native kernels one operator at a time (each a torch.autograd.Function carrying the reference's gradient registration).

  custom_padded                   piso_helpers.py:35-55
  arrange_rhs_term                piso_helpers.py:169-172   (arrange_rhs_term_tf)
  finite_volume_gradient_tensor   piso_helpers.py:236-274   backward: autodiff of pad/slice/multiply, periodic axes use
                                                            circular_padded_gradient's registered gradient (:226-233)
  finite_volume_divergence        piso_helpers.py:277-310   backward: registered gradient (:291-305)
  explicit_H_csr                  piso_helpers.py:209-223   backward: transposed product (gather/segment_sum autodiff)
"""
import numpy as np
import torch

from . import ops
from .grids import (CenteredGrid, StaggeredGrid, as_tensor, extrapolation_codes, flatten_staggered_data,
                    stack_staggered_components, stagger_flattened_data, unstack_staggered_tensor)
from .pressure_solver import _periodic_flags


def _geom_for(sim, ny, nx, device):
    per_y, per_x = _periodic_flags(sim) if sim is not None else (False, False)
    return ops.Geometry.get(ny, nx, per_y, per_x, device)


def _spacing(field):
    return float(np.float32(field.dx[0])), float(np.float32(field.dx[1]))


def custom_padded(staggered_field, widths=1, bool_periodic=(False, False)):
    """Pad each velocity component by one cell: periodic axes wrap (dropping the duplicated last face along the
    component's own axis and padding (1, 2) there), all other modes replicate the edge (width 1).
    Returns [v_padded [B, ny+3, nx+2], u_padded [B, ny+2, nx+3]]."""
    if widths != 1:
        raise NotImplementedError("the PISO path pads by exactly one cell")
    v, u = unstack_staggered_tensor(staggered_field.staggered_tensor())
    v, u = v[..., 0], u[..., 0]
    per_y, per_x = bool(bool_periodic[0]), bool(bool_periodic[1])

    def pad_axis(t, dim, periodic, own):
        n = t.shape[dim]
        if periodic:
            if own:
                t = t.narrow(dim, 0, n - 1)
                n -= 1
                idx = [(i - 1) % n for i in range(n + 3)]
            else:
                idx = [(i - 1) % n for i in range(n + 2)]
        else:
            idx = [min(max(i - 1, 0), n - 1) for i in range(n + 2)]
        return t.index_select(dim, torch.as_tensor(idx, device=t.device))
    v = pad_axis(pad_axis(v, 1, per_y, True), 2, per_x, False)
    u = pad_axis(pad_axis(u, 1, per_y, False), 2, per_x, True)
    return [v, u]


def arrange_rhs_term(rhs, dirichlet_mask, dirichlet_values, beta=None, coord_flip=False):
    """(1 - mask) * rhs + mask * values * -1, flattened (piso_helpers.py:169-172)."""
    rhs, m, dv = as_tensor(rhs), as_tensor(dirichlet_mask).to(rhs.device), as_tensor(dirichlet_values).to(rhs.device)
    return flatten_staggered_data((1 - m) * rhs + m * dv * -1, coord_flip=coord_flip)


class _FvGradientFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, geom, access, dy, dx, pbc):
        ctx.args = (geom, access, dy, dx, pbc)
        return ops.fv_gradient(geom, p.contiguous(), access, dy, dx, pbc)

    @staticmethod
    def backward(ctx, gs):
        geom, access, dy, dx, pbc = ctx.args
        return ops.fv_gradient_adj(geom, gs.contiguous(), access, dy, dx, pbc), None, None, None, None, None


def finite_volume_gradient_tensor(centered_field, sim_physics=None):
    """Pressure-gradient influence on the staggered grid: (p+ - p-) * dx*dy / d on faces with ghost cells from the field's
    extrapolation, times min(accessible+, accessible-).  Returns the staggered tensor [B, ny+1, nx+1, 2]."""
    assert isinstance(centered_field, CenteredGrid)
    data = centered_field.data
    b, ny, nx = data.shape[0], data.shape[1], data.shape[2]
    geom = _geom_for(sim_physics, ny, nx, data.device)
    if sim_physics is not None:
        access = ops.to_device_masks(sim_physics, geom)["access"]
    else:
        access = torch.ones((ny + 2) * (nx + 2), device=data.device)
    dy, dx = _spacing(centered_field)
    flat = _FvGradientFn.apply(data.reshape(b, ny * nx), geom, access, dy, dx,
                               extrapolation_codes(centered_field.extrapolation))
    return stagger_flattened_data(flat, (b, ny + 1, nx + 1, 2), coord_flip=True)


class _FvDivergenceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vel_flat, geom, dy, dx):
        ctx.args = (geom, dy, dx)
        return ops.fv_divergence(geom, vel_flat.contiguous(), dy, dx)

    @staticmethod
    def backward(ctx, gc):
        geom, dy, dx = ctx.args
        return ops.fv_divergence_adj(geom, gc.contiguous(), dy, dx), None, None, None


def finite_volume_divergence(staggered_field, bool_periodic=(False, False)):
    """sum_d (vel_d+ - vel_d-) * dx*dy / d on cell centres, [B, ny, nx, 1]; the backward is the reference's registered
    gradient, whose periodic branch follows `bool_periodic` = (periodic_y, periodic_x)."""
    assert isinstance(staggered_field, StaggeredGrid)
    ny, nx = staggered_field.resolution
    flat = staggered_field.flat
    geom = ops.Geometry.get(ny, nx, bool(bool_periodic[0]), bool(bool_periodic[1]), flat.device)
    dy, dx = _spacing(staggered_field)
    div = _FvDivergenceFn.apply(flat, geom, dy, dx)
    return div.reshape(flat.shape[0], ny, nx, 1)


class _HApplyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d_flat, values, a_flat, geom, beta):
        ctx.geom, ctx.beta = geom, beta
        ctx.save_for_backward(values, a_flat)
        zero = torch.zeros_like(d_flat)
        return ops.h_apply(geom, values, a_flat, zero, d_flat.contiguous(), beta)

    @staticmethod
    def backward(ctx, gh):
        values, a_flat = ctx.saved_tensors
        return ops.h_apply_adj(ctx.geom, values, a_flat, gh.contiguous(), ctx.beta), None, None, None, None


def explicit_H_csr(matrix_values, row_pointers, column_indices, velocity, staggered_shape, A, beta=0, bool_periodic=(False, False)):
    """H v = M v - (A - beta) v for the u and v advection-diffusion matrices (second PISO corrector).
    matrix_values [B, nnz]; velocity a StaggeredGrid; A the staggered tensor (or flat [B, n_u+n_v]) of the matrix
    diagonal.  row_pointers / column_indices are accepted for interface parity (the layout is implied by the grid).
    Returns the staggered tensor of H v; differentiable w.r.t. the velocity only, like the reference."""
    ny, nx = int(staggered_shape[1]) - 1, int(staggered_shape[2]) - 1
    flat = velocity.flat if isinstance(velocity, StaggeredGrid) else flatten_staggered_data(as_tensor(velocity), True).contiguous()
    geom = ops.Geometry.get(ny, nx, bool(bool_periodic[0]), bool(bool_periodic[1]), flat.device)
    a = as_tensor(A)
    a_flat = flatten_staggered_data(a, True).contiguous() if a.dim() == 4 else a.reshape(flat.shape).contiguous()
    values = as_tensor(matrix_values).reshape(flat.shape[0], -1).contiguous()
    h = _HApplyFn.apply(flat, values, a_flat, geom, float(np.float32(beta)))
    return stagger_flattened_data(h, (flat.shape[0], ny + 1, nx + 1, 2), coord_flip=True)
