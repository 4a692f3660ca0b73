"""Mask / boundary-value builders of the reference's setups, with the reference's names, argument
meaning and array shapes: `compute_mixingLayer_masks` (diffpiso/piso_helpers.py:73-133), `temporal_mixing_layer_masks`
(:136-166), `update_dirichlet_values` (:58-70).  Staggered arrays are [1, ny+1, nx+1, 2] (channel 0 = v, 1 = u), centred
masks [1, ny+2, nx+2, 1].  numpy in, numpy out -- except `update_dirichlet_values`, which also accepts torch tensors
(it sits inside the unrolled graph: the inflow profile changes every step)."""
import numpy as np
import torch


def _cat(parts, axis):
    if any(isinstance(p, torch.Tensor) for p in parts):
        ref = next(p for p in parts if isinstance(p, torch.Tensor))
        return torch.cat([p if isinstance(p, torch.Tensor) else torch.as_tensor(np.asarray(p), dtype=ref.dtype, device=ref.device)
                          for p in parts], dim=axis)
    return np.concatenate(parts, axis)


def _unstack(t):
    return [t[:, :, :-1, 0:1], t[:, :-1, :, 1:2]]


def _stack(v, u):
    if isinstance(v, torch.Tensor) or isinstance(u, torch.Tensor):
        v, u = torch.as_tensor(v), torch.as_tensor(u)
        return torch.cat([torch.nn.functional.pad(v, (0, 0, 0, 1)), torch.nn.functional.pad(u, (0, 0, 0, 0, 0, 1))], dim=-1)
    return np.concatenate([np.pad(v, ((0, 0), (0, 0), (0, 1), (0, 0))), np.pad(u, ((0, 0), (0, 1), (0, 0), (0, 0)))], -1)


def update_dirichlet_values(dirichlet_values, update_bool, dirichlet_array):
    """piso_helpers.py:58-70: replace the boundary rows/columns selected by update_bool = ((y_lo, y_hi), (x_lo, x_hi))
    with dirichlet_array[dim][side] ([1, 1, nx+2, 1] for y sides, [1, ny+2, 1, 1] for x sides; the two corner entries
    are dropped)."""
    v, u = _unstack(dirichlet_values)
    if update_bool[0][0]:
        v = _cat([dirichlet_array[0][0][..., 1:-1, :], v[:, 1:]], 1)
    if update_bool[0][1]:
        v = _cat([v[:, :-1], dirichlet_array[0][1][..., 1:-1, :]], 1)
    if update_bool[1][0]:
        u = _cat([dirichlet_array[1][0][:, 1:-1], u[:, :, 1:]], 2)
    if update_bool[1][1]:
        u = _cat([u[:, :, :-1], dirichlet_array[1][1][:, 1:-1]], 2)
    return _stack(v, u)


def compute_mixingLayer_masks(staggered_shape, dirichlet_bool, dirichlet_array, dtype=np.float32):
    """piso_helpers.py:73-133 -> (dirichlet_mask, dirichlet_values, neumann_mask, active_mask, accessible_mask)."""
    _, s1, s2, _ = [int(k) for k in staggered_shape]
    ny, nx = s1 - 1, s2 - 1
    interior = [np.zeros((1, s1 - 2, nx, 1)), np.zeros((1, ny, s2 - 2, 1))]
    edge = [(1, 1, nx, 1), (1, ny, 1, 1)]
    mask, neumann, values = [], [], []
    for dim in (0, 1):
        cm, cn, cv = [], [], []
        for side in (0, 1):
            if dirichlet_bool[dim][side]:
                arr = dirichlet_array[dim][side]
                cv.append(np.asarray(arr[..., 1:-1, :] if dim == 0 else arr[:, 1:-1]))
                cm.append(np.ones(edge[dim], dtype))
                cn.append(np.zeros(edge[dim], dtype))
            else:
                cv.append(np.zeros(edge[dim], dtype))
                cm.append(np.zeros(edge[dim], dtype))
                cn.append(np.ones(edge[dim], dtype) * (1 + side))
        mask.append(np.concatenate([cm[0], interior[dim], cm[1]], dim + 1))
        neumann.append(np.concatenate([cn[0], interior[dim], cn[1]], dim + 1))
        values.append(np.concatenate([cv[0], interior[dim], cv[1]], dim + 1))
    accessible = np.ones((s1 + 1, s2 + 1))
    accessible[:, 0] = 0
    accessible[0, :] = 0
    accessible[-1, :] = 0
    active = np.pad(np.ones((ny, nx)), ((1, 1), (1, 1)))
    return (_stack(*mask), _stack(*values), _stack(*neumann), active[None, :, :, None], accessible[None, :, :, None])


def temporal_mixing_layer_masks(staggered_shape, dirichlet_bool, dirichlet_array, dtype=np.float32):
    """piso_helpers.py:136-166 -> (dirichlet_mask, dirichlet_values, [boundary_bool_x, boundary_bool_y], active_mask,
    accessible_mask); walls in y (v Dirichlet from dirichlet_array[0]), periodic in x (no u Dirichlet faces)."""
    assert dirichlet_bool == ((True, True), (False, False))
    _, s1, s2, _ = [int(k) for k in staggered_shape]
    ny, nx = s1 - 1, s2 - 1
    ones = np.ones((1, 1, nx, 1))
    mask_v = np.concatenate([ones, np.zeros((1, s1 - 2, nx, 1)), ones], 1)
    values_v = np.concatenate([np.asarray(dirichlet_array[0][0][..., 1:-1, :]), np.zeros((1, s1 - 2, nx, 1)),
                               np.asarray(dirichlet_array[0][1][..., 1:-1, :])], 1)
    zeros_u = np.zeros((1, ny, s2, 1))
    bx = np.zeros([1, ny, s2, 4], dtype=bool)
    bx[:, 0, :, 2] = True
    bx[:, -1, :, 3] = True
    by = np.zeros([1, s1, nx, 4], dtype=bool)
    by[:, 0, :, 2] = True
    by[:, -1, :3] = True
    accessible = np.concatenate([np.zeros((1, s2 + 1)), np.ones((ny, s2 + 1)), np.zeros((1, s2 + 1))], axis=0)
    accessible = accessible[None, :, :, None]
    return _stack(mask_v, zeros_u), _stack(values_v, zeros_u.copy()), [bx, by], accessible, accessible
