"""Training objectives of the reference (diffpiso/losses.py:6-148) on torch tensors: same signatures, same return
convention `(accumulated loss, this loss's contribution)`.  `fields` / `velocity_fields` = [list of StaggeredGrid per
unrolled step]; `ground_truths` = [tensor [B, steps, ny+1, nx+1, 2]].  Everything is differentiable torch; the
gradients reach the PISO steps through `piso_step`'s autograd.Function."""
import numpy as np
import torch

from .grids import StaggeredGrid, pad_with_extrapolation
from .statistics import EK_spectrum_2D_torch


def _ranges(step_range, loss_factor):
    if not isinstance(step_range, list):
        step_range = [0, step_range]
    if not isinstance(loss_factor, list):
        loss_factor = [loss_factor for _ in range(step_range[1])]
    return step_range, loss_factor


def _total(parts):
    return torch.stack([torch.as_tensor(p) for p in parts]).sum() if len(parts) else torch.zeros(())


def _crop(t, buffer_width, sponge_start):
    n1 = int(t.shape[1])
    return t[:, buffer_width[0][0]:n1 - buffer_width[0][1], buffer_width[1][0]:int(sponge_start) - buffer_width[1][1], :]


def L2_field_loss(loss, fields, ground_truths, step_range, buffer_width, loss_factor, sponge_start, box=None,
                  sum_steps=True, loss_influence_range=None, **kwargs):
    """losses.py:6-35: sum over steps of factor * 0.5 * ||staggered - target||^2 (tf.nn.l2_loss) on the cropped window."""
    step_range, loss_factor = _ranges(step_range, loss_factor)
    contrib = [[] for _ in range(step_range[1] - step_range[0])]
    for i in range(len(fields)):
        for s in range(step_range[0], step_range[1]):
            stag = fields[i][s].staggered_tensor()
            target = ground_truths[i][:, s, ...]
            if buffer_width is not None:
                if sponge_start == 0:
                    sponge_start = stag.shape[2]
                diff = _crop(stag, buffer_width, sponge_start) - _crop(target, buffer_width, sponge_start)
            else:
                diff = stag - target
            contrib[s - step_range[0]].append(loss_factor[s] * 0.5 * (diff ** 2).sum())
    if sum_steps:
        c = _total([x for row in contrib for x in row])
        return loss + c, c
    per = [_total(row) for row in contrib]
    r = loss_influence_range
    groups = [_total(per[i * r:min((i + 1) * r, len(per))]) for i in range((len(per) - 1) // r + 1)]
    return [loss[i] + groups[i // r] for i in range(step_range[1] - step_range[0])], groups


def spectral_energy_loss(loss, velocity_fields, ground_truths, step_range, buffer_width=[[0, 0], [0, 0]], loss_factor=1,
                         sponge_start=0, log_distance=True, start_wavenumber=0, sum_steps=True,
                         loss_influence_range=None, **kwargs):
    """losses.py:38-67: distance between the shell-summed energy spectra of sample 0 and its target."""
    step_range, loss_factor = _ranges(step_range, loss_factor)
    contrib = []
    for s in range(step_range[0], step_range[1]):
        central = velocity_fields[0][s].at_centers().data
        if sponge_start == 0:
            sponge_start = central.shape[2]
        e = EK_spectrum_2D_torch(_crop(central, buffer_width, sponge_start)[0])
        gt = StaggeredGrid(ground_truths[0][:, s, ...]).at_centers().data
        g = EK_spectrum_2D_torch(_crop(gt, buffer_width, sponge_start)[0])
        if log_distance:
            d = torch.log(g[:e.shape[0]] / e) ** 2
            contrib.append(torch.sqrt(d[1 + start_wavenumber:].sum()) * loss_factor[s])
        else:
            contrib.append(torch.abs(g[:e.shape[0]] - e)[1:].sum() * loss_factor[s])
    return _combine(loss, contrib, sum_steps, loss_influence_range, step_range)


def _combine(loss, contrib, sum_steps, loss_influence_range, step_range):
    if sum_steps:
        c = _total(contrib)
        return loss + c, c
    r = loss_influence_range
    return [loss[i] + _total(contrib[i:min(i + r, len(contrib))]) for i in range(step_range[1] - step_range[0])], contrib


def _forward_gradient(t, dx):
    """phi math.gradient(t, dx, 'forward') with the default replicate padding (math/nd.py:186-216): [B, n0, n1, 1] ->
    [B, n0, n1, 2], channel k divided by dx[k]."""
    p = pad_with_extrapolation(t, [[0, 1], [0, 1]], "boundary")
    n0, n1 = t.shape[1:3]
    g = torch.cat([p[:, 1:n0 + 1, :n1] - p[:, :n0, :n1], p[:, :n0, 1:n1 + 1] - p[:, :n0, :n1]], dim=-1)
    return g / torch.as_tensor(np.asarray(dx, np.float32), device=t.device)


def _strain(grid, dx):
    g = [_forward_gradient(c.data, dx) for c in grid.data]
    shear = (g[0][:, 1:-1, 0:-1, 1] + g[1][:, 0:-1, 1:-1, 0]) / 2
    return [g[0][:, :-1, :, 0], shear, shear, g[1][:, :, :-1, 1]]


def strain_rate_loss(loss, velocity_fields, ground_truths, step_range, buffer_width, loss_factor=1, sponge_start=0,
                     box=None, sum_steps=True, loss_influence_range=None, **kwargs):
    """losses.py:69-96: L1 distance of the four strain-rate components built from forward differences of the faces."""
    step_range, loss_factor = _ranges(step_range, loss_factor)
    contrib = []
    for s in range(step_range[0], step_range[1]):
        vel = velocity_fields[0][s]
        gt = StaggeredGrid(ground_truths[0][:, s, ...], box=vel.box)
        a, b = _strain(vel, vel.dx), _strain(gt, vel.dx)
        contrib.append(sum(torch.abs(a[i] - b[i]).sum() for i in range(4)) * loss_factor[s])
    return _combine(loss, contrib, sum_steps, loss_influence_range, step_range)


def multistep_averaging_loss(loss, velocity_fields, ground_truths, step_range, buffer_width, loss_factor=1,
                             sponge_start=0, box=None, sum_steps=True, loss_influence_range=None, **kwargs):
    """losses.py:98-148: L1 distance between running means (window = loss_influence_range) of the face velocities and
    of the targets; windows are clamped at both ends of the unroll."""
    if not isinstance(step_range, list):
        step_range = [0, step_range]
    n = step_range[1] - step_range[0]

    def crop(t):
        return t[:, buffer_width[0][0]:t.shape[1] - buffer_width[0][1], buffer_width[1][0]:t.shape[2] - buffer_width[1][1], 0]
    du, dv, gu, gv = [], [], [], []
    for s in range(step_range[0], step_range[1]):
        v, u = velocity_fields[0][s].data
        tv, tu = StaggeredGrid(ground_truths[0][:, s, ...]).data
        du.append(crop(u.data)); dv.append(crop(v.data)); gu.append(crop(tu.data)); gv.append(crop(tv.data))
    r = n if loss_influence_range is None else loss_influence_range
    du, dv, gu, gv = [torch.cat(x, dim=0) for x in (du, dv, gu, gv)]
    windows = range(n - r + 1)
    dist = [torch.abs(du[i:i + r].mean(0) - gu[i:i + r].mean(0)).sum() + torch.abs(dv[i:i + r].mean(0) - gv[i:i + r].mean(0)).sum()
            for i in windows]
    contrib = []
    for i in range(n):
        if i < r // 2:
            contrib.append(dist[0] * loss_factor)
        elif i >= r // 2 + n - r:
            contrib.append(dist[-1] * loss_factor)
        else:
            contrib.append(dist[i - r // 2] * loss_factor)
    if sum_steps:
        c = _total(contrib)
        return loss + c, c
    return [loss[i] + contrib[i] for i in range(n)], contrib
