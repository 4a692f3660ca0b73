"""Tensor holders and layout helpers mirroring the slice of PhiFlow the PISO path touches.

Only layouts and padding semantics are kept: a staggered tensor is [B, ny+1, nx+1, 2] with
channel 0 = v (y-velocity, (ny+1) x nx valid), channel 1 = u (ny x (nx+1) valid), zero padded
(PhiFlow/phi/physics/field/staggered_grid.py:33-46); centred data is [B, ny, nx, 1].  Flat face vectors handed to the
native ops are [B, n_u + n_v] = [u rows..., v rows...] (diffpiso/piso_helpers.py:175-185 with coord_flip=True).
"""
import numpy as np
import torch


def as_tensor(x, dtype=torch.float32, device=None):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if device is not None and t.device != torch.device(device):
        t = t.to(device)
    return t


def unstack_staggered_tensor(t):
    """-> [v [B,ny+1,nx,1], u [B,ny,nx+1,1]]  (staggered_grid.py:33-39)"""
    return [t[:, :, :-1, 0:1], t[:, :-1, :, 1:2]]


def stack_staggered_components(tensors):
    """[v [B,ny+1,nx,1], u [B,ny,nx+1,1]] -> [B,ny+1,nx+1,2] zero padded (staggered_grid.py:42-46)"""
    v, u = tensors
    v = torch.nn.functional.pad(v, (0, 0, 0, 1))
    u = torch.nn.functional.pad(u, (0, 0, 0, 0, 0, 1))
    return torch.cat([v, u], dim=-1)


def flatten_staggered_data(data, coord_flip=False):
    """diffpiso/piso_helpers.py:175-185, batched: [B, n] instead of the reference's batch-of-one 1-D vector."""
    t = data.staggered_tensor() if isinstance(data, StaggeredGrid) else data
    v, u = unstack_staggered_tensor(t)
    parts = [v.reshape(v.shape[0], -1), u.reshape(u.shape[0], -1)]
    if coord_flip:
        parts = parts[::-1]
    return torch.cat(parts, dim=1)


def stagger_flattened_data(flat, staggered_shape, coord_flip=False):
    """Inverse of flatten_staggered_data (piso_helpers.py:188-206)."""
    ny, nx = int(staggered_shape[1]) - 1, int(staggered_shape[2]) - 1
    if flat.dim() == 1:
        flat = flat[None]
    n_u, n_v = ny * (nx + 1), (ny + 1) * nx
    if coord_flip:
        u, v = flat[:, :n_u], flat[:, n_u:n_u + n_v]
    else:
        v, u = flat[:, :n_v], flat[:, n_v:n_v + n_u]
    b = flat.shape[0]
    return stack_staggered_components([v.reshape(b, ny + 1, nx, 1), u.reshape(b, ny, nx + 1, 1)])


def _side_modes(extrapolation, rank=2):
    """PhiFlow extrapolation struct -> [(lo, hi)] per spatial dim (y, x)."""
    if isinstance(extrapolation, str):
        return [(extrapolation, extrapolation)] * rank
    out = []
    for e in extrapolation:
        out.append((e, e) if isinstance(e, str) else (e[0], e[1]))
    if len(out) != rank:
        raise ValueError("extrapolation must be a string or one entry per spatial dimension")
    return out


def extrapolation_codes(extrapolation):
    """-> [y_lo, y_hi, x_lo, x_hi] ghost-cell rules for the native ops ('boundary' replicate, 'constant' zero,
    'periodic' wrap; PhiFlow/phi/physics/material.py:70-108, backend pad modes)."""
    codes = []
    for lo, hi in _side_modes(extrapolation):
        for m in (lo, hi):
            if m not in ("boundary", "constant", "periodic"):
                raise ValueError("unknown extrapolation %r" % (m,))
            codes.append({"boundary": 0, "constant": 1, "periodic": 2}[m])
    return codes


_TORCH_PAD = {"boundary": "replicate", "constant": "constant", "periodic": "circular"}


def pad_with_extrapolation(data, widths, extrapolation):
    """Pad [B, ny, nx, C] along y and x by widths [[lo, hi], [lo, hi]], each side with its own PhiFlow extrapolation
    ('boundary' -> replicate, 'constant' -> 0, 'periodic' -> wrap; grid.py:257-281, backend pad modes)."""
    modes = _side_modes(extrapolation)
    t = data
    for axis in (0, 1):
        dim = 1 + axis
        n = t.shape[dim]
        for side, width in enumerate(widths[axis]):
            if width == 0:
                continue
            mode = modes[axis][side]
            if mode not in _TORCH_PAD:
                raise ValueError("unknown extrapolation %r" % (mode,))
            if mode == "constant":
                shape = list(t.shape)
                shape[dim] = width
                piece = t.new_zeros(shape)
            elif mode == "boundary":
                edge = t.narrow(dim, 0 if side == 0 else t.shape[dim] - 1, 1)
                piece = edge.expand(*[width if d == dim else -1 for d in range(t.dim())])
            else:
                # wrap relative to the unpadded extent (lower side was possibly padded already)
                off = widths[axis][0] if side == 1 else 0
                start = (n - width) if side == 0 else off
                piece = t.narrow(dim, start, width)
            t = torch.cat([piece, t] if side == 0 else [t, piece], dim=dim)
    return t


class _Grid(object):
    """As in PhiFlow the box size is an fp32 number (AABox stores fp32) and dx = size / resolution is formed from it in
    fp64 (fp32 array / int64 array); every constant of the step derives from this dx (piso_tf.py:26,53,96-97), pinned
    against the reference's Python by tests/golden/ref_python/step_*.npz.  A dx given explicitly is taken as is."""

    def _init_box(self, resolution, box, dx):
        res = np.asarray(resolution, dtype=np.float64)
        if dx is not None:
            self.dx = np.asarray(dx, dtype=np.float64) * np.ones(2)
            self.box = res * self.dx if box is None else box
        else:
            if box is None:
                size = res
            elif isinstance(box, (list, tuple, np.ndarray, float, int)):
                size = np.asarray(box, dtype=np.float64) * np.ones(2)
            else:
                size = np.asarray(box.size, dtype=np.float64) * np.ones(2)    # AABox-like
            self.box = size
            self.dx = size.astype(np.float32).astype(np.float64) / res


class CenteredGrid(_Grid):
    """data [B, ny, nx, 1]; dx = (dy, dx); extrapolation as in PhiFlow ('boundary' | 'constant' | 'periodic',
    optionally per dimension / side)."""

    def __init__(self, data, box=None, extrapolation="boundary", dx=None, name=None):
        self.data = as_tensor(data)
        if self.data.dim() != 4:
            raise ValueError("centred data must be [B, ny, nx, C]")
        self.extrapolation = extrapolation
        self.name = name
        self._init_box(self.data.shape[1:3], box, dx)

    @property
    def resolution(self):
        return tuple(self.data.shape[1:3])

    def copied_with(self, data):
        return CenteredGrid(data, box=self.box, dx=self.dx, extrapolation=self.extrapolation)

    def __add__(self, other):
        o = other.data if isinstance(other, CenteredGrid) else other
        return self.copied_with(self.data + o)

    def padded(self, widths):
        """PhiFlow/phi/physics/field/grid.py:188-194: pad the spatial axes with the grid's extrapolation."""
        if isinstance(widths, int):
            widths = [[widths, widths]] * 2
        return CenteredGrid(pad_with_extrapolation(self.data, widths, self.extrapolation), dx=self.dx,
                            extrapolation=self.extrapolation)

    def gradient(self, physical_units=True, difference="central"):
        """grid.py:218-223 / math/nd.py:186-216: finite-difference gradient of a scalar field, channels (d/dy, d/dx);
        the closure network's pressure input (combined_training_integrated.py:401-405)."""
        if self.data.shape[-1] != 1:
            raise ValueError("gradient needs a scalar field")
        if not np.allclose(self.dx, np.mean(self.dx)):
            raise NotImplementedError("Only cubic cells supported.")
        lo, hi, div = {"central": (1, 1, 2.0), "forward": (0, 1, 1.0), "backward": (1, 0, 1.0)}[difference.lower()]
        t = pad_with_extrapolation(self.data, [[lo, hi], [lo, hi]], self.extrapolation)
        ny, nx = self.data.shape[1:3]
        gy = t[:, lo + hi:lo + hi + ny, lo:lo + nx] - t[:, 0:ny, lo:lo + nx]
        gx = t[:, lo:lo + ny, lo + hi:lo + hi + nx] - t[:, lo:lo + ny, 0:nx]
        scale = float(np.mean(self.dx)) * div if physical_units else div
        return CenteredGrid(torch.cat([gy, gx], dim=-1) / scale, dx=self.dx, extrapolation=self.extrapolation)

    def at_faces(self, component):
        """`CenteredGrid.at(velocity.data[component])` (grid.py:108-130 -> linear resampling with this grid's
        extrapolation): values on the v faces [B, ny+1, nx, C] (component 0) or the u faces [B, ny, nx+1, C]
        (component 1) as the mean of the two adjacent cells; how the closure forcing reaches the faces
        (combined_training_integrated.py:407-410)."""
        w = [[1, 1], [0, 0]] if component == 0 else [[0, 0], [1, 1]]
        t = pad_with_extrapolation(self.data, w, self.extrapolation)
        if component == 0:
            return 0.5 * (t[:, 1:] + t[:, :-1])
        return 0.5 * (t[:, :, 1:] + t[:, :, :-1])


class StaggeredGrid(_Grid):
    """Staggered velocity holder.  Built either from a staggered tensor [B, ny+1, nx+1, 2] or (internally) from the
    flat [B, n_u+n_v] vector; the other representation is produced on demand."""

    def __init__(self, data=None, box=None, extrapolation="boundary", dx=None, name=None, flat=None, resolution=None,
                 pad_periodic=None):
        # pad_periodic: None = custom_padded pads the velocity as the simulation's periodic flags say; (y, x) booleans
        # override that (False = replicate on a periodic axis: the state of the reference's unrolled steps, Q21)
        self.pad_periodic = pad_periodic
        if data is None and flat is None:
            raise ValueError("need a staggered tensor or a flat vector")
        self._staggered = None if data is None else as_tensor(data)
        self._flat = flat
        if self._staggered is not None:
            if self._staggered.dim() != 4 or self._staggered.shape[-1] != 2:
                raise ValueError("staggered tensor must be [B, ny+1, nx+1, 2]")
            resolution = (self._staggered.shape[1] - 1, self._staggered.shape[2] - 1)
        if resolution is None:
            raise ValueError("resolution required when building from a flat vector")
        self._resolution = (int(resolution[0]), int(resolution[1]))
        self.extrapolation = extrapolation
        self.name = name
        self._init_box(self._resolution, box, dx)

    @property
    def resolution(self):
        return self._resolution

    @property
    def staggered_shape(self):
        b = self._staggered.shape[0] if self._staggered is not None else self._flat.shape[0]
        return (b, self._resolution[0] + 1, self._resolution[1] + 1, 2)

    def staggered_tensor(self):
        if self._staggered is None:
            self._staggered = stagger_flattened_data(self._flat, self.staggered_shape, coord_flip=True)
        return self._staggered

    @property
    def flat(self):
        if self._flat is None:
            self._flat = flatten_staggered_data(self._staggered, coord_flip=True).contiguous()
        return self._flat

    def copied_with(self, data=None, flat=None):
        return StaggeredGrid(data, box=self.box, dx=self.dx, extrapolation=self.extrapolation, flat=flat,
                             resolution=self._resolution, pad_periodic=self.pad_periodic)

    @property
    def data(self):
        """[v, u] component grids as in PhiFlow (`velocity.data[1].data` = u faces [B, ny, nx+1, 1])."""
        v, u = unstack_staggered_tensor(self.staggered_tensor())
        return [CenteredGrid(v, dx=self.dx, extrapolation=self.extrapolation),
                CenteredGrid(u, dx=self.dx, extrapolation=self.extrapolation)]

    def at_centers(self):
        """staggered_grid.py:150-151: both components linearly resampled at the cell centres -> CenteredGrid with
        data [B, ny, nx, 2] (channel 0 = v, 1 = u)."""
        v, u = unstack_staggered_tensor(self.staggered_tensor())
        c = torch.cat([0.5 * (v[:, 1:] + v[:, :-1]), 0.5 * (u[:, :, 1:] + u[:, :, :-1])], dim=-1)
        return CenteredGrid(c, dx=self.dx, extrapolation=self.extrapolation)
