"""Turbulence statistics used by the reference's evaluation and losses (diffpiso/evaluation_tools.py:92-113 numpy
energy spectrum, :163-186 differentiable energy spectrum; vorticity as in :52-55).  Plain torch on whatever device the
fields live on (these are diagnostics around the hot path); no native code involved."""
import numpy as np
import torch

from .grids import StaggeredGrid, pad_with_extrapolation, unstack_staggered_tensor


def _shell_index(n0, n1, device):
    """round(|k|) of every fft-shifted mode, integer wavenumbers relative to the centre (evaluation_tools.py:108-111)."""
    i = torch.arange(n0, device=device, dtype=torch.float64) - n0 / 2
    j = torch.arange(n1, device=device, dtype=torch.float64) - n1 / 2
    return torch.round(torch.sqrt(i[:, None] ** 2 + j[None, :] ** 2)).to(torch.int64)


def _shell_energy(centered, fft_norm, floor_shift=False):
    """0.5 * sum over shells of |u_hat|^2 + |v_hat|^2 for [B, ny, nx, 2] (or [ny, nx, 2]) centred velocities.
    floor_shift: the reference's own `tf_fftshift` (evaluation_tools.py:157-161) moves n//2 entries to the back, which
    differs from numpy's fftshift for odd sizes."""
    c = centered if centered.dim() == 4 else centered[None]
    n0, n1 = c.shape[1:3]
    f = torch.fft.fft2(c.to(torch.complex64 if c.dtype in (torch.float32, torch.complex64) else torch.complex128),
                       dim=(1, 2))
    e = (f.real ** 2 + f.imag ** 2).sum(dim=-1) * fft_norm              # [B, n0, n1]
    if floor_shift:
        e = torch.roll(e, shifts=(-(n0 // 2), -(n1 // 2)), dims=(1, 2))
    else:
        e = torch.fft.fftshift(e, dim=(1, 2))
    shells = _shell_index(n0, n1, c.device).reshape(-1)
    count = int(np.ceil((n0 ** 2 + n1 ** 2) ** 0.5 * 0.5)) + 1
    out = torch.zeros(c.shape[0], count, dtype=e.dtype, device=c.device)
    out.index_add_(1, shells, e.reshape(c.shape[0], -1))
    return 0.5 * out


def EK_spectrum_2D(velocity_centered, domain_size=None):
    """evaluation_tools.py:92-113 -> (wavenumbers [N//2], E(k) [N//2]) as numpy arrays for ONE centred velocity field
    [ny, nx, 2] (channel 1 = u, 0 = v; as called from :139-145; the cut-off is shape[1]//2 = nx//2)."""
    c = torch.as_tensor(velocity_centered)
    if c.dim() == 4:
        if c.shape[0] != 1:
            raise ValueError("one field at a time")
        c = c[0]
    n = c.shape[1]
    c = c[None]
    e = _shell_energy(c.double(), 1.0 / float(c.shape[1] * c.shape[2]) ** 2)[0] + 1e-20
    return np.arange(e.numel(), dtype=np.float64)[:n // 2], e[:n // 2].cpu().numpy()


def EK_spectrum_2D_torch(velocity_centered):
    """Differentiable spectrum of evaluation_tools.py:163-186 (`EK_spectrum_2D_tf`) for one field [ny, nx, 2]: shell
    sums of 0.5(|u_hat|^2+|v_hat|^2) / (ny nx)^2, first min(ny, nx)//2 shells."""
    c = velocity_centered
    if c.dim() == 4:
        if c.shape[0] != 1:
            raise ValueError("one field at a time (the reference slices [0, ...])")
        c = c[0]
    ny, nx = c.shape[0], c.shape[1]
    cutoff = min(ny, nx) // 2
    e = _shell_energy(c.real if c.is_complex() else c, 1.0 / float(ny * nx) ** 2, floor_shift=True)[0]
    return e[:cutoff]


EK_spectrum_2D_tf = EK_spectrum_2D_torch     # reference name


def vorticity(velocity):
    """The reference's vorticity expression (evaluation_tools.py:52-55): the staggered tensor padded by one with the grid's extrapolation, differences of v along x and u along y divided by
    dx[0] -> [B, ny+1, nx+1]."""
    if not isinstance(velocity, StaggeredGrid):
        raise TypeError("vorticity expects a StaggeredGrid")
    t = pad_with_extrapolation(velocity.staggered_tensor(), [[1, 1], [1, 1]], velocity.extrapolation)
    d = float(velocity.dx[0])
    return (t[:, 1:-1, 1:-1, 0] - t[:, 1:-1, :-2, 0]) / d - (t[:, 1:-1, 1:-1, 1] - t[:, :-2, 1:-1, 1]) / d


def kinetic_energy(velocity):
    """Mean of 0.5 |u|^2 over the cell centres, per sample."""
    c = velocity.at_centers().data.double()
    return 0.5 * (c ** 2).sum(dim=-1).mean(dim=(1, 2))


def enstrophy(velocity):
    """Mean of 0.5 w^2 over the periodic/interior vorticity points, per sample."""
    w = vorticity(velocity)[:, :-1, :-1].double()
    return 0.5 * (w ** 2).mean(dim=(1, 2))
