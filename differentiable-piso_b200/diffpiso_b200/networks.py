"""The closure network of the reference (diffpiso/networks.py:3-73): seven convolutions 4 -> 16 -> 16 -> 32 -> 64 -> 64
-> 64 -> 2 with kernels 7, 5, 5, 3, 3, 1, 1 and leaky-ReLU (slope 0.2, TF default) between them, applied to the centred
input [B, ny, nx, 4] = (v, u, dp/dy, dp/dx) cropped by `buffer_width`, zero-padded back to the input shape.

Weights are kept in TF's HWIO layout [kh, kw, c_in, c_out] so checkpoints of the reference map one to one.  This is the
caller of the PISO path, plain torch (cuDNN); only `piso_step` runs on the native kernels."""
import math

import numpy as np
import torch

from .grids import StaggeredGrid

N_FEAT = [16, 16, 32, 64, 64, 64]
KERNELS = [7, 5, 5, 3, 3, 1, 1]


def _conv(x_nchw, w_hwio, padding):
    w = w_hwio.permute(3, 2, 0, 1)
    pad = (w_hwio.shape[0] // 2, w_hwio.shape[1] // 2) if padding == "SAME" else 0
    return torch.nn.functional.conv2d(x_nchw, w, padding=pad)


def fullyconv_network(staggered_fields, w, buffer_width, padding="SAME", restore_shape=False):
    """networks.py:3-52 for padding 'SAME' or 'VALID' (the per-axis padding list variant of the reference is not
    executable as written and is rejected)."""
    if isinstance(padding, (list, tuple)):
        raise NotImplementedError("per-axis padding lists")
    x = staggered_fields
    if isinstance(x, StaggeredGrid):
        x = x.at_centers().data
    target_shape = None
    if buffer_width is not None:
        shape = x.shape
        x = x[:, buffer_width[0][0]:shape[1] - buffer_width[0][1], buffer_width[1][0]:shape[2] - buffer_width[1][1], :]
        target_shape = x.shape
    f = x.permute(0, 3, 1, 2)
    for i in range(6):
        f = torch.nn.functional.leaky_relu(_conv(f, w[i], padding), 0.2)
    f = _conv(f, w[6], padding).permute(0, 2, 3, 1)
    if padding == "VALID" and buffer_width is not None and restore_shape is True:
        n = int(sum(k.shape[0] - 1 for k in w) // 2)
        f = torch.nn.functional.pad(f, (0, 0, n, target_shape[2] - f.shape[2] - n, n, target_shape[1] - f.shape[1] - n))
    if buffer_width is not None:
        f = torch.nn.functional.pad(f, (0, 0, buffer_width[1][0], buffer_width[1][1], buffer_width[0][0], buffer_width[0][1]))
    return f


def initialise_fullyconv_network(buffer_width, padding="SAME", restore_shape=False, initialiser=None, device=None,
                                 generator=None):
    """networks.py:55-73 -> (network callable, weights, reduced_buffer_width).  Default initialiser: Glorot normal
    (stddev = sqrt(2 / (fan_in + fan_out)), tf.glorot_normal_initializer draws a truncated normal; a plain normal is
    used here).  `initialiser(shape) -> tensor` overrides it."""
    chans = [4] + N_FEAT + [2]
    weights = []
    for i, k in enumerate(KERNELS):
        shape = (k, k, chans[i], chans[i + 1])
        if initialiser is not None:
            t = torch.as_tensor(initialiser(shape), dtype=torch.float32)
        else:
            std = math.sqrt(2.0 / (k * k * (chans[i] + chans[i + 1])))
            t = torch.randn(shape, generator=generator) * std
        weights.append(t.to(device).requires_grad_(True))
    reduced = int(np.sum([k // 2 for k in [7, 5, 5, 3, 3]]))
    if buffer_width is not None:
        reduced = [[i + reduced for i in j] for j in buffer_width]
    return (lambda vel: fullyconv_network(vel, weights, buffer_width, padding, restore_shape)), weights, reduced
