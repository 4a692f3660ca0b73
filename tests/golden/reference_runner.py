"""Runs the reference's OWN Python for the PISO step -- forward and registered gradients -- in this container.

TEST INFRASTRUCTURE (build container only; needs /root/reference, never imported by tests or the product).

tensorflow 1.14 is not installable here, so `diffpiso` cannot be imported as a package.  But the vendored PhiFlow ships a
PyTorch backend (PhiFlow/phi/torch), and the step's Python (`piso_step`, `advection_matrix_cuda`, the solver classes'
`solve`, `finite_volume_*`, `explicit_H_csr`, `custom_padded`, ..., and every `grad` closure registered with
`tf.custom_gradient`) only touches TensorFlow through a dozen entry points.  This module

* imports PhiFlow from where it lies (a few aliases removed from numpy>=1.24 / Python>=3.10 are restored first),
* extracts the reference's function/class definitions with `ast` from the files under /root/reference/diffpiso and
  executes them UNMODIFIED in a namespace whose `tf` is a small shim over torch (`tf.custom_gradient` becomes a
  `torch.autograd.Function` calling the reference's own `grad` closure),
* substitutes the three custom-op libraries (`cd_csr_op`, `multi_bicg_op`, `pressure_op`), which are CUDA-only and need
  CUDA-10 cuSPARSE, by callables with the ops' exact argument lists that forward to the CPU oracle kernels (assembly and
  Laplace are pinned bit-exactly against the reference's CUDA kernels by tests/golden/ref_kernels).

What this pins: everything the reference does in Python around the ops -- padding, flattening orders, signs and scalings
of the predictor right-hand side and the two correctors, H application, pressure accumulation, the argument plumbing into
the ops, and the backward pass TF would assemble from the registered gradients (transposed predictor solve with the
forward initial guess, pressure solves on the incoming gradient, the periodic-axis conventions Q19/Q20 of the divergence
and circular-gradient registrations)."""
import ast
import collections
import collections.abc
import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _restore_aliases():
    for n, t in (("float", float), ("int", int), ("object", object), ("complex", complex)):
        if not hasattr(np, n):
            setattr(np, n, t)
    for n in ("Iterable", "Mapping", "Sequence", "Callable"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))


class _Shape(tuple):
    def as_list(self):
        return list(self)


class _TfTensor(torch.Tensor):
    """torch.Tensor that behaves like a tf.Tensor where the reference's Python relies on it: `.shape.as_list()` exists,
    and numpy operands of arithmetic are converted to the tensor's dtype (tf.convert_to_tensor with the dtype of the
    tensor operand) instead of dragging the tensor into numpy."""

    @property
    def shape(self):
        return _Shape(super().shape)

    def set_shape(self, shape):          # static-shape hint in TF, nothing to do
        assert list(self.shape) == list(shape)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        ref = next((a for a in args if isinstance(a, torch.Tensor)), None)

        def conv(a):
            if isinstance(a, np.ndarray):
                t = torch.from_numpy(np.ascontiguousarray(a))
                return t.to(ref.dtype) if ref is not None else t
            return a
        return super().__torch_function__(func, types, tuple(conv(a) for a in args), kwargs or {})


def _binary(name):
    base = getattr(torch.Tensor, name)

    def op(self, other):
        if isinstance(other, np.ndarray):
            other = torch.from_numpy(np.ascontiguousarray(other)).to(self.dtype)
        return base(self, other)
    return op


for _n in ("__mul__", "__rmul__", "__add__", "__radd__", "__sub__", "__rsub__", "__truediv__", "__rtruediv__"):
    setattr(_TfTensor, _n, _binary(_n))


def tf_tensor(x, requires_grad=False):
    t = torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x).as_subclass(_TfTensor)
    return t.requires_grad_(True) if requires_grad else t


def _plain(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, _TfTensor) else t


class _TfNn(object):
    @staticmethod
    def l2_loss(x):
        return (x ** 2).sum() / 2

    @staticmethod
    def conv2d(x, w, strides, padding):
        """NHWC input, HWIO filter, stride 1, 'SAME' (odd kernels: symmetric zero padding) or 'VALID'."""
        assert list(strides) == [1, 1, 1, 1] and padding in ("SAME", "VALID")
        pad = (w.shape[0] // 2, w.shape[1] // 2) if padding == "SAME" else 0
        y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=pad)
        return y.permute(0, 2, 3, 1)

    @staticmethod
    def leaky_relu(x, alpha=0.2):
        return torch.nn.functional.leaky_relu(x, alpha)


class _TfMath(object):
    @staticmethod
    def segment_sum(data, segment_ids):
        ids = segment_ids.long()
        out = torch.zeros(int(ids.max()) + 1, dtype=data.dtype)
        return out.index_add(0, ids, data)


def _stack_nested(x):
    if isinstance(x, (list, tuple)):
        parts = [_stack_nested(k) for k in x]
        parts = [p if isinstance(p, torch.Tensor) else torch.as_tensor(p) for p in parts]
        return torch.stack([p.to(torch.float32) if not p.is_floating_point() else p for p in parts])
    return x


class TfShim(object):
    float32, float64, int32, bool, complex64 = torch.float32, torch.float64, torch.int32, torch.bool, torch.complex64
    nn = _TfNn()
    math = _TfMath()

    def reduce_sum(self, x, axis=None):
        x = _stack_nested(x)
        return x.sum() if axis is None else x.sum(dim=tuple(int(a) for a in np.atleast_1d(axis)))

    def convert_to_tensor(self, x):
        return _stack_nested(x)

    def fft2d(self, x):
        return torch.fft.fft2(x)

    def conj(self, x):
        return torch.conj(x)

    def abs(self, x):
        return torch.abs(x)

    def log(self, x):
        return torch.log(x)

    def sqrt(self, x):
        return torch.sqrt(x)

    def round(self, x):
        return torch.round(x)

    def matmul(self, a, b):
        return torch.matmul(a, b)

    def expand_dims(self, x, axis):
        return torch.unsqueeze(x, axis)

    def reshape(self, x, shape):
        return torch.reshape(x, tuple(int(k) for k in shape))

    def argsort(self, x):
        return torch.argsort(x, stable=True)

    @staticmethod
    def _shape(shape):
        if isinstance(shape, (int, float, np.integer, np.floating)) or (isinstance(shape, (torch.Tensor, np.ndarray)) and shape.ndim == 0):
            return (int(shape),)
        return tuple(int(s) for s in shape)

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(self._shape(shape), dtype=dtype).as_subclass(_TfTensor)

    def ones(self, shape, dtype=torch.float32):
        return torch.ones(self._shape(shape), dtype=dtype).as_subclass(_TfTensor)

    def zeros_like(self, x, dtype=None):
        return torch.zeros_like(torch.as_tensor(x), dtype=dtype).as_subclass(_TfTensor)

    def constant(self, value, dtype=None, shape=None):
        t = torch.as_tensor(np.asarray(value))
        if dtype is None and t.dtype == torch.float64:
            dtype = torch.float32          # tf.constant(python float) is float32
        if dtype is not None:
            t = t.to(dtype)
        if shape is not None:
            t = t.reshape(self._shape(shape)) if t.numel() > 1 else t.reshape(-1)[:1].expand(self._shape(shape)).clone()
        return t

    def cast(self, x, dtype):
        return torch.as_tensor(x).to(dtype).as_subclass(_TfTensor)

    def identity(self, x):
        return x.clone()

    def is_tensor(self, x):
        return isinstance(x, torch.Tensor)

    def stop_gradient(self, x):
        return x.detach()

    def gather(self, params, indices):
        return params[torch.as_tensor(indices).long()]

    def range(self, n, dtype=torch.int32):
        return torch.arange(int(n), dtype=dtype).as_subclass(_TfTensor)

    def searchsorted(self, sorted_sequence, values, side="left"):
        return torch.searchsorted(sorted_sequence.contiguous(), values.contiguous(), right=(side == "right")).to(torch.int32)

    def segment_sum(self, data, segment_ids):
        ids = segment_ids.long()
        out = torch.zeros(int(ids.max()) + 1, dtype=data.dtype)
        return out.index_add(0, ids, data)

    def concat(self, values, axis):
        return torch.cat(list(values), dim=axis)

    def pad(self, x, paddings):
        flat = []
        for lo, hi in reversed([tuple(p) for p in paddings]):
            flat += [int(lo), int(hi)]
        return torch.nn.functional.pad(x, flat)

    def custom_gradient(self, f):
        """tf.custom_gradient: f(*args) -> (outputs, grad_fn); d(outputs)/d(args) comes ONLY from grad_fn."""
        def wrapper(*args):
            idx = [i for i, a in enumerate(args) if isinstance(a, torch.Tensor) and a.is_floating_point()]

            class Fn(torch.autograd.Function):
                @staticmethod
                def forward(ctx, *targs):
                    full = list(args)
                    for k, i in enumerate(idx):
                        full[i] = targs[k].detach()
                    out, grad = f(*full)
                    ctx.grad_fn_ref = grad
                    ctx.multi = isinstance(out, (list, tuple))
                    outs = tuple(out) if ctx.multi else (out,)
                    outs = tuple((o if isinstance(o, torch.Tensor) else torch.as_tensor(np.asarray(o))).as_subclass(_TfTensor) for o in outs)
                    ctx.meta = [(o.shape, o.dtype) for o in outs]
                    nd = [o for o in outs if not o.is_floating_point()]
                    if nd:
                        ctx.mark_non_differentiable(*nd)
                    return outs if ctx.multi else outs[0]

                @staticmethod
                def backward(ctx, *douts):
                    douts = [torch.zeros(s, dtype=d) if g is None else g for g, (s, d) in zip(douts, ctx.meta)]
                    douts = [g.as_subclass(_TfTensor) for g in douts]
                    with torch.no_grad():
                        g = ctx.grad_fn_ref(*douts)
                    if not isinstance(g, (list, tuple)):
                        g = [g]
                    g = list(g) + [None] * (len(args) - len(g))
                    res = []
                    for i in idx:
                        gi = g[i]
                        if gi is not None and not isinstance(gi, torch.Tensor):
                            gi = getattr(gi, "data", gi)          # a CenteredGrid
                        res.append(None if gi is None else _plain(gi).to(args[i].dtype))
                    return tuple(res)
            return Fn.apply(*[args[i] for i in idx])
        return wrapper


def _np(x, dtype=None):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    a = np.asarray(x)
    return a if dtype is None else np.ascontiguousarray(a, dtype=dtype)


class OpLog(object):
    """Arguments the reference's Python handed to the ops during the last step (so tests can pin them too)."""

    def __init__(self):
        self.calls = []

    def add(self, name, **kw):
        self.calls.append((name, kw))

    def last(self, name, nth=-1):
        return [kw for n, kw in self.calls if n == name][nth]


def make_op_shims(oracle, log):
    """Callables with the custom ops' argument lists (central_difference_csr_op.cc:9-32, multi_bicgstab_ilu_linear_solve
    _op.cc:9-48, pressure_solve_op.cc:8-46) forwarding to the CPU oracle."""
    O = oracle

    class CdCsr(object):
        @staticmethod
        def central_difference_matrix_csr(velocity_padded, csr_val, csr_col_ind, csr_row_ptr, diag_comp, dirichlet_mask,
                                          active_mask, accessible_mask, viscosity, dimensions, pad_depth, cell_area,
                                          grid_spacing, no_slip_mask, bool_periodic, beta, unrolling_step):
            dims = _np(dimensions)                       # [Nx+1, Ny, Nx, Ny+1]
            nx, ny = int(dims[2]), int(dims[1])
            per_x, per_y = bool(_np(bool_periodic)[0]), bool(_np(bool_periodic)[1])
            vp = _np(velocity_padded, np.float32)
            n_up = (ny + 2) * (nx + 3)
            assert vp.size == n_up + (ny + 3) * (nx + 2)
            rp, ci = O.csr_structure(ny, nx, per_x, per_y)
            assert csr_val.numel() == ci.size and csr_row_ptr.numel() == rp.size and diag_comp.numel() == dims[0] * dims[1] + dims[2] * dims[3]
            gs, ca = _np(grid_spacing, np.float32), _np(cell_area, np.float32)
            noslip = _np(no_slip_mask).astype(np.uint8).ravel()
            values, a_diag = O.assemble(ny, nx, per_x, per_y, float(gs[1]), float(gs[0]), float(np.float32(beta)), vp[:n_up],
                                        vp[n_up:], _np(dirichlet_mask).astype(np.uint8).ravel(), _np(active_mask, np.float32).ravel(),
                                        noslip, _np(viscosity, np.float32).ravel(), rp, areas=(float(ca[0]), float(ca[1])))
            log.add("assemble", velocity_padded=vp, cell_area=ca, grid_spacing=gs, dimensions=dims, beta=float(beta),
                    bool_periodic=np.array([per_x, per_y]))
            return (torch.from_numpy(values), torch.from_numpy(ci.astype(np.int32)), torch.from_numpy(rp.astype(np.int32)),
                    torch.from_numpy(a_diag))

    class MultiBicg(object):
        @staticmethod
        def multi_bicgstab_ilu_linear_solve(values, row_ptr, col_ind, rhs, x0, s, shat, p, phat, r, rhat, v, t, z, x_buffer,
                                            warn, matrix_sizes, accuracy, batch, max_iterations, transpose, unrolling_step):
            sizes = [int(k) for k in _np(matrix_sizes)]
            rp_all, ci_all, val = _np(row_ptr, np.int32), _np(col_ind, np.int32), _np(values, np.float32)
            b, x_init = _np(rhs, np.float32), _np(x0, np.float32)
            tol = float(_np(accuracy))
            out, stats, w = [], [], 0
            row0, nz0, r0 = 0, 0, 0
            for n in sizes:
                rp = rp_all[r0:r0 + n + 1]
                nz = int(rp[-1])
                x, st = O.bicgstab_ilu(rp, ci_all[nz0:nz0 + nz], val[nz0:nz0 + nz], b[row0:row0 + n], x_init[row0:row0 + n],
                                       tol, int(max_iterations), transpose=bool(transpose))
                out.append(x)
                stats.append(st)
                w |= int(st["warn"])
                row0, nz0, r0 = row0 + n, nz0 + nz, r0 + n + 1
            log.add("bicgstab", transpose=bool(transpose), stats=stats, tol=tol, x0=x_init.copy(), rhs=b.copy())
            warn_out = torch.tensor([bool(w) or bool(_np(warn).ravel()[0])])
            return [values, row_ptr, col_ind, torch.from_numpy(np.concatenate(out)), warn_out]

    class Pressure(object):
        @staticmethod
        def pressure_solve_op(dimensions, mask_dimensions, active_mask, accessible_mask, laplace_matrix, divergence, p, r, z,
                              guess, advection_influence, staggered_dimensions, rank_deficient, accuracy, max_iterations,
                              bool_periodic, init_with_zeros, residual_reset, randomized_restarts, unrolling_step):
            nx, ny = int(dimensions[0]), int(dimensions[1])
            assert init_with_zeros is True and randomized_restarts == 0
            fp64 = divergence.dtype == torch.float64
            dt = np.float64 if fp64 else np.float32
            lap = O.laplace(ny, nx, _np(active_mask, np.float32).ravel(), _np(accessible_mask, np.float32).ravel(),
                            _np(advection_influence, np.float32), dt)
            per_x, per_y = bool(bool_periodic[0]), bool(bool_periodic[1])
            div = _np(divergence, dt)
            xs, its = [], []
            for b in range(div.shape[0]):
                x, it = O.pressure_cg(ny, nx, per_x, per_y, lap, div[b].ravel(), float(accuracy), int(max_iterations),
                                      int(residual_reset), bool(_np(rank_deficient).ravel()[0]))
                xs.append(x)
                its.append(it)
            log.add("pressure", iterations=its, k_faces=_np(advection_influence, np.float32).copy(), div=div.copy(),
                    rank_deficient=bool(_np(rank_deficient).ravel()[0]), fp64=fp64)
            x = torch.from_numpy(np.stack(xs).astype(dt)).reshape(divergence.shape)
            return x, torch.tensor(its[-1:], dtype=torch.int32), torch.from_numpy(np.asarray(lap))

    return CdCsr, MultiBicg, Pressure


def _definitions(path, names=None):
    tree = ast.parse(open(path).read())
    out = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and (names is None or node.name in names):
            out.append(node)
    return out


def _complete_torch_backend():
    """PhiFlow's torch backend refuses mixed tensor/numpy lists where its TF backend (tf.concat) converts the numpy
    entries; give it the TF behaviour."""
    from phi.torch.torch_backend import TorchBackend

    def concat(self, values, axis):
        ref = next(v for v in values if isinstance(v, torch.Tensor))
        vals = [v if isinstance(v, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(v)).to(ref.dtype) for v in values]
        return torch.cat(vals, dim=axis)
    TorchBackend.concat = concat

    def split(self, tensor, num_or_size, axis):          # tf.split (tf_backend.py:415-416): sizes, one may be -1
        if isinstance(num_or_size, int):
            return list(torch.chunk(tensor, num_or_size, dim=axis))
        sizes = [int(k) for k in num_or_size]
        if -1 in sizes:
            sizes[sizes.index(-1)] = int(tensor.shape[axis]) - (sum(sizes) + 1)
        return list(torch.split(tensor, sizes, dim=axis))
    TorchBackend.split = split

    def roll(self, tensor, shift, axis):                 # tf.roll (tf_backend.py:418-419)
        return torch.roll(tensor, shift, axis)
    TorchBackend.roll = roll


def load_reference(oracle):
    """-> (namespace holding the reference's definitions, OpLog)."""
    _restore_aliases()
    warnings.filterwarnings("ignore")
    if REF + "/PhiFlow" not in sys.path:
        sys.path.insert(0, REF + "/PhiFlow")
    ns = {}
    exec("from phi.torch.flow import *\n"
         "from phi.physics.field.staggered_grid import *\n"
         "from phi.physics.field.grid import *\n"
         "from phi.physics.pressuresolver.solver_api import PoissonSolver\n"
         "import six, os, sys, scipy", ns)
    _complete_torch_backend()
    log = OpLog()
    ns["tf"] = TfShim()
    ns["np"] = np
    ns["cd_csr_op"], ns["multi_bicg_op"], ns["pressure_op"] = make_op_shims(oracle, log)
    d = REF + "/diffpiso/"
    wanted = [
        (d + "piso_helpers.py", None),
        (d + "linear_solver.py", {"LinearSolver", "LinearSolverCudaMultiBicgstabILU"}),
        (d + "piso_cuda_pressure_solver.py", {"PisoPressureSolverCudaCustom"}),
        (d + "piso_tf.py", {"piso_step", "advection_matrix_cuda", "pressure_extrapolation", "SimulationParameters"}),
        (d + "combined_training_integrated.py", {"zero_gradient_op", "run_piso_steps"}),
        (d + "networks.py", {"fullyconv_network"}),
        (d + "evaluation_tools.py", {"tf_fftshift", "EK_spectrum_2D_tf", "EK_spectrum_1D_tf"}),
        (d + "losses.py", None),
        (REF + "/spatial_mixing_layer_differentiable_training.py", {"neural_network_wrapper"}),
    ]
    for path, names in wanted:
        for node in _definitions(path, names):
            if path.endswith("piso_cuda_pressure_solver.py") and isinstance(node, ast.ClassDef):
                ns.setdefault("SimulationParameters", object)      # annotation in the solve() signature
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns, log
