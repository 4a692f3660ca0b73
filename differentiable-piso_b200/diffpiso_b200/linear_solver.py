"""Velocity-predictor solvers with the reference's plug-in surface (diffpiso/linear_solver.py:15-30, 114-178).

`LinearSolverCudaMultiBicgstabILU.solve` keeps the reference signature and return value `[x, warn]`; the TF-1.14
`tf.custom_gradient` registration (":169-173") becomes a `torch.autograd.Function` whose backward is the transposed
solve multiplied by (1 - warn).  The solve itself is one launch of the batched persistent kernel in csrc/bicgstab.cu.
"""
import torch

from . import ops
from .grids import as_tensor


class LinearSolver(object):
    """diffpiso/linear_solver.py:15-30"""

    def __init__(self, name, supported_devices, supports_guess, supports_batch, solver_type, input_format):
        self.name = name
        self.supported_devices = supported_devices
        self.supports_guess = supports_guess
        self.supports_batch = supports_batch
        self.solver_type = solver_type
        self.input_format = input_format

    def solve(self, *args):
        raise NotImplementedError(self.__class__)

    def __repr__(self):
        return self.name


def _infer_periodic(ny, nx, nnz_total):
    """Periodic flags from the total entry count (piso_tf.py:102-106) when the caller gave no structure."""
    nf = ny * (nx + 1) + (ny + 1) * nx
    hits = [(py, px) for py in (False, True) for px in (False, True)
            if 5 * nf - 2 * (1 - px) * (2 * ny + 1) - 2 * (1 - py) * (2 * nx + 1) == nnz_total]
    if len(hits) != 1:
        raise ValueError("cannot infer the periodic flags from the matrix size; pass structure=Geometry")
    return hits[0]


class _BicgSolveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rhs, values, x0, solver, geom, transpose):
        pivots = None
        if ctx.needs_input_grad[0] and solver.reuse_factors and not solver.cast_to_double and ops.factor_reuse_supported(geom):
            pivots = torch.empty_like(rhs)
        x, stats, warn = ops.bicgstab_ilu(geom, values, rhs, x0, solver.accuracy, solver.max_iterations, transpose,
                                          pivots_out=pivots, fp64=solver.cast_to_double)
        solver.last_stats = stats
        ctx.solver, ctx.geom, ctx.transpose, ctx.has_pivots = solver, geom, transpose, pivots is not None
        ctx.save_for_backward(*([values, x0] + ([pivots] if pivots is not None else [])))
        ctx.mark_non_differentiable(warn)
        return x, warn

    @staticmethod
    def backward(ctx, gx, gwarn):
        saved = ctx.saved_tensors
        values, x0 = saved[0], saved[1]
        solver = ctx.solver
        # linear_solver.py:169-173: same op on ds with `not transpose`, same initial-guess tensor, times (1 - warn);
        # the ILU(0) pivots of the forward solve are reused where that is exact (structurally symmetric components)
        df, stats, warn = ops.bicgstab_ilu(ctx.geom, values, gx.contiguous(), x0, solver.accuracy, solver.max_iterations,
                                           not ctx.transpose, pivots_in=saved[2] if ctx.has_pivots else None,
                                           fp64=solver.cast_to_double)
        solver.last_adjoint_stats = stats
        # per-sample NaN guard (stats[:, :, 2]); the batch-wide OR is only the returned `warn` value
        keep = 1.0 - stats[:, :, 2].amax(dim=1, keepdim=True).to(torch.float32)
        return df * keep, None, None, None, None, None


class LinearSolverCudaMultiBicgstabILU(LinearSolver):
    """Batched ILU(0)-BiCGStab for the u and v momentum systems (diffpiso/linear_solver.py:114-178)."""

    _dpiso_native = True

    def __init__(self, accuracy=1e-5, max_iterations=2000, cast_to_double=False, reuse_factors=True):
        LinearSolver.__init__(self, 'CUDA dual iLU-preconditioned BiCGStab solve', supported_devices=('GPU',),
                              supports_guess=True, supports_batch=True, solver_type='iterative', input_format='csr')
        self.max_iterations = int(max_iterations)
        # True: the fp64 variant (linear_solver.py:130-133): values / rhs are cast to fp64, the solution back to fp32;
        # here the casts are fused into the kernel (dpiso_bicgstab_ilu_f64)
        self.cast_to_double = bool(cast_to_double)
        self.accuracy = float(accuracy)
        # adjoint solves take the forward ILU(0) pivots instead of factorising A^T again, per component and only where
        # that is the reference's preconditioner up to rounding (structurally symmetric pattern); components
        # that are periodic along their staggered axis (Q18) always re-factorise like the reference
        self.reuse_factors = bool(reuse_factors)
        self.last_stats = None
        self.last_adjoint_stats = None

    def solve_native(self, geom, values, rhs, x0, transpose, negate=False, pivots_out=None, pivots_in=None,
                     adjoint=False):
        """The solve without autograd bookkeeping, for callers that own forward and backward themselves (piso_step):
        -> (x, warn float32 [1], stats int32 [B, 2, 4])."""
        x, stats, warn = ops.bicgstab_ilu(geom, values, rhs, x0, self.accuracy, self.max_iterations, transpose,
                                          negate=negate, pivots_out=None if self.cast_to_double else pivots_out,
                                          pivots_in=None if self.cast_to_double else pivots_in, fp64=self.cast_to_double)
        if adjoint:
            self.last_adjoint_stats = stats
        else:
            self.last_stats = stats
        return x, warn, stats

    def solve(self, matrix_values, row_ptr, col_indices, rhs, staggered_shape, initial_guess=None, offset=0,
              transpose=False, unrolling_step=0, warn=None, structure=None):
        """matrix_values [B, nnz] (or 1-D for B = 1), rhs / initial_guess [B, n_u+n_v] flattened [u, v].
        row_ptr / col_indices are accepted for interface parity; the pattern is implied by the grid (`structure`,
        a `Geometry`, or inferred from `staggered_shape` and the entry count).  Returns [x, warn]."""
        rhs = as_tensor(rhs)
        one_d = rhs.dim() == 1
        values = as_tensor(matrix_values)
        if values.dim() == 1:
            values = values[None]
        if one_d:
            rhs = rhs[None]
        ny, nx = int(staggered_shape[1]) - 1, int(staggered_shape[2]) - 1
        if structure is None:
            per_y, per_x = _infer_periodic(ny, nx, values.shape[-1])
            structure = ops.Geometry.get(ny, nx, per_y, per_x, rhs.device)
        if initial_guess is None:
            x0 = torch.zeros_like(rhs)
        else:
            x0 = as_tensor(initial_guess).reshape(rhs.shape).detach()
        x, warn_f = _BicgSolveFn.apply(rhs, values.detach(), x0, self, structure, bool(transpose))
        if warn is not None:
            warn_f = torch.maximum(warn_f, as_tensor(warn).to(warn_f.device).reshape(-1)[:1].to(torch.float32))
        if one_d:
            x = x[0]
        return [x, warn_f]


class LinearSolverCudaBicgstabILU(LinearSolverCudaMultiBicgstabILU):
    """The single-matrix predecessor (diffpiso/linear_solver.py:60-111): same algorithm, but its own call surface --
    `solve(matrix_values, row_ptr, col_indices, rhs, initial_guess=None, offset=0, transpose=False)` returning the
    solution tensor only.  The matrix is one momentum component pair in the layout of the Multi solver; the grid is
    named with the keyword `structure` (a `Geometry`) or `staggered_shape`."""

    def solve(self, matrix_values, row_ptr, col_indices, rhs, initial_guess=None, offset=0, transpose=False,
              structure=None, staggered_shape=None):
        if structure is None and staggered_shape is None:
            raise ValueError("LinearSolverCudaBicgstabILU.solve needs structure=Geometry or staggered_shape=[B, ny+1, nx+1, 2]")
        if staggered_shape is None:
            staggered_shape = (1, structure.ny + 1, structure.nx + 1, 2)
        x, _ = LinearSolverCudaMultiBicgstabILU.solve(self, matrix_values, row_ptr, col_indices, rhs, staggered_shape,
                                                      initial_guess, offset=offset, transpose=transpose,
                                                      structure=structure)
        return x
