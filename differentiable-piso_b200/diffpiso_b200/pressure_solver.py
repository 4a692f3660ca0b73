"""Pressure-Poisson solver with the reference's plug-in surface (diffpiso/piso_cuda_pressure_solver.py:36-114).

`PisoPressureSolverCudaCustom.solve(scaling_field, divergence, guess, enable_backprop, simulation_physics, ...)`
returns `(pressure, iteration, laplace)` like the reference.  Differences that are visible: `iteration` has one entry
per sample (the reference has batch 1), and every sample follows the reference's control flow on its own.
The adjoint registration (":97-107") -- the same solve applied to the incoming gradient -- is a torch.autograd.Function.
"""
import torch

from . import ops
from .grids import as_tensor, flatten_staggered_data


class PoissonSolver(object):
    """PhiFlow/phi/physics/pressuresolver/solver_api.py:10-46 (constructor surface only)"""

    def __init__(self, name, supported_devices, supports_guess, supports_loop_counter, supports_continuous_masks):
        self.name = name
        self.supported_devices = supported_devices
        self.supports_guess = supports_guess
        self.supports_loop_counter = supports_loop_counter
        self.supports_continuous_masks = supports_continuous_masks

    def solve(self, *args, **kwargs):
        raise NotImplementedError(self.__class__)

    def __repr__(self):
        return self.name


class ScalingFromDiagonal(object):
    """Lazy form of the scaling field `1/(beta - A) * dx_factor` (piso_tf.py:53-54): carries the flat matrix diagonal so
    that the Laplace kernel forms the face coefficients on the fly instead of reading a materialised staggered tensor."""

    def __init__(self, a_diag_flat, beta, dx_factor):
        self.a_diag, self.beta, self.dx_factor = a_diag_flat, float(beta), float(dx_factor)
        self._lap = {}      # pressure matrix built from this diagonal, per precision: the two solves of a pass share it
                            # (the reference rebuilds the identical matrix for every solve)


class _PressureSolveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, div, lap, solver, geom, rank_deficient):
        p, its = ops.pressure_cg(geom, lap, div, solver.accuracy, solver.max_iterations, solver.residual_reset,
                                 rank_deficient)
        ctx.solver, ctx.geom, ctx.rank_deficient = solver, geom, rank_deficient
        ctx.save_for_backward(lap)
        ctx.mark_non_differentiable(its)
        solver.last_iterations = its
        return p, its

    @staticmethod
    def backward(ctx, gp, gits):
        (lap,) = ctx.saved_tensors
        s = ctx.solver
        # piso_cuda_pressure_solver.py:97-107: same op on dp, zero initial guess, no transpose (symmetric matrix)
        g, its = ops.pressure_cg(ctx.geom, lap, gp.contiguous(), s.accuracy, s.max_iterations, s.residual_reset,
                                 ctx.rank_deficient)
        s.last_adjoint_iterations = its
        return g, None, None, None, None


class PisoPressureSolverCudaCustom(PoissonSolver):
    """CG on the 5-point variable-coefficient PISO pressure matrix (diffpiso/piso_cuda_pressure_solver.py:36-114)."""

    _dpiso_native = True

    def __init__(self, dx, accuracy=1e-5, max_iterations=2000, residual_reset=10, randomized_restarts=0,
                 cast_to_double=True):
        PoissonSolver.__init__(self, 'CUDA Conjugate Gradient', supported_devices=('GPU',), supports_loop_counter=False,
                               supports_guess=True, supports_continuous_masks=False)
        assert randomized_restarts >= 0
        if randomized_restarts != 0:
            raise NotImplementedError("randomized_restarts > 0 is dead code in the reference (SURVEY Q3) and not built")
        self.accuracy = float(accuracy)
        self.max_iterations = int(max_iterations)
        self.dx = dx
        self.residual_reset = int(residual_reset)
        self.randomized_restarts = 0
        self.cast_to_double = bool(cast_to_double)
        self.laplace_rank_deficient = None
        self.last_iterations = None
        self.last_adjoint_iterations = None

    def solve(self, scaling_field, divergence, guess, enable_backprop, simulation_physics, offset=0, unrolling_step=0):
        """divergence [B, ny, nx, 1]; scaling_field: staggered tensor [B, ny+1, nx+1, 2] of 1/(beta-A)*dx_factor, or a
        `ScalingFromDiagonal`.  `guess` is accepted and ignored exactly like the reference (init_with_zeros=True,
        ":95").  Returns (pressure [B, ny, nx, 1] float32, iterations int32 [B], laplace [B, 5*ny*nx])."""
        div = as_tensor(divergence)
        b, ny, nx = div.shape[0], div.shape[1], div.shape[2]
        per_y, per_x = _periodic_flags(simulation_physics)
        geom = ops.Geometry.get(ny, nx, per_y, per_x, div.device)
        masks = ops.to_device_masks(simulation_physics, geom)
        if self.laplace_rank_deficient is None:                      # ":83-87" (cached on first use, like the reference)
            self.laplace_rank_deficient = masks["rank_deficient"]
        if isinstance(scaling_field, ScalingFromDiagonal):
            lap = scaling_field._lap.get(self.cast_to_double)
            if lap is None:
                lap = ops.laplace(geom, masks["active"], masks["access"], scaling_field.a_diag, 1, scaling_field.beta,
                                  scaling_field.dx_factor, fp64=self.cast_to_double)
                scaling_field._lap[self.cast_to_double] = lap
        else:
            k = flatten_staggered_data(as_tensor(scaling_field), coord_flip=False).contiguous()   # ":70"
            if k.shape[0] == 1 and b > 1:
                k = k.expand(b, -1).contiguous()
            lap = ops.laplace(geom, masks["active"], masks["access"], k, 0, 0.0, 1.0, fp64=self.cast_to_double)
        p, its = _PressureSolveFn.apply(div.reshape(b, ny * nx), lap, self, geom, bool(self.laplace_rank_deficient))
        return p.reshape(b, ny, nx, 1), its, lap.reshape(b, -1)


def _periodic_flags(sim):
    bp = getattr(sim, "bool_periodic", None)
    if bp is None:
        return False, False
    return bool(bp[0]), bool(bp[1])
