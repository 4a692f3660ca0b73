"""diffpiso_b200 -- B200-native PISO step (hot path of tum-pbs/differentiable-piso) behind the reference's operator
surface.  Importing this package loads libdpiso.so (sm_100a kernels); there is no CPU or library fallback."""
from . import _native
from .grids import (CenteredGrid, StaggeredGrid, flatten_staggered_data, stack_staggered_components,
                    stagger_flattened_data, unstack_staggered_tensor)
from .helpers import (arrange_rhs_term, custom_padded, explicit_H_csr, finite_volume_divergence,
                      finite_volume_gradient_tensor)
from .linear_solver import LinearSolver, LinearSolverCudaBicgstabILU, LinearSolverCudaMultiBicgstabILU
from .ops import Geometry
from .piso import SimulationParameters, advection_matrix_cuda, piso_step, pressure_extrapolation
from .pressure_solver import PisoPressureSolverCudaCustom, PoissonSolver
from .masks import compute_mixingLayer_masks, temporal_mixing_layer_masks, update_dirichlet_values
from .networks import fullyconv_network, initialise_fullyconv_network
from .losses import L2_field_loss, multistep_averaging_loss, spectral_energy_loss, strain_rate_loss
from .training import boundary_perturbation_fun, inference_rollout, run_piso_steps, zero_gradient_op
from .datamanagement import create_base_dir, data_path_assembler, load_function
from .sharding import SampleGroups

__all__ = ["CenteredGrid", "StaggeredGrid", "flatten_staggered_data", "stagger_flattened_data",
           "stack_staggered_components", "unstack_staggered_tensor", "LinearSolver", "LinearSolverCudaBicgstabILU",
           "LinearSolverCudaMultiBicgstabILU", "PisoPressureSolverCudaCustom", "PoissonSolver", "SimulationParameters",
           "piso_step", "advection_matrix_cuda", "pressure_extrapolation", "Geometry", "custom_padded", "arrange_rhs_term",
           "finite_volume_gradient_tensor", "finite_volume_divergence", "explicit_H_csr",
           "compute_mixingLayer_masks", "temporal_mixing_layer_masks", "update_dirichlet_values", "fullyconv_network",
           "initialise_fullyconv_network", "L2_field_loss", "spectral_energy_loss", "strain_rate_loss",
           "multistep_averaging_loss", "run_piso_steps", "zero_gradient_op", "inference_rollout",
           "boundary_perturbation_fun", "create_base_dir", "data_path_assembler",
           "load_function", "SampleGroups"]
