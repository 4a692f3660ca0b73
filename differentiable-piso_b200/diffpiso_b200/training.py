"""The caller of the PISO path in training and inference: `run_piso_steps` unrolls `step_count`
steps, optionally with the closure network's forcing (diffpiso/combined_training_integrated.py:396-478), and the
data-parallel pieces around it -- gradient all-reduce of the closure weights over NCCL (the only collective of the
whole workload: samples are independent) and an Adam step (`training_run`, :71-76).

Everything here is torch glue; each `piso_step` inside runs on the native kernels and brings its own adjoint."""
import numpy as np
import torch

from .grids import CenteredGrid, StaggeredGrid, stack_staggered_components
from .piso import piso_step


class _ZeroGradient(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * 0


def zero_gradient_op(centered_data):
    """combined_training_integrated.py:386-393: identity whose gradient is zero (used instead of stop_gradient on the
    pressure so that the backward graph stays connected)."""
    return _ZeroGradient.apply(centered_data)


def closure_input(velocity, pressure, pressure_included=True):
    """Network input [B, ny, nx, 2 or 4]: face velocities averaged to the centres (channel 0 = v, 1 = u) and, if
    requested, the central-difference pressure gradient (d/dy, d/dx) (:401-405)."""
    nn_in = velocity.at_centers().data
    if pressure_included:
        nn_in = torch.cat([nn_in, pressure.gradient().data], dim=-1)
    return nn_in


def closure_forcing(nn_out, velocity):
    """Network output [B, ny, nx, 2] -> staggered forcing tensor: channel 0 resampled to the v faces, channel 1 to the
    u faces, each as a CenteredGrid with the default 'boundary' extrapolation (:407-410)."""
    fv = CenteredGrid(nn_out[..., 0:1], dx=velocity.dx).at_faces(0)
    fu = CenteredGrid(nn_out[..., 1:2], dx=velocity.dx).at_faces(1)
    return stack_staggered_components([fv, fu])


def spatial_mixing_layer_network_wrapper(neural_network, input, fluid, physical_parameters, simulation_parameters,
                                         loss_buffer_width, buffer_width):
    """spatial_mixing_layer_differentiable_training.py:6-10: the network sees the region upstream of the sponge layer
    only; its output is zero-padded back to the full width."""
    sponge_start = int(simulation_parameters["HRres"][1] * simulation_parameters["sponge_ratio"]) // simulation_parameters["dx_ratio"]
    out = neural_network(input[:, :, :sponge_start, :])
    return torch.nn.functional.pad(out, (0, 0, 0, int(fluid.resolution[1]) - sponge_start))


def run_piso_steps(velocity, pressure, domain, physical_parameters, simulation_parameters, training_dict, neural_network,
                   neural_network_wrapper, sim_physics, viscosity_field, bcx, bc_placeholders,
                   dirichlet_placeholder_update=None, loss_buffer_width=None):
    """combined_training_integrated.py:396-478, same arguments and the same 9-tuple
    (velocity_all_steps, pressure_all_steps, nn_all_steps, velnew, pnew, NN_out, warn, velocity_all_arrays,
    pressure_all_arrays).  `domain` only needs `.resolution`; `bc_placeholders` is the per-step inflow perturbation
    [step_count, 1, ny+2, 1, 1] (a tensor instead of a TF placeholder); every sample of the batch is unrolled at once."""
    step_count = training_dict["step_count"] if training_dict is not None else 1
    dt = simulation_parameters["dt"] * simulation_parameters["dt_ratio"]
    dirichlet_values = sim_physics.dirichlet_values
    device = velocity.flat.device
    nn_all_steps, nn_out = [], []
    velocity_all_steps, pressure_all_steps, velocity_all_arrays, pressure_all_arrays = [], [], [], []
    warn = [None] * step_count
    if neural_network is not None:
        buffer_width = [[i // simulation_parameters["dx_ratio"] for i in j] for j in training_dict["HR_buffer_width"]]
    velnew, pnew = velocity, pressure
    for i in range(step_count):
        if i > 0:
            if i % training_dict["loss_influence_range"] == 0:            # :436-438
                velnew = StaggeredGrid(velnew.staggered_tensor().detach(), dx=velnew.dx, extrapolation=velnew.extrapolation,
                                       pad_periodic=velnew.pad_periodic)
                pnew = CenteredGrid(zero_gradient_op(pnew.data), dx=pnew.dx, extrapolation=pnew.extrapolation)
            if dirichlet_placeholder_update is not None:                  # :440-441
                bc = torch.as_tensor(np.asarray(bcx), dtype=bc_placeholders[i].dtype, device=device) + bc_placeholders[i]
                dirichlet_values = dirichlet_placeholder_update(sim_physics.dirichlet_values, (([], []), (bc, [])))
        residual_force = None
        if neural_network is not None:
            nn_in = closure_input(velnew, pnew, training_dict["pressure_included"])
            nn_out = neural_network_wrapper(neural_network, nn_in, domain, physical_parameters, simulation_parameters,
                                            loss_buffer_width, buffer_width)
            residual_force = closure_forcing(nn_out, velocity)
            nn_all_steps.append(nn_out)
        # the increments only carry box / extrapolation (their values are ignored by the solver); the
        # reference builds them without an extrapolation argument, i.e. 'boundary' (:419-420)
        inc1 = CenteredGrid(torch.zeros_like(pressure.data) + 5e-13, dx=pressure.dx)
        inc2 = CenteredGrid(torch.zeros_like(pressure.data) + 1e-12, dx=pressure.dx)
        vel_piso, p_piso, warn[i] = piso_step(velnew, pnew, inc1, inc2, dt, sim_physics, dirichlet_values,
                                              viscosity_field=viscosity_field, forcing_term=residual_force,
                                              unrolling_step=i)
        velocity_all_steps.append(vel_piso)
        pressure_all_steps.append(p_piso)
        velocity_all_arrays.append(vel_piso.staggered_tensor())
        pressure_all_arrays.append(p_piso.data)
        # Q21: the reference writes StaggeredGrid(array, vel_piso.box, vel_piso.extrapolation) (:431-432, :473-474), but the
        # third positional parameter of PhiFlow's StaggeredGrid is `name`: the re-wrapped state has the DEFAULT
        # extrapolation 'boundary'.  From the second unrolled step on custom_padded therefore replicates the velocity on
        # periodic axes (the matrix keeps its periodic structure) and finite_volume_divergence registers its non-circular
        # gradient.  Reproduced, because it changes the states (1e-4 per step on a 256-wide grid) and the gradients.
        velnew = StaggeredGrid(velocity_all_arrays[i], dx=vel_piso.dx, extrapolation="boundary", pad_periodic=(False, False))
        pnew = CenteredGrid(pressure_all_arrays[i], dx=p_piso.dx, extrapolation=p_piso.extrapolation)
    return (velocity_all_steps, pressure_all_steps, nn_all_steps, velnew, pnew, nn_out, warn, velocity_all_arrays,
            pressure_all_arrays)


def allreduce_gradients(parameters, world_size=None, average=True):
    """Sum (or mean) of the closure-network gradients over the ranks: ONE flat NCCL all-reduce per training iteration
    (81 856 parameters = 0.33 MB; latency-bound, so a single bucket).  No-op without an initialised process group."""
    import torch.distributed as dist
    params = [p for p in parameters if p.grad is not None]
    if not (dist.is_available() and dist.is_initialized()) or not params:
        return
    world_size = dist.get_world_size() if world_size is None else world_size
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= world_size
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def training_iteration(optimizer, weights, loss_fn):
    """One optimisation step as in `training_run` (:61-76, :300-330): loss -> gradients w.r.t. the closure weights
    through the unrolled PISO steps -> all-reduce over the data-parallel ranks -> Adam."""
    optimizer.zero_grad(set_to_none=True)
    loss = loss_fn()
    loss.backward()
    allreduce_gradients(weights)
    optimizer.step()
    return loss.detach()


def boundary_perturbation_fun(domain, average_velocity, shape, time, perturbation_amplitudes):
    """combined_training_integrated.py:7-14 (after J. Ko et al.): time-dependent perturbation of the inflow profile of
    the spatial mixing layer, one value per row of the padded grid.  `domain` needs `.resolution` and the box height
    (`.box` = (ly, lx) or an object with `.size`)."""
    ny = int(domain.resolution[0])
    box = domain.box
    ly = float(np.asarray(box.size if hasattr(box, "size") and not isinstance(box, np.ndarray) else box).ravel()[0])
    y_disc = np.linspace(0, ly, ny + 2) - ly / 2
    eps = [perturbation_amplitudes[0] * average_velocity, perturbation_amplitudes[1] * average_velocity]
    n = [.4 * np.pi, .3 * np.pi]
    omeg = [.22, .11]
    u_perturb = np.sum([eps[i] * np.cos(n[i] * y_disc) * (1 - np.tanh(y_disc / 2) ** 2) * np.sin(omeg[i] * time)
                        for i in range(len(eps))], axis=0)
    return np.reshape(u_perturb, shape)


def inference_rollout(velocity, pressure, timesteps, domain, physical_parameters, simulation_parameters, sim_physics,
                      viscosity_field, bcx=None, perturbation_fun=None, dirichlet_placeholder_update=None,
                      neural_network=None, neural_network_wrapper=None, training_dict=None, save_dir=None,
                      on_step=None):
    """Forward roll-out as in spatial_mixing_layer_differentiable_inference.py:100-165 (and the data generator
    spatial_mixing_layer.py:52-92): `timesteps - 1` PISO steps, the inflow perturbation re-evaluated every step, the
    closure network's forcing if a network is given; frames are written as velocity_/pressure_/nn_forcing_%06d.npz
    (batch 1) when `save_dir` is set.  Unlike the reference the state stays on the GPU between steps.
    -> (velocity, pressure) grids after the last step."""
    from .datamanagement import frame_path, save_frame
    device = velocity.flat.device
    td = dict(step_count=1, HR_buffer_width=[[0, 0], [0, 0]], pressure_included=True, loss_influence_range=1)
    if training_dict is not None:
        td.update({k: training_dict[k] for k in ("HR_buffer_width", "pressure_included") if k in training_dict})
    if save_dir is not None:
        save_frame(save_dir, 0, velocity, pressure)
    dt = simulation_parameters["dt"] * simulation_parameters["dt_ratio"]
    base_values = sim_physics.dirichlet_values
    with torch.no_grad():
        for i in range(1, timesteps):
            if perturbation_fun is not None and dirichlet_placeholder_update is not None:
                ny = int(velocity.resolution[0])
                pert = torch.as_tensor(np.asarray(perturbation_fun((1, ny + 2, 1, 1), dt * i), np.float32), device=device)
                bc = torch.as_tensor(np.asarray(bcx, np.float32), device=device) + pert
                sim_physics.dirichlet_values = dirichlet_placeholder_update(base_values, (([], []), (bc, [])))
            out = run_piso_steps(velocity, pressure, domain, physical_parameters, simulation_parameters, td, neural_network,
                                 neural_network_wrapper, sim_physics, viscosity_field, bcx, None, None, None)
            velocity, pressure = out[3], out[4]
            if save_dir is not None:
                save_frame(save_dir, i, velocity, pressure)
                if neural_network is not None:
                    np.savez(frame_path(save_dir, "nn_forcing", i), out[5].detach().cpu().numpy())
            if on_step is not None:
                on_step(i, velocity, pressure)
    sim_physics.dirichlet_values = base_values
    return velocity, pressure
