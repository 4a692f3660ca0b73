"""GPU parity against golden vectors produced by the REFERENCE'S OWN PYTHON (tests/golden/reference_runner.py):
one `piso_step` forward + backward per setup, and a 3-step `run_piso_steps` unroll with the closure network, the
per-step inflow update and a stop-gradient window.  Nothing here touches /root/reference at run time."""
import os

import numpy as np
import pytest
import torch

from common import SMALL_SETUPS, cg_iteration_slack, rel_l2
from test_gpu_piso_step import DEV, build_sim, extrap

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python")


def _t(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
    return t.requires_grad_(True) if grad else t


@pytest.mark.parametrize("name", ["ldc8", "periodic16", "periodic24x20", "tml16x24", "sml16x48", "obstacle16x24"])
def test_piso_step_forward_backward_matches_reference_python(name):
    """Integer/assembly outputs bit-exact; fields within the north_star 1e-5 relative L2; iteration counts +-1
    (BiCGStab) / quantisation slack (CG); gradients within 1e-4 (three nested iterative solves)."""
    import diffpiso_b200 as dp
    g, s = np.load(os.path.join(GOLD, "step_%s.npz" % name)), SMALL_SETUPS[name]()
    sim = build_sim(s)
    ny, nx = s["ny"], s["nx"]
    nc = ny * nx
    dxy = (s["dy"], s["dx"])
    tv, tp, tf = _t(g["vel"][None], True), _t(g["pres"][None], True), _t(g["forcing"][None], True)
    td = _t(s["dirichlet_values"][None], True)
    velocity = dp.StaggeredGrid(flat=tv, resolution=(ny, nx), dx=dxy)
    pressure = dp.CenteredGrid(tp.reshape(1, ny, nx, 1), dx=dxy, extrapolation=extrap(s["pbc"]))
    inc = dp.CenteredGrid(torch.zeros(1, ny, nx, 1, device=DEV), dx=dxy, extrapolation=extrap(s["pbc_inc"]))
    visc_field = _t(s["visc"]) if np.atleast_1d(s["visc"]).size > 1 else None
    out = dp.piso_step(velocity, pressure, inc, inc, s["dt"], sim, td, viscosity_field=visc_field, forcing_term=tf,
                       full_output=True)
    assert np.array_equal(out[6].cpu().numpy(), g["row_ptr"]) and np.array_equal(out[5].cpu().numpy(), g["col_ind"])
    assert np.array_equal(out[4][0].cpu().numpy(), g["values"])
    assert np.array_equal(out[9][0].cpu().numpy(), g["a_diag"])
    assert np.array_equal(out[10][0].cpu().numpy(), g["rhs"])
    assert np.array_equal(out[14][0].cpu().numpy().ravel(), g["lap1"].ravel())
    n_u = ny * (nx + 1)
    u_star = torch.cat([out[7][:, :-1, :, 1].reshape(1, -1), out[7][:, :, :-1, 0].reshape(1, -1)], dim=1)[0].cpu().numpy()
    assert rel_l2(u_star, g["u_star"]) < 1e-5
    assert rel_l2(out[13].detach().cpu().numpy().ravel(), g["div1"]) < 1e-4
    assert rel_l2(out[2].data.detach().cpu().numpy().ravel(), g["p1"]) < 1e-4
    assert rel_l2(out[0].flat[0].detach().cpu().numpy(), g["vel_next"]) < 1e-5
    assert rel_l2(out[1].data.detach().cpu().numpy().ravel(), g["pres_next"]) < 1e-4
    bicg = sim.linear_solver.last_stats.cpu().numpy()
    assert abs(int(bicg[0, 0, 0]) - int(g["bicg_iterations"][0])) <= 1
    assert abs(int(bicg[0, 1, 0]) - int(g["bicg_iterations"][1])) <= 1
    it2 = int(sim.pressure_solver.last_iterations[0])
    assert abs(it2 - int(g["cg_iterations"][1])) <= cg_iteration_slack(s, int(g["cg_iterations"][1]))
    loss = (out[0].flat * _t(g["w_u"][None])).sum() + (out[1].data.reshape(1, nc) * _t(g["w_p"][None])).sum()
    loss.backward()
    tol = 1e-4
    assert rel_l2(tv.grad[0].cpu().numpy(), g["g_vel"]) < tol
    assert rel_l2(tp.grad[0].cpu().numpy(), g["g_pres"]) < tol
    assert rel_l2(tf.grad[0].cpu().numpy(), g["g_forcing"]) < tol
    if s["dirichlet"].any():
        assert rel_l2(td.grad[0].cpu().numpy(), g["g_dvals"]) < (5e-4 if name == "ldc8" else tol)


def test_run_piso_steps_unroll_matches_reference_python():
    """combined_training_integrated.py:396-478 on the spatial mixing layer 16x48: three unrolled steps with closure
    forcing, inflow perturbation per step, gradients stopped after step 2; all step states, network outputs and the
    gradients w.r.t. the closure weights and the initial state against the reference's own run."""
    import diffpiso_b200 as dp
    from diffpiso_b200 import masks as M, networks as N, setups as SU, training as T
    g = np.load(os.path.join(GOLD, "unroll_sml16x48.npz"))
    s = SMALL_SETUPS["sml16x48"]()
    ny, nx = s["ny"], s["nx"]
    steps = g["velocities"].shape[0]
    torch.backends.cudnn.allow_tf32 = False          # fp32 convolutions for the comparison (TF32 is ~1e-3)
    dxy = (s["dy"], s["dx"])
    sim = build_sim(s)
    bcx = g["bcx"]
    sim.dirichlet_values = _t(M.update_dirichlet_values(s["dirichlet_values_staggered"], ((False, False), (True, False)),
                                                        (([], []), (bcx + g["bc_pert"][0], []))).astype(np.float32))
    w = [_t(g["w%d" % i], True) for i in range(7)]
    tv = _t(SU.stagger_flat(g["vel"][None], ny, nx), True)
    tp = _t(g["pres"].reshape(1, ny, nx, 1), True)
    velocity = dp.StaggeredGrid(tv, dx=dxy)
    pressure = dp.CenteredGrid(tp, dx=dxy, extrapolation=extrap(s["pbc"]))
    simulation_parameters = dict(dx_ratio=1, dt=s["dt"], dt_ratio=1, HRres=[ny, nx], sponge_ratio=0.875)
    training_dict = dict(step_count=steps, HR_buffer_width=[[0, 0], [0, 0]], pressure_included=True, loss_influence_range=2)
    network = lambda x: N.fullyconv_network(x, w, [[0, 0], [0, 0]], "SAME", False)
    update = lambda dv, pl: M.update_dirichlet_values(dv, ((False, False), (True, False)), pl)
    out = T.run_piso_steps(velocity, pressure, velocity, {}, simulation_parameters, training_dict, network,
                           T.spatial_mixing_layer_network_wrapper, sim, _t(np.asarray(s["visc"], np.float32)), bcx,
                           _t(g["bc_pert"]), update, None)
    for k in range(steps):
        assert rel_l2(out[2][k].detach().cpu().numpy(), g["nn_out"][k]) < 2e-5, k
        assert rel_l2(out[7][k].detach().cpu().numpy(), g["velocities"][k]) < 1e-5 * (k + 1), k
        assert rel_l2(out[8][k].detach().cpu().numpy(), g["pressures"][k]) < 1e-4 * (k + 1), k
    loss = sum((out[7][k] * _t(g["w_loss"][k])).sum() for k in range(steps))
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    loss.backward()
    for i in range(7):
        assert rel_l2(w[i].grad.cpu().numpy(), g["g_w%d" % i]) < 1e-3, i
    assert rel_l2(tv.grad.cpu().numpy(), g["g_vel"]) < 1e-3
    assert rel_l2(tp.grad.cpu().numpy(), g["g_pres"]) < 1e-3


@pytest.mark.parametrize("name", ["ldc8", "periodic16", "periodic24x20", "tml16x24", "sml16x48", "obstacle16x24"])
def test_bicgstab_kernel_matches_reference_cpu_solver(name):
    """CUDA ILU0-BiCGStab (forward and transposed) against the reference's CPU solver path (spsolve on the CSR matrices
    built by the reference's convert_to_scipy_csr), stored in the step goldens."""
    from diffpiso_b200 import ops
    g, s = np.load(os.path.join(GOLD, "step_%s.npz" % name)), SMALL_SETUPS[name]()
    geo = ops.Geometry.get(s["ny"], s["nx"], s["per_y"], s["per_x"], torch.device(DEV))
    neg = _t(-g["values"][None])
    x, st, w = ops.bicgstab_ilu(geo, neg, _t(g["rhs"][None]), _t(g["vel"][None]), s["bicg_tol"], s["bicg_max_it"], False)
    assert rel_l2(x[0].cpu().numpy(), g["u_star_spsolve"]) < 1e-6
    xt, st, w = ops.bicgstab_ilu(geo, neg, _t(g["bwd_bicg_rhs"][None]), _t(g["bwd_bicg_x0"][None]), s["bicg_tol"],
                                 s["bicg_max_it"], True)
    assert rel_l2(xt[0].cpu().numpy(), g["bicg_adj_spsolve"]) < 2e-5
    assert int(w[0]) == 0
