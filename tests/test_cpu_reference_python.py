"""CPU suite: the oracle against golden vectors produced by the REFERENCE'S OWN PYTHON executed in the build container
(tests/golden/reference_runner.py: `diffpiso.piso_tf.piso_step`, the solver classes' `solve`, the helper functions and
every registered `grad` closure run unmodified on PhiFlow's torch backend; only the three CUDA op libraries are replaced
by callables with the ops' argument lists).  This pins

* the Python glue of the step (padding, flattening orders, signs/scalings of predictor rhs and both correctors, H
  application, pressure accumulation, the constants beta / cell_area / grid_spacing / dx_factor): BIT-EXACT,
* the backward pass TF assembles from the registered gradients (op order, transposed predictor solve started from the
  forward initial guess, the periodic-axis conventions Q19/Q20): gradients within solver tolerance,
* the numpy helpers either side of the path (energy spectrum, training-sample file lists)."""
import json
import os

import numpy as np
import pytest

from common import SMALL_SETUPS, rel_l2
from oracle import adjoint as A
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python")
STEP_SETUPS = ["ldc8", "periodic16", "periodic24x20", "tml16x24", "sml16x48", "obstacle16x24"]


def load_step(name):
    return np.load(os.path.join(GOLD, "step_%s.npz" % name))


@pytest.mark.parametrize("name", STEP_SETUPS)
def test_constants_and_padding_equal_reference_python(name):
    g, s = load_step(name), SMALL_SETUPS[name]()
    ny, nx = s["ny"], s["nx"]
    n_u = ny * (nx + 1)
    assert s["dy"] == float(g["dy"]) and s["dx"] == float(g["dx"])                  # Domain.dx (fp32 box / resolution)
    assert O.cell_areas(s["dy"], s["dx"]) == tuple(float(x) for x in g["cell_area"])    # piso_tf.py:97
    assert [float(np.float32(s["dx"])), float(np.float32(s["dy"]))] == [float(x) for x in g["grid_spacing"]]
    assert O.step_constants(s["dy"], s["dx"], s["dt"])["beta"] == float(np.float32(g["beta"]))      # piso_tf.py:26
    up, vp = O.pad_velocity(ny, nx, s["per_x"], s["per_y"], g["vel"][:n_u].reshape(ny, nx + 1),
                            g["vel"][n_u:].reshape(ny + 1, nx))
    assert np.array_equal(np.concatenate([up.ravel(), vp.ravel()]), g["velocity_padded"])   # custom_padded + flatten
    rp, ci = O.csr_structure(ny, nx, s["per_x"], s["per_y"])
    assert np.array_equal(rp, g["row_ptr"]) and np.array_equal(ci, g["col_ind"])


@pytest.mark.parametrize("name", STEP_SETUPS)
def test_oracle_step_is_bit_identical_to_reference_python(name):
    """Same inputs, same native-op substitutes: every intermediate of the step equals the reference's bit for bit."""
    g, s = load_step(name), SMALL_SETUPS[name]()
    vel_next, pres_next, st, ex = O.piso_step(s, g["vel"], g["pres"], forcing=g["forcing"], full_output=True)
    for key in ("values", "a_diag", "rhs", "u_star", "div1", "p1", "u_s2", "div2", "p2"):
        assert np.array_equal(np.asarray(ex[key]).ravel(), g[key].ravel()), key
    assert np.array_equal(vel_next, g["vel_next"]) and np.array_equal(pres_next, g["pres_next"])
    assert np.array_equal(np.asarray(ex["lap"]).ravel(), g["lap1"].ravel())
    assert [st["cg1"], st["cg2"]] == g["cg_iterations"].tolist()
    assert [st["bicg_u"][0], st["bicg_v"][0]] == g["bicg_iterations"].tolist()
    assert float(g["warn"].max()) == 0.0


@pytest.mark.parametrize("name", STEP_SETUPS)
def test_oracle_adjoint_matches_reference_registered_gradients(name):
    g, s = load_step(name), SMALL_SETUPS[name]()
    # what TF's backward executes: two pressure solves (second corrector first), then the transposed predictor solve
    assert g["bwd_ops"].tolist() == ["pressure", "pressure", "bicgstab:T"]
    # ... which the reference starts from the FORWARD initial guess (linear_solver.py:164-167 passes flat_x again)
    assert np.array_equal(g["bwd_bicg_x0"], g["vel"])
    ref = A.piso_step_adjoint(s, g["vel"], g["pres"], g["w_u"], g["w_p"], forcing=g["forcing"])
    tol = 1e-5
    assert rel_l2(ref["g_vel"], g["g_vel"]) < tol
    assert rel_l2(ref["g_pres"], g["g_pres"]) < tol
    assert rel_l2(ref["g_forcing"], g["g_forcing"]) < tol
    if s["dirichlet"].any():
        # ldc8 runs the adjoint CG into its 1000-iteration cap (restart every 10): both sides stop unconverged
        assert rel_l2(ref["g_dvals"], g["g_dvals"]) < (1e-4 if name == "ldc8" else tol)
    else:
        assert not g["g_dvals"].any() and not ref["g_dvals"].any()


@pytest.mark.parametrize("tag", ["16x16", "12x20", "9x14"])
def test_energy_spectrum_matches_reference_numpy(tag):
    """diffpiso/evaluation_tools.py:92-113 executed from source vs diffpiso_b200.statistics.EK_spectrum_2D."""
    from diffpiso_b200 import statistics as S
    g = np.load(os.path.join(GOLD, "ek_spectrum_%s.npz" % tag))
    k, e = S.EK_spectrum_2D(g["field"], None)
    assert np.array_equal(k, g["k"])
    assert np.allclose(e, g["e"], rtol=1e-5, atol=1e-12)


def test_data_path_assembler_matches_reference():
    """diffpiso/datamanagement.py:35-48 executed from source vs diffpiso_b200.datamanagement.data_path_assembler."""
    from diffpiso_b200 import datamanagement as D
    j = json.load(open(os.path.join(GOLD, "data_path_assembler.json")))
    out = D.data_path_assembler(**j["args"])
    assert json.loads(json.dumps(out)) == j["out"]


def test_frame_files_round_trip(tmp_path):
    """velocity_%06d.npz / pressure_%06d.npz with `arr_0` (spatial_mixing_layer.py:60-75) through save_frame,
    load_frame and the training-sample loader of datamanagement.py:51-58."""
    from diffpiso_b200 import datamanagement as D
    rng = np.random.RandomState(0)
    d = str(tmp_path) + "/"
    frames = [(rng.randn(1, 5, 7, 2).astype(np.float32), rng.randn(1, 4, 6, 1).astype(np.float32)) for _ in range(6)]
    for i, (v, p) in enumerate(frames):
        D.save_frame(d, i, v, p)
    assert sorted(os.listdir(d))[0] == "pressure_000000.npz"
    v, p = D.load_frame(d, 3)
    assert np.array_equal(v, frames[3][0]) and np.array_equal(p, frames[3][1])
    files = D.data_path_assembler([d], ["velocity", "pressure"], [[float(i) for i in range(6)]], [0], [6], [2], dt_ratio=2)
    assert len(files[0]) == 2 and files[0][1][-1].endswith("velocity_000005.npz")
    vel, pres, ch = D.load_function(files[0][1], files[1][1], files[2][1])
    assert vel.shape == (1, 3, 5, 7, 2) and pres.shape == (1, 3, 4, 6, 1) and ch.tolist() == [1.0]
    assert np.array_equal(vel[0, 1], frames[3][0][0])
    ds = D.FrameDataset(files, rank=1, world_size=2)
    assert len(ds) == 1 and tuple(ds[0][0].shape) == (3, 5, 7, 2)


def test_mask_builders_match_reference():
    """compute_mixingLayer_masks / temporal_mixing_layer_masks / update_dirichlet_values (piso_helpers.py:58-166)."""
    import torch
    from diffpiso_b200 import masks as M
    g = np.load(os.path.join(GOLD, "masks.npz"))
    shape = tuple(int(k) for k in g["shape"])
    arr = ((g["bcy"], g["bcy"] * 2), (g["bcx"], g["bcx"] * 3))
    for k, bb in enumerate(g["mixing_cases"]):
        bb = tuple(tuple(bool(x) for x in row) for row in bb)
        for j, a in enumerate(M.compute_mixingLayer_masks(shape, bb, arr)):
            ref = g["mixing%d_%d" % (k, j)]
            assert a.shape == ref.shape and np.array_equal(a, ref), (bb, j)
    r = M.temporal_mixing_layer_masks(shape, ((True, True), (False, False)), arr)
    for j, a in enumerate((r[0], r[1], r[2][0], r[2][1], r[3], r[4])):
        assert np.array_equal(a, g["temporal_%d" % j]), j
    for k, ub in enumerate(g["update_cases"]):
        ub = tuple(tuple(bool(x) for x in row) for row in ub)
        assert np.array_equal(M.update_dirichlet_values(g["dv"], ub, arr), g["update%d" % k])
        t = M.update_dirichlet_values(torch.from_numpy(g["dv"]), ub, arr)                  # in-graph variant
        assert np.array_equal(t.numpy(), g["update%d" % k].astype(np.float32))


def _weights(g):
    import torch
    return [torch.from_numpy(g["w%d" % i]) for i in range(7)]


def test_closure_network_matches_reference():
    """diffpiso/networks.py:3-52 executed from source (tf.nn.conv2d -> torch conv) vs diffpiso_b200.networks."""
    import torch
    from diffpiso_b200 import networks as N
    g = np.load(os.path.join(GOLD, "network.npz"))
    x, w = torch.from_numpy(g["x"]), _weights(g)
    same = N.fullyconv_network(x, w, None, "SAME", False)
    valid = N.fullyconv_network(x, w, [[1, 2], [0, 3]], "VALID", True)
    assert same.shape == g["same"].shape and valid.shape == g["valid"].shape
    assert rel_l2(same.numpy(), g["same"]) < 1e-6 and rel_l2(valid.numpy(), g["valid"]) < 1e-6
    assert np.array_equal(valid.numpy() == 0, g["valid"] == 0)          # same zero frame


def test_closure_coupling_of_first_unrolled_step_matches_reference():
    """Network input (face->centre averages + central pressure gradient with the pressure's mixed extrapolation) and the
    sponge-cropping wrapper, as run_piso_steps evaluates them for step 0 (combined_training_integrated.py:399-410)."""
    import torch
    from diffpiso_b200 import networks as N, setups as SU, training as T
    from diffpiso_b200.grids import CenteredGrid, StaggeredGrid
    from test_gpu_piso_step import extrap
    g = np.load(os.path.join(GOLD, "unroll_sml16x48.npz"))
    s = SMALL_SETUPS["sml16x48"]()
    ny, nx = s["ny"], s["nx"]
    velocity = StaggeredGrid(torch.from_numpy(SU.stagger_flat(g["vel"][None], ny, nx)), dx=(s["dy"], s["dx"]))
    pressure = CenteredGrid(torch.from_numpy(g["pres"].reshape(1, ny, nx, 1)), dx=(s["dy"], s["dx"]),
                            extrapolation=extrap(s["pbc"]))
    w = _weights(g)
    nn_in = T.closure_input(velocity, pressure, True)
    sim_par = dict(dx_ratio=1, HRres=[ny, nx], sponge_ratio=0.875)
    out = T.spatial_mixing_layer_network_wrapper(lambda x: N.fullyconv_network(x, w, [[0, 0], [0, 0]], "SAME", False),
                                                 nn_in, velocity, {}, sim_par, None, [[0, 0], [0, 0]])
    assert rel_l2(out.numpy(), g["nn_out"][0]) < 1e-6
    assert not out[:, :, 42:].any()


@pytest.mark.parametrize("name,factor", [("L2_field_loss", 50), ("spectral_energy_loss", 0.5), ("strain_rate_loss", 2),
                                         ("multistep_averaging_loss", 0.5)])
def test_losses_match_reference(name, factor):
    """diffpiso/losses.py executed from source (summed and per-step variants, cropped window, sponge cut) vs
    diffpiso_b200.losses."""
    import torch
    from diffpiso_b200 import StaggeredGrid, losses as L
    g = np.load(os.path.join(GOLD, "losses.npz"))
    steps = g["fields"].shape[0]
    grids = [StaggeredGrid(torch.from_numpy(g["fields"][k]), dx=float(g["dx"])) for k in range(steps)]
    gt, bw, sponge = torch.from_numpy(g["gt"]), g["buffer_width"].tolist(), int(g["sponge_start"])
    fn = getattr(L, name)
    total, _ = fn(0, [grids], [gt], steps, bw, factor, sponge, sum_steps=True, loss_influence_range=2)
    per, _ = fn([0.0] * steps, [grids], [gt], steps, bw, factor, sponge, sum_steps=False, loss_influence_range=2)
    assert np.isclose(float(total), float(g[name + "_sum"]), rtol=2e-5)
    assert np.allclose([float(x) for x in per], g[name + "_steps"], rtol=2e-5)


@pytest.mark.parametrize("name", STEP_SETUPS)
def test_oracle_bicgstab_matches_reference_cpu_solver(name):
    """The reference's CPU solver path (LinearSolverScipy, linear_solver.py:33-57: CSR from convert_to_scipy_csr,
    piso_helpers.py:326-343, solved with spsolve) on the matrices and right-hand sides of the step, forward and
    transposed: the oracle's ILU0-BiCGStab agrees to fp32 level."""
    g, s = load_step(name), SMALL_SETUPS[name]()
    ny, nx = s["ny"], s["nx"]
    n_u = ny * (nx + 1)
    rp, ci, neg = g["row_ptr"], g["col_ind"], -g["values"]
    z_u = int(rp[n_u])
    for lo, hi, rps, zs in ((0, n_u, rp[:n_u + 1], slice(0, z_u)), (n_u, None, rp[n_u + 1:], slice(z_u, None))):
        x, st = O.bicgstab_ilu(rps, ci[zs], neg[zs], g["rhs"][lo:hi], g["vel"][lo:hi], s["bicg_tol"], s["bicg_max_it"], False)
        assert rel_l2(x, g["u_star_spsolve"][lo:hi]) < 1e-6
        xt, st = O.bicgstab_ilu(rps, ci[zs], neg[zs], g["bwd_bicg_rhs"][lo:hi], g["bwd_bicg_x0"][lo:hi], s["bicg_tol"],
                                s["bicg_max_it"], True)
        # (absolute stopping tolerance: small right-hand sides -- the obstacle case uses 1e-3 loss weights -- end with a
        #  larger relative error)
        assert rel_l2(xt, g["bicg_adj_spsolve"][lo:hi]) < 2e-5
    assert rel_l2(g["u_star"], g["u_star_spsolve"]) < 1e-6


def test_oracle_unroll_on_a_periodic_axis_reproduces_the_rewrapped_state_quirk():
    """run_piso_steps (combined_training_integrated.py:396-478) on the temporal mixing layer (periodic in x), 5 unrolled
    steps executed by the reference's own Python.  From the second step on the reference re-wraps the state with
    StaggeredGrid(array, box, extrapolation) -- the extrapolation lands in the `name` parameter -- so custom_padded
    replicates the velocity on the periodic axis while the matrix stays periodic (quirk Q21), and the increments have the
    default 'boundary' extrapolation.  The oracle with `vel_pad_periodic=(False, False)` from step 2 on reproduces every
    state; without the quirk the second state is already off by ~1e-3."""
    import torch
    from diffpiso_b200 import networks as N, setups as SU, training as T
    from diffpiso_b200.grids import CenteredGrid, StaggeredGrid
    from test_gpu_piso_step import extrap
    g = np.load(os.path.join(GOLD, "unroll_tml16x24.npz"))
    s = SMALL_SETUPS["tml16x24"]()
    ny, nx = s["ny"], s["nx"]
    steps = g["velocities"].shape[0]
    quirk = dict(s, pbc_inc=[SU.REPLICATE] * 4, vel_pad_periodic=(False, False))
    first = dict(s, pbc_inc=[SU.REPLICATE] * 4)
    w = _weights(g)
    flat = lambda t: SU.flatten_staggered(t)[0].astype(np.float32)
    for use_quirk in (True, False):
        vel, pres = g["vel"].copy(), g["pres"].copy()
        worst = 0.0
        for k in range(steps):
            velocity = StaggeredGrid(torch.from_numpy(SU.stagger_flat(vel[None], ny, nx)), dx=(s["dy"], s["dx"]))
            forcing = flat(T.closure_forcing(torch.from_numpy(g["nn_out"][k]), velocity).numpy())
            setup = first if (k == 0 or not use_quirk) else quirk
            vel, pres, _ = O.piso_step(setup, vel, pres, forcing=forcing)
            worst = max(worst, rel_l2(vel, flat(g["velocities"][k])))
        if use_quirk:
            assert worst < 2e-6, worst
        else:
            assert worst > 1e-4, worst                                   # the quirk is visible: it has to be reproduced
