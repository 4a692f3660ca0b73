"""Golden vectors of ONE PISO step, forward and backward, produced by the reference's own Python (see
reference_runner.py for how it is executed without tensorflow).  Build container only:

    python tests/golden/make_reference_step_goldens.py

Writes tests/golden/ref_python/step_<setup>.npz: the seeded inputs, the 17 `full_output` results of
`diffpiso.piso_tf.piso_step`, the arguments its Python handed to the three native ops, and the gradients of
loss = <w_u, u_next> + <w_p, p_next> w.r.t. velocity, pressure, forcing term and Dirichlet values that the reference's
registered gradients produce."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "differentiable-piso_b200"), HERE]

from oracle import oracle as O                      # noqa: E402
import reference_runner as RR                       # noqa: E402

OUT = os.path.join(HERE, "ref_python")
# The pressure solver stops on an ABSOLUTE residual (1e-8).  With O(1) loss weights the adjoint right-hand sides are
# O(10..100) and their fp32 rounding leaves a non-zero mean of ~1e-5 over the fluid cells, which a rank-deficient system
# can never reduce below the tolerance: the reference's own adjoint CG then runs into max_iterations and returns
# garbage (seen for ldc8's Dirichlet gradient and the obstacle case).  Gradients are linear in the weights, so the
# obstacle case uses small ones.
WEIGHT_SCALE = {"obstacle16x24": 1e-3}
SETUPS = ["ldc8", "periodic16", "periodic24x20", "tml16x24", "sml16x48", "obstacle16x24"]


def _material(ns, code):
    return {0: ns["OPEN"], 1: ns["CLOSED"], 2: ns["PERIODIC"]}[code]


def reference_objects(ns, s):
    """Domain, SimulationParameters and solver objects of the reference for one of our setup dicts."""
    pbc = s["pbc"]
    boundaries = ((_material(ns, pbc[0]), _material(ns, pbc[1])), (_material(ns, pbc[2]), _material(ns, pbc[3])))
    ny, nx = s["ny"], s["nx"]
    domain = ns["Domain"]([ny, nx], box=ns["box"][0:ny * s["dy"], 0:nx * s["dx"]], boundaries=boundaries)
    ls = ns["LinearSolverCudaMultiBicgstabILU"](accuracy=s["bicg_tol"], max_iterations=s["bicg_max_it"], cast_to_double=False)
    ps = ns["PisoPressureSolverCudaCustom"](dx=[], accuracy=s["cg_tol"], max_iterations=s["cg_max_it"],
                                             residual_reset=s["cg_reset"], randomized_restarts=0,
                                             cast_to_double=s.get("cg_fp64", True))
    visc = s["visc"]
    sim = ns["SimulationParameters"](dirichlet_mask=s["dirichlet_mask"].astype(bool),
                                     dirichlet_values=s["dirichlet_values_staggered"],
                                     active_mask=s["active_mask"], accessible_mask=s["accessible_mask"],
                                     bool_periodic=(s["per_y"], s["per_x"]), no_slip_mask=s["no_slip_mask"],
                                     viscosity=float(np.atleast_1d(visc)[0]), linear_solver=ls, pressure_solver=ps)
    return domain, sim


def run_reference_step(ns, log, s, vel_flat, pres, forcing_flat, w_u, w_p):
    from diffpiso_b200 import setups as SU
    ny, nx = s["ny"], s["nx"]
    domain, sim = reference_objects(ns, s)
    vel_t = RR.tf_tensor(SU.stagger_flat(vel_flat[None], ny, nx), True)
    pres_t = RR.tf_tensor(pres.reshape(1, ny, nx, 1).copy(), True)
    force_t = RR.tf_tensor(SU.stagger_flat(forcing_flat[None], ny, nx), True)
    dvals_t = RR.tf_tensor(s["dirichlet_values_staggered"].copy(), True)
    velocity = ns["StaggeredGrid"].sample(vel_t, domain=domain)
    pressure = ns["CenteredGrid"](pres_t, box=domain.box, extrapolation=ns["pressure_extrapolation"](domain.boundaries))
    # increments as in run_piso_steps (combined_training_integrated.py:419-420); their values are ignored (Q1)
    # (there the grids get the default 'boundary' extrapolation, lid_driven_cavity_2d.py:55-56 passes the pressure's;
    # the setup's `pbc_inc` says which one the case uses)
    mode = {0: "boundary", 1: "constant", 2: "periodic"}
    pi = s["pbc_inc"]
    inc_ext = ((mode[pi[0]], mode[pi[1]]), (mode[pi[2]], mode[pi[3]]))
    inc1 = ns["CenteredGrid"](torch.zeros_like(pres_t) + 5e-13, pressure.box, extrapolation=inc_ext)
    inc2 = ns["CenteredGrid"](torch.zeros_like(pres_t) + 1e-12, pressure.box, extrapolation=inc_ext)
    visc_field = None
    if np.atleast_1d(s["visc"]).size > 1:
        visc_field = RR.tf_tensor(np.asarray(s["visc"], np.float32))
    log.calls.clear()
    out = ns["piso_step"](velocity, pressure, inc1, inc2, s["dt"], sim, dvals_t, viscosity_field=visc_field,
                          forcing_term=force_t, unrolling_step=0, full_output=True)
    v_next = out[0].staggered_tensor()
    p_next = out[1].data
    fwd_calls = list(log.calls)
    wu_t = RR.tf_tensor(SU.stagger_flat(w_u[None], ny, nx))
    wp_t = RR.tf_tensor(w_p.reshape(1, ny, nx, 1))
    loss = (v_next * wu_t).sum() + (p_next * wp_t).sum()
    loss.backward()
    bwd_calls = log.calls[len(fwd_calls):]

    def flat(t):
        return SU.flatten_staggered(t.detach().numpy())[0].astype(np.float32)

    def tens(x):
        return x.detach().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    asm = [kw for n, kw in fwd_calls if n == "assemble"][0]
    # the reference's CPU solver path (LinearSolverScipy, linear_solver.py:33-57): CSR matrices assembled by the reference's
    # convert_to_scipy_csr (piso_helpers.py:326-343) from the arrays piso_step hands to the solver (-M, [u, v] order) and
    # solved directly with scipy.sparse.linalg.spsolve -- forward system and the transposed system of the adjoint
    import scipy.sparse.linalg as spla
    mats = ns["convert_to_scipy_csr"](-tens(out[4]).astype(np.float64), tens(out[5]), tens(out[6]),
                                      np.array([1, ny + 1, nx + 1, 2]))
    n_u = ny * (nx + 1)
    rhs_np = tens(out[10]).astype(np.float64)
    adj_rhs = [kw for n, kw in bwd_calls if n == "bicgstab"][0]["rhs"].astype(np.float64)
    sp_fwd = np.concatenate([spla.spsolve(mats[0].tocsc(), rhs_np[:n_u]), spla.spsolve(mats[1].tocsc(), rhs_np[n_u:])])
    sp_adj = np.concatenate([spla.spsolve(mats[0].T.tocsc(), adj_rhs[:n_u]), spla.spsolve(mats[1].T.tocsc(), adj_rhs[n_u:])])
    adj_sol = [kw for n, kw in bwd_calls if n == "bicgstab"][0]
    res = dict(
        u_star_spsolve=sp_fwd.astype(np.float32), bicg_adj_spsolve=sp_adj.astype(np.float32),
        vel=vel_flat, pres=pres, forcing=forcing_flat, w_u=w_u, w_p=w_p,
        vel_next=flat(v_next), pres_next=tens(p_next).ravel(), p1=tens(out[2].data).ravel(), p2=tens(out[3].data).ravel(),
        values=tens(out[4]), col_ind=tens(out[5]), row_ptr=tens(out[6]), u_star=flat(out[7]), u_s2=flat(out[8]),
        a_diag=tens(out[9]), rhs=tens(out[10]), div1=tens(out[13]).ravel(), lap1=tens(out[14]), lap2=tens(out[15]),
        warn=tens(out[16]).astype(np.float32),
        velocity_padded=asm["velocity_padded"], cell_area=asm["cell_area"], grid_spacing=asm["grid_spacing"],
        beta=np.float64(asm["beta"]), dy=np.float64(domain.dx[0]), dx=np.float64(domain.dx[1]),
        pressure_extrapolation=np.array(str(pressure.extrapolation)), velocity_extrapolation=np.array(str(velocity.extrapolation)),
        k_faces=[kw for n, kw in fwd_calls if n == "pressure"][0]["k_faces"],
        div2=[kw for n, kw in fwd_calls if n == "pressure"][1]["div"].ravel(),
        cg_iterations=np.array([kw["iterations"][0] for n, kw in fwd_calls if n == "pressure"]),
        bicg_iterations=np.array([st["iterations"] for n, kw in fwd_calls if n == "bicgstab" for st in kw["stats"]]),
        g_vel=flat(vel_t.grad), g_pres=tens(pres_t.grad).ravel(), g_forcing=flat(force_t.grad), g_dvals=flat(dvals_t.grad),
        bwd_ops=np.array([n + (":T" if kw.get("transpose") else "") for n, kw in bwd_calls]),
        bwd_bicg_x0=[kw for n, kw in bwd_calls if n == "bicgstab"][0]["x0"],
        bwd_bicg_rhs=[kw for n, kw in bwd_calls if n == "bicgstab"][0]["rhs"],
        bwd_cg_rhs=np.stack([kw["div"].ravel() for n, kw in bwd_calls if n == "pressure"]),
        bwd_cg_iterations=np.array([kw["iterations"][0] for n, kw in bwd_calls if n == "pressure"]),
    )
    return res


def inputs_for(s, seed=50):
    """Same recipe as tests/test_gpu_adjoint.py::test_piso_step_backward_matches_oracle (sample 0)."""
    from common import random_fields
    ny, nx = s["ny"], s["nx"]
    nf, nc = ny * (nx + 1) + (ny + 1) * nx, ny * nx
    rng = np.random.RandomState(5)
    vel, pres = random_fields(s, seed)
    forcing = (rng.randn(2, nf) * 0.01).astype(np.float32)[0]
    w_u, w_p = rng.randn(2, nf).astype(np.float32)[0], rng.randn(2, nc).astype(np.float32)[0]
    if s["rank_deficient"]:
        act = s["active"].reshape(ny + 2, nx + 2)[1:-1, 1:-1].ravel() != 0
        w_p[~act] = 0
        w_p[act] -= w_p[act].mean()
    return vel, pres, forcing, w_u, w_p


def mask_goldens(ns):
    """compute_mixingLayer_masks / temporal_mixing_layer_masks / update_dirichlet_values (piso_helpers.py:58-166)."""
    ny, nx = 6, 9
    shape = (1, ny + 1, nx + 1, 2)
    rng = np.random.RandomState(3)
    bcy, bcx = rng.randn(1, 1, nx + 2, 1), rng.randn(1, ny + 2, 1, 1)
    dv = rng.randn(*shape).astype(np.float32)
    arr = ((bcy, bcy * 2), (bcx, bcx * 3))
    out = dict(bcy=bcy, bcx=bcx, dv=dv, shape=np.array(shape))
    cases = [((True, True), (True, False)), ((False, True), (True, True)), ((True, False), (False, False))]
    out["mixing_cases"] = np.array(cases)
    for k, bb in enumerate(cases):
        for j, a in enumerate(ns["compute_mixingLayer_masks"](shape, bb, arr)):
            out["mixing%d_%d" % (k, j)] = np.asarray(a)
    r = ns["temporal_mixing_layer_masks"](shape, ((True, True), (False, False)), arr)
    for j, a in enumerate((r[0], r[1], r[2][0], r[2][1], r[3], r[4])):
        out["temporal_%d" % j] = np.asarray(a)
    upd = [((False, False), (True, False)), ((True, True), (True, True)), ((False, True), (False, False))]
    out["update_cases"] = np.array(upd)
    for k, ub in enumerate(upd):
        out["update%d" % k] = np.asarray(ns["update_dirichlet_values"](dv, ub, arr))
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **out)


def closure_weights(seed=11, scale=0.05):
    """Small seeded closure weights in TF's HWIO layout (networks.py:57-65 shapes)."""
    rng = np.random.RandomState(seed)
    chans = [4, 16, 16, 32, 64, 64, 64, 2]
    ks = [7, 5, 5, 3, 3, 1, 1]
    return [(rng.randn(k, k, chans[i], chans[i + 1]) * scale / k).astype(np.float32) for i, k in enumerate(ks)]


def network_goldens(ns):
    """fullyconv_network (networks.py:3-52): SAME, and VALID with restore_shape and a buffer (the training default)."""
    rng = np.random.RandomState(2)
    w = closure_weights()
    x = rng.randn(2, 30, 44, 4).astype(np.float32)
    tw = [RR.tf_tensor(k) for k in w]
    same = ns["fullyconv_network"](RR.tf_tensor(x), tw, None, "SAME", False)
    valid = ns["fullyconv_network"](RR.tf_tensor(x), tw, [[1, 2], [0, 3]], "VALID", True)
    np.savez_compressed(os.path.join(OUT, "network.npz"), x=x, same=same.detach().numpy(), valid=valid.detach().numpy(),
                        **{"w%d" % i: k for i, k in enumerate(w)})


def unroll_goldens(ns, log, name="sml16x48", steps=3, influence=2):
    """run_piso_steps (combined_training_integrated.py:396-478) with the closure network, the per-step inflow
    perturbation and a stop-gradient window of 2 steps; loss = sum_s <w_s, velocity_s>; gradients w.r.t. the closure
    weights and the initial state."""
    from common import SMALL_SETUPS, random_fields
    from diffpiso_b200 import setups as SU
    s = SMALL_SETUPS[name]()
    ny, nx = s["ny"], s["nx"]
    domain, sim = reference_objects(ns, s)
    rng = np.random.RandomState(21)
    vel, pres = random_fields(s, 60)
    w = closure_weights()
    tw = [RR.tf_tensor(k, True) for k in w]
    inflow = "inlet_profile" in s                      # spatial mixing layer: per-step inflow perturbation; periodic setups: none
    bcx = s["inlet_profile"].reshape(1, ny + 2, 1, 1).astype(np.float32) if inflow else np.zeros((1, ny + 2, 1, 1), np.float32)
    bc_pert = (rng.randn(steps, 1, ny + 2, 1, 1) * (0.01 if inflow else 0.0)).astype(np.float32)
    if inflow:
        sim.dirichlet_values = ns["update_dirichlet_values"](s["dirichlet_values_staggered"], ((False, False), (True, False)),
                                                           (([], []), (bcx + bc_pert[0], [])))
    vel_t = RR.tf_tensor(SU.stagger_flat(vel[None], ny, nx), True)
    pres_t = RR.tf_tensor(pres.reshape(1, ny, nx, 1).copy(), True)
    velocity = ns["StaggeredGrid"].sample(vel_t, domain=domain)
    pressure = ns["CenteredGrid"](pres_t, box=domain.box, extrapolation=ns["pressure_extrapolation"](domain.boundaries))
    simulation_parameters = dict(dx_ratio=1, dt=s["dt"], dt_ratio=1, HRres=[ny, nx], sponge_ratio=0.875)
    training_dict = dict(step_count=steps, HR_buffer_width=[[0, 0], [0, 0]], pressure_included=True, loss_influence_range=influence)
    network = lambda x: ns["fullyconv_network"](x, tw, [[0, 0], [0, 0]], "SAME", False)
    visc_field = RR.tf_tensor(np.asarray(s["visc"], np.float32)) if np.atleast_1d(s["visc"]).size > 1 else None
    update = (lambda dv, pl: ns["update_dirichlet_values"](dv, ((False, False), (True, False)), pl)) if inflow else None
    wrapper = ns["neural_network_wrapper"] if inflow else (lambda net, x, *a: net(x))
    log.calls.clear()
    out = ns["run_piso_steps"](velocity, pressure, domain, {}, simulation_parameters, training_dict, network,
                               wrapper, sim, visc_field, bcx, RR.tf_tensor(bc_pert), update, None)
    w_loss = rng.randn(steps, 1, ny + 1, nx + 1, 2).astype(np.float32)
    loss = sum((out[7][k] * RR.tf_tensor(w_loss[k])).sum() for k in range(steps))
    loss.backward()
    res = dict(vel=vel, pres=pres, bcx=bcx, bc_pert=bc_pert, w_loss=w_loss, loss=np.float64(loss.detach().numpy()),
               velocities=np.stack([t.detach().numpy() for t in out[7]]),
               pressures=np.stack([t.detach().numpy() for t in out[8]]),
               nn_out=np.stack([t.detach().numpy() for t in out[2]]),
               g_vel=vel_t.grad.numpy(), g_pres=pres_t.grad.numpy(),
               ops=np.array([n + (":T" if kw.get("transpose") else "") for n, kw in log.calls]))
    for i, k in enumerate(w):
        res["w%d" % i] = k
        res["g_w%d" % i] = tw[i].grad.numpy()
    np.savez_compressed(os.path.join(OUT, "unroll_%s.npz" % name), **res)
    print("unroll", name, "loss", float(loss), "ops", len(log.calls), "|g_w0|", float(np.abs(res["g_w0"]).sum()))


def unroll_c3_goldens(ns, log, steps=16):
    """BASELINE configs[2] at full size: temporally evolving mixing layer 256 x 128, 16-step unrolled run_piso_steps with
    the closure network and gradients through all 16 steps (loss_influence_range = 16), executed by the reference's own
    Python.  To keep the fixture small only steps 0, 7 and 15 are stored in full, plus fp64 checksums (sum, L2 norm) of
    every step's velocity and pressure, the loss, and the gradients w.r.t. the closure weights and the initial velocity."""
    from common import random_fields
    from diffpiso_b200 import setups as SU
    s = SU.temporal_mixing_layer(ny=128, nx=256, visc=2e-3, dt=0.05)
    ny, nx = s["ny"], s["nx"]
    domain, sim = reference_objects(ns, s)
    rng = np.random.RandomState(23)
    vel, pres = random_fields(s, 61)
    w = closure_weights()
    tw = [RR.tf_tensor(k, True) for k in w]
    bcx = np.zeros((1, ny + 2, 1, 1), np.float32)
    bc_pert = np.zeros((steps, 1, ny + 2, 1, 1), np.float32)
    vel_t = RR.tf_tensor(SU.stagger_flat(vel[None], ny, nx), True)
    pres_t = RR.tf_tensor(pres.reshape(1, ny, nx, 1).copy(), True)
    velocity = ns["StaggeredGrid"].sample(vel_t, domain=domain)
    pressure = ns["CenteredGrid"](pres_t, box=domain.box, extrapolation=ns["pressure_extrapolation"](domain.boundaries))
    simulation_parameters = dict(dx_ratio=1, dt=s["dt"], dt_ratio=1, HRres=[ny, nx], sponge_ratio=0.875)
    training_dict = dict(step_count=steps, HR_buffer_width=[[0, 0], [0, 0]], pressure_included=True, loss_influence_range=steps)
    network = lambda x: ns["fullyconv_network"](x, tw, [[0, 0], [0, 0]], "SAME", False)
    wrapper = lambda net, x, *a: net(x)
    log.calls.clear()
    out = ns["run_piso_steps"](velocity, pressure, domain, {}, simulation_parameters, training_dict, network,
                               wrapper, sim, None, bcx, RR.tf_tensor(bc_pert), None, None)
    w_loss = rng.randn(1, ny + 1, nx + 1, 2).astype(np.float32)          # the same cotangent for every step
    loss = sum((out[7][k] * RR.tf_tensor(w_loss)).sum() for k in range(steps))
    loss.backward()
    vs = [t.detach().numpy() for t in out[7]]
    ps = [t.detach().numpy() for t in out[8]]
    res = dict(vel=vel, pres=pres, w_loss=w_loss, loss=np.float64(loss.detach().numpy()), keep=np.array([0, 7, 15]),
               velocities=np.stack([vs[k] for k in (0, 7, 15)]), pressures=np.stack([ps[k] for k in (0, 7, 15)]),
               vel_sum=np.array([v.astype(np.float64).sum() for v in vs]), vel_l2=np.array([np.linalg.norm(v.astype(np.float64)) for v in vs]),
               pres_l2=np.array([np.linalg.norm((p - p.mean()).astype(np.float64)) for p in ps]),
               g_vel=vel_t.grad.numpy(), ops=np.array([n + (":T" if kw.get("transpose") else "") for n, kw in log.calls]))
    for i, k in enumerate(w):
        res["g_w%d" % i] = tw[i].grad.numpy()
    np.savez_compressed(os.path.join(OUT, "unroll_c3_tml256x128.npz"), **res)
    print("unroll c3", "loss", float(loss), "ops", len(log.calls), "|g_w0|", float(np.abs(res["g_w0"]).sum()))


def loss_goldens(ns):
    """The four training objectives (diffpiso/losses.py:6-148) on seeded fields, summed and per-step variants."""
    rng = np.random.RandomState(8)
    steps, ny, nx = 4, 12, 20
    fields = rng.randn(steps, 1, ny + 1, nx + 1, 2).astype(np.float32)
    fields[:, :, -1, :, 1] = 0
    fields[:, :, :, -1, 0] = 0
    gt = (fields + 0.1 * rng.randn(*fields.shape)).astype(np.float32).transpose(1, 0, 2, 3, 4).copy()
    gt[:, :, -1, :, 1] = 0
    gt[:, :, :, -1, 0] = 0
    box = ns["box"][0:ny * 0.5, 0:nx * 0.5]
    grids = [ns["StaggeredGrid"](RR.tf_tensor(fields[k]), box) for k in range(steps)]
    gt_t = RR.tf_tensor(gt)
    bw = [[1, 2], [0, 3]]
    out = dict(fields=fields, gt=gt, buffer_width=np.array(bw), dx=np.float64(0.5), sponge_start=np.int64(17))
    for fn_name, factor in (("L2_field_loss", 50), ("spectral_energy_loss", 0.5), ("strain_rate_loss", 2),
                            ("multistep_averaging_loss", 0.5)):
        fn = ns[fn_name]
        total, contrib = fn(0, [grids], [gt_t], steps, bw, factor, 17, sum_steps=True, loss_influence_range=2)
        out[fn_name + "_sum"] = np.float64(float(total))
        per, contrib = fn([0.0] * steps, [grids], [gt_t], steps, bw, factor, 17, sum_steps=False, loss_influence_range=2)
        out[fn_name + "_steps"] = np.array([float(x) for x in per])
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)
    print("losses", {k: v for k, v in out.items() if k.endswith("_sum")})


def main():
    from common import SMALL_SETUPS
    os.makedirs(OUT, exist_ok=True)
    ns, log = RR.load_reference(O)
    if "--c3" in sys.argv:                                   # only the full-size configs[2] unroll (minutes of CPU)
        unroll_c3_goldens(ns, log)
        return
    if "--unroll-periodic" in sys.argv:                      # run_piso_steps on a periodic axis: 5 steps, gradients through all
        unroll_goldens(ns, log, name="tml16x24", steps=5, influence=5)
        return
    mask_goldens(ns)
    network_goldens(ns)
    loss_goldens(ns)
    unroll_goldens(ns, log)
    for name in SETUPS:
        s = SMALL_SETUPS[name]()
        vel, pres, forcing, w_u, w_p = inputs_for(s)
        k = np.float32(WEIGHT_SCALE.get(name, 1.0))
        res = run_reference_step(ns, log, s, vel, pres, forcing, w_u * k, w_p * k)
        np.savez_compressed(os.path.join(OUT, "step_%s.npz" % name), **res)
        print(name, "cg", res["cg_iterations"], "bicg", res["bicg_iterations"], "bwd", list(res["bwd_ops"]))


if __name__ == "__main__":
    main()
