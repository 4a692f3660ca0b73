"""GPU parity of the backward pass: the transposed pointwise kernels against their numpy restatements (incl. the
reference's periodic-axis conventions Q19/Q20), exact-transpose dot-product tests on non-periodic grids, and the whole
`piso_step` backward (torch.autograd through the public API) against oracle/adjoint.py."""
import numpy as np
import pytest
import torch

from common import ALL_SETUPS, SMALL_SETUPS, field_tolerances, random_fields, record, rel_l2
from oracle import adjoint as A
from oracle import oracle as O
from test_gpu_piso_step import DEV, build_sim, extrap

pytestmark = pytest.mark.gpu


def _t(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dtype).to(DEV)


@pytest.mark.parametrize("name", ["ldc8", "periodic16", "periodic24x20", "tml16x24", "sml16x48", "obstacle16x24"])
def test_pointwise_adjoint_kernels(name):
    from diffpiso_b200 import ops
    s = SMALL_SETUPS[name]()
    g = ops.Geometry.get(s["ny"], s["nx"], s["per_y"], s["per_x"], DEV)
    rng = np.random.RandomState(11)
    access = _t(s["access"])
    gs = rng.randn(2, g.nf).astype(np.float32)
    gc = rng.randn(2, g.nc).astype(np.float32)
    base_c = rng.randn(2, g.nc).astype(np.float32)
    a_diag = (-rng.rand(2, g.nf)).astype(np.float32)
    beta = 3.0
    prod = np.float32(np.float64(np.float32(s["dy"])) * np.float64(np.float32(s["dx"])))
    for pbc in (s["pbc"], s["pbc_inc"]):
        out = ops.fv_gradient_adj(g, _t(gs), access, s["dy"], s["dx"], pbc).cpu().numpy()
        out2 = ops.fv_gradient_adj(g, _t(gs), access, s["dy"], s["dx"], pbc, a_diag=_t(a_diag), beta=beta,
                                   divisor=float(prod), negate=True, base=_t(base_c)).cpu().numpy()
        for i in range(2):
            ref = A.fv_gradient_adj(s["ny"], s["nx"], s["dy"], s["dx"], pbc, s["access"], gs[i])
            assert rel_l2(out[i], ref) < 1e-6
            t = -((gs[i] / (np.float32(beta) - a_diag[i])) / prod)
            ref2 = base_c[i] + A.fv_gradient_adj(s["ny"], s["nx"], s["dy"], s["dx"], pbc, s["access"], t)
            assert rel_l2(out2[i], ref2) < 1e-6
    dv = ops.fv_divergence_adj(g, _t(gc), s["dy"], s["dx"]).cpu().numpy()
    dv2 = ops.fv_divergence_adj(g, _t(gc), s["dy"], s["dx"], base=_t(gs), a_diag=_t(a_diag), beta=beta).cpu().numpy()
    for i in range(2):
        ref = A.fv_divergence_adj(s["ny"], s["nx"], s["per_x"], s["per_y"], s["dy"], s["dx"], gc[i])
        assert np.array_equal(dv[i], ref)
        assert rel_l2(dv2[i], (gs[i] + ref) / (np.float32(beta) - a_diag[i])) < 1e-6
    # H^T against scipy's transposed product on assembled matrices
    vels = np.stack([random_fields(s, 30 + i)[0] for i in range(2)])
    m = dict(dirichlet=_t(s["dirichlet"], torch.uint8), active=_t(s["active"]), noslip=_t(s["noslip"], torch.uint8))
    c = O.step_constants(s["dy"], s["dx"], s["dt"])
    values, adg = ops.assemble(g, _t(vels), m["dirichlet"], m["active"], m["noslip"], _t(np.atleast_1d(s["visc"])),
                               s["dy"], s["dx"], c["beta"])
    ht = ops.h_apply_adj(g, values, adg, _t(gs), c["beta"]).cpu().numpy()
    for i in range(2):
        ref = A.h_apply_adj(s, values[i].cpu().numpy(), adg[i].cpu().numpy(), c["beta"], gs[i])
        assert rel_l2(ht[i], ref) < 1e-6
    # forward H and H^T are transposes of each other: <H d, w> == <d, H^T w>
    d = rng.randn(2, g.nf).astype(np.float32)
    hd = ops.h_apply(g, values, adg, torch.zeros(2, g.nf, device=DEV), _t(d), c["beta"]).cpu().numpy().astype(np.float64)
    assert abs((hd * gs).sum() - (d.astype(np.float64) * ht).sum()) < 1e-4 * abs((hd * gs).sum()) + 1e-6


@pytest.mark.parametrize("name", ["ldc8", "sml16x48", "obstacle16x24"])
def test_nonperiodic_adjoints_are_exact_transposes(name):
    """<G p, s> == <p, G^T s> and <D v, g> == <v, D^T g> on grids without periodic axes."""
    from diffpiso_b200 import ops
    s = SMALL_SETUPS[name]()
    g = ops.Geometry.get(s["ny"], s["nx"], s["per_y"], s["per_x"], DEV)
    rng = np.random.RandomState(2)
    access = _t(s["access"])
    p, sf = _t(rng.randn(1, g.nc)), _t(rng.randn(1, g.nf))
    for pbc in (s["pbc"], s["pbc_inc"]):
        lhs = (ops.fv_gradient(g, p, access, s["dy"], s["dx"], pbc).double() * sf.double()).sum()
        rhs = (p.double() * ops.fv_gradient_adj(g, sf, access, s["dy"], s["dx"], pbc).double()).sum()
        assert abs(lhs - rhs) < 1e-5 * abs(lhs) + 1e-7
    v, gc = _t(rng.randn(1, g.nf)), _t(rng.randn(1, g.nc))
    lhs = (ops.fv_divergence(g, v, s["dy"], s["dx"]).double() * gc.double()).sum()
    rhs = (v.double() * ops.fv_divergence_adj(g, gc, s["dy"], s["dx"]).double()).sum()
    assert abs(lhs - rhs) < 1e-5 * abs(lhs) + 1e-7


@pytest.mark.parametrize("name", ["periodic16", "periodic24x20", "tml16x24", "sml16x48", "ldc8", "periodic64"])
def test_piso_step_backward_matches_oracle(name):
    """loss = <w_u, u_next> + <w_p, p_next>; gradients w.r.t. velocity, pressure, forcing and Dirichlet values from
    torch.autograd through piso_step against oracle/adjoint.py: north_star's 1e-5 relative L2 at the paper's 1e-8 solver
    tolerance, 1.5e-4 (2x the measured worst case, common.field_tolerances) where the setup runs the solvers at 1e-6."""
    import diffpiso_b200 as dp
    s = ALL_SETUPS[name]()
    sim = build_sim(s)
    ny, nx = s["ny"], s["nx"]
    nf, nc = ny * (nx + 1) + (ny + 1) * nx, ny * nx
    rng = np.random.RandomState(5)
    states = [random_fields(s, 50 + i) for i in range(2)]
    vel = np.stack([v for v, _ in states])
    pres = np.stack([p for _, p in states])
    forcing = (rng.randn(2, nf) * 0.01).astype(np.float32)
    w_u, w_p = rng.randn(2, nf).astype(np.float32), rng.randn(2, nc).astype(np.float32)
    if s["rank_deficient"]:      # keep the adjoint pressure right-hand sides compatible (zero mean on active cells)
        act = s["active"].reshape(ny + 2, nx + 2)[1:-1, 1:-1].ravel() != 0
        w_p[:, ~act] = 0
        w_p[:, act] -= w_p[:, act].mean(axis=1, keepdims=True)
    dxy = (s["dy"], s["dx"])
    tv = _t(vel).requires_grad_(True)
    tp = _t(pres).requires_grad_(True)
    tf = _t(forcing).requires_grad_(True)
    td = _t(s["dirichlet_values"])[None].clone().requires_grad_(True)
    velocity = dp.StaggeredGrid(flat=tv, resolution=(ny, nx), dx=dxy)
    pressure = dp.CenteredGrid(tp.reshape(2, ny, nx, 1), dx=dxy, extrapolation=extrap(s["pbc"]))
    inc = dp.CenteredGrid(torch.zeros(2, ny, nx, 1, device=DEV), dx=dxy, extrapolation=extrap(s["pbc_inc"]))
    visc_field = _t(s["visc"]) if np.atleast_1d(s["visc"]).size > 1 else None
    v_new, p_new, warn = dp.piso_step(velocity, pressure, inc, inc, s["dt"], sim, td, viscosity_field=visc_field,
                                      forcing_term=tf)
    loss = (v_new.flat * _t(w_u)).sum() + (p_new.data.reshape(2, nc) * _t(w_p)).sum()
    loss.backward()
    gd_total = np.zeros(nf, np.float32)
    for i in range(2):
        ref = A.piso_step_adjoint(s, vel[i], pres[i], w_u[i], w_p[i], forcing=forcing[i])
        record("adjoint", setup=name, sample=i, cg_tol=s["cg_tol"], g_vel=rel_l2(tv.grad[i].cpu().numpy(), ref["g_vel"]),
               g_pres=rel_l2(tp.grad[i].cpu().numpy(), ref["g_pres"]),
               g_forcing=rel_l2(tf.grad[i].cpu().numpy(), ref["g_forcing"]),
               cg_adj_it=int(sim.pressure_solver.last_iterations[i]), cg_adj_it_oracle=ref["stats"]["cg_adj"][1])
        gtol = field_tolerances(s)["grad"]       # 1e-5 at the paper's solver tolerance, 2x measured at 1e-6
        assert rel_l2(tv.grad[i].cpu().numpy(), ref["g_vel"]) < gtol, (name, i, "vel")
        assert rel_l2(tp.grad[i].cpu().numpy(), ref["g_pres"]) < gtol, (name, i, "pres")
        assert rel_l2(tf.grad[i].cpu().numpy(), ref["g_forcing"]) < gtol, (name, i, "forcing")
        gd_total += ref["g_dvals"]
        # the last pressure solve issued by backward is the first-corrector adjoint
        oit = ref["stats"]["cg_adj"][1]
        from common import cg_iteration_slack
        assert abs(int(sim.pressure_solver.last_iterations[i]) - oit) <= cg_iteration_slack(s, oit)
    if s["dirichlet"].any():
        # the Dirichlet-value gradient is the transposed predictor solution restricted to the Dirichlet rows, summed over
        # the batch: measured 5.4e-5 on the 9 x 8 cavity (restart-10 CG in the chain) -> 1.5e-4
        e_dv = rel_l2(td.grad[0].cpu().numpy(), gd_total)
        record("adjoint_dvals", setup=name, g_dvals=e_dv)
        assert e_dv < 1.5e-4, (name, e_dv)


def test_forward_and_adjoint_step_replay_as_one_cuda_graph():
    """Forward + backward of a step enqueue native kernels only (no torch arithmetic between the solver launches, no
    host<->device copy, no synchronisation): the pair is captured once as ONE CUDA graph and its replays reproduce the
    eager gradients and state bit for bit."""
    import diffpiso_b200 as dp
    s = ALL_SETUPS["periodic32"]()
    sim = build_sim(s)
    ny, nx = s["ny"], s["nx"]
    nf, nc = ny * (nx + 1) + (ny + 1) * nx, ny * nx
    b = 3
    states = [random_fields(s, 70 + i) for i in range(b)]
    vel0 = _t(np.stack([v for v, _ in states]))
    pres0 = _t(np.stack([p for _, p in states]))
    rng = np.random.RandomState(3)
    w_u = _t(rng.randn(b, nf).astype(np.float32))
    w_p0 = rng.randn(b, nc).astype(np.float32)
    w_p = _t(w_p0 - w_p0.mean(axis=1, keepdims=True))
    dvals = _t(s["dirichlet_values"])[None]
    dxy = (s["dy"], s["dx"])

    def fwd_bwd(v, p):
        v = v.detach().requires_grad_(True)
        p = p.detach().requires_grad_(True)
        velocity = dp.StaggeredGrid(flat=v, resolution=(ny, nx), dx=dxy, extrapolation="periodic")
        pressure = dp.CenteredGrid(p.reshape(b, ny, nx, 1), dx=dxy, extrapolation="periodic")
        vn, pn, _ = dp.piso_step(velocity, pressure, pressure, pressure, s["dt"], sim, dvals)
        gv, gp = torch.autograd.grad([vn.flat, pn.data.reshape(b, nc)], [v, p], [w_u, w_p])
        return vn.flat.detach(), pn.data.reshape(b, nc).detach(), gv, gp

    eager = [t.clone() for t in fwd_bwd(vel0, pres0)]
    sv, sp = vel0.clone(), pres0.clone()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fwd_bwd(sv, sp)                                    # warm-up on the capture stream (tables, scratch, allocator)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        outs = fwd_bwd(sv, sp)
    torch.cuda.synchronize()
    for o in outs:
        o.zero_()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    for got, want in zip(outs, eager):
        assert torch.isfinite(want).all() and torch.equal(got, want)


@pytest.mark.parametrize("name", ["ldc8", "sml16x48", "obstacle16x24", "tml16x24", "sml32x128"])
def test_adjoint_solve_reuses_forward_factorisation(name):
    """north_star: "the adjoint solves reuse the forward factorisation".  For structurally symmetric patterns
    ILU(0)(A^T) = (U^T D^-1)(D L^T): the transposed solve that takes the forward pivots must agree with the
    re-factorising solve (the reference: csr2csc + a fresh csrilu02, multi_bicgstab_ilu_linear_solve_op.cu.cc:113-134)
    in iteration count (+-1) and solution; components that are periodic along their staggered axis (Q18, here the u
    component of tml16x24) keep re-factorising, so they agree bit for bit."""
    from diffpiso_b200 import ops
    from oracle import oracle as O
    s = ALL_SETUPS[name]()
    ny, nx = s["ny"], s["nx"]
    g = ops.Geometry.get(ny, nx, s["per_y"], s["per_x"], DEV)
    assert ops.factor_reuse_supported(g)
    states = [random_fields(s, 80 + i) for i in range(2)]
    vel = np.stack([v for v, _ in states])
    values = []
    for i in range(2):
        _, _, _, ex = O.piso_step(s, vel[i], states[i][1], full_output=True)
        values.append(ex["values"])
    values = _t(np.stack(values))
    rng = np.random.RandomState(4)
    gbar = _t((rng.randn(2, g.nf) * 1e-2).astype(np.float32))
    rhs = _t(vel) * 0.5
    piv = torch.empty_like(rhs)
    _, st_f, _ = ops.bicgstab_ilu(g, values, rhs, _t(vel), s["bicg_tol"], s["bicg_max_it"], False, negate=True, pivots_out=piv)
    assert torch.isfinite(piv).all() and (piv != 0).all()
    x_ref, st_ref, _ = ops.bicgstab_ilu(g, values, gbar, _t(vel), s["bicg_tol"], s["bicg_max_it"], True, negate=True)
    x_reu, st_reu, _ = ops.bicgstab_ilu(g, values, gbar, _t(vel), s["bicg_tol"], s["bicg_max_it"], True, negate=True,
                                        pivots_in=piv)
    st_ref, st_reu = st_ref.cpu().numpy(), st_reu.cpu().numpy()
    assert np.all(np.abs(st_ref[:, :, 0] - st_reu[:, :, 0]) <= 1), (st_ref[:, :, 0], st_reu[:, :, 0])
    assert np.array_equal(st_ref[:, :, 1:], st_reu[:, :, 1:])
    err = rel_l2(x_reu.cpu().numpy(), x_ref.cpu().numpy())
    record("factor_reuse", setup=name, rel_l2=err, its_refactor=st_ref[:, :, 0].tolist(), its_reuse=st_reu[:, :, 0].tolist())
    assert err < 2e-5, err
    n_u = ny * (nx + 1)
    if s["per_x"]:      # u is periodic along its staggered axis: not reused -> identical arithmetic
        assert torch.equal(x_reu[:, :n_u], x_ref[:, :n_u])
    if s["per_y"]:
        assert torch.equal(x_reu[:, n_u:], x_ref[:, n_u:])
