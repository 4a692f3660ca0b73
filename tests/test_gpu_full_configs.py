"""BASELINE.json configs at their FULL sizes on the GPU.

configs[1] (periodic 128x128, batch 64), configs[2] (temporal mixing layer 256x128) and configs[3] (spatial mixing layer
512x128, Dirichlet inflow, sponge viscosity field) are small enough for the C oracle, so one full-size step is compared
sample by sample; configs[4] (periodic 1024^2 / 2048^2 solver sweep) is checked through size-independent properties:
true residuals of the three linear solves recomputed independently with torch, discrete continuity of the result,
quantised CG iteration counts, and identical samples giving identical answers."""
import numpy as np
import pytest
import torch

from common import cg_iteration_slack, random_fields, rel_l2
from oracle import oracle as O
from test_gpu_piso_step import DEV, build_sim, run_step

pytestmark = pytest.mark.gpu


def _setups():
    from diffpiso_b200 import setups as SU
    return {
        "c2_periodic128_b64": (lambda: SU.periodic_box(128, 128, visc=1e-3), 64, (0, 31, 63)),
        "c3_tml256x128_b8": (lambda: SU.temporal_mixing_layer(ny=128, nx=256, visc=2e-3, dt=0.05, bicg_tol=1e-8,
                                                              cg_tol=1e-8), 8, (0, 7)),
        "c4_sml512x128_b4": (lambda: SU.spatial_mixing_layer(ny=128, nx=512, box=(64.0, 256.0), dt=0.05), 4, (0, 3)),
    }


@pytest.mark.parametrize("name", ["c2_periodic128_b64", "c3_tml256x128_b8", "c4_sml512x128_b4"])
def test_full_size_step_matches_oracle(name):
    make, batch, check = _setups()[name]
    s = make()
    sim = build_sim(s)
    states = [random_fields(s, 300 + i) for i in range(batch)]
    vel = np.stack([v for v, _ in states])
    pres = np.stack([p for _, p in states])
    out = run_step(s, sim, vel, pres, full_output=True)
    v_new = out[0].flat.cpu().numpy()
    p_new = out[1].data.reshape(batch, -1).cpu().numpy()
    bicg = sim.linear_solver.last_stats.cpu().numpy()
    assert float(out[16].max()) == 0.0
    for i in check:
        ov, op, st, ex = O.piso_step(s, vel[i], pres[i], full_output=True)
        assert np.array_equal(out[4][i].cpu().numpy(), ex["values"])
        assert np.array_equal(out[9][i].cpu().numpy(), ex["a_diag"])
        assert np.array_equal(out[10][i].cpu().numpy(), ex["rhs"])
        assert abs(int(bicg[i, 0, 0]) - st["bicg_u"][0]) <= 1 and abs(int(bicg[i, 1, 0]) - st["bicg_v"][0]) <= 1
        assert rel_l2(v_new[i], ov) < 1e-5, (name, i, rel_l2(v_new[i], ov))
        assert rel_l2(p_new[i], op) < 1e-4, (name, i, rel_l2(p_new[i], op))
    # the last pressure solve of the step is the second corrector
    for i in check[-1:]:
        it2 = int(sim.pressure_solver.last_iterations[i])
        assert abs(it2 - st["cg2"]) <= cg_iteration_slack(s, st["cg2"]), (it2, st["cg2"])
    if s["dirichlet"].any():       # Dirichlet faces (walls, inflow profile) keep their prescribed values exactly
        d = s["dirichlet"].astype(bool)
        assert np.array_equal(v_new[:, d], np.broadcast_to(s["dirichlet_values"][d], (batch, int(d.sum()))))


def _lap_apply(lap, x, per_x, per_y):
    """L x for the 5-coefficient rows [y-, x-, c, x+, y+] in fp64 (torch), missing neighbours contribute nothing."""
    def sh(a, dy, dx):
        r = torch.roll(a, shifts=(-dy, -dx), dims=(1, 2))
        if dy == -1 and not per_y: r[:, 0, :] = 0
        if dy == 1 and not per_y: r[:, -1, :] = 0
        if dx == -1 and not per_x: r[:, :, 0] = 0
        if dx == 1 and not per_x: r[:, :, -1] = 0
        return r
    return (lap[..., 0] * sh(x, -1, 0) + lap[..., 1] * sh(x, 0, -1) + lap[..., 2] * x + lap[..., 3] * sh(x, 0, 1) +
            lap[..., 4] * sh(x, 1, 0))


@pytest.mark.parametrize("n,batch", [(1024, 2), (2048, 1)])
def test_config5_large_periodic_grid_properties(n, batch):
    """Periodic n x n, one step of `batch` copies of one seeded state + one different state."""
    from diffpiso_b200 import setups as SU
    s = SU.periodic_box(n, n, visc=1e-3)
    sim = build_sim(s)
    v0, p0 = random_fields(s, 4321)
    vel = np.stack([v0] * batch)
    pres = np.stack([p0] * batch)
    out = run_step(s, sim, vel, pres, full_output=True)
    g_nu = n * (n + 1)
    assert float(out[16].max()) == 0.0
    v_new, values, rhs = out[0].flat, out[4], out[10]
    rp, ci = out[6].long(), out[5].long()
    u_star = torch.cat([out[7][:, :-1, :, 1].reshape(batch, -1), out[7][:, :, :-1, 0].reshape(batch, -1)], dim=1)
    bicg = sim.linear_solver.last_stats.cpu().numpy()
    nnz_u = int(rp[g_nu])
    for i in range(batch):
        # predictor: (-M) u* = rhs, true fp32 residual relative to ||rhs||
        for rows, vals, cols, sl in ((rp[:g_nu + 1], values[i, :nnz_u], ci[:nnz_u], slice(0, g_nu)),
                                     (rp[g_nu + 1:], values[i, nnz_u:], ci[nnz_u:], slice(g_nu, None))):
            m = torch.sparse_csr_tensor(rows, cols, -vals.double(), size=(rows.numel() - 1, rows.numel() - 1))
            r = rhs[i, sl].double() - m @ u_star[i, sl].double()
            assert float(r.norm() / rhs[i, sl].double().norm()) < 2e-6
        assert 1 <= int(bicg[i, 0, 0]) <= 6 and 1 <= int(bicg[i, 1, 0]) <= 6
    # pressure: both solves end with the (shifted, rank-deficient) true residual at the solver tolerance
    its2 = sim.pressure_solver.last_iterations.cpu().numpy()
    assert np.all(its2 % 5 == 0) and np.all(its2 < s["cg_max_it"])
    for lap, x, b in ((out[14], out[2].data, out[13]),):
        l = lap.reshape(batch, n, n, 5)
        xx = x.reshape(batch, n, n).double()
        shift = 0.1 / (n * n) * l[..., 2].abs().sum(dim=(1, 2))
        res = b.reshape(batch, n, n).double() - (_lap_apply(l, xx, True, True) +
                                                 (shift * xx.sum(dim=(1, 2)))[:, None, None])
        assert float(res.abs().max()) < 10 * s["cg_tol"], float(res.abs().max())
    # continuity: the corrected field is discretely divergence free to the level the solves allow
    from diffpiso_b200 import ops
    geo = ops.Geometry.get(n, n, True, True, v_new.device)
    div_star = out[13].abs().max()
    div_new = ops.fv_divergence(geo, v_new, s["dy"], s["dx"]).abs().max()
    assert float(div_new) < 1e-3 * float(div_star) + 1e-7, (float(div_new), float(div_star))
    assert torch.isfinite(v_new).all() and torch.isfinite(out[1].data).all()
    for i in range(1, batch):       # identical samples, identical answers (samples never interact)
        assert torch.equal(v_new[i], v_new[0])
