"""Diagnostic: intermediates of one forward step at large periodic sizes against the oracle (fixed CG budget)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from common import random_fields, rel_l2
from diffpiso_b200 import setups as SU
from oracle import oracle as O
import test_gpu_piso_step as TS
N_, cap = int(sys.argv[1]), int(sys.argv[2])
s = SU.periodic_box(N_, N_, visc=1e-3, cg_max_it=cap)
sim = TS.build_sim(s)
v0, p0 = random_fields(s, 4321)
out = TS.run_step(s, sim, v0[None], p0[None], full_output=True)
ov, op, st, ex = O.piso_step(s, v0, p0, full_output=True)
g_nu = s["ny"] * (s["nx"] + 1)
g = lambda p: np.asarray(p, np.float64) - np.asarray(p, np.float64).mean()
print("N", N_, "cap", cap, "oracle its", st, "gpu cg its", sim.pressure_solver.last_iterations.tolist(), "bicg", sim.linear_solver.last_stats.cpu().numpy()[0, :, 0].tolist())
print("values equal", np.array_equal(out[4][0].cpu().numpy(), ex["values"]), "rhs equal", np.array_equal(out[10][0].cpu().numpy(), ex["rhs"]))
print("u_star", rel_l2(out[7][0].cpu().numpy()[:-1, :, 1].ravel(), ex["u_star"][:g_nu]))
print("div1", rel_l2(out[13][0].cpu().numpy().ravel(), ex["div1"]), "max|div1|", np.abs(ex["div1"]).max())
print("p1", rel_l2(g(out[2].data[0].cpu().numpy().ravel()), g(ex["p1"])), "norm p1", np.linalg.norm(ex["p1"]))
print("vel", rel_l2(out[0].flat[0].cpu().numpy(), ov), "pres", rel_l2(g(out[1].data.reshape(-1).cpu().numpy()), g(op)))
# the oracle's own CG on the GPU's divergence: separates the CG kernel from its input
lap = ex["lap"]
x_o, it_o = O.pressure_cg(N_, N_, 1, 1, lap, out[13][0].cpu().numpy().ravel().astype(np.float64), s["cg_tol"], cap, s["cg_reset"], 1)
print("oracle CG on GPU div1 vs GPU p1:", rel_l2(g(out[2].data[0].cpu().numpy().ravel()), g(x_o.astype(np.float32))), "vs oracle p1:", rel_l2(g(x_o), g(ex["p1"])))
