"""Diagnostic: pressure CG with a fixed iteration budget against the oracle (unconverged iterates must agree to rounding)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from common import rel_l2
from diffpiso_b200 import ops, setups as SU, _native as N
from oracle import oracle as O
import test_gpu_kernels as TK
ny, nx, batch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cluster, variant = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, -1)
s = SU.periodic_box(ny, nx, visc=1e-3)
g, m, a_diag, beta, dx_factor, div = TK._cg_problem(s, 9, 1)
a_diag = np.repeat(a_diag, batch, 0); div = np.repeat(div, batch, 0)
lap = ops.laplace(g, m["active"], m["access"], TK._t(a_diag), 1, beta, dx_factor, fp64=True)
lap_h = lap.cpu().numpy()
N.lib.dpiso_pressure_cg_set_tuning(cluster, variant)
N.lib.dpiso_pressure_cg_set_reduction_order(int(os.environ.get('TWO_RED', '0')))
gm = lambda p: np.asarray(p, np.float64) - np.asarray(p, np.float64).mean()
for cap in (5, 10, 25, 50, 100, 300):
    x, its = ops.pressure_cg(g, lap, TK._t(div), 1e-8, cap, 1000, True)
    ox, oit = O.pressure_cg(ny, nx, True, True, lap_h[0].ravel(), div[0].astype(np.float64), 1e-8, cap, 1000, True)
    xs = x.cpu().numpy()
    print("grid", ny, nx, "batch", batch, "cap", cap, "its", its.cpu().tolist()[:3], oit, "rel", [float("%.3g" % rel_l2(gm(xs[i]), gm(ox))) for i in range(min(batch, 3))],
          "cfg", ops.pressure_cg_config())
