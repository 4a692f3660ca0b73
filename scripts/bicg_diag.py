"""Diagnostic: where do two predictor kernels differ (pivots and solution), run to run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from common import ALL_SETUPS, random_fields
from diffpiso_b200 import ops, _native as N
import test_gpu_kernels as TK
name, dbg, cluster = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
s = ALL_SETUPS[name]()
dev = "cuda:0"
g = ops.Geometry.get(s["ny"], s["nx"], s["per_y"], s["per_x"], dev)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
vels = np.stack([random_fields(s, 50 + i)[0] for i in range(2)])
mm = TK._masks(s); beta = TK._beta(s)
values, _ = ops.assemble(g, t(vels), mm["dirichlet"], mm["active"], mm["noslip"], t(np.atleast_1d(s["visc"])), s["dy"], s["dx"], beta)
rhs = t((vels * beta).astype(np.float32))
def run(d, c, max_it):
    N.lib.dpiso_bicgstab_set_debug(d); N.lib.dpiso_bicgstab_set_band_cluster(c)
    piv = torch.zeros(2, g.nf, device=dev)
    x, st, w = ops.bicgstab_ilu(g, values, rhs, t(vels), s["bicg_tol"], max_it, False, negate=True, pivots_out=piv)
    return x.cpu().numpy(), piv.cpu().numpy(), st.cpu().numpy()[:, :, 0]
for max_it in (3, 100):
    xr, pr, ir = run(128, 0, max_it)
    for k in range(4):
        xb, pb, ib = run(dbg, cluster, max_it)
        dp = np.argwhere(pb != pr); dxx = np.argwhere(xb != xr)
        print("max_it", max_it, "run", k, "pivot diffs", len(dp), "x diffs", len(dxx), "its", ib.tolist(), ir.tolist())
        if len(dxx):
            for smp in range(2):
                idx = dxx[dxx[:, 0] == smp][:, 1]
                u = idx[idx < g.n_u]; v = idx[idx >= g.n_u] - g.n_u
                if len(u): print("   sample", smp, "u: rows", np.unique(u // (s["nx"] + 1))[:12], "... cols min/max", (u % (s["nx"] + 1)).min(), (u % (s["nx"] + 1)).max(), "count", len(u),
                                 "max abs diff", np.abs(xb[smp, u] - xr[smp, u]).max())
                if len(v): print("   sample", smp, "v: rows", np.unique(v // s["nx"])[:12], "... cols min/max", (v % s["nx"]).min(), (v % s["nx"]).max(), "count", len(v),
                                 "max abs diff", np.abs(xb[smp, g.n_u + v] - xr[smp, g.n_u + v]).max())
