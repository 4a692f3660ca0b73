"""Determinism stress of a predictor kernel: the same solve N times, results compared bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from common import ALL_SETUPS, random_fields
from diffpiso_b200 import ops, _native as N
name, dbg, cluster, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
s = ALL_SETUPS[name]()
dev = "cuda:0"
g = ops.Geometry.get(s["ny"], s["nx"], s["per_y"], s["per_x"], dev)
t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
vels = np.stack([random_fields(s, 50 + i)[0] for i in range(2)])
import test_gpu_kernels as TK
mm = TK._masks(s)
beta = TK._beta(s)
values, _ = ops.assemble(g, t(vels), mm["dirichlet"], mm["active"], mm["noslip"], t(np.atleast_1d(s["visc"])), s["dy"], s["dx"], beta)
rhs = t((vels * beta).astype(np.float32))
N.lib.dpiso_bicgstab_set_debug(dbg); N.lib.dpiso_bicgstab_set_band_cluster(cluster); N.lib.dpiso_bicgstab_set_tile_cluster(cluster)
for tr in (False, True):
    ref = None; bad = 0; its = set()
    for k in range(reps):
        ops.POISON_SCRATCH = (k % 2 == 0)
        x, st, w = ops.bicgstab_ilu(g, values, rhs, t(vels), s["bicg_tol"], s["bicg_max_it"], tr, negate=True)
        its.add(tuple(st[:, :, 0].flatten().tolist()))
        if ref is None: ref = x.clone()
        elif not torch.equal(x, ref): bad += 1
    print(name, "dbg", dbg, "cluster", cluster, "transpose", tr, "mismatching runs", bad, "of", reps, "iteration sets", its)
