"""Micro-benchmark of the batched BiCGStab kernel on the bench workload (periodic 128x128, batch 64): launch time and
the in-kernel cycle breakdown of system 0 (dpiso_bicgstab_set_timing)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200")]
import numpy as np
import torch
from diffpiso_b200 import ops, setups as SU, _native as N

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    ny = nx = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = "cuda:0"
    s = SU.periodic_box(ny, nx, visc=1e-3)
    g = ops.Geometry.get(ny, nx, True, True, dev)
    base = SU.solenoidal_field(ny, nx, seed=100)
    vel = torch.as_tensor(np.stack([base if ny > 256 else SU.solenoidal_field(ny, nx, seed=100 + i) for i in range(B)])).to(dev)
    ones = torch.ones((ny + 2) * (nx + 2), device=dev)
    dm = torch.zeros(g.nf, dtype=torch.uint8, device=dev); ns = torch.zeros((ny + 2) * (nx + 2), dtype=torch.uint8, device=dev)
    beta = float(np.float32(s["dy"] * s["dx"] / s["dt"]))
    values, a_diag = ops.assemble(g, vel, dm, ones, ns, torch.tensor([1e-3], device=dev), s["dy"], s["dx"], beta)
    rhs = vel * beta
    neg = torch.neg(values)
    out = {}
    max_it = int(os.environ.get('BICG_MAXIT', '10000'))
    dbg = int(os.environ.get('BICG_DBG', '-1'))       # 16: per-level barrier sweeps, 32: shallow ring, 8: level-major kernel
    ref = {}
    if dbg >= 0:                                       # results of the default configuration, for a bit-for-bit comparison
        for tr in (False, True):
            ref[tr] = ops.bicgstab_ilu(g, neg, rhs, vel, 1e-8, max_it, tr)[0].clone()
    N.lib.dpiso_bicgstab_set_debug(dbg)
    N.lib.dpiso_bicgstab_set_band_cluster(int(os.environ.get('BAND_CLUSTER', '0')))
    for tr in (False, True):
        cnt = torch.zeros(8, dtype=torch.int64, device=dev)
        x, st, w = ops.bicgstab_ilu(g, neg, rhs, vel, 1e-8, max_it, tr)
        torch.cuda.synchronize()
        N.lib.dpiso_bicgstab_set_timing(cnt.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); x, st, w = ops.bicgstab_ilu(g, neg, rhs, vel, 1e-8, max_it, tr); e1.record()
        torch.cuda.synchronize()
        N.lib.dpiso_bicgstab_set_timing(None)
        c = cnt.cpu().numpy()
        out["transpose" if tr else "forward"] = {"ms": e0.elapsed_time(e1), "iterations": st.cpu().numpy()[0, :, 0].tolist(),
            "cycles_setup": int(c[0]), "cycles_ilu": int(c[1]), "cycles_sweeps": int(c[2]), "cycles_stream": int(c[4])}
        if dbg >= 0:
            out["transpose" if tr else "forward"]["bit_equal_to_default"] = bool(torch.equal(x, ref[tr]))
    print(json.dumps(out))
main()
