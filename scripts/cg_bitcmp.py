"""Bitwise regression harness of the pressure CG: solve a fixed set of pressure systems (strip / general layout, one CTA /
cluster, walls / periodic, frequent resets, fp64 / fp32, both reduction orders) with the library named by DPISO_LIBRARY
(default: the product build) and write x and the iteration counts to an .npz; `--compare a.npz b.npz` reports the first
difference.  Used to prove that a kernel restructuring left every bit of the result alone.

    python scripts/cg_bitcmp.py --out gpurun_out/cg_new.npz
    DPISO_LIBRARY=.../libdpiso_old.so python scripts/cg_bitcmp.py --out gpurun_out/cg_old.npz
    python scripts/cg_bitcmp.py --compare gpurun_out/cg_old.npz gpurun_out/cg_new.npz"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402


def solve_all(out):
    import torch
    from common import ALL_SETUPS
    from test_gpu_kernels import _cg_problem, _t
    from diffpiso_b200 import _native as N, ops
    res = {}
    names = ["ldc8", "ldc32", "periodic16", "periodic24x20", "periodic32", "tml16x24", "sml16x48", "obstacle16x24",
             "periodic64", "periodic128", "periodic64x32", "tml64x128", "sml32x128", "ldc_like64"]
    for name in names:
        s = ALL_SETUPS[name]()
        g, m, a_diag, beta, dx_factor, div = _cg_problem(s, 5, 3)
        for fp64 in (True, False):
            lap = ops.laplace(g, m["active"], m["access"], _t(a_diag), 1, beta, dx_factor, fp64=fp64)
            tol = s["cg_tol"] if fp64 else 1e-5
            for two in (0, 1):
                N.lib.dpiso_pressure_cg_set_reduction_order(two)
                x, its = ops.pressure_cg(g, lap, _t(div), tol, min(s["cg_max_it"], 3000), s["cg_reset"], s["rank_deficient"])
                N.lib.dpiso_pressure_cg_set_reduction_order(0)
                key = "%s_%s_%d" % (name, "f64" if fp64 else "f32", two)
                res[key + "_x"] = x.cpu().numpy()
                res[key + "_it"] = its.cpu().numpy()
                res[key + "_cfg"] = np.array(list(ops.pressure_cg_config().values()))
    # the bench workload: periodic 128^2, 16 samples, and adjoint-like right-hand sides (mean only zero to fp32 rounding)
    import bench as B
    from diffpiso_b200 import setups as SU
    s = B.setup_case()
    vel_h, _ = B.initial_state(s, 16, 1234)
    g = ops.Geometry.get(B.NY, B.NX, True, True, "cuda:0")
    tv = torch.as_tensor(vel_h).to("cuda:0")
    ones = torch.ones((B.NY + 2) * (B.NX + 2), device="cuda:0")
    dm = torch.zeros(g.nf, dtype=torch.uint8, device="cuda:0")
    ns = torch.zeros((B.NY + 2) * (B.NX + 2), dtype=torch.uint8, device="cuda:0")
    beta = float(np.float32(s["dy"] * s["dx"] / s["dt"]))
    values, a_diag = ops.assemble(g, tv, dm, ones, ns, torch.tensor([1e-3], device="cuda:0"), s["dy"], s["dx"], beta)
    div = ops.fv_divergence(g, tv, s["dy"], s["dx"])
    lap = ops.laplace(g, ones, ones, a_diag, 1, beta, float(np.float32(s["dx"] / s["dy"])), fp64=True)
    rng = np.random.RandomState(3)
    rhs2 = rng.randn(16, B.NY * B.NX).astype(np.float32)
    rhs2 = torch.as_tensor(rhs2 - rhs2.mean(axis=1, keepdims=True)).to("cuda:0")
    for k, rhs in (("bench_div", div), ("bench_random", rhs2)):
        x, its = ops.pressure_cg(g, lap, rhs, 1e-8, 2000, 1000, True)
        res[k + "_x"], res[k + "_it"] = x.cpu().numpy(), its.cpu().numpy()
    np.savez(out, **res)
    print("wrote", out, len(res), "arrays")


def compare(a, b):
    A, Bz = np.load(a), np.load(b)
    bad = 0
    for k in A.files:
        if k.endswith("_cfg"):
            continue
        same = A[k].shape == Bz[k].shape and np.array_equal(A[k].view(np.uint32) if A[k].dtype == np.float32 else A[k],
                                                            Bz[k].view(np.uint32) if Bz[k].dtype == np.float32 else Bz[k])
        if not same:
            bad += 1
            d = np.abs(A[k].astype(np.float64) - Bz[k].astype(np.float64)).max()
            print("DIFF", k, "max abs", d, "its" if k.endswith("_it") else "", A[k].ravel()[:4], Bz[k].ravel()[:4])
    print("compared %d arrays: %d differ" % (len([k for k in A.files if not k.endswith('_cfg')]), bad))
    return bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out")
    ap.add_argument("--compare", nargs=2)
    a = ap.parse_args()
    if a.compare:
        sys.exit(1 if compare(*a.compare) else 0)
    solve_all(a.out)
