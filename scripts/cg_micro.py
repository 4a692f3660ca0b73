"""Micro-benchmark of the pressure-CG kernel on the bench workload (periodic 128x128, batch 64, fp64, tol 1e-8).

    python scripts/cg_micro.py [--cluster C] [--variant V] [--reps R] [--batch B] [--ny NY --nx NX]

Prints one JSON line: launch time (CUDA events), iterations, time per iteration, launch configuration.  Used under
`ncu` to profile the dominant kernel and by hand to sweep the tuning knobs (dpiso_pressure_cg_set_tuning)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ny", type=int, default=128)
    ap.add_argument("--nx", type=int, default=128)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--max-it", type=int, default=10000)
    ap.add_argument("--check", type=int, default=0, help="compare the first N samples with the CPU oracle")
    ap.add_argument("--static-nx", type=int, default=1, help="0: never use the compile-time-nx kernel instantiations")
    args = ap.parse_args()
    from diffpiso_b200 import ops, setups as SU
    from diffpiso_b200 import _native as N
    dev = "cuda:0"
    s = SU.periodic_box(args.ny, args.nx, visc=1e-3)
    g = ops.Geometry.get(args.ny, args.nx, True, True, dev)
    rng = np.random.RandomState(0)
    vel = np.stack([SU.solenoidal_field(args.ny, args.nx, seed=100 + i) for i in range(args.batch)])
    vel = vel + 0.01 * rng.randn(*vel.shape).astype(np.float32)
    tv = torch.as_tensor(vel).to(dev)
    ones = torch.ones((args.ny + 2) * (args.nx + 2), device=dev)
    dm = torch.zeros(g.nf, dtype=torch.uint8, device=dev)
    ns = torch.zeros((args.ny + 2) * (args.nx + 2), dtype=torch.uint8, device=dev)
    prod = s["dy"] * s["dx"]
    beta = float(np.float32(prod / s["dt"]))
    values, a_diag = ops.assemble(g, tv, dm, ones, ns, torch.tensor([1e-3], device=dev), s["dy"], s["dx"], beta)
    div = ops.fv_divergence(g, tv, s["dy"], s["dx"])
    div = div - div.mean(dim=1, keepdim=True)
    lap = ops.laplace(g, ones, ones, a_diag, 1, beta, float(np.float32(s["dx"] / s["dy"])), fp64=not args.fp32)
    N.lib.dpiso_pressure_cg_set_tuning(args.cluster, args.variant)
    N.lib.dpiso_pressure_cg_set_static_nx(args.static_nx)
    tol = 1e-8 if not args.fp32 else 1e-5
    x, its = ops.pressure_cg(g, lap, div, tol, args.max_it, 1000, True)
    torch.cuda.synchronize()
    times = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x, its = ops.pressure_cg(g, lap, div, tol, args.max_it, 1000, True)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    its = its.cpu().numpy()
    ms = float(np.median(times))
    check = None
    if args.check:
        from oracle import oracle as O
        lap_h, div_h, x_h = lap.cpu().numpy(), div.cpu().numpy(), x.cpu().numpy()
        check = []
        for i in range(args.check):
            ox, oit = O.pressure_cg(args.ny, args.nx, True, True, lap_h[i].ravel(), div_h[i].astype(lap_h.dtype), tol,
                                    args.max_it, 1000, True)
            rel = float(np.linalg.norm(x_h[i] - ox) / np.linalg.norm(ox))
            check.append({"gpu_it": int(its[i]), "oracle_it": int(oit), "rel_l2": rel})
    print(json.dumps({"ms": ms, "mean_it": float(its.mean()), "max_it": int(its.max()), "min_it": int(its.min()),
                      "us_per_iteration_whole_batch": 1e3 * ms / float(its.mean()), "config": ops.pressure_cg_config(),
                      "batch": args.batch, "grid": [args.ny, args.nx], "finite": bool(torch.isfinite(x).all()),
                      "its_hist": np.unique(its, return_counts=True)[0].tolist(), "check": check}))


if __name__ == "__main__":
    main()
