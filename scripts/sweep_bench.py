"""BASELINE config #5 (solver sweep): forward PISO steps on periodic 1024^2 / 2048^2 grids, per-solver launch times and
the algorithmic-HBM roofline of the step (SURVEY 8(d) model).  Not the driver's bench line; writes one JSON line per case.

    python scripts/sweep_bench.py [--cases 1024x8,2048x4] [--steps 2]"""
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
    ap.add_argument("--cases", default="1024x8,2048x4")
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    import diffpiso_b200 as dp
    from diffpiso_b200 import ops, setups as SU
    dev = "cuda:0"
    peak = 6460.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ev = {"cg": [], "bicg": []}

    def hook(name, fn, idx):
        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            ev[name].append((e0, e1, out[idx]))
            return out
        return wrapped
    ops.pressure_cg = hook("cg", ops.pressure_cg, 1)
    ops.bicgstab_ilu = hook("bicg", ops.bicgstab_ilu, 1)
    for case in args.cases.split(","):
        n, b = [int(k) for k in case.split("x")]
        s = SU.periodic_box(n, n, visc=1e-3)
        ls = dp.LinearSolverCudaMultiBicgstabILU(accuracy=s["bicg_tol"], max_iterations=s["bicg_max_it"])
        ps = dp.PisoPressureSolverCudaCustom(dx=s["dx"], accuracy=s["cg_tol"], max_iterations=s["cg_max_it"],
                                             residual_reset=s["cg_reset"])
        sim = dp.SimulationParameters(s["dirichlet_mask"], s["dirichlet_values_staggered"], s["active_mask"],
                                      s["accessible_mask"], bool_periodic=(True, True), no_slip_mask=s["no_slip_mask"],
                                      viscosity=float(s["visc"]), linear_solver=ls, pressure_solver=ps)
        v0 = SU.solenoidal_field(n, n, seed=4321)
        nf, nc = v0.size, n * n
        vel = torch.as_tensor(np.stack([v0] * b)).to(dev)
        pres = torch.zeros(b, nc, device=dev)
        dvals = torch.zeros(1, nf, device=dev)
        dxy = (s["dy"], s["dx"])

        def step(vel, pres):
            velocity = dp.StaggeredGrid(flat=vel, resolution=(n, n), dx=dxy, extrapolation="periodic")
            pressure = dp.CenteredGrid(pres.reshape(b, n, n, 1), dx=dxy, extrapolation="periodic")
            with torch.no_grad():
                v_new, p_new, _ = dp.piso_step(velocity, pressure, pressure, pressure, s["dt"], sim, dvals)
            return v_new.flat, p_new.data.reshape(b, nc)
        vel, pres = step(vel, pres)                       # warm-up
        torch.cuda.synchronize()
        ev["cg"].clear(); ev["bicg"].clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            vel, pres = step(vel, pres)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        cg_ms = [a.elapsed_time(c) for a, c, _ in ev["cg"]]
        cg_it = np.concatenate([it.cpu().numpy() for _, _, it in ev["cg"]]).astype(np.float64)
        bi_ms = [a.elapsed_time(c) for a, c, _ in ev["bicg"]]
        bi_it = np.concatenate([st.cpu().numpy()[:, :, 0].ravel() for _, _, st in ev["bicg"]]).astype(np.float64)
        n_cg = len(cg_ms) // args.steps
        step_bytes = b * nc * (56 + 80 + 120 + 2 * 48 + 448 * bi_it.mean() + 168 * cg_it.mean() * n_cg)
        cg_bytes = b * nc * 168 * cg_it.mean()
        print(json.dumps({
            "workload": "periodic_%dx%d_batch%d_forward" % (n, n, b), "ms_per_step": ms,
            "cell_updates_per_s": b * nc / (ms * 1e-3), "finite": bool(torch.isfinite(vel).all()),
            "pressure_cg": {"launch_ms": float(np.mean(cg_ms)), "mean_iterations": float(cg_it.mean()),
                            "us_per_iteration": 1e3 * float(np.mean(cg_ms)) / float(cg_it.mean()),
                            "algorithmic_gbs": cg_bytes / (float(np.mean(cg_ms)) * 1e-3) / 1e9,
                            "frac_of_hbm_peak": cg_bytes / (float(np.mean(cg_ms)) * 1e-3) / 1e9 / peak,
                            "config": ops.pressure_cg_config()},
            "bicgstab": {"launch_ms": float(np.mean(bi_ms)), "mean_iterations": float(bi_it.mean())},
            "step_roofline": {"algorithmic_bytes": step_bytes, "achieved_gbs": step_bytes / (ms * 1e-3) / 1e9,
                              "frac_of_hbm_peak": step_bytes / (ms * 1e-3) / 1e9 / peak, "peak_gbs": peak}}))
        del vel, pres
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
