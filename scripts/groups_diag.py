"""Diagnostics of the in-API stream groups: host enqueue time vs GPU time per fwd / fwd+adjoint step, groups 1/2/4."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200")]
import numpy as np
import torch
import bench as B


def main():
    import diffpiso_b200 as dp
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    s = B.setup_case()
    NY, NX, BATCH = B.NY, B.NX, B.BATCH
    nf, nc = NY * (NX + 1) + (NY + 1) * NX, NY * NX
    ls = dp.LinearSolverCudaMultiBicgstabILU(accuracy=s["bicg_tol"], max_iterations=s["bicg_max_it"])
    ps = dp.PisoPressureSolverCudaCustom(dx=s["dx"], accuracy=s["cg_tol"], max_iterations=s["cg_max_it"], residual_reset=s["cg_reset"])
    sim = dp.SimulationParameters(s["dirichlet_mask"], s["dirichlet_values_staggered"], s["active_mask"], s["accessible_mask"],
                                  bool_periodic=(True, True), no_slip_mask=s["no_slip_mask"], viscosity=float(s["visc"]),
                                  linear_solver=ls, pressure_solver=ps)
    vel_h, pres_h = B.initial_state(s, BATCH, 1234)
    dxy = (s["dy"], s["dx"])
    dvals = torch.zeros(1, nf, device=dev)
    rng = np.random.RandomState(99)
    w_u = torch.as_tensor(rng.randn(BATCH, nf).astype(np.float32)).to(dev)
    w_p = rng.randn(BATCH, nc).astype(np.float32)
    w_p = torch.as_tensor(w_p - w_p.mean(axis=1, keepdims=True)).to(dev)

    def step(vel, pres, bwd):
        vel = vel.detach().requires_grad_(bwd)
        pres = pres.detach().requires_grad_(bwd)
        velocity = dp.StaggeredGrid(flat=vel, resolution=(NY, NX), dx=dxy, extrapolation="periodic")
        pressure = dp.CenteredGrid(pres.reshape(BATCH, NY, NX, 1), dx=dxy, extrapolation="periodic")
        v_new, p_new, warn = dp.piso_step(velocity, pressure, pressure, pressure, s["dt"], sim, dvals)
        if bwd:
            loss = (v_new.flat * w_u).sum() + (p_new.data.reshape(BATCH, nc) * w_p).sum()
            gv, gp = torch.autograd.grad(loss, (vel, pres))
        return v_new.flat.detach(), p_new.data.reshape(BATCH, nc).detach()

    for bwd in (False, True):
        for G in (1, 2, 4):
            sim.stream_groups = G
            vel, pres = torch.as_tensor(vel_h).to(dev), torch.as_tensor(pres_h).to(dev)
            for _ in range(3):
                vel, pres = step(vel, pres, bwd)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            n = 10
            for _ in range(n):
                vel, pres = step(vel, pres, bwd)
            e1.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(json.dumps({"bwd": bwd, "groups": G, "gpu_ms_per_step": e0.elapsed_time(e1) / n,
                              "host_enqueue_ms_per_step": (t1 - t0) * 1e3 / n, "host_total_ms_per_step": (t2 - t0) * 1e3 / n,
                              "mem_reserved_gb": torch.cuda.memory_reserved() / 1e9}))


main()
