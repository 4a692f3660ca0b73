"""Sweep of diffpiso_b200.SampleGroups on the bench workload: groups x {eager, graph} -> ms per fwd+adjoint step."""
import json, os, sys
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

    def fn(vel, pres, wu, wp):
        nb = vel.shape[0]
        vel = vel.detach().requires_grad_(True)
        pres = pres.detach().requires_grad_(True)
        velocity = dp.StaggeredGrid(flat=vel, resolution=(NY, NX), dx=dxy, extrapolation="periodic")
        pressure = dp.CenteredGrid(pres.reshape(nb, NY, NX, 1), dx=dxy, extrapolation="periodic")
        v_new, p_new, warn = dp.piso_step(velocity, pressure, pressure, pressure, s["dt"], sim, dvals)
        loss = (v_new.flat * wu).sum() + (p_new.data.reshape(nb, nc) * wp).sum()
        gv, gp = torch.autograd.grad(loss, (vel, pres))
        return v_new.flat.detach(), p_new.data.reshape(nb, nc).detach(), gv, gp

    vel, pres = torch.as_tensor(vel_h).to(dev), torch.as_tensor(pres_h).to(dev)
    ref = None
    cfgs = [(int(a.split(":")[0]), a.split(":")[1] == "g") for a in sys.argv[1:]] or [(1, False), (1, True), (4, False), (4, True), (8, True), (16, True)]
    for G, graph in cfgs:
        runner = dp.SampleGroups(fn, (vel, pres, w_u, w_p), groups=G, graph=graph)
        for _ in range(3):
            runner.step({0: 0, 1: 1})
        runner.join()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in runner.streams:
            st.wait_event(e0)
        n = 10
        for _ in range(n):
            runner.step({0: 0, 1: 1})
        runner.join()
        e1.record()
        torch.cuda.synchronize()
        v = runner.gather(0)
        if ref is None:
            ref = v.clone()
        print(json.dumps({"groups": G, "graph": graph, "ms_per_step": e0.elapsed_time(e1) / n,
                          "bit_identical_to_first": bool(torch.equal(v, ref)), "mem_reserved_gb": torch.cuda.memory_reserved() / 1e9}))
        del runner


main()
