"""Probe: does running sample groups of the bench batch on separate CUDA streams shorten the fwd+adjoint step?
    python scripts/split_probe.py [groups ...]       (default 1 2 4)
Same workload as bench.py (periodic 128x128, batch 64); every group is an independent piso_step + adjoint on its own
stream (no data dependency between samples), so solver launches of one group overlap the tails of the other's."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench as B  # noqa: E402


def main():
    import diffpiso_b200 as dp
    groups_list = [int(a) for a in sys.argv[1:]] or [1, 2, 4]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    s = B.setup_case()
    NY, NX, BATCH = B.NY, B.NX, B.BATCH
    nf, nc = NY * (NX + 1) + (NY + 1) * NX, NY * NX
    ls = dp.LinearSolverCudaMultiBicgstabILU(accuracy=s["bicg_tol"], max_iterations=s["bicg_max_it"])
    ps = dp.PisoPressureSolverCudaCustom(dx=s["dx"], accuracy=s["cg_tol"], max_iterations=s["cg_max_it"],
                                         residual_reset=s["cg_reset"])
    sim = dp.SimulationParameters(s["dirichlet_mask"], s["dirichlet_values_staggered"], s["active_mask"],
                                  s["accessible_mask"], bool_periodic=(True, True), no_slip_mask=s["no_slip_mask"],
                                  viscosity=float(s["visc"]), linear_solver=ls, pressure_solver=ps)
    vel_h, pres_h = B.initial_state(s, BATCH, 1234)
    dxy = (s["dy"], s["dx"])
    dvals = torch.zeros(1, nf, device=dev)
    rng = np.random.RandomState(99)
    w_u = torch.as_tensor(rng.randn(BATCH, nf).astype(np.float32)).to(dev)
    w_p = rng.randn(BATCH, nc).astype(np.float32)
    w_p = torch.as_tensor(w_p - w_p.mean(axis=1, keepdims=True)).to(dev)

    def step(vel, pres, wu, wp):
        b = vel.shape[0]
        vel = vel.detach().requires_grad_(True)
        pres = pres.detach().requires_grad_(True)
        velocity = dp.StaggeredGrid(flat=vel, resolution=(NY, NX), dx=dxy, extrapolation="periodic")
        pressure = dp.CenteredGrid(pres.reshape(b, NY, NX, 1), dx=dxy, extrapolation="periodic")
        v_new, p_new, warn = dp.piso_step(velocity, pressure, pressure, pressure, s["dt"], sim, dvals)
        loss = (v_new.flat * wu).sum() + (p_new.data.reshape(b, nc) * wp).sum()
        gv, gp = torch.autograd.grad(loss, (vel, pres))
        return v_new.flat.detach(), p_new.data.reshape(b, nc).detach(), gv, gp

    ref = None
    for G in groups_list:
        streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
        per = BATCH // G
        vel, pres = torch.as_tensor(vel_h).to(dev), torch.as_tensor(pres_h).to(dev)
        state = [(vel[i * per:(i + 1) * per].clone(), pres[i * per:(i + 1) * per].clone()) for i in range(G)]
        wus = [w_u[i * per:(i + 1) * per].contiguous() for i in range(G)]
        wps = [w_p[i * per:(i + 1) * per].contiguous() for i in range(G)]
        torch.cuda.synchronize()
        grads = [None] * G

        def run(n):
            for _ in range(n):
                for i in range(G):
                    with torch.cuda.stream(streams[i]):
                        v, p, gv, gp = step(state[i][0], state[i][1], wus[i], wps[i])
                        state[i] = (v, p)
                        grads[i] = gv
        run(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        steps = 10
        run(steps)
        for st in streams:
            e = torch.cuda.Event()
            e.record(st)
            torch.cuda.current_stream().wait_event(e)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        v_all = torch.cat([st_[0] for st_ in state])
        if ref is None:
            ref = v_all.clone()
        print(json.dumps({"groups": G, "ms_per_step": ms, "cell_updates_per_s": BATCH * nc / (ms * 1e-3),
                          "max_abs_diff_vs_first": float((v_all - ref).abs().max()),
                          "finite": bool(torch.isfinite(v_all).all())}))


if __name__ == "__main__":
    main()
