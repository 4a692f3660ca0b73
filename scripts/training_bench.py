"""BASELINE configs #3 / #4 as a throughput measurement of the training iteration around the PISO path: unrolled
`run_piso_steps` with the closure network (forward), the reference's four losses, backward through every PISO step
(adjoint solves), NCCL all-reduce of the 81 856 closure gradients, Adam.  Synthetic initial states and targets.

    python scripts/training_bench.py --config tml --batch 8 --unroll 16 --iters 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        scripts/training_bench.py --config tml --batch 8

Prints one JSON line on rank 0 (ms per training iteration = max over ranks, whole-job cell-updates/s, weak scaling)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "differentiable-piso_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="tml", choices=["tml", "sml"])
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU")
    ap.add_argument("--unroll", type=int, default=0, help="unrolled steps (default: 16 for tml, 10 for sml)")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    out = measure(args, dev, rank, world)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def measure(args, dev, rank, world):
    """One measurement on an initialised process group (every rank calls; rank 0 gets the dict, the others None).
    `args`: config, batch, unroll, iters, warmup."""
    dev = str(dev)
    import diffpiso_b200 as dp
    from diffpiso_b200 import losses as L, masks as M, networks as N, setups as SU, training as T
    from common import random_fields
    if args.config == "tml":
        s = SU.temporal_mixing_layer(ny=128, nx=256, visc=2e-3, dt=0.05)            # solver tolerances 1e-6 (training)
        unroll = args.unroll or 16
        ext_p = (("constant", "constant"), ("periodic", "periodic"))
        wrapper = lambda net, x, *a: net(x)
        sponge_start, padding = 0, "SAME"
    else:
        s = SU.spatial_mixing_layer(ny=128, nx=512, box=(64.0, 256.0), dt=0.05, solver_precision=1e-6)
        unroll = args.unroll or 10
        ext_p = (("boundary", "boundary"), ("boundary", "constant"))
        wrapper = T.spatial_mixing_layer_network_wrapper
        sponge_start, padding = int(512 * 0.875), "SAME"
    ny, nx = s["ny"], s["nx"]
    B = args.batch
    ls = dp.LinearSolverCudaMultiBicgstabILU(accuracy=s["bicg_tol"], max_iterations=s["bicg_max_it"])
    ps = dp.PisoPressureSolverCudaCustom(dx=s["dx"], accuracy=s["cg_tol"], max_iterations=s["cg_max_it"],
                                         residual_reset=s["cg_reset"])
    sim = dp.SimulationParameters(s["dirichlet_mask"], torch.as_tensor(s["dirichlet_values_staggered"]).to(dev),
                                  s["active_mask"], s["accessible_mask"], bool_periodic=(s["per_y"], s["per_x"]),
                                  no_slip_mask=s["no_slip_mask"], viscosity=float(np.atleast_1d(s["visc"])[0]),
                                  linear_solver=ls, pressure_solver=ps)
    visc_field = torch.as_tensor(np.asarray(s["visc"], np.float32)).to(dev) if np.atleast_1d(s["visc"]).size > 1 else None
    from diffpiso_b200 import sharding
    first_sample, _ = sharding.shard_bounds(B * world, world, rank)             # weak scaling: B samples per rank
    states = [random_fields(s, 500 + first_sample + i) for i in range(B)]
    vel0 = torch.as_tensor(SU.stagger_flat(np.stack([v for v, _ in states]), ny, nx)).to(dev)
    pres0 = torch.as_tensor(np.stack([p for _, p in states]).reshape(B, ny, nx, 1)).to(dev)
    dxy = (s["dy"], s["dx"])
    gen = torch.Generator().manual_seed(42)                     # same initial weights on every rank
    net, weights, _ = N.initialise_fullyconv_network([[0, 0], [0, 0]], padding, device=dev, generator=gen)
    opt = torch.optim.Adam(weights, lr=1e-5)
    sim_par = dict(dx_ratio=1, dt=s["dt"], dt_ratio=1, HRres=[ny, nx], sponge_ratio=0.875)
    tr = dict(step_count=unroll, HR_buffer_width=[[0, 0], [0, 0]], pressure_included=True, loss_influence_range=unroll)
    target = (vel0[:, None] + 0.01 * torch.randn(B, unroll, ny + 1, nx + 1, 2, device=dev))
    bcx = s.get("inlet_profile", np.zeros(ny + 2, np.float32)).reshape(1, ny + 2, 1, 1).astype(np.float32)
    bc_pert = torch.zeros(unroll, 1, ny + 2, 1, 1, device=dev)
    update = (lambda dv, pl: M.update_dirichlet_values(dv, ((False, False), (True, False)), pl)) if args.config == "sml" else None
    its = {"cg": [], "bicg": []}

    def loss_fn():
        velocity = dp.StaggeredGrid(vel0, dx=dxy)
        pressure = dp.CenteredGrid(pres0, dx=dxy, extrapolation=ext_p)
        out = T.run_piso_steps(velocity, pressure, velocity, {}, sim_par, tr, net, wrapper, sim, visc_field, bcx, bc_pert,
                               update, None)
        loss = 0
        for fn, factor in ((L.L2_field_loss, 50), (L.spectral_energy_loss, 0.5), (L.strain_rate_loss, 2),
                           (L.multistep_averaging_loss, 0.5)):
            loss, _ = fn(loss, [out[0]], [target], unroll, [[0, 0], [0, 0]], factor, sponge_start, sum_steps=True,
                         loss_influence_range=unroll)
        return loss / B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        T.training_iteration(opt, weights, loss_fn)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        loss = T.training_iteration(opt, weights, loss_fn)
    e1.record()
    barrier()
    ms = sharding.max_over_ranks([e0.elapsed_time(e1) / args.iters], device=dev)[0]
    # the collective alone: the flat closure-gradient bucket this workload all-reduces once per iteration
    ar_us = None
    if world > 1:
        bucket = torch.zeros(sum(w.numel() for w in weights), device=dev)
        for _ in range(3):
            dist.all_reduce(bucket)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            dist.all_reduce(bucket)
        a1.record()
        torch.cuda.synchronize()
        t_ar = torch.tensor([a0.elapsed_time(a1) / 20 * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t_ar, op=dist.ReduceOp.MAX)
        ar_us = float(t_ar)
    wsum = torch.stack([w.detach().double().sum() for w in weights]).sum()
    if world > 1:                                               # replicas must stay identical
        lo, hi = wsum.clone(), wsum.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(lo == hi)
    else:
        in_sync = True
    if rank != 0:
        return None
    cells = B * ny * nx * unroll * world
    nparam = int(sum(w.numel() for w in weights))
    return {"workload": "%s_%dx%d_unroll%d_batch%d_per_gpu_training_iteration" % (args.config, nx, ny, unroll, B),
            "n_gpus": world, "ms_per_iteration": float(ms), "cell_updates_per_s": cells / (float(ms) * 1e-3),
            "scaling": "weak", "loss": float(loss), "finite": bool(torch.isfinite(loss)),
            "replicas_in_sync": in_sync, "closure_parameters": nparam,
            "collective": {"what": "NCCL all-reduce of the flat closure-gradient bucket, once per iteration",
                           "bytes": 4 * nparam, "us_per_allreduce_alone": ar_us},
            "last_cg_iterations": float(ps.last_iterations.float().mean()),
            "conv_precision": "tf32" if torch.backends.cudnn.allow_tf32 else "fp32"}


if __name__ == "__main__":
    main()
