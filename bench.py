"""bench.py -- PISO cell-updates/s (forward + adjoint) on B200, BASELINE.json's metric on configs[1]:
2-D decaying turbulence, periodic 128x128, batch 64 per GPU, fp32 predictor / fp64 pressure CG at the paper's 1e-8
tolerances (residual_reset 1000).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one PISO step of the whole batch, forward AND adjoint (loss gradient back to the step's inputs), advancing
a rollout.  The batch of a GPU runs as `--groups` independent sample pipelines (diffpiso_b200.SampleGroups: one stream and
one CUDA graph per group; bit-identical to the one-stream batch).  `value` = batch * ny * nx * n_gpus * steps / time with
inputs resident in HBM; `e2e` = the same with HOST buffers (pinned H2D of the state, D2H of the new state and gradients,
every step of every group).  A per-kernel pass (the same steps eagerly on one stream) feeds the rooflines.  The batch is
sharded over GPUs with no data-path collective (weak scaling: 64 samples per GPU).
`--impl reference` times the reference's CPU path = the oracle port (the reference has no CPU implementation of the
step and its CUDA path cannot be built here) on all host cores, bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

# one hardware queue per sample-group stream (the default of 8 is enough for the default 8 groups; more groups alias
# queues and serialise: 16 groups 8.9 ms per step with 8 queues, 8.0 ms with 32).  Read at CUDA initialisation.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "differentiable-piso_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

NY = NX = 128
BATCH = 64
WORKLOAD = "decaying_turbulence_periodic_128x128_batch64_fwd+adjoint"
# the workload both arms (--impl ours / reference) run: identical `config` in their JSON lines; what is specific to the GPU
# implementation (launch geometry, sample groups) goes into the `execution` key of our line
WORKLOAD_CONFIG = {"workload": WORKLOAD, "grid": [NY, NX], "batch_per_gpu": BATCH, "visc": 1e-3, "cfl": 0.5,
                   "bicgstab": "fp32 tol 1e-8", "pressure_cg": "fp64 tol 1e-8 reset 1000",
                   "l2": "inputs larger than L2: per-step working set (~0.4 GB of solver workspace + state) against 126 MB; "
                         "the rollout state changes every step"}
# algorithmic bytes per cell per CG iteration (fp64, 5 stored coefficients) and per reset
CG_BYTES_PER_CELL_ITER = 168
CG_BYTES_PER_CELL_RESET = 88
# essential fp64 flops per cell and CG iteration: stencil 1 mul + 4 fma = 9, inner products p.r, p.Lp, r.Lp,
# Lp.Lp = 4 fma = 8 plus the three plain sums (rank-deficiency shift) = 3, update x (fma 2), z + shift (1), r (fma 2),
# p = beta p + r (mul + add 2) = 7
CG_FLOPS_PER_CELL_ITER = 27
FP64_FMA_LANES_PER_SM = 64          # B200: 64 DFMA / clk / SM (16 per SM sub-partition), the issue roofline of the fp64 pipe


def setup_case():
    from diffpiso_b200 import setups as SU
    return SU.periodic_box(NY, NX, visc=1e-3, cfl=0.5, umax=1.0, bicg_tol=1e-8, bicg_max_it=10000, cg_tol=1e-8,
                           cg_max_it=10000, cg_reset=1000, cg_fp64=True)


def initial_state(s, batch, seed0):
    from diffpiso_b200 import setups as SU
    vel = np.stack([SU.solenoidal_field(NY, NX, seed=seed0 + i) for i in range(batch)])
    return vel.astype(np.float32), np.zeros((batch, NY * NX), np.float32)


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, steps, adjoint = args
    from oracle import adjoint as A
    from oracle import oracle as O
    s = setup_case()
    vel, pres = initial_state(s, 1, seed)
    vel, pres = vel[0], pres[0]
    rng = np.random.RandomState(seed)
    w_u = rng.randn(vel.size).astype(np.float32)
    w_p = rng.randn(pres.size).astype(np.float32)
    w_p -= w_p.mean()
    t0 = time.perf_counter()
    for _ in range(steps):
        if adjoint:
            out = A.piso_step_adjoint(s, vel, pres, w_u, w_p)      # runs the forward step inside
            vel, pres = out["vel_next"], out["pres_next"]
        else:
            vel, pres, _ = O.piso_step(s, vel, pres)
    return time.perf_counter() - t0


def usable_cores():
    """Host cores this process may actually use: scheduler affinity, capped by the cgroup CPU quota if one is set."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


def cpu_throughput(cores, steps_per_core, adjoint=True):
    """cell-updates/s of the oracle port: `cores` independent samples in parallel, `steps_per_core` steps each."""
    import multiprocessing as mp
    from oracle import oracle as O
    O.lib()     # build once before forking
    t0 = time.perf_counter()
    if cores == 1:
        _cpu_worker((1234, steps_per_core, adjoint))
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_worker, [(1234 + i, steps_per_core, adjoint) for i in range(cores)])
    dt = time.perf_counter() - t0
    return cores * steps_per_core * NY * NX / dt, dt


REF_CHUNK = 8      # fwd+adjoint steps every core advances its sample by, per reference-arm "step"


def run_reference(args):
    """Reference arm: the CPU implementation of the path (the oracle port; the reference itself has no CPU
    implementation of the step) on all usable host cores.  One "step" = every core advances its own sample of the
    128x128 case by REF_CHUNK fwd+adjoint PISO steps (a bounded sample of the 64-sample batch step); the worker pool
    is created once, outside the timed steps."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.lib()     # build once before forking
    cores = usable_cores()
    total = args.warmup + args.steps
    per_step = []
    with mp.get_context("fork").Pool(cores) as pool:
        for i in range(total):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(1234 + c + 1000 * i, REF_CHUNK, True) for c in range(cores)])
            if i >= args.warmup:
                per_step.append(time.perf_counter() - t0)
    t = sum(per_step)
    value = cores * REF_CHUNK * args.steps * NY * NX / t
    line = {"impl": "reference", "metric": "piso_cell_updates_per_s_fwd_adjoint", "value": value, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": dict(WORKLOAD_CONFIG),
            "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                             "sample": "%d samples (one per core) x %d fwd+adjoint steps of the 128x128 case per timed step, %d timed steps"
                                       % (cores, REF_CHUNK, args.steps)},
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING a timed region: an NVML thread in this process (a sample every 5 ms,
    the first one immediately; the regions last ~80 ms), or -- without the NVML bindings -- an `nvidia-smi -lms` child."""

    _REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.sm, self.mx, self.reasons = [], [], set()
        self.thread = self.p = None
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
            self.stop_flag = threading.Event()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self._sample()
            self.thread.start()
        except Exception:
            self.thread = None
            self._start_smi(index)

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        for name, bit in self._REASONS:
            if mask & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self.stop_flag.wait(0.005):
            try:
                self._sample()
            except Exception:
                break

    def _start_smi(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join()
            try:
                self._sample()
            except Exception:
                pass
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for k, nme in enumerate(names):
                if len(r) > 5 + k and "Active" in r[5 + k] and "Not" not in r[5 + k]:
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def run_config5(dev, hbm_peak, n=1024, batch=8, steps=2, cg_max_it=20000):
    """BASELINE configs[4]: periodic n x n, `batch` samples on one GPU, forward + adjoint PISO steps at the paper's solver
    tolerances.  Here the solver state does not fit on chip: the pressure CG streams its vectors through HBM
    (pressure_cg_global_kernel) -- this is the HBM-bound configuration of the path."""
    import torch
    import diffpiso_b200 as dp
    from diffpiso_b200 import ops, setups as SU
    s = SU.periodic_box(n, n, visc=1e-3, cfl=0.5, umax=1.0, bicg_tol=1e-8, bicg_max_it=10000, cg_tol=1e-8, cg_max_it=cg_max_it,
                        cg_reset=1000, cg_fp64=True)
    nf, nc = n * (n + 1) + (n + 1) * n, n * n
    ls = dp.LinearSolverCudaMultiBicgstabILU(accuracy=s["bicg_tol"], max_iterations=s["bicg_max_it"])
    ps = dp.PisoPressureSolverCudaCustom(dx=s["dx"], accuracy=s["cg_tol"], max_iterations=s["cg_max_it"],
                                         residual_reset=s["cg_reset"])
    sim = dp.SimulationParameters(s["dirichlet_mask"], s["dirichlet_values_staggered"], s["active_mask"],
                                  s["accessible_mask"], bool_periodic=(True, True), no_slip_mask=s["no_slip_mask"],
                                  viscosity=float(s["visc"]), linear_solver=ls, pressure_solver=ps)
    base = SU.solenoidal_field(n, n, seed=4321)
    rng = np.random.RandomState(7)
    vel = torch.as_tensor(np.stack([base * (1.0 + 0.05 * i) for i in range(batch)]).astype(np.float32)).to(dev)
    pres = torch.zeros(batch, nc, device=dev)
    dvals = torch.zeros(1, nf, device=dev)
    # smooth adjoint seeds (a loss on the large scales), zero-mean for the pressure
    w_u = torch.as_tensor(np.stack([SU.solenoidal_field(n, n, seed=99 + i) for i in range(batch)]).astype(np.float32)).to(dev)
    w_p = torch.zeros(batch, nc, device=dev)
    dxy = (s["dy"], s["dx"])
    ev = {"cg": [], "bicg": []}
    orig_cg, orig_bicg = ops.pressure_cg, ops.bicgstab_ilu

    def timed(kind, fn):
        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            ev[kind].append((e0, e1, out[1]))
            return out
        return wrapped
    ops.pressure_cg, ops.bicgstab_ilu = timed("cg", orig_cg), timed("bicg", orig_bicg)
    try:
        def step(v, p):
            v = v.detach().requires_grad_(True)
            p = p.detach().requires_grad_(True)
            velocity = dp.StaggeredGrid(flat=v, resolution=(n, n), dx=dxy, extrapolation="periodic")
            pressure = dp.CenteredGrid(p.reshape(batch, n, n, 1), dx=dxy, extrapolation="periodic")
            vn, pn, _ = dp.piso_step(velocity, pressure, pressure, pressure, s["dt"], sim, dvals)
            gv, gp = torch.autograd.grad([vn.flat, pn.data.reshape(batch, nc)], [v, p], [w_u, w_p])
            return vn.flat.detach(), pn.data.reshape(batch, nc).detach(), gv
        vel, pres, gv = step(vel, pres)                       # warm-up (tables, workspaces)
        torch.cuda.synchronize()
        ev["cg"].clear(); ev["bicg"].clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            vel, pres, gv = step(vel, pres)
        e1.record()
        torch.cuda.synchronize()
    finally:
        ops.pressure_cg, ops.bicgstab_ilu = orig_cg, orig_bicg
    ms = e0.elapsed_time(e1) / steps
    cg_ms = [a.elapsed_time(b) for a, b, _ in ev["cg"]]
    cg_it = [float(it.max()) for _, _, it in ev["cg"]]       # a launch lasts as long as its slowest sample
    bicg_ms = [a.elapsed_time(b) for a, b, _ in ev["bicg"]]
    us_per_it = 1e3 * sum(cg_ms) / max(sum(cg_it), 1.0)
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_cg_global.json")))
    except Exception:
        prof = {}
    model_gbs = batch * nc * CG_BYTES_PER_CELL_ITER / (us_per_it * 1e-6) / 1e9
    meas = prof.get("dram_bytes_per_cell_iteration")
    return {"workload": "periodic_%dx%d_batch%d_fwd+adjoint" % (n, n, batch), "steps": steps, "ms_per_step": ms,
            "cell_updates_per_s": batch * nc / (ms * 1e-3), "finite": bool(torch.isfinite(gv).all()),
            "pressure_cg": {"launches_per_step": len(cg_ms) // steps, "ms_per_launch": float(np.mean(cg_ms)),
                            "max_iterations_per_launch": cg_it, "us_per_iteration_whole_batch": us_per_it,
                            "share_of_step": sum(cg_ms) / (ms * steps), "launch": ops.pressure_cg_config(),
                            "hbm_model": {"bytes_per_cell_iteration": CG_BYTES_PER_CELL_ITER, "achieved_gbs": model_gbs,
                                          "ratio_to_hbm_peak": model_gbs / hbm_peak},
                            "hbm_measured": {"bytes_per_cell_iteration": meas,
                                             "achieved_gbs": (batch * nc * meas / (us_per_it * 1e-6) / 1e9) if meas else None,
                                             "frac": (batch * nc * meas / (us_per_it * 1e-6) / 1e9 / hbm_peak) if meas else None,
                                             "source": prof.get("source")}},
            "bicgstab": {"launches_per_step": len(bicg_ms) // steps, "ms_per_launch": float(np.mean(bicg_ms)),
                         "iterations": [int(v) for v in ev["bicg"][0][2].cpu().numpy()[0, :, 0]] if ev["bicg"] else None,
                         "share_of_step": sum(bicg_ms) / (ms * steps)}}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import diffpiso_b200 as dp
    from diffpiso_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    s = setup_case()
    nf, nc = NY * (NX + 1) + (NY + 1) * NX, NY * NX
    ls = dp.LinearSolverCudaMultiBicgstabILU(accuracy=s["bicg_tol"], max_iterations=s["bicg_max_it"])
    ps = dp.PisoPressureSolverCudaCustom(dx=s["dx"], accuracy=s["cg_tol"], max_iterations=s["cg_max_it"],
                                         residual_reset=s["cg_reset"])
    sim = dp.SimulationParameters(s["dirichlet_mask"], s["dirichlet_values_staggered"], s["active_mask"],
                                  s["accessible_mask"], bool_periodic=(True, True), no_slip_mask=s["no_slip_mask"],
                                  viscosity=float(s["visc"]), linear_solver=ls, pressure_solver=ps,
                                  stream_groups=args.stream_groups)
    # --same-seeds: every rank solves the same samples, which separates host launch jitter from iteration imbalance
    from diffpiso_b200 import sharding
    # weak scaling: the job has BATCH * world samples, rank r owns the contiguous block shard_bounds() gives it
    first_sample, n_local = sharding.shard_bounds(BATCH * world, world, rank)
    assert n_local == BATCH
    seed_rank = 0 if args.same_seeds else rank
    vel_h, pres_h = initial_state(s, BATCH, 1234 + (0 if args.same_seeds else first_sample))
    dxy = (s["dy"], s["dx"])
    dvals = torch.zeros(1, nf, device=dev)
    rng = np.random.RandomState(99 + seed_rank)
    w_u = torch.as_tensor(rng.randn(BATCH, nf).astype(np.float32)).to(dev)
    w_p = rng.randn(BATCH, nc).astype(np.float32)
    w_p = torch.as_tensor(w_p - w_p.mean(axis=1, keepdims=True)).to(dev)

    def group_step(vel, pres, w_u, w_p):
        """forward + adjoint of one PISO step for a block of samples -> new state and input gradients (what the sample
        groups run; the loss weights are per sample, so the block's gradients are those of the whole-batch loss)"""
        nb = vel.shape[0]
        vel = vel.detach().requires_grad_(True)
        pres = pres.detach().requires_grad_(True)
        velocity = dp.StaggeredGrid(flat=vel, resolution=(NY, NX), dx=dxy, extrapolation="periodic")
        pressure = dp.CenteredGrid(pres.reshape(nb, NY, NX, 1), dx=dxy, extrapolation="periodic")
        v_new, p_new, warn = dp.piso_step(velocity, pressure, pressure, pressure, s["dt"], sim, dvals)
        # gradient of the loss sum(w_u * velocity) + sum(w_p * pressure) with respect to the step's inputs, as the
        # vector-Jacobian product with the loss weights (no elementwise / reduction kernels around the native ones)
        gv, gp = torch.autograd.grad([v_new.flat, p_new.data.reshape(nb, nc)], [vel, pres], [w_u, w_p])
        return v_new.flat.detach(), p_new.data.reshape(nb, nc).detach(), gv, gp

    def step(vel, pres):
        return group_step(vel, pres, w_u, w_p)

    # launch counting / per-kernel event timing hooks (eager passes only: a graph replay does not come through Python)
    launches = {"n": 0}
    cg_events = []
    orig_check = ops.N.check

    def counting_check(rc, what):
        launches["n"] += 1
        return orig_check(rc, what)
    ops.N.check = counting_check
    orig_cg = ops.pressure_cg

    def timed_cg(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_cg(*a, **k)
        e1.record()
        cg_events.append((e0, e1, out[1]))
        return out
    bicg_events = []
    orig_bicg = ops.bicgstab_ilu

    def timed_bicg(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_bicg(*a, **k)
        e1.record()
        bicg_events.append((e0, e1, out[1]))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    vel, pres = torch.as_tensor(vel_h).to(dev), torch.as_tensor(pres_h).to(dev)
    # ---- headline: the batch as `groups` independent sample pipelines (diffpiso_b200.SampleGroups), each captured once
    # as a CUDA graph; a step = one launch of every group, the new state fed back as the next step's input -------------
    groups = max(1, min(args.groups, BATCH))
    def make_runner(fn, inputs):
        """Sample groups with one CUDA graph each; if capture fails on this box the same groups run eagerly (recorded)."""
        if args.graph:
            try:
                return dp.SampleGroups(fn, inputs, groups=groups, graph=True)
            except Exception as e:                               # never lose the bench line to a capture problem
                sys.stderr.write("CUDA graph capture failed (%r): running the sample groups eagerly\n" % (e,))
                torch.cuda.synchronize()
                args.graph = False
        return dp.SampleGroups(fn, inputs, groups=groups, graph=False)
    runner = make_runner(group_step, (vel, pres, w_u, w_p))
    feedback = {0: 0, 1: 1}
    for _ in range(args.warmup):
        runner.step(feedback)
    runner.join()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st_ in runner.streams:
        st_.wait_event(e0)
    for _ in range(args.steps):
        runner.step(feedback)
    runner.join()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    vel, pres, gv = runner.gather(0), runner.gather(1), runner.gather(2)
    finite = bool(torch.isfinite(vel).all() and torch.isfinite(gv).all())

    # ---- per-kernel pass: the same steps, eagerly, with the whole batch on ONE stream, so that every solver launch runs
    # alone and its CUDA-event duration is the kernel's own (in the headline region the launches of the sample groups
    # overlap each other on purpose); the rooflines below are computed from these durations and say so.  The native
    # launches of one eager step are counted here: every group's graph holds the same kernel sequence.
    ops.pressure_cg, ops.bicgstab_ilu = timed_cg, timed_bicg
    vs_, ps_ = vel.clone(), pres.clone()
    for _ in range(2):
        vs_, ps_, _, _ = step(vs_, ps_)
    cg_events.clear()
    bicg_events.clear()
    launches["n"] = 0
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        vs_, ps_, _, _ = step(vs_, ps_)
    k1.record()
    barrier()
    ms_serial = k0.elapsed_time(k1)
    n_launch = launches["n"] * groups                              # launches inside the headline region (all groups)
    cfg = ops.pressure_cg_config()                                 # launch geometry of the headline grid
    ops.pressure_cg, ops.bicgstab_ilu = orig_cg, orig_bicg
    cg_ms = [a.elapsed_time(b) for a, b, _ in cg_events]
    cg_its = np.concatenate([it.cpu().numpy() for _, _, it in cg_events]).astype(np.float64)
    bicg_ms = [a.elapsed_time(b) for a, b, _ in bicg_events]
    bicg_its = np.concatenate([st.cpu().numpy()[:, :, 0].ravel() for _, _, st in bicg_events]).astype(np.float64)

    # ---- forward-only rollout (reported beside the headline): the same sample groups, forward step only -------------------
    def group_forward(vel, pres):
        nb = vel.shape[0]
        with torch.no_grad():
            velocity = dp.StaggeredGrid(flat=vel, resolution=(NY, NX), dx=dxy, extrapolation="periodic")
            pressure = dp.CenteredGrid(pres.reshape(nb, NY, NX, 1), dx=dxy, extrapolation="periodic")
            v_new, p_new, _ = dp.piso_step(velocity, pressure, pressure, pressure, s["dt"], sim, dvals)
        return v_new.flat, p_new.data.reshape(nb, nc)
    fwd_runner = make_runner(group_forward, (vel, pres))
    for _ in range(args.warmup):
        fwd_runner.step(feedback)
    fwd_runner.join()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for st_ in fwd_runner.streams:
        st_.wait_event(f0)
    for _ in range(args.steps):
        fwd_runner.step(feedback)
    fwd_runner.join()
    f1.record()
    barrier()
    ms_fwd = f0.elapsed_time(f1)
    finite = finite and bool(torch.isfinite(fwd_runner.gather(0)).all())
    del fwd_runner
    import gc
    gc.collect()                                         # graph pools are released here, not inside the next timed loop
    # the same forward-only steps eagerly on one stream
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    vf, pf = vel.clone(), pres.clone()
    vf, pf = group_forward(vf, pf)                       # untimed: allocator warm-up of this loop
    barrier()
    f0.record()
    for _ in range(args.steps):
        vf, pf = group_forward(vf, pf)
    f1.record()
    barrier()
    ms_fwd_serial = f0.elapsed_time(f1)

    # ---- end-to-end: host buffers in, host buffers out, every step of every group ---------------------------------------
    # The same runner, driven with HOST buffers: every step a group uploads its state from pinned host memory, runs forward
    # + adjoint, and downloads the new state and the gradients into pinned host memory; the host waits for the group's
    # download before it hands the step's output buffers back as the next step's input buffers (pointer swap, no host
    # memcpy).  Groups are independent, so one group's copies overlap the others' kernels.  Every copy of every step is
    # inside the timed region.
    h_in = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in runner.inputs(i)[:2]] for i in range(groups)]
    h_out = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in runner.outputs(i)] for i in range(groups)]
    for i in range(groups):
        for h, d in zip(h_in[i], runner.outputs(i)[:2]):
            h.copy_(d)
    torch.cuda.synchronize()

    def e2e_steps(n):
        """n steps of every group.  A group's next step is issued as soon as the host has its previous results (state +
        gradients downloaded: the group's `done` event), whichever group that is -- the host polls the events instead of
        waiting for the groups in a fixed order."""
        left = [n] * groups
        done = [None] * groups
        while any(left) or any(d is not None for d in done):
            for i in range(groups):
                if done[i] is not None:
                    if not done[i].query():
                        continue
                    done[i] = None                   # the host owns group i's results from here on
                if left[i]:
                    runner.load(i, h_in[i][0], h_in[i][1], None, None)
                    runner.launch(i)
                    runner.fetch(i, *h_out[i])
                    done[i] = torch.cuda.Event()
                    done[i].record(runner.streams[i])
                    h_in[i][0], h_out[i][0] = h_out[i][0], h_in[i][0]       # next step reads what this step writes
                    h_in[i][1], h_out[i][1] = h_out[i][1], h_in[i][1]
                    left[i] -= 1
    e2e_steps(2)                                         # untimed: warm-up of this loop
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for st_ in runner.streams:
        st_.wait_event(g0)
    e2e_steps(args.steps)
    runner.join()
    g1.record()
    barrier()
    ms_e2e = g0.elapsed_time(g1)
    ops.N.check = orig_check
    times = torch.tensor([ms, ms_fwd, ms_e2e, ms_serial], dtype=torch.float64, device=dev)
    # per-rank diagnostics of the device-resident loop: step time, sum of CG iterations, mean of the per-launch maxima
    # (a launch lasts as long as its slowest sample), time inside the CG / BiCGStab launches
    launch_max = [float(it.max()) for _, _, it in cg_events[:len(cg_ms)]]
    mine = torch.tensor([ms / args.steps, float(cg_its.sum()), float(np.mean(launch_max)), float(sum(cg_ms)) / args.steps,
                         float(sum(bicg_ms)) / args.steps], dtype=torch.float64, device=dev)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    per_rank = [[float(v) for v in t_.cpu()] for t_ in per_rank]
    ms, ms_fwd, ms_e2e, ms_serial_all = sharding.max_over_ranks([float(x) for x in times.cpu()], device=dev)   # slowest rank
    # BASELINE configs[2]: one training iteration around the path (16-step unroll with the closure network, backward through
    # every step, NCCL all-reduce of the closure gradients -- the only collective of the workload -- and Adam) at N GPUs
    training = None
    if args.training:
        import types
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import training_bench
        try:
            training = training_bench.measure(types.SimpleNamespace(config="tml", batch=8, unroll=16, iters=1, warmup=1),
                                              dev, rank, world)
        except Exception as e:                                   # never lose the headline line to the extra key
            training = {"error": repr(e)[:200]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cells = BATCH * nc * world
    value = cells * args.steps / (ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def prof(name):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            return {}
    cg_prof, bicg_prof = prof("cg_dram_traffic.json"), prof("bicg_dram_traffic.json")
    # ---- dominant kernel: the cluster-resident pressure CG (rank 0's launches) -------------------------------------
    mean_it = float(cg_its.mean())
    cg_avg_ms = float(np.mean(cg_ms))
    cell_its = float(BATCH * nc) * mean_it                         # cell-iterations per launch
    # (i) what binds it: fp64 issue.  x, r, p, z and the matrix stay in registers / shared memory for the whole solve, so
    #     HBM only carries the compulsory read of (lap, div) and the write of x; the roofline is the fp64 pipe.
    props = torch.cuda.get_device_properties(dev)
    sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
    fp64_peak = props.multi_processor_count * FP64_FMA_LANES_PER_SM * 2 * sm_hz / 1e12
    fp64_achieved = CG_FLOPS_PER_CELL_ITER * cell_its / (cg_avg_ms * 1e-3) / 1e12
    # (ii) the HBM model of SURVEY 8(d) (what a streaming implementation such as the reference moves), for comparison
    resets = math.floor((mean_it + 1) / s["cg_reset"])
    model_bytes = BATCH * nc * (CG_BYTES_PER_CELL_ITER * mean_it + CG_BYTES_PER_CELL_RESET * resets)
    model_gbs = model_bytes / (cg_avg_ms * 1e-3) / 1e9
    traffic = cg_prof.get("bytes_per_launch")
    roofline = {
        "bound": "fp64",
        "kernel": "pressure_cg_kernel<double,float,%d threads,%d cells/thread> cluster %d (4 launches per fwd+adjoint step)"
                  % (cfg["threads"], cfg["cells_per_thread"], cfg["cluster"]),
        "achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
        "peak_source": "%d SMs x %d DFMA lanes x 2 flop x %.0f MHz (SM clock sampled during the timed region)"
                       % (props.multi_processor_count, FP64_FMA_LANES_PER_SM, sm_hz / 1e6),
        "flops_per_cell_iteration": CG_FLOPS_PER_CELL_ITER, "mean_cg_iterations": mean_it,
        "cg_iterations_min_max": [float(cg_its.min()), float(cg_its.max())],
        "mean_of_per_launch_max_iterations": float(np.mean(launch_max)),
        "avg_launch_ms": cg_avg_ms, "cg_share_of_step": float(sum(cg_ms) / ms_serial),
        "timed": "per-kernel pass of this run: %d steps with the whole batch on one stream (%.3f ms per step), each launch "
                 "bracketed by CUDA events; in the headline region the launches of %d sample groups overlap"
                 % (args.steps, ms_serial / args.steps, groups),
        "traffic": traffic,
        "hbm_actual": {"achieved": (traffic / (cg_avg_ms * 1e-3) / 1e9) if traffic else None, "peak": peak, "unit": "GB/s",
                       "frac": (traffic / (cg_avg_ms * 1e-3) / 1e9 / peak) if traffic else None,
                       "note": "ncu dram__bytes_read+write per launch (profiles/cg_dram_traffic.json): the compulsory "
                               "read of lap + div and write of x"},
        "hbm_model": {"achieved": model_gbs, "peak": peak, "unit": "GB/s", "ratio_to_hbm_peak": model_gbs / peak,
                      "algorithmic_bytes_per_launch": model_bytes, "peak_source": peak_src,
                      "note": "SURVEY 8(d) streaming model (168 B per cell and iteration): what an implementation that "
                              "keeps the vectors in HBM would move; > 1 because this kernel does not stream them, so it is "
                              "NOT the binding roofline here"},
        "ncu": {k: cg_prof.get(k) for k in ("fp64_pipe_pct", "issue_active_pct", "warps_active_pct", "source") if k in cg_prof},
        "note": "frac = essential fp64 flops (27 per cell and iteration) / launch time / fp64 peak; the kernel issues more "
                "than the essential instructions (index arithmetic, conversions, reductions), see profiles/",
    }
    # ---- second kernel: BiCGStab + ILU(0), HBM-nominal (setup 36 + ILU 40 + 224 per iteration, bytes per row)
    bicg_it = float(bicg_its.mean())
    bicg_bytes = BATCH * nf * (36.0 + 40.0 + 224.0 * bicg_it)
    bicg_avg_ms = float(np.mean(bicg_ms))
    bicg_gbs = bicg_bytes / (bicg_avg_ms * 1e-3) / 1e9
    # ---- whole step against the step model of SURVEY 8(d) --------------------------------------------------------------
    n_cg = len(cg_ms) // args.steps
    step_bytes = BATCH * nc * (56 + 2 * 80 + 2 * 120 + 4 * 48 + 448 * 2 * bicg_it + 168 * mean_it * n_cg)
    step_achieved = step_bytes / (ms / args.steps * 1e-3) / 1e9
    line = {
        "metric": "piso_cell_updates_per_s_fwd_adjoint", "value": value, "unit": "cell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": dict(WORKLOAD_CONFIG),
        "execution": {"cg_launch": cfg, "same_seeds": bool(args.same_seeds), "sample_groups": groups,
                      "cuda_graph_per_group": bool(args.graph)},
        "single_stream": {"ms_per_step": ms_serial_all / args.steps, "value": cells * args.steps / (ms_serial_all * 1e-3),
                          "note": "the same steps run eagerly with the whole batch on one stream (the per-kernel pass the "
                                  "rooflines are taken from)"},
        "forward_only": {"value": cells * args.steps / (ms_fwd * 1e-3), "unit": "cell-updates/s",
                         "ms_per_step": ms_fwd / args.steps, "single_stream_ms_per_step": ms_fwd_serial / args.steps},
        "e2e": {"value": cells * args.steps / (ms_e2e * 1e-3), "unit": "cell-updates/s",
                "h2d_bytes_per_step": int(BATCH * (nf + nc) * 4), "d2h_bytes_per_step": int(2 * BATCH * (nf + nc) * 4)},
        "gpu_launches": n_launch,
        "roofline": roofline,
        "roofline_bicgstab": {"bound": "hbm", "kernel": "bicgstab_rows_kernel (one 512-thread CTA per system, 2 launches per step)",
                              "achieved": bicg_gbs, "peak": peak, "unit": "GB/s", "frac": bicg_gbs / peak,
                              "traffic": bicg_prof.get("bytes_per_launch"), "mean_iterations": bicg_it,
                              "avg_launch_ms": bicg_avg_ms, "share_of_step": float(sum(bicg_ms) / ms_serial),
                              "note": "algorithmic bytes (SURVEY 8(d)) / launch time; latency-bound: 13 triangular sweeps x "
                                      "~260 dependent wavefront levels per solve, one CTA per system"},
        "roofline_step": {"bound": "hbm", "achieved": step_achieved, "peak": peak, "unit": "GB/s",
                          "ratio_to_hbm_peak": step_achieved / peak, "algorithmic_bytes_per_step": step_bytes,
                          "note": "SURVEY 8(d) streaming step model with the iteration counts of this run; exceeds 1 because the "
                                  "pressure solves (98 % of the model's bytes) run out of registers / shared memory"},
        "per_rank": {"ms_per_step": [r_[0] for r_ in per_rank], "cg_iterations_sum": [r_[1] for r_ in per_rank],
                     "cg_mean_of_launch_max_iterations": [r_[2] for r_ in per_rank],
                     "cg_ms_per_step": [r_[3] for r_ in per_rank], "bicgstab_ms_per_step": [r_[4] for r_ in per_rank]},
        "clocks": clocks, "finite": finite,
    }
    if training is not None:
        line["training_c3"] = training
    if args.config5 and world == 1:
        line["config5"] = run_config5(dev, peak)
    if args.cpu_baseline:
        cores = usable_cores()      # the CPU baseline is the all-core figure (one sample per core), as in the reference arm
        v, dt = cpu_throughput(cores, args.cpu_steps, adjoint=True)
        line["cpu_baseline"] = {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                "sample": "%d samples (one per core) x %d fwd+adjoint steps of the 128x128 case (%.1f s)"
                                          % (cores, args.cpu_steps, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-steps", type=int, default=60)
    ap.add_argument("--same-seeds", action="store_true", help="every rank solves the same samples (scaling diagnostics)")
    ap.add_argument("--groups", type=int, default=8,
                    help="independent sample pipelines the batch is run as (diffpiso_b200.SampleGroups; 1 = one stream)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="run the sample groups eagerly (no CUDA graphs)")
    ap.add_argument("--stream-groups", type=int, default=None,
                    help="sample groups forked INSIDE every piso_step call (SimulationParameters.stream_groups)")
    ap.add_argument("--no-config5", dest="config5", action="store_false",
                    help="skip the extra BASELINE configs[4] measurement (periodic 1024^2, batch 8)")
    ap.add_argument("--no-training", dest="training", action="store_false",
                    help="skip the extra BASELINE configs[2] training-iteration measurement (closure network + all-reduce)")
    ap.add_argument("--config5-only", action="store_true", help="run only the configs[4] measurement (used under ncu)")
    ap.add_argument("--config5-maxit", type=int, default=20000, help="CG iteration cap of the configs[4] run (profiling)")
    ap.add_argument("--config5-n", type=int, default=1024, help="grid of the --config5-only run (1024 or 2048)")
    ap.add_argument("--config5-batch", type=int, default=8, help="batch of the --config5-only run")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.config5_only:
        import torch
        torch.cuda.set_device(0)
        print(json.dumps(run_config5(torch.device("cuda", 0), 6460.9, n=args.config5_n, batch=args.config5_batch, steps=1,
                                     cg_max_it=args.config5_maxit)))
    else:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            args.cpu_baseline = args.cpu_baseline and int(os.environ.get("RANK", "0")) == 0
        run_ours(args)


if __name__ == "__main__":
    main()
