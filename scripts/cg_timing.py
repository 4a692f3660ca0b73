"""Phase timing of the cluster-resident pressure CG (diagnostics build: python differentiable-piso_b200/build.py --timing).

    DPISO_LIBRARY=differentiable-piso_b200/diffpiso_b200/libdpiso_timing.so python scripts/cg_timing.py [batch ...]

For every batch size: launch time by CUDA events, and the mean %clock deltas between the stamps thread 0 of cluster rank 0
takes in iterations 4..63 (pressure_cg.cu, CG_T): loop top -> halo wait + halo update -> barrier -> stencil + dot products
-> partials in smem + barrier -> warp tree + st.async -> mbarrier wait -> rank-order sum + broadcast -> alpha, beta ->
update pass."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "differentiable-piso_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

NAMES = ["halo_wait+update", "barrier_top", "stencil+dots", "sts+barrier", "tree+st.async", "mbar_wait", "sum+bcast",
         "alpha_beta", "update", "loop_back"]


def main():
    from diffpiso_b200 import ops, setups as SU
    from diffpiso_b200 import _native as N
    batches = [int(a) for a in sys.argv[1:]] or [1, 8, 16, 32, 33, 64]
    ny = nx = 128
    dev = "cuda:0"
    s = SU.periodic_box(ny, nx, visc=1e-3)
    g = ops.Geometry.get(ny, nx, True, True, dev)
    fn = N.lib.dpiso_pressure_cg_set_timing
    fn.argtypes = [C.c_void_p]
    fn.restype = C.c_int
    for batch in batches:
        rng = np.random.RandomState(0)
        vel = np.stack([SU.solenoidal_field(ny, nx, seed=100 + i) for i in range(batch)])
        vel = vel + 0.01 * rng.randn(*vel.shape).astype(np.float32)
        tv = torch.as_tensor(vel).to(dev)
        ones = torch.ones((ny + 2) * (nx + 2), device=dev)
        dm = torch.zeros(g.nf, dtype=torch.uint8, device=dev)
        ns = torch.zeros((ny + 2) * (nx + 2), dtype=torch.uint8, device=dev)
        beta = float(np.float32(s["dy"] * s["dx"] / s["dt"]))
        values, a_diag = ops.assemble(g, tv, dm, ones, ns, torch.tensor([1e-3], device=dev), s["dy"], s["dx"], beta)
        div = ops.fv_divergence(g, tv, s["dy"], s["dx"])
        div = div - div.mean(dim=1, keepdim=True)
        lap = ops.laplace(g, ones, ones, a_diag, 1, beta, float(np.float32(s["dx"] / s["dy"])), fp64=True)
        stamps = torch.zeros(batch * 64 * 12, dtype=torch.int32, device=dev)
        fn(None)
        for _ in range(2):
            x, its = ops.pressure_cg(g, lap, div, 1e-8, 10000, 1000, True)
        torch.cuda.synchronize()
        times = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            x, its = ops.pressure_cg(g, lap, div, 1e-8, 10000, 1000, True)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        fn(stamps.data_ptr())
        x, its = ops.pressure_cg(g, lap, div, 1e-8, 10000, 1000, True)
        torch.cuda.synchronize()
        fn(None)
        st = stamps.cpu().numpy().astype(np.uint32).reshape(batch, 64, 12)
        itn = its.cpu().numpy()
        out = {"batch": batch, "ms": float(np.median(times)), "mean_it": float(itn.mean()), "max_it": int(itn.max()),
               "us_per_it_of_longest": 1e3 * float(np.median(times)) / float(itn.max())}
        for which in sorted({0, batch // 2, batch - 1}):
            t = st[which, 4:63].astype(np.int64)
            d = {}
            for k in range(9):
                d[NAMES[k]] = float(((t[:, k + 1] - t[:, k]) % (1 << 32)).mean())
            nxt = st[which, 5:64, 0].astype(np.int64)
            d[NAMES[9]] = float(((nxt - t[:, 9]) % (1 << 32)).mean())
            d["iteration"] = float(((nxt - t[:, 0]) % (1 << 32)).mean())
            d["smid"] = int(st[which, 10, 10])
            out["sample_%d" % which] = {k: round(v, 1) if isinstance(v, float) else v for k, v in d.items()}
        print(json.dumps(out))


if __name__ == "__main__":
    main()
