"""Where a configs[2] training iteration spends its time: every solver launch timed with a host sync."""
import os, sys, time, types, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "differentiable-piso_b200"), os.path.join(ROOT, "scripts")]
import torch
import training_bench
from diffpiso_b200 import ops
log = []
def wrap(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        st = out[1]
        info = st.float().max().item() if name == "cg" else st[:, :, 0].max().item()
        extra = "" if name == "cg" else " restarts %d warn %d kind %s transpose %s" % (st[:, :, 1].max().item(), st[:, :, 2].max().item(), st[:, :, 3].unique().tolist(), k.get("transpose", a[6] if len(a) > 6 else None))
        log.append("%s %.2f ms max_it %d%s" % (name, dt, info, extra))
        return out
    return w
ops.pressure_cg = wrap("cg", ops.pressure_cg)
ops.bicgstab_ilu = wrap("bicg", ops.bicgstab_ilu)
torch.cuda.set_device(0)
t0 = time.perf_counter()
out = training_bench.measure(types.SimpleNamespace(config="tml", batch=8, unroll=int(sys.argv[1]) if len(sys.argv) > 1 else 4, iters=1, warmup=0), "cuda:0", 0, 1)
print("total s", time.perf_counter() - t0, out["ms_per_iteration"])
print("\n".join(log))
