"""torch.profiler table of one configs[2] training iteration (after one warm-up iteration)."""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "differentiable-piso_b200"), os.path.join(ROOT, "scripts")]
import torch
import training_bench
from torch.profiler import profile, ProfilerActivity
torch.cuda.set_device(0)
unroll = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ns = types.SimpleNamespace(config="tml", batch=8, unroll=unroll, iters=1, warmup=1)
t0 = time.perf_counter()
out = training_bench.measure(ns, "cuda:0", 0, 1)
print("unprofiled: total s", time.perf_counter() - t0, "ms_per_iteration", out["ms_per_iteration"])
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    out = training_bench.measure(types.SimpleNamespace(config="tml", batch=8, unroll=unroll, iters=1, warmup=0), "cuda:0", 0, 1)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))
