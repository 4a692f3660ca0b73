mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-config5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; tail -c 300 gpurun_out/scale_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; tail -c 300 gpurun_out/scale_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --same-seeds --no-training > gpurun_out/scale_n2_same.json 2> gpurun_out/scale_n2_same.err
python - <<'PY'
import json
for f in ("scale_n1", "scale_n2", "scale_n2_same"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.3f per_rank ms %s cg_max_it %s training %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["per_rank"]["ms_per_step"], d["per_rank"]["cg_mean_of_launch_max_iterations"], json.dumps(d.get("training_c3"))[:400]))
    except Exception as e:
        print(f, "failed", e)
PY
