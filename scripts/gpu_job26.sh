python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-config5 --no-training > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N2 value %.4g e2e %.4g ms %.3f serial %.3f per_rank %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['per_rank']['ms_per_step']))
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-config5 --no-training --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('N1 value %.4g e2e %.4g ms %.3f serial %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['single_stream']['ms_per_step']))
print('cg', d['roofline']['avg_launch_ms'], d['roofline']['frac'])
PY
