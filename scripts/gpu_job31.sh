python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-config5 --no-training --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -3 gpurun_out/bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1])
print('N4 value %.4g e2e %.4g ms %.3f serial %.3f per_rank %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], [round(x,2) for x in d['per_rank']['ms_per_step']]))
PY
