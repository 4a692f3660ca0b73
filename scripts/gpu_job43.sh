timeout 900 python bench.py --steps 200 --warmup 3 --no-config5 --no-training --no-cpu-baseline > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err; tail -2 gpurun_out/bench_long.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_long.json').read().strip().splitlines()[-1])
print('steps 200: value %.4g e2e %.4g ms %.3f serial %.3f fwd %.3f finite %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['forward_only']['ms_per_step'], d['finite']))
print(d['clocks'], d['roofline']['cg_iterations_min_max'], d['roofline']['mean_cg_iterations'])
PY
