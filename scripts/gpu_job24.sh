timeout 900 python bench.py --steps 10 --warmup 3 --no-config5 --no-training --no-cpu-baseline > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; tail -3 gpurun_out/bench_g8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_g8.json').read().strip().splitlines()[-1])
print('value %.4g e2e %.4g ms %.3f serial %.3f fwd %.3f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['forward_only']['ms_per_step'], d['gpu_launches']))
print('e2e ms/step', 64*16384/d['e2e']['value']*1e3, 'cg', d['roofline']['avg_launch_ms'], d['roofline']['frac'], 'bicg', d['roofline_bicgstab']['avg_launch_ms'], d['clocks'], d.get('cpu_baseline'))
PY
