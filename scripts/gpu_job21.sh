timeout 900 python -m pytest tests/test_gpu_piso_step.py tests/test_gpu_adjoint.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "stream_groups" 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-config5 --no-training --no-cpu-baseline > gpurun_out/bench_groups.json 2> gpurun_out/bench_groups.err; tail -3 gpurun_out/bench_groups.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_groups.json').read().strip().splitlines()[-1])
print('value %.4g e2e %.4g ms %.3f serial %.3f fwd %.3f groups %s launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['forward_only']['ms_per_step'], d['config']['stream_groups'], d['gpu_launches']))
print('cg', d['roofline']['avg_launch_ms'], d['roofline']['frac'], 'bicg', d['roofline_bicgstab']['avg_launch_ms'])
PY
