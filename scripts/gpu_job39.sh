timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; tail -2 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null; echo "reference exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/final_bench_reference.json').read().strip().splitlines()[-1])
print('value %.4g e2e %.4g ms %.3f serial %.3f fwd %.3f (%.3f) launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['forward_only']['ms_per_step'], d['forward_only']['single_stream_ms_per_step'], d['gpu_launches']))
print('cg', d['roofline']['avg_launch_ms'], d['roofline']['frac'], 'bicg', d['roofline_bicgstab']['avg_launch_ms'], d['roofline_bicgstab']['frac'], d['clocks'])
print('cpu', d.get('cpu_baseline')); print('train', d.get('training_c3',{}).get('ms_per_iteration')); print('c5', d['config5']['ms_per_step'])
print('reference arm', r['value'], r['cpu_baseline']['cores'], 'ratio e2e', d['e2e']['value']/r['value'])
PY
