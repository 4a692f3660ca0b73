for k in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-config5 --no-cpu-baseline --no-training 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g e2e %.4g ms/step %.3f  e2e ms/step %.3f finite %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], 64*16384/d['e2e']['value']*1e3, d['finite']))"; done
