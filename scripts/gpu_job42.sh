L=differentiable-piso_b200/diffpiso_b200
DPISO_LIBRARY=$L/libdpiso_old.so timeout 600 python scripts/cg_bitcmp.py --out gpurun_out/cg_old.npz 2>&1 | tail -1
timeout 600 python scripts/cg_bitcmp.py --out gpurun_out/cg_new.npz 2>&1 | tail -1
python scripts/cg_bitcmp.py --compare gpurun_out/cg_old.npz gpurun_out/cg_new.npz 2>&1 | tail -5
for i in 1 2; do DPISO_LIBRARY=$L/libdpiso_old.so timeout 120 python scripts/cg_micro.py --batch 64 2>&1 | tail -1 | cut -c1-60; timeout 120 python scripts/cg_micro.py --batch 64 2>&1 | tail -1 | cut -c1-60; done
DPISO_LIBRARY=$L/libdpiso_old.so timeout 120 python scripts/cg_micro.py --batch 33 2>&1 | tail -1 | cut -c1-60; timeout 120 python scripts/cg_micro.py --batch 33 2>&1 | tail -1 | cut -c1-60
DPISO_LIBRARY=$L/libdpiso_timing.so timeout 300 python scripts/cg_timing.py 1 33 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['batch'], d['ms'], d['sample_0'])"
rm -f gpurun_out/cg_old.npz gpurun_out/cg_new.npz
