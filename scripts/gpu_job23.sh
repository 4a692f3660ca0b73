timeout 600 python -m pytest tests/test_gpu_groups.py tests/test_gpu_piso_step.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "groups" 2>&1 | tail -6
timeout 600 python scripts/groups_sweep.py 1:e 1:g 2:g 4:e 4:g 6:g 8:g 12:g 16:g 32:g 2>&1 | tail -12
