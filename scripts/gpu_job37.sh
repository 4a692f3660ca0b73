timeout 120 python scripts/bicg_micro.py 64 128 | cut -c1-420
timeout 120 python scripts/bicg_micro.py 8 128 | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_adjoint.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "bicgstab or factor or linear_solver" 2>&1 | tail -3
timeout 200 python scripts/groups_sweep.py 1:e 8:g | tail -2
