timeout 120 python scripts/bicg_micro.py 64 128
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "bicgstab" 2>&1 | tail -3
