mkdir -p gpurun_out
for k in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -p no:cacheprovider -k "bicgstab" 2>&1 | tail -4; done
timeout 300 python scripts/bicg_micro.py 8 1024
