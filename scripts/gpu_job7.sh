mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -k "cluster_kernel" 2>&1 | tail -15
echo "--- 128^2 x 64: default rows kernel vs cluster kernel C=1"
timeout 120 python scripts/bicg_micro.py 64 128
BICG_DBG=64 timeout 120 python scripts/bicg_micro.py 64 128
BICG_DBG=96 timeout 120 python scripts/bicg_micro.py 64 128
echo "--- 1024^2 x 8"
timeout 300 python scripts/bicg_micro.py 8 1024
BAND_CLUSTER=4 timeout 300 python scripts/bicg_micro.py 8 1024
BICG_DBG=32 timeout 300 python scripts/bicg_micro.py 8 1024
echo "--- 2048^2 x 4"
timeout 300 python scripts/bicg_micro.py 4 2048
