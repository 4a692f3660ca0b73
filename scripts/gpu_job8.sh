mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "tile_kernel" 2>&1 | tail -25
echo "--- 128^2 x 64: tile (default) vs rows kernel"
timeout 120 python scripts/bicg_micro.py 64 128
BICG_DBG=128 timeout 120 python scripts/bicg_micro.py 64 128
echo "--- 1024^2 x 8"
timeout 300 python scripts/bicg_micro.py 8 1024
echo "--- 2048^2 x 4"
timeout 300 python scripts/bicg_micro.py 4 2048
