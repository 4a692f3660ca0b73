mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -p no:cacheprovider -k "tile_kernel or deterministic" 2>&1 | tail -5
echo "--- 2048^2 x 4 tile: TMA (default) vs cp.async"
timeout 300 python scripts/bicg_micro.py 4 2048
BICG_DBG=1280 timeout 300 python scripts/bicg_micro.py 4 2048
echo "--- 1024^2 x 8 tile: TMA vs cp.async vs band (default)"
BICG_DBG=256 timeout 300 python scripts/bicg_micro.py 8 1024
BICG_DBG=1280 timeout 300 python scripts/bicg_micro.py 8 1024
timeout 300 python scripts/bicg_micro.py 8 1024
echo "--- 128^2 x 64 tile TMA"
BICG_DBG=256 timeout 300 python scripts/bicg_micro.py 64 128
