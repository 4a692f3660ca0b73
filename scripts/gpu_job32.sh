echo "default connections"; timeout 300 python scripts/groups_sweep.py 8:g 16:g 2>&1 | tail -2
echo "CUDA_DEVICE_MAX_CONNECTIONS=32"; CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 python scripts/groups_sweep.py 8:g 12:g 16:g 21:g 32:g 2>&1 | tail -5
