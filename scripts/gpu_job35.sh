timeout 120 python scripts/bicg_micro.py 64 128
timeout 120 python scripts/bicg_micro.py 8 128
bash scripts/gpu_job24.sh
