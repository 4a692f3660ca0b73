for d in 16 80; do echo "BICG_DBG=$d"; BICG_MAXIT=3 BICG_DBG=$d timeout 120 python scripts/bicg_micro.py; done 2>&1 | tee gpurun_out/bicg_fake.txt
