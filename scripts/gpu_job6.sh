mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_adjoint.py tests/test_gpu_reference_python.py -m gpu -q --timeout 400 -p no:cacheprovider -k "bicgstab or factor or backward or reference" 2>&1 | tail -5
for d in -1 16 32 48; do echo "BICG_DBG=$d"; BICG_DBG=$d timeout 120 python scripts/bicg_micro.py; done 2>&1 | tee gpurun_out/bicg_ab.txt
