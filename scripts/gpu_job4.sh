mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_adjoint.py -m gpu -q --timeout 400 -p no:cacheprovider 2>&1 | grep -v "^  File\|^    " | tail -80 > gpurun_out/job4_tests.log
tail -5 gpurun_out/job4_tests.log
for d in -1 16 32; do echo "BICG_DBG=$d"; BICG_DBG=$d timeout 120 python scripts/bicg_micro.py; done 2>&1 | tee gpurun_out/bicg_ab.txt
