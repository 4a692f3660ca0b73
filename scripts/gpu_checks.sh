mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
rm -f gpurun_out/summary.txt
for f in test_cabi_consumer test_gpu_kernels test_gpu_groups test_gpu_piso_step test_gpu_adjoint test_gpu_reference_pin test_gpu_helpers test_gpu_full_configs test_gpu_reference_python; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 240 --timeout-method thread -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/$f.log | tail -n 30
done
cat gpurun_out/summary.txt
