mkdir -p gpurun_out
for b in 64 32 37 74 128; do
 for snx in 0 1; do
  echo "batch=$b static_nx=$snx"; timeout 120 python scripts/cg_micro.py --batch $b --static-nx $snx --reps 7 --check 2 | cut -c1-330
 done
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -k "pressure_cg" 2>&1 | tail -3
