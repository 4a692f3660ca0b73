mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_adjoint.py tests/test_gpu_piso_step.py tests/test_gpu_reference_python.py -m gpu -q --timeout 400 -p no:cacheprovider 2>&1 | tail -15
for d in -1 16 32 48; do echo "BICG_DBG=$d"; BICG_DBG=$d timeout 120 python scripts/bicg_micro.py; done 2>&1 | tee gpurun_out/bicg_ab.txt
rm -f gpurun_out/cg_sweep3.txt
for cfg in "0 -1 64" "16 8 64" "16 9 64" "16 8 33" "16 9 33" "16 8 128" "16 9 128" "0 -1 128"; do
  set -- $cfg
  echo "cluster=$1 variant=$2 batch=$3" >> gpurun_out/cg_sweep3.txt
  timeout 120 python scripts/cg_micro.py --cluster $1 --variant $2 --batch $3 --reps 5 --check 2 2>&1 | tail -1 | cut -c1-700 >> gpurun_out/cg_sweep3.txt
done
cat gpurun_out/cg_sweep3.txt
