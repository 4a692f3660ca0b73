mkdir -p gpurun_out
rm -f gpurun_out/cg_sweep2.txt
for cfg in "0 -1 64" "16 8 64" "16 9 64" "0 -1 32" "16 8 32" "16 9 32" "0 -1 128" "16 8 128" "16 9 128" "0 -1 37" "0 -1 33"; do
  set -- $cfg
  echo "cluster=$1 variant=$2 batch=$3" >> gpurun_out/cg_sweep2.txt
  timeout 120 python scripts/cg_micro.py --cluster $1 --variant $2 --batch $3 --reps 5 --check 2 >> gpurun_out/cg_sweep2.txt 2>&1
done
cat gpurun_out/cg_sweep2.txt
