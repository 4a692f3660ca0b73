mkdir -p gpurun_out
rm -f gpurun_out/cg_sweep.txt
for cfg in "0 -1" "8 4" "4 0" "8 1" "8 2" "4 3" "16 2"; do
  set -- $cfg
  echo "cluster=$1 variant=$2" >> gpurun_out/cg_sweep.txt
  timeout 120 python scripts/cg_micro.py --cluster $1 --variant $2 --reps 5 --check 3 >> gpurun_out/cg_sweep.txt 2>&1
done
cat gpurun_out/cg_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pressure_cg -s 1 -c 1 -f -o gpurun_out/prof_cg python scripts/cg_micro.py --reps 1 > gpurun_out/ncu_cg.log 2>&1
tail -3 gpurun_out/ncu_cg.log
