mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bicgstab -s 2 -c 1 -f -o gpurun_out/prof_bicg python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bicg.log 2>&1
tail -2 gpurun_out/ncu_bicg.log
