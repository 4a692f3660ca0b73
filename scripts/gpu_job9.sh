mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bicgstab_tile -s 1 -c 1 -f -o gpurun_out/prof_tile python scripts/bicg_micro.py 64 128 > gpurun_out/ncu_tile.log 2>&1; tail -3 gpurun_out/ncu_tile.log
