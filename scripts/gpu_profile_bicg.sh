mkdir -p gpurun_out
BICG_DBG=${BICG_DBG:--1} timeout 600 ncu --set full --clock-control none --import-source on -k regex:bicgstab_rows -s 2 -c 1 -f -o gpurun_out/prof_bicg_r02 python scripts/bicg_micro.py > gpurun_out/ncu_bicg.log 2>&1
tail -3 gpurun_out/ncu_bicg.log
