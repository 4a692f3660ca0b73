# one pass over everything the round's evidence is made of: tests, bench (both arms), launch list, per-kernel DRAM,
# full ncu captures of the two solver kernels, config #5 DRAM traffic and sweep
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
bash scripts/gpu_checks.sh
bash scripts/gpu_bench.sh
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
bash scripts/gpu_profile_all.sh > /dev/null 2>&1; head -14 gpurun_out/kernels_dram.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pressure_cg -s 1 -c 1 -f -o gpurun_out/prof_cg_r02 python scripts/cg_micro.py --reps 1 > gpurun_out/ncu_cg.log 2>&1; tail -2 gpurun_out/ncu_cg.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bicgstab_rows -s 2 -c 1 -f -o gpurun_out/prof_bicg_r02 python scripts/bicg_micro.py > gpurun_out/ncu_bicg.log 2>&1; tail -2 gpurun_out/ncu_bicg.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pressure_cg|bicgstab" --csv --log-file gpurun_out/c5_kernels.csv python bench.py --config5-only --config5-maxit 200 > gpurun_out/c5_ncu.log 2>&1; tail -c 300 gpurun_out/c5_ncu.log
bash scripts/gpu_c5_sweep.sh 2>&1 | tail -12
