mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-steps 2 > gpurun_out/bench.log 2>&1; echo "bench exit $?"
tail -2 gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config5 --no-training > gpurun_out/bench_ncu.log 2>&1; echo "launch list exit $?"
