# per-kernel duration + DRAM bytes of ONE bench invocation (all warm-up and timed steps), summarised per kernel name
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/kernels_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config5 --no-training > gpurun_out/kernels_dram.log 2>&1
echo "ncu exit $?"
python scripts/summarize_ncu.py gpurun_out/kernels_dram.csv > gpurun_out/kernels_dram.md
cat gpurun_out/kernels_dram.md
