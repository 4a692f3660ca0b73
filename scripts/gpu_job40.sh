mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "pressure_cg_matches_oracle and (tml64x128 or periodic64x32 or ldc_like64 or periodic24x20) and True" \
    > gpurun_out/sanitize_cluster_$tool.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/sanitize_cluster_$tool.log | sort | uniq -c | sort -rn | head -8
done
