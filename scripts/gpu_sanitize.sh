# compute-sanitizer over a small, representative subset of the GPU tests (memcheck: out-of-bounds / misaligned global,
# shared and DSMEM accesses; racecheck: shared-memory hazards; synccheck: barrier misuse; initcheck: uninitialised reads)
mkdir -p gpurun_out
SEL=${SEL:-'ldc8 or periodic16 or tml16x24 or sml16x48 or periodic24x20'}
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_adjoint.py tests/test_gpu_piso_step.py -m gpu -q -p no:cacheprovider -k "$SEL" \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | sort -rn | head -12
done
