timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_groups.py tests/test_gpu_piso_step.py -m gpu -q -p no:cacheprovider -k "groups" > gpurun_out/sanitize_groups_memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitize_groups_memcheck.log | sort | uniq -c | head
