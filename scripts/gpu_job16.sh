timeout 900 python -m pytest tests/test_gpu_reference_python.py tests/test_gpu_piso_step.py tests/test_gpu_adjoint.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -8
grep "unroll_vs_reference_python\|c3_tml256x128" gpurun_out/parity_records.jsonl | tail -3 | cut -c1-1500
