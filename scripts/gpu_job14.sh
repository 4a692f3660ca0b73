timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -k "fp64_matches or cast_to_double" 2>&1 | tail -15
