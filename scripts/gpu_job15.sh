timeout 900 python -m pytest tests/test_gpu_reference_python.py -m gpu -q -s --timeout 600 -p no:cacheprovider -k "c3_full_size" 2>&1 | tail -8
timeout 300 python bench.py --steps 3 --warmup 3 --no-config5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
