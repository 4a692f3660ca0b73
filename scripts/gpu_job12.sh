timeout 900 python -m pytest tests/test_gpu_full_configs.py -m gpu -q --timeout 600 -p no:cacheprovider -k "c5_1024" 2>&1 | tail -8
