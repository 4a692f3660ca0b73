timeout 300 python scripts/split_probe.py 1 2 4 8 2>&1 | tail -6
