timeout 300 python scripts/training_diag.py 4 2>&1 | tail -60
