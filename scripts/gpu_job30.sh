timeout 500 python scripts/training_diag.py 16 2>&1 | tail -100 | sort -t" " -k2 -n -r | head -30
