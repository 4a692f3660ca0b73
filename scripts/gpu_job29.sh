timeout 500 python scripts/training_prof.py 8 2>&1 | cut -c1-220 | tail -75
