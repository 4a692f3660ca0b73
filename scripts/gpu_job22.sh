timeout 300 python scripts/groups_diag.py 2>&1 | tail -8
