DPISO_LIBRARY=differentiable-piso_b200/diffpiso_b200/libdpiso_timing.so timeout 300 python scripts/cg_timing.py 1 8 16 32 33 64 2>&1 | tail -8
