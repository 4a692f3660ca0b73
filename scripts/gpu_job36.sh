cat > /tmp/rb.py <<'PY'
import ctypes, os, sys
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "differentiable-piso_b200")]
from diffpiso_b200 import _native as N
N.lib.dpiso_bicgstab_set_rows_block.argtypes = [ctypes.c_int]
N.lib.dpiso_bicgstab_set_rows_block(int(os.environ.get("ROWS_BLOCK", "0")))
sys.argv = sys.argv[1:]
exec(open(sys.argv[0]).read())
PY
for rb in 0 256 384; do echo "rows block $rb"; ROWS_BLOCK=$rb timeout 120 python /tmp/rb.py scripts/bicg_micro.py 64 128 | cut -c1-200; ROWS_BLOCK=$rb timeout 200 python /tmp/rb.py scripts/groups_sweep.py 8:g | tail -1; done
