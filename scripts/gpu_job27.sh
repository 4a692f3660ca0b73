L=differentiable-piso_b200/diffpiso_b200
for v in "0 -1" "8 10"; do set -- $v; for b in 64 33; do echo "cluster=$1 variant=$2 batch=$b"; timeout 120 python scripts/cg_micro.py --cluster $1 --variant $2 --batch $b --check 2 2>&1 | tail -1 | cut -c1-420; done; done
echo OLD; DPISO_LIBRARY=$L/libdpiso_old.so timeout 400 python scripts/training_bench.py --config tml --batch 8 --unroll 16 --iters 1 --warmup 0 2>&1 | tail -1 | cut -c1-600
echo NEW; timeout 400 python scripts/training_bench.py --config tml --batch 8 --unroll 16 --iters 1 --warmup 0 2>&1 | tail -1 | cut -c1-600
