for cfg in "64 1" "64 2" "64 4"; do set -- $cfg; timeout 200 python scripts/bicg_stress.py ldc_like64 $1 $2 30; done
timeout 200 python scripts/bicg_stress.py periodic264x256 64 8 10
timeout 200 python scripts/bicg_stress.py periodic264x256 256 5 10
timeout 200 python scripts/bicg_stress.py tml64x128 64 2 20
timeout 200 python scripts/bicg_stress.py tml64x128 256 3 20
timeout 200 python scripts/bicg_stress.py tml64x128 0 0 20
