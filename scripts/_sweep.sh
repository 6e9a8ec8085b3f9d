timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for p in dilu ilu0; do timeout 100 python scripts/quick_bench.py C3 1.0 tiles 0 $p 2>&1 | cut -c1-200; done
timeout 100 python scripts/quick_bench.py C3 1.0 levels 6 dilu 2>&1 | cut -c1-200
