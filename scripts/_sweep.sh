timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "replayed or spmv_parity" 2>&1 | tail -3
CPS=1 POLL=4 timeout 200 python scripts/quick_bench.py C4slab 1.0 tiles 0 dilu 2>&1 | grep -E "variant|lower|upper|spmv|prec_update" | cut -c1-160
CPS=2 POLL=3 timeout 200 python scripts/quick_bench.py C4slab 1.0 tiles 0 dilu 2>&1 | grep -E "variant|lower|upper" | cut -c1-160
timeout 200 python scripts/quick_bench.py C4slab 1.0 levels 6 dilu 2>&1 | grep -E "variant|lower|upper|spmv|prec_update" | cut -c1-160
timeout 100 python scripts/quick_bench.py C2 1.0 auto 0 dilu 2>&1 | grep -E "variant|lower|upper|spmv|prec_update|solve" | cut -c1-200
