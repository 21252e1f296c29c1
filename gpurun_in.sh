timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "symmshe" 2>&1 | tail -5
timeout 100 python tools/run_she.py 4096 0 10
