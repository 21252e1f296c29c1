timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "symmshe" 2>&1 | tail -3
timeout 120 python tools/run_she.py 4096 0 10
timeout 120 python tools/run_she.py 4096 1024 10
