timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "plain or int64 or double_and" 2>&1 | tail -3
timeout 120 python tools/run_plain.py 14400 65536 2>&1 | grep Norm
