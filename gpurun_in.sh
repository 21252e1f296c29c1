timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "complex or double_and or plain or golden" 2>&1 | tail -3
