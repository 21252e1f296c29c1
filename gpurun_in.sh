timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "host_pipeline or dataflow or config_b" 2>&1 | tail -4
