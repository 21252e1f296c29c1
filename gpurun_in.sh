timeout 800 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow or power_of_two or non_canonical or batched_rq or config_b or full_size" 2>&1 | tail -3
