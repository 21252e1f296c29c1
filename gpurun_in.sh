for mb in 3 4; do for op in CRT CRTInv; do LOLB_FUSED_A_K2_MB=$mb timeout 120 python tools/run_op.py 14400 1008001,1065601 32768 $op 20; done; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "batched_rq or full_size or golden or fused_crt_mul" 2>&1 | tail -2
