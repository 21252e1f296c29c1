timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "batched_rq or full_size or golden or fused_crt_mul or dropin_zq or non_canonical" 2>&1 | tail -2
for op in CRT CRTInv; do timeout 120 python tools/run_op.py 14400 14401 65536 $op 30; done
timeout 120 python tools/run_op.py 14400 14401,1008001,429336001 16384 CRT 10
