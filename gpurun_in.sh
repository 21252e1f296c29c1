timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "batched_rq or fused_crt_mul or non_canonical or host_pipeline" 2>&1 | tail -3
Q3=14401,1008001,429336001
Q4=43201,57601,100801,115201
for v in 0 1 2 3; do for op in CRT CRTInv; do LOLB_FUSED_A_KN=$v timeout 120 python tools/run_op.py 14400 $Q3 16384 $op 10; done; done
for v in 0 1 2 3; do for op in CRT CRTInv; do LOLB_FUSED_A_KN=$v timeout 120 python tools/run_op.py 14400 $Q4 16384 $op 10; done; done
