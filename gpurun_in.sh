timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow or power_of_two or config_b or non_canonical" 2>&1 | tail -3
Q4=537133057,537591809,537722881,538116097
for sch in paired unpaired; do
for op in CRT CRTInv; do
  LOLB_DF_SCHEDULE=$sch timeout 120 python tools/run_op.py 65536 $Q4 1024 $op 20
  LOLB_DF_SCHEDULE=$sch timeout 120 python tools/run_op.py 65536 537133057,537591809 2048 $op 20
  LOLB_DF_SCHEDULE=$sch timeout 120 python tools/run_op.py 65536 537133057 4096 $op 20
done
done
