timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow or power_of_two or batched_rq or dropin_zq or non_canonical" 2>&1 | tail -4
for cfg in "1024 12289 262144" "2048 12289 131072" "2048 537133057,537591809 65536" "2048 537133057,537591809,537722881,538116097 32768" "1024 537133057,537591809,537722881,538116097 65536"; do
  set -- $cfg
  for op in CRT CRTInv; do timeout 120 python tools/run_op.py $1 $2 $3 $op 10; done
done
