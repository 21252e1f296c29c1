timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow or power_of_two or non_canonical or batched_rq" 2>&1 | tail -3
for cfg in "4096 40961 65535" "4096 537133057,537591809 32768"; do
  set -- $cfg
  for op in CRT CRTInv; do timeout 120 python tools/run_op.py $1 $2 $3 $op 10; done
done
