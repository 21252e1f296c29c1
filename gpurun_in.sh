for cfg in "16 4 23 CRTInv 3 1" "16 4 23 CRTInv 3 1" "16 2 23 CRTInv 3 1" "16 4 23 CRTInv 8 1" "16 4 23 CRT 3 1"; do
  set -- $cfg
  echo "== $cfg"
  LOLB_DF_RING=$5 LOLB_DF_LAG=$6 timeout 40 python tools/df_probe.py $1 $2 $3 $4 2>&1 | tail -4
done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow or power_of_two or config_b or non_canonical" 2>&1 | tail -3
Q4=537133057,537591809,537722881,538116097
for op in CRT CRTInv; do
  timeout 120 python tools/run_op.py 65536 $Q4 1024 $op 20
  timeout 120 python tools/run_op.py 65536 537133057,537591809 2048 $op 20
  timeout 120 python tools/run_op.py 65536 537133057 4096 $op 20
done
