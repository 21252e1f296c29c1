Q4=537133057,537591809,537722881,538116097
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow or power_of_two or config_b or non_canonical" 2>&1 | tail -3
for cfg in "0 0" "12 48" "20 44" "30 64" "40 96" "60 128"; do
  set -- $cfg
  echo "== lag=$1 ring=$2"
  for op in CRT CRTInv; do
    LOLB_DF_LAG=$1 LOLB_DF_RING=$2 timeout 120 python tools/run_op.py 65536 $Q4 1024 $op 20
  done
done
for cfg in "0 0" "60 128" "113 230" "160 400"; do
  set -- $cfg
  echo "== k1 lag=$1 ring=$2"
  for op in CRT CRTInv; do
  LOLB_DF_LAG=$1 LOLB_DF_RING=$2 timeout 120 python tools/run_op.py 65536 537133057 4096 $op 20
  done
done
timeout 120 python tools/run_op.py 65536 537133057,537591809 2048 CRT 20
timeout 120 python tools/run_op.py 65536 537133057,537591809 2048 CRTInv 20
