export LOLB_DF_SCHEDULE=d16
for cfg in "16 4 23 CRT 0 0" "16 4 23 CRTInv 0 0" "16 1 23 CRT 3 1" "16 2 23 CRTInv 3 1" "14 4 23 CRT 5 4" "13 2 23 CRTInv 5 4" "15 1 23 CRT 0 0"; do
  set -- $cfg
  echo "== $cfg"
  LOLB_DF_RING=$5 LOLB_DF_LAG=$6 timeout 40 python tools/df_probe.py $1 $2 $3 $4 2>&1 | tail -4
done
Q4=537133057,537591809,537722881,538116097
for op in CRT CRTInv; do
  timeout 120 python tools/run_op.py 65536 $Q4 1024 $op 20
  timeout 120 python tools/run_op.py 65536 537133057,537591809 2048 $op 20
  timeout 120 python tools/run_op.py 65536 537133057 4096 $op 20
done
