Q4=537133057,537591809,537722881,538116097
V=lol_b200/csrc/build/variants
for lib in $V/df_minb6.so $V/df_nw8.so $V/df_nw8b.so; do
  echo "== lib=$lib"
  for op in CRT CRTInv; do
    LOLB_LIBRARY=$lib timeout 120 python tools/run_op.py 65536 $Q4 1024 $op 20
    LOLB_LIBRARY=$lib timeout 120 python tools/run_op.py 65536 537133057 4096 $op 20
  done
done
