Q4=537133057,537591809,537722881,538116097
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow or config_b" 2>&1 | tail -3
V=lol_b200/csrc/build/variants
for lib in "" $V/df_nopf.so; do
  echo "== lib=$lib"
  for op in CRT CRTInv; do
    LOLB_LIBRARY=$lib timeout 120 python tools/run_op.py 65536 $Q4 1024 $op 20
    LOLB_LIBRARY=$lib timeout 120 python tools/run_op.py 65536 537133057 4096 $op 20
  done
done
