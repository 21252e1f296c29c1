V=lol_b200/csrc/build/variants
for lib in "" $V/a_l2pf.so; do
  for op in CRT CRTInv; do LOLB_LIBRARY=$lib timeout 120 python tools/run_op.py 14400 14401 65536 $op 30; done
done
