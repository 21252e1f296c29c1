timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for cfg in "1728 3457 131072" "5184 10369 65536" "42 8191 1000000" "64 257 1000000" "2016 12097 100000"; do
  set -- $cfg
  for op in CRT CRTInv; do timeout 120 python tools/run_op.py $1 $2 $3 $op 5; done
done
timeout 120 python tools/run_plain.py 1728 65536 2>&1 | grep CRT
