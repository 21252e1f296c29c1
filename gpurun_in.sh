timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
Q4=537133057,537591809,537722881,538116097
timeout 120 python tools/run_op.py 65536 $Q4 1024 CRT 20
timeout 120 python tools/run_op.py 2048 12289 131072 CRT 10
timeout 120 python tools/run_op.py 8192 40961 32768 CRTInv 10
