#!/bin/bash
# Next GPU session: ncu evidence for ext_stream / coeff_stream (none was captured in round 1, the budget ran out).
# Run under gpurun on ONE GPU, only after `python tools/run_ext.py` itself has exited 0 without ncu:
#   gpurun --timeout 300 -- 'bash tools/ncu_ext.sh'
# Outputs land in gpurun_out/ (copy the summaries into profiles/ afterwards with tools/ncu_summary.py).
set -e
mkdir -p gpurun_out
python tools/run_ext.py 576 14400 16384 10 > gpurun_out/ext_bench.jsonl
# launch list (cold-cache, serialised durations: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ext_launches.csv \
    python tools/run_ext.py 576 14400 4096 1 > gpurun_out/ext_ncu_list.log 2>&1
# one full capture each of the two kernels that are below 80 % of the HBM roofline
ncu --set full --clock-control none --import-source on -k regex:k_ext_twace_crt_zq -c 1 -o gpurun_out/ext_twace_crt \
    python tools/run_ext.py 576 14400 4096 1 > gpurun_out/ext_ncu_twace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_coeff_stream -c 12 -o gpurun_out/coeff_stream \
    python tools/run_ext.py 576 14400 4096 1 > gpurun_out/ext_ncu_coeff.log 2>&1
