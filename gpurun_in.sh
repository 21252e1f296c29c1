timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; echo bench rc=$?
python tools/bsum.py gpurun_out/bench_r01_final.json
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2>/dev/null; echo ref rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list.log 2>&1; echo list rc=$?
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_fused_a -s 6 -c 2 -o gpurun_out/prof_a_final -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-per-op > gpurun_out/ncu_full.log 2>&1; echo full rc=$?
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
