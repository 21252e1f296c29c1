set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; tail -c 300 gpurun_out/bench_r01c.err
python tools/bsum.py gpurun_out/bench_r01c.json | tail -34
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_a -s 6 -c 2 -o gpurun_out/prof_a_final -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-per-op > gpurun_out/ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
