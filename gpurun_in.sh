set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; tail -c 300 gpurun_out/bench_r01d.err
python tools/bsum.py gpurun_out/bench_r01d.json | tail -40
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
