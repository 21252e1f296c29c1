set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; tail -c 300 gpurun_out/bench_r01e.err
python tools/bsum.py gpurun_out/bench_r01e.json | head -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_a -s 6 -c 2 -o gpurun_out/prof_a_final -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-per-op > gpurun_out/ncu_full.log 2>&1
