timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 200 gpurun_out/bench_n8.err
python tools/bsum.py gpurun_out/bench_n8.json 2>/dev/null | head -6
