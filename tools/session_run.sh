mkdir -p gpurun_out/n4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4/bench_n4.json 2> gpurun_out/n4/bench_n4.err; tail -c 400 gpurun_out/n4/bench_n4.json; tail -2 gpurun_out/n4/bench_n4.err
