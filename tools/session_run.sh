mkdir -p gpurun_out/s12
python bench.py --steps 20 --warmup 5 > gpurun_out/s12/bench_n1.json 2> gpurun_out/s12/bench_n1.err; tail -c 1900 gpurun_out/s12/bench_n1.json; tail -3 gpurun_out/s12/bench_n1.err
python bench.py --workload pirl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s12/bench_pirl.json 2> gpurun_out/s12/bench_pirl.err; head -c 600 gpurun_out/s12/bench_pirl.json
