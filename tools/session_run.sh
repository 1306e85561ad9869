set -x
mkdir -p gpurun_out/fin
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin/smoke.log 2>&1; tail -2 gpurun_out/fin/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/fin/bench_n1.json 2> gpurun_out/fin/bench_n1.err; tail -c 1500 gpurun_out/fin/bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/fin/bench_reference.json 2> gpurun_out/fin/bench_reference.err; tail -c 300 gpurun_out/fin/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/fin/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/fin/ncu_bench.log 2>&1; tail -2 gpurun_out/fin/ncu_bench.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fk_jac -s 4 -c 2 -o gpurun_out/fin/prof_fk python tools/prof_kernels.py fk --iters 2 > gpurun_out/fin/ncu_fk.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:torm_ -s 8 -c 3 -o gpurun_out/fin/prof_torm python tools/prof_kernels.py torm --iters 1 > gpurun_out/fin/ncu_torm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pirl_env -s 3 -c 1 -o gpurun_out/fin/prof_pirl_1M python tools/l2_fetch_probe.py > gpurun_out/fin/ncu_p1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pirl_env -s 3 -c 1 -o gpurun_out/fin/prof_pirl_16k python tools/prof_kernels.py pirl --envB 16384 --iters 3 > gpurun_out/fin/ncu_p2.log 2>&1
python tools/prof_tsgd.py 2>&1 | grep -v "stochastic=0" > gpurun_out/fin/prof_tsgd.log; tail -6 gpurun_out/fin/prof_tsgd.log
ls -la gpurun_out/fin
