mkdir -p gpurun_out/s16
python tools/torm_check.py > gpurun_out/s16/torm_check.log 2>&1; grep "bench shape\|^torm \|cost\[" gpurun_out/s16/torm_check.log
for k in 0 4 6; do RLITG_TEST_SEED=$k python -m pytest tests/test_torm_gpu.py tests/test_torm_schedule_gpu.py tests/test_torm_init_gpu.py -q -m gpu > gpurun_out/s16/seed$k.log 2>&1; echo "seed $k: $(tail -1 gpurun_out/s16/seed$k.log)"; done
python bench.py --workload torm --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s16/bench_torm.json 2> gpurun_out/s16/bench_torm.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/s16/bench_torm.json').read().strip().splitlines()[-1])
print(json.dumps(d['summary']['torm']))
print(json.dumps(d['also']['torm_cost_grad'])[:900])
P
