set -x
mkdir -p gpurun_out/s4
python -m pytest tests/test_torm_schedule_gpu.py -q -m gpu -x > gpurun_out/s4/pytest_sched.log 2>&1; tail -5 gpurun_out/s4/pytest_sched.log
python tools/prof_tsgd.py --parts-only > gpurun_out/s4/prof_parts.log 2>&1; tail -20 gpurun_out/s4/prof_parts.log
