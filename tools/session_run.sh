set -x
mkdir -p gpurun_out/s8
python -m pytest tests -q -m gpu -x > gpurun_out/s8/pytest_gpu.log 2>&1; tail -3 gpurun_out/s8/pytest_gpu.log
python tools/prof_update.py > gpurun_out/s8/upd.log 2>&1; cat gpurun_out/s8/upd.log
python tools/prof_tsgd.py 2>&1 | grep -v "stochastic=0\|block=1 " > gpurun_out/s8/prof_tsgd.log; grep "mode" gpurun_out/s8/prof_tsgd.log
