mkdir -p gpurun_out/s19
python -m pytest tests/test_torm_gpu.py -q -m gpu > gpurun_out/s19/pytest.log 2>&1; tail -2 gpurun_out/s19/pytest.log
RLITG_TEST_SEED=1 python -m pytest tests/test_torm_gpu.py -q -m gpu -k benchmark_shape > gpurun_out/s19/pytest1.log 2>&1; tail -1 gpurun_out/s19/pytest1.log
RLITG_TEST_SEED=5 python -m pytest tests/test_torm_gpu.py -q -m gpu -k benchmark_shape > gpurun_out/s19/pytest5.log 2>&1; tail -1 gpurun_out/s19/pytest5.log
