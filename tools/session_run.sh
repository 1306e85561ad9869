mkdir -p gpurun_out/s11
for k in 4 0 6 12 13 14; do RLITG_TEST_SEED=$k python -m pytest tests/test_torm_gpu.py -q -m gpu > gpurun_out/s11/seed$k.log 2>&1; echo "seed $k: $(tail -1 gpurun_out/s11/seed$k.log)"; done
