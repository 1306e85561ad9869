mkdir -p gpurun_out/s17
for k in 0 1 2 3 4 5 6 7; do RLITG_TEST_SEED=$k python -m pytest tests -q -m gpu > gpurun_out/s17/seed$k.log 2>&1; echo "seed $k: $(tail -1 gpurun_out/s17/seed$k.log)"; done
