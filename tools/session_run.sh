mkdir -p gpurun_out/s18
python tools/coll_seed_stats.py 0 1 2 3 4 5 6 7 > gpurun_out/s18/coll_seed_stats.txt 2>&1; cat gpurun_out/s18/coll_seed_stats.txt | tail -10
