python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
AMH_BENCH_DIMS=65,72,80,96,100,112,128 python tools/bench_configs.py c2 2>&1 | tee gpurun_out/r3b_c2_dims_128.txt
AMH_BENCH_LONG=1 AMH_BENCH_DIMS=32,16 python tools/bench_configs.py c2 2>&1 | tee -a gpurun_out/r3b_c2_dims_128.txt
