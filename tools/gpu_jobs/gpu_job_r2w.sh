set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "rwmh_mvnormal or padded or tensor_core or static_mh or config2" 2>&1 | tail -n 15
AMH_BENCH_DIMS=33,36,40,44,48,50,56,60,64,32 python tools/bench_configs.py c2 2>&1 | tee gpurun_out/r2w_c2_dims_wide.txt
AMH_BENCH_LONG=1 AMH_BENCH_DIMS=32,64,48 python tools/bench_configs.py c2 2>&1 | tee -a gpurun_out/r2w_c2_dims_wide.txt
