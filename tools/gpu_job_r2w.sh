set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "rwmh_mvnormal or padded or tensor_core or static_mh or config2" 2>&1 | tail -n 15
AMH_BENCH_DIMS=7,9,11,13,14,15,17,18,19,21,22,23,25,26,27,28,29,30,31,32,16 python tools/bench_configs.py c2 2>&1 | tee gpurun_out/r2w_c2_dims_padded.txt
AMH_BENCH_LONG=1 AMH_BENCH_DIMS=32,28 python tools/bench_configs.py c2 2>&1 | tee -a gpurun_out/r2w_c2_dims_padded.txt
