set -x
timeout 300 python -m pytest tests/test_parity_baseline_gpu.py -m gpu -x -q -k "tensor" 2>&1 | tail -25 > gpurun_out/r2i_pytest_tensor.log
timeout 200 python tools/bench_configs.py c4t c4 > gpurun_out/r2i_c4t.txt 2>&1
