python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "logistic" 2>&1 | tail -n 12
timeout 500 python tools/dim_cliff_probe.py logistic 2>&1 | tee gpurun_out/r3j_logistic.txt
