python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "static" 2>&1 | tail -n 8
python tools/dim_cliff_probe.py static 2>&1 | tee gpurun_out/r3a_static_dims.txt
