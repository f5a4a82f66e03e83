python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "mala" 2>&1 | tail -n 8
python tools/dim_cliff_probe.py mala 2>&1 | tee gpurun_out/r2y_mala_dims.txt
