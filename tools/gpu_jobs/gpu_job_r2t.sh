set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "stretch" 2>&1 | tail -n 5
bash tools/c3_shape_sweep3.sh > gpurun_out/r2t_c3_shape_sweep3.txt 2>&1
cat gpurun_out/r2t_c3_shape_sweep3.txt
