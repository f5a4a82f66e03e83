python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
python tools/dim_cliff_probe.py rwros rwgp 2>&1 | tee gpurun_out/r2z_rw_dims.txt
