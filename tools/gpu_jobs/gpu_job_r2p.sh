python tools/job_e2e_bisect.py 2 2>&1 | grep -v "^c5\|^parity" | tail -n 20 | tee gpurun_out/r2p_bisect.txt
python tools/job_e2e_bisect.py 2 --torch-first 2>&1 | grep -v "^c5\|^parity" | tail -n 20 | tee gpurun_out/r2p_bisect_torch_first.txt
