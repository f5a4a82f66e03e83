python tools/blas_spin_probe.py 2 2>&1 | tail -n 5 | tee gpurun_out/r2q_blas_spin_probe.txt
AMH_JOB_SPIN_US=0 python tools/blas_spin_probe.py 2 2>&1 | tail -n 5 | tee gpurun_out/r2q_blas_spin_probe_sleepers.txt
python tools/job_check.py 2 2>/dev/null | grep "^e2e"
