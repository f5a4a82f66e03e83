AMH_TRACE=1 python tools/job_e2e_trace.py 2 3 > gpurun_out/r2o_job_trace.txt 2>&1
tail -n 60 gpurun_out/r2o_job_trace.txt
