ncu --set full --clock-control none --import-source on -k regex:mala_logistic_kernel -s 2 -c 1 -f -o gpurun_out/prof_r3_k3l_rw python tools/dim_cliff_probe.py logistic > gpurun_out/r3z_ncu_k3l_rw.log 2>&1
ncu -i gpurun_out/prof_r3_k3l_rw.ncu-rep --page raw --csv > gpurun_out/prof_r3_k3l_rw_raw.csv
ncu -i gpurun_out/prof_r3_k3l_rw.ncu-rep --page source --csv > gpurun_out/prof_r3_k3l_rw_src.csv
tail -n 3 gpurun_out/r3z_ncu_k3l_rw.log
