timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "stretch" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_c3_k2r.csv python tools/bench_configs.py c3 > gpurun_out/c3_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stretch_sweep_res -s 3 -c 1 -f -o gpurun_out/prof_r2_k2r python tools/bench_configs.py c3 > gpurun_out/c3_ncu.log 2>&1
ncu -i gpurun_out/prof_r2_k2r.ncu-rep --page raw --csv > gpurun_out/prof_r2_k2r_raw.csv
ncu -i gpurun_out/prof_r2_k2r.ncu-rep --page source --csv > gpurun_out/prof_r2_k2r_src.csv
ncu -i gpurun_out/prof_r2_k2r.ncu-rep --page details > gpurun_out/prof_r2_k2r_details.txt
