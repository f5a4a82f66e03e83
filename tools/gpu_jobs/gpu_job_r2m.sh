set -x
ncu --set full --clock-control none --import-source on -k regex:stretch_plan_res -s 3 -c 1 -f -o gpurun_out/prof_r2_plan python tools/bench_configs.py c3 > gpurun_out/r2m_ncu_plan.log 2>&1
ncu -i gpurun_out/prof_r2_plan.ncu-rep --page raw --csv > gpurun_out/prof_r2_plan_raw.csv
ncu -i gpurun_out/prof_r2_plan.ncu-rep --page source --csv > gpurun_out/prof_r2_plan_src.csv
