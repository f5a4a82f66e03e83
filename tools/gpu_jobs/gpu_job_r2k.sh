set -x
tools/run_sanitizers.sh r2b
ncu --set full --clock-control none --import-source on -k regex:mala_tensor_kernel -s 70 -c 1 -f -o gpurun_out/prof_r2_k3t python tools/bench_configs.py c4t > gpurun_out/r2k_ncu_k3t.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_launches_k3t.csv python tools/bench_configs.py c4t > gpurun_out/r2k_k3t_launch.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
