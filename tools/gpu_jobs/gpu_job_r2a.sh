set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2a_pytest.log
tools/run_sanitizers.sh r2a
# K3L: launch list and one full capture
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_k3l.csv python tools/bench_configs.py c4 > gpurun_out/r2_k3l_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mala_logistic -s 20 -c 1 -f -o gpurun_out/prof_r2_k3l python tools/bench_configs.py c4 > gpurun_out/r2_k3l_full.log 2>&1
python tools/bench_configs.py c2 c3 c4 c5 > gpurun_out/r2a_configs.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt
