timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_baseline_gpu.py -m gpu -x -q -k "stretch or c3" 2>&1 | tail -2
AMH_STRETCH_RES=1 AMH_STRETCH_WIN=7 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "stretch" 2>&1 | tail -1
echo "384: $(timeout 300 python tools/bench_configs.py c3 2>&1 | tail -1)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file gpurun_out/r2_launches_c3_res.csv python tools/bench_configs.py c3 > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/r2_launches_c3_res.csv | awk -F'","' '{print substr($5,1,40), $NF}'
