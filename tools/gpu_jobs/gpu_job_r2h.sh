set -x
python -m pytest tests -m gpu -x -q -k "ram or RAM or c5 or contract" 2>&1 | tail -6 > gpurun_out/r2h_pytest_ram.log
for v in p8 p8m2 p16m2; do AMH_LIB=tools/ubench/lib_rams_$v.so python tools/bench_configs.py c5 2>&1 | sed "s/^/stream_$v /" >> gpurun_out/r2h_c5.txt; done
