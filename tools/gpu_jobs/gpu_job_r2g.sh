set -x
python -m pytest tests -m gpu -x -q -k "ram or RAM or golden or job or c5 or contract" 2>&1 | tail -15 > gpurun_out/r2g_pytest_ram.log
AMH_RAMW_FORCE_REDO=1 python -m pytest tests -m gpu -x -q -k "ram_bit_exact or c5" 2>&1 | tail -5 >> gpurun_out/r2g_pytest_ram.log
for p in warp stream; do AMH_RAM_PATH=$p python tools/bench_configs.py c5 2>&1 | sed "s/^/$p /" >> gpurun_out/r2g_c5.txt; done
AMH_LIB=tools/ubench/lib_rams4.so python tools/bench_configs.py c5 2>&1 | sed "s/^/stream_minb4 /" >> gpurun_out/r2g_c5.txt
AMH_RAMS_CTAS=2 python tools/bench_configs.py c5 2>&1 | sed "s/^/stream_2ctas /" >> gpurun_out/r2g_c5.txt
