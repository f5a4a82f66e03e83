set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
# A/B: explicit warp barrier before the in-place overwrite of the K1T16 tile
AMH_BENCH_LONG=1 AMH_BENCH_DIMS=32 python tools/bench_configs.py c2 > gpurun_out/r2b_ab_default.txt 2>&1
AMH_LIB=tools/ubench/lib_syncwarp.so AMH_BENCH_LONG=1 AMH_BENCH_DIMS=32 python tools/bench_configs.py c2 > gpurun_out/r2b_ab_syncwarp.txt 2>&1
AMH_BENCH_LONG=1 AMH_BENCH_DIMS=32 python tools/bench_configs.py c2 >> gpurun_out/r2b_ab_default.txt 2>&1
AMH_LIB=tools/ubench/lib_syncwarp.so AMH_BENCH_LONG=1 AMH_BENCH_DIMS=32 python tools/bench_configs.py c2 >> gpurun_out/r2b_ab_syncwarp.txt 2>&1
# DRAM traffic of the timed launch shape: 65536 chains, 500 fused steps
ncu --set full --clock-control none --import-source on -k regex:mh_step -s 2 -c 1 -f -o gpurun_out/prof_r2_k1t16_500 tools/c_driver/amh_c_driver 32 65536 2 500 2 > gpurun_out/r2b_ncu_k1t16_500.log 2>&1
