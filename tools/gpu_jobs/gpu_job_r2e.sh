set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2e_pytest.log
for cv in 1 2; do
  AMH_CONTRACT=$cv AMH_BENCH_LONG=1 AMH_BENCH_DIMS=32,24,16,10,2 python tools/bench_configs.py c2 c2iso > gpurun_out/r2e_c2_v$cv.txt 2>&1
  AMH_CONTRACT=$cv python tools/bench_configs.py c4 c5 > gpurun_out/r2e_c45_v$cv.txt 2>&1
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
ncu --set full --clock-control none --import-source on -k regex:mh_step -s 2 -c 1 -f -o gpurun_out/prof_r2_k1t16_v2 tools/c_driver/amh_c_driver 32 65536 2 500 2 > gpurun_out/r2e_ncu_k1t16_v2.log 2>&1
