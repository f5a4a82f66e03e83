set -x
for rep in 1 2; do
for lib in default nosplit nopace; do
  if [ $lib = default ]; then unset AMH_LIB; else export AMH_LIB=tools/ubench/lib_$lib.so; fi
  AMH_BENCH_LONG=1 AMH_BENCH_DIMS=32,16 python tools/bench_configs.py c2 2>&1 | sed "s/^/$lib /" >> gpurun_out/r2f_ab.txt
done
done
unset AMH_LIB
ncu --set full --clock-control none --import-source on -k regex:mh_step -s 2 -c 1 -f -o gpurun_out/prof_r2_k1t16_v2 tools/c_driver/amh_c_driver 32 65536 2 500 2 > gpurun_out/r2f_ncu_k1t16_v2.log 2>&1
