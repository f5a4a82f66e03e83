set -x
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r3d_gputest.log 2>&1
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r3d_smoke.log 2>&1
( time python bench.py ) > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err
( time python bench.py --impl reference ) > gpurun_out/r3d_bench_ref.json 2> gpurun_out/r3d_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3d_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r3d_bench_under_ncu.json 2> gpurun_out/r3d_bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:mh_step_tc16 -s 2 -c 1 -f -o gpurun_out/prof_r3_k1t16_pad64 tools/c_driver/amh_c_driver 60 65536 2 500 2 > gpurun_out/r3d_ncu_pad64.log 2>&1
ncu -i gpurun_out/prof_r3_k1t16_pad64.ncu-rep --page raw --csv > gpurun_out/prof_r3_k1t16_pad64_raw.csv
ncu -i gpurun_out/prof_r3_k1t16_pad64.ncu-rep --page source --csv > gpurun_out/prof_r3_k1t16_pad64_src.csv
for f in gpurun_out/r3d_gputest.log gpurun_out/r3d_smoke.log gpurun_out/r3d_bench.err gpurun_out/r3d_bench_ref.err gpurun_out/r3d_ncu_pad64.log; do tail -n 4 $f; done
