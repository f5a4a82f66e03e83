for n in 256 1024 4096 16384 65536; do AMH_BENCH_N=$n AMH_BENCH_DIMS=32,10 python tools/bench_configs.py c2 2>&1; done | tee gpurun_out/r3c_c2_nchains.txt
