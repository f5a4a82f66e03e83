python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
for n in 256 1024 4096 16384 32768 49151 49152 65536; do AMH_BENCH_N=$n AMH_BENCH_DIMS=32,13 python tools/bench_configs.py c2 2>&1; done | tee gpurun_out/r3h_c2_nchains_small_cta.txt
python -c "import __graft_entry__ as g; g.smoke()"
