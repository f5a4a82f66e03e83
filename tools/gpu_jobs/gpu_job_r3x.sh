python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
run() { echo "$1 d=$2 nw=$3 ne=$4: $(AMH_C3_TARGET=$1 AMH_C3_D=$2 AMH_C3_NW=$3 AMH_C3_NE=$4 timeout 300 python tools/bench_configs.py c3 2>&1 | tail -1 | cut -c1-120)"; }
( run ros 14 4096 64; run ros 32 2048 64; run mvn 32 2048 64 ) | tee gpurun_out/r3x_stretch_dims.txt
