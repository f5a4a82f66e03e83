run() { echo "$1 d=$2 nw=$3 ne=$4 res=$5: $(AMH_C3_TARGET=$1 AMH_C3_D=$2 AMH_C3_NW=$3 AMH_C3_NE=$4 AMH_STRETCH_RES=$5 timeout 300 python tools/bench_configs.py c3 2>&1 | tail -1 | cut -c1-110)"; }
for r in 1 0; do run ros 16 2048 64 $r; done
for r in 1 0; do run ros 16 4096 64 $r; done
for r in 1 0; do run ros 12 4096 64 $r; done
for r in 1 0; do run mvn 12 2048 64 $r; done
for r in 1 0; do run mvn 10 4096 64 $r; done
for r in 1 0; do run ros 20 2048 64 $r; done
