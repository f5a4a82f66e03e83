run() { echo "$1 d=$2 nw=$3 ne=$4 res=$5: $(AMH_C3_TARGET=$1 AMH_C3_D=$2 AMH_C3_NW=$3 AMH_C3_NE=$4 AMH_STRETCH_RES=$5 AMH_C3_SPL=${6:-16} timeout 300 python tools/bench_configs.py c3 2>&1 | tail -1 | cut -c1-190)"; }
for d in 6 7 9 12 20 24; do run ros $d 4096 64 1; done
run ros 20 2048 64 1; run ros 20 2048 64 0; run ros 24 2048 64 0; run mvn 12 2048 64 1; run mvn 12 2048 64 0
for s in 8 16 32 64; do run ros 10 4096 64 1 $s; done
