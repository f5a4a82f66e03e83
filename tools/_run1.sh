for cfg in "256 1024" "1024 256" "2048 128" "512 148" "128 4096"; do
  set -- $cfg
  for res in 1 0; do
    echo "nw=$1 ne=$2 res=$res: $(AMH_C3_NW=$1 AMH_C3_NE=$2 AMH_STRETCH_RES=$res timeout 300 python tools/bench_configs.py c3 2>&1 | tail -1)"
  done
done
