set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
run() { echo "$1 d=$2 nw=$3 ne=$4 res=$5: $(AMH_C3_TARGET=$1 AMH_C3_D=$2 AMH_C3_NW=$3 AMH_C3_NE=$4 AMH_C3_SPL=${6:-0} timeout 300 python tools/bench_configs.py c3 2>&1 | tail -1 | cut -c1-190)"; }
( run ros 10 4096 64 auto; run ros 10 4096 64 auto 128; run mvn 12 2048 64 auto; run mvn 16 2048 64 auto; run ros 10 2048 74 auto; run ros 10 2048 128 auto; run ros 5 4096 64 auto ) > gpurun_out/r2u_c3_defaults.txt 2>&1
cat gpurun_out/r2u_c3_defaults.txt
python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; tail -n 3 gpurun_out/r2u_bench.err
python -c "
import json
j=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['traffic'], j['configs']['c3']['value'], j['configs']['c3']['frac'])
"
