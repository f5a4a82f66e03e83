set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r3n_bench_n2.json 2> gpurun_out/r3n_bench_n2.err
tail -n 4 gpurun_out/r3n_bench_n2.err
python -c "
import json
j=json.loads(open('gpurun_out/r3n_bench_n2.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['n_gpus'])
for k,v in j['configs'].items(): print(k, v.get('value'), v.get('n_gpus'))
"
python tools/job_check.py 2 2>/dev/null | grep "^e2e"
