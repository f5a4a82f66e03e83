set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r3f_bench_n8.json 2> gpurun_out/r3f_bench_n8.err
tail -n 4 gpurun_out/r3f_bench_n8.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5 --no-configs ) > gpurun_out/r3f_bench_n4.json 2> gpurun_out/r3f_bench_n4.err
tail -n 4 gpurun_out/r3f_bench_n4.err
