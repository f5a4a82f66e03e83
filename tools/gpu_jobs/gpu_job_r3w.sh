set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 30 --warmup 5 ) > gpurun_out/r3w_bench_n8.json 2> gpurun_out/r3w_bench_n8.err
tail -n 4 gpurun_out/r3w_bench_n8.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r3w_bench_ref_n8.json 2> gpurun_out/r3w_bench_ref_n8.err
tail -n 4 gpurun_out/r3w_bench_ref_n8.err
