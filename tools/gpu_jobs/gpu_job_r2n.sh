set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r2n_bench_ref_n2.json 2> gpurun_out/r2n_bench_ref_n2.err
python tools/job_check.py 2 > gpurun_out/r2n_job_check_2gpu.txt 2>&1
tail -n 3 gpurun_out/r2n_bench_n2.err gpurun_out/r2n_bench_ref_n2.err gpurun_out/r2n_job_check_2gpu.txt
