set -x
nproc
python tools/job_check.py 8 > gpurun_out/r2r_job_check_8gpu.txt 2> gpurun_out/r2r_job_check_8gpu.err
grep "^e2e\|^c5\|^bcast" gpurun_out/r2r_job_check_8gpu.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err
tail -n 4 gpurun_out/r2r_bench_n8.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r2r_bench_ref_n8.json 2> gpurun_out/r2r_bench_ref_n8.err
tail -n 4 gpurun_out/r2r_bench_ref_n8.err
