set -x
N=${1:-8}
python tools/job_check.py $N > gpurun_out/r2d_job_check_n$N.txt 2> gpurun_out/r2d_job_check_n$N.err
tail -5 gpurun_out/r2d_job_check_n$N.err
# phase trace of one in-process e2e call at N GPUs
AMH_TRACE=1 python - > gpurun_out/r2d_trace_n$N.txt 2>&1 <<PY
import sys, numpy as np
sys.path.insert(0, ".")
import amh_b200 as amh, bench
N = $N
d, per, spl = 32, 65536, 500
t, s, Sg = bench.make_problem(amh, d)
eng = amh.default_engine(0)
n = per * N
hinit = eng.pinned_empty((d, n)); hinit[...] = np.random.default_rng(0).normal(size=(d, n))
pout = eng.pinned_empty((2, d + 1, n)); pacc = eng.pinned_empty((2, n), dtype=np.uint8)
for i in range(3):
    print("---- call", i, file=sys.stderr)
    amh.sample(amh.DensityModel(t), s, amh.MCMCB200(ngpus=N), 2, n, initial_params=hinit, thinning=spl, chain_type=amh.Chains, seed=i, out=(pout, pacc))
PY
# the torchrun (one process per GPU) bench line for comparison
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2d_bench_n$N.json 2> gpurun_out/r2d_bench_n$N.err
tail -3 gpurun_out/r2d_bench_n$N.err
