"""Do step kernels of two contexts (streams) on one GPU overlap?  Times k shards of the C2 workload launched back to back
from one host thread, against one unsharded launch."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amh_b200 as amh
import bench
d, n, spl = 32, 65536, 500
target, sampler, Sigma = bench.make_problem(amh, d)
seeds = np.random.default_rng(0).integers(0, 2 ** 64, size=n, dtype=np.uint64)
init = np.linalg.cholesky(Sigma) @ np.random.default_rng(1).normal(size=(d, n))
for k in (1, 2, 4, 8):
    engs = [amh.Engine(device=0) for _ in range(k)]
    runs = []
    for j, e in enumerate(engs):
        a, b = j * n // k, (j + 1) * n // k
        runs.append(e.run(e.target_of(target), sampler.lower(e, d), b - a, seeds[a:b], init[:, a:b], chain_offset=a))
    for r in runs: r.steps(spl, steps_per_launch=spl)
    for r in runs: r.sync()
    t0 = time.perf_counter()
    for rep in range(3):
        for r in runs: r.steps(spl, steps_per_launch=spl)
    for r in runs: r.sync()
    dt = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    for rep in range(3):
        for r in runs: r.steps(spl, steps_per_launch=spl)
        for r in runs: r.sync()                      # no overlap between the reps: every rep pays its own tail
    dt2 = (time.perf_counter() - t0) / 3
    print(f"shards {k}: {dt * 1e3:.3f} ms per {spl} steps of all {n} chains -> {n * spl / dt:.4g} chain-steps/s;  "
          f"with a barrier after every rep: {dt2 * 1e3:.3f} ms -> {n * spl / dt2:.4g}", flush=True)
    for r in runs: r.close()
