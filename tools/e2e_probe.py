import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import amh_b200 as amh
import bench
d, n, spl = 32, 65536, 500
target, sampler, Sigma = bench.make_problem(amh, d)
eng = amh.default_engine(0)
L = np.linalg.cholesky(Sigma)
hinit = eng.pinned_empty((d, n)); hinit[...] = L @ np.random.default_rng(1).normal(size=(d, n))
pout = eng.pinned_empty((2, d + 1, n)); pacc = eng.pinned_empty((2, n), dtype=np.uint8)
model = amh.DensityModel(target)
def call(i):
    return amh.sample(model, sampler, amh.MCMCB200(device=0, gather=False), 2, n, initial_params=hinit, thinning=spl, chain_type=amh.Chains, seed=i, out=(pout, pacc))
call(0)
t0 = time.perf_counter()
for i in range(5): call(i)
print("e2e ms/call", (time.perf_counter() - t0) / 5 * 1e3)
# phases
import advancedmh_jl_b200.sampling as S
seeds = np.random.default_rng(0).integers(0, 2**64, size=n, dtype=np.uint64)
for rep in range(3):
    t = [time.perf_counter()]
    th = eng.target_of(target); sh = sampler.lower(eng, d); t.append(time.perf_counter())
    run = eng.run(th, sh, n, seeds, hinit); t.append(time.perf_counter())
    eng.sync(); t.append(time.perf_counter())
    out, acc, summ = run.sample(2, 0, spl, 0, store=True, store_accepted=True, summary=False, chain_means=False, out=pout, acc=pacc); t.append(time.perf_counter())
    run.close(); sh.close(); th.close(); t.append(time.perf_counter())
    print("handles %.3f  run_create %.3f  sync %.3f  run_sample %.3f  close %.3f ms" % tuple((b - a) * 1e3 for a, b in zip(t, t[1:])))
t0 = time.perf_counter(); x = S._initial_matrix(hinit, sampler, d, n, True); print("initial_matrix ms", (time.perf_counter() - t0) * 1e3, x is hinit)
