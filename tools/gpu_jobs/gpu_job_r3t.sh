python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8
python - <<'P'
import sys, numpy as np
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import amh_b200 as amh
from bench_configs import timed, spd
eng = amh.default_engine(0)
seeds = lambda n, s: np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)
for d, nrows, n in ((128, 10000, 16384), (20, 2000, 16384)):
    rng = np.random.default_rng(128)
    X = rng.normal(size=(nrows, d)) / np.sqrt(d)
    y = (rng.random(nrows) < 1 / (1 + np.exp(-X @ rng.normal(size=d)))).astype(float)
    t = amh.LogisticRegressionTarget(X, y, tau=10.0)
    s = amh.RWMH(amh.MvNormal(np.zeros(d), (0.05 ** 2 / 1.0) * spd(d, 7, 0.5, 2.0)))
    run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, seeds(n, 1), np.zeros((d, n)))
    ms = timed(run, 8, spl=4)
    st = run.state()
    print(f"RWMH full-covariance proposal x logistic d={d} rows={nrows} chains={n}: {n * 8 / (ms * 1e-3):.4g} chain-steps/s accept={st['naccept'].sum() / (n * st['step']):.3f}", flush=True)
    run.close()
P
