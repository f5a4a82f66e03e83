for np in "" 1; do
echo "== AMH_LOGISTIC_NO_PAD=$np"
AMH_LOGISTIC_NO_PAD=$np python - <<'P'
import sys, os, numpy as np
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
if not os.environ.get("AMH_LOGISTIC_NO_PAD"): os.environ.pop("AMH_LOGISTIC_NO_PAD", None)
import amh_b200 as amh
from bench_configs import timed
eng = amh.default_engine(0)
seeds = lambda n, s: np.random.default_rng(s).integers(0, 2 ** 64, size=n, dtype=np.uint64)
for d, nrows, n in ((20, 2000, 16384), (50, 5000, 16384), (100, 10000, 16384)):
    rng = np.random.default_rng(128)
    X = rng.normal(size=(nrows, d)) / np.sqrt(d)
    y = (rng.random(nrows) < 1 / (1 + np.exp(-X @ rng.normal(size=d)))).astype(float)
    t = amh.LogisticRegressionTarget(X, y, tau=10.0)
    sg2 = 0.002
    for tag, s in (("RWMH", amh.RWMH(amh.MvNormal(np.zeros(d), (0.05 ** 2) * amh.I))), ("MALA", amh.MALA(lambda g: amh.MvNormal(0.5 * sg2 * g, sg2 * amh.I)))):
        run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, seeds(n, 1), np.zeros((d, n)))
        ms = timed(run, 4, spl=2)
        print(f"{tag} x logistic d={d} rows={nrows} chains={n}: {n * 4 / (ms * 1e-3):.4g} chain-steps/s", flush=True)
        run.close()
P
done
