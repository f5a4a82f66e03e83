"""Multi-GPU check on real GPUs (torchrun, NCCL): every rank runs its contiguous block of chains on its own B200, the
final all_gather rebuilds the full [N, d+1, nchains] array on every rank, and the result equals the run of ALL chains on
one GPU bit for bit (global chain identity, no per-step collective).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_gather_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amh_b200 as amh

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
d, n, N = 8, 4096 + 64 + 5, 6
rng = np.random.default_rng(1)
Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
Sigma = (Q * np.linspace(1, 20, d)) @ Q.T
Sigma = (Sigma + Sigma.T) / 2
model = amh.DensityModel(amh.MvNormalTarget(None, Sigma))
cases = {
    "rwmh": (amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma)), n, dict(thinning=5)),
    "stretch": (amh.Ensemble(128, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I))), 7, dict(thinning=3)),
    "ram": (amh.RobustAdaptiveMetropolis(), 333, dict(num_warmup=20, initial_params=np.zeros((d, 333)))),
}
ok = True
for name, (spl, nch, kw) in cases.items():
    full = amh.sample(model, spl, amh.MCMCB200(device=local, gather=True), N, nch, seed=7, chain_type=amh.Chains, **kw)
    # the same job on this rank's GPU alone (the process group is bypassed by handing sample() a plain engine and MCMCThreads)
    single = amh.sample(model, spl, amh.MCMCThreads(), N, nch, seed=7, chain_type=amh.Chains, engine=amh.default_engine(local), **kw)
    same = np.array_equal(full.value, single.value) and np.array_equal(full.accepted, single.accepted)
    ok = ok and same
    if rank == 0:
        print(f"{name}: gathered {full.value.shape} from {world} ranks == single-GPU run: {same}", flush=True)
t = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ALL OK" if int(t.item()) == 1 else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
