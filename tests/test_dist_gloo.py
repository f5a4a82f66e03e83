"""N>1 path on CPU: world_size-2 `gloo` processes shard the chains (contiguous blocks, global chain identity,
no data-path collective) and gather the samples at the end.  The engine here is the CPU oracle (no GPU in
this container); on the B200 box bench.py runs the same host logic with the CUDA engine over NCCL."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import torch.distributed as dist
import amh_b200 as amh
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
orc = amh.Engine(lib_path=os.path.join({root!r}, "oracle", "libamh_oracle.so"), prefix="amho_")
out = {{}}
# independent chains: 7 chains over 2 ranks (uneven split 4 + 3)
target = amh.MvNormalTarget(None, np.array([[2.0, 0.3], [0.3, 1.0]]))
ch = amh.sample(np.random.default_rng(3), target, amh.RWMH(2), amh.MCMCB200(), 25, 7, chain_type=amh.Chains, engine=orc,
                discard_initial=5, thinning=2)
out["mh"] = ch.value; out["mh_acc"] = ch.accepted
out["mh_chains"] = np.array(ch.info["chains"])
# ensembles are indivisible: 3 ensembles of 8 walkers over 2 ranks (2 + 1)
spl = amh.Ensemble(8, amh.StretchProposal(amh.MvNormal(np.zeros(2), amh.I)))
ch = amh.sample(np.random.default_rng(4), target, spl, amh.MCMCB200(), 10, 3, chain_type=amh.Chains, engine=orc)
out["st"] = ch.value
out["st_chains"] = np.array(ch.info["chains"])
# no gather: each rank keeps its shard
ch = amh.sample(np.random.default_rng(3), target, amh.RWMH(2), amh.MCMCB200(gather=False), 25, 7, chain_type=amh.Chains,
                engine=orc, discard_initial=5, thinning=2)
out["mh_local"] = ch.value
# resume across ranks: the gathered state of all chains continues the run on a sharded call
a = amh.sample(np.random.default_rng(3), target, amh.RWMH(2), amh.MCMCB200(), 10, 7, chain_type=amh.Chains, engine=orc,
               save_state=True)
b = amh.sample(np.random.default_rng(5), target, amh.RWMH(2), amh.MCMCB200(), 8, 7, chain_type=amh.Chains, engine=orc,
               initial_state=a.info["state"])
out["resume"] = np.concatenate([a.value, b.value])
out["state_x"] = a.info["state"]["x"]
np.savez(os.path.join({tmp!r}, f"rank{{rank}}.npz"), **out)
dist.destroy_process_group()
"""


@pytest.mark.timeout(300)
def test_world_size_2_gloo_sharding_and_gather(tmp_path, amh, oracle):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, tmp=str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    r0 = np.load(tmp_path / "rank0.npz"); r1 = np.load(tmp_path / "rank1.npz")
    # single-process result of the same call (same rng => same seeds for ALL chains)
    target = amh.MvNormalTarget(None, np.array([[2.0, 0.3], [0.3, 1.0]]))
    ref = amh.sample(np.random.default_rng(3), target, amh.RWMH(2), amh.MCMCSerial(), 25, 7, chain_type=amh.Chains,
                     engine=oracle, discard_initial=5, thinning=2)
    for r in (r0, r1):
        assert np.array_equal(r["mh"], ref.value) and np.array_equal(r["mh_acc"], ref.accepted)
    assert list(r0["mh_chains"]) == [0, 4] and list(r1["mh_chains"]) == [4, 7]
    assert np.array_equal(r0["mh_local"], ref.value[:, :, :4]) and np.array_equal(r1["mh_local"], ref.value[:, :, 4:])
    ref = amh.sample(np.random.default_rng(3), target, amh.RWMH(2), amh.MCMCSerial(), 18, 7, chain_type=amh.Chains, engine=oracle)
    for r in (r0, r1):
        assert np.array_equal(r["resume"], ref.value) and r["state_x"].shape == (2, 7)
    spl = amh.Ensemble(8, amh.StretchProposal(amh.MvNormal(np.zeros(2), amh.I)))
    ref = amh.sample(np.random.default_rng(4), target, spl, amh.MCMCSerial(), 10, 3, chain_type=amh.Chains, engine=oracle)
    assert np.array_equal(r0["st"], ref.value) and np.array_equal(r1["st"], ref.value)
    assert list(r0["st_chains"]) == [0, 16] and list(r1["st_chains"]) == [16, 24]
