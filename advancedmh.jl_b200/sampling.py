"""`sample` -- the AbstractMCMC call surface of the multi-chain path
(README.md:135-148, test/runtests.jl:96-110; loop semantics SURVEY.md A.1):

    sample([rng,] model, sampler, N; kw...)
    sample([rng,] model, sampler, parallel, N, nchains; kw...)

All chains advance in lock-step on the GPU(s); `parallel` only says how many.
Keyword arguments keep the reference's names: initial_params, discard_initial,
thinning, num_warmup, chain_type, param_names, progress (ignored), callback."""
from __future__ import annotations

import os
import sys
import time
from dataclasses import dataclass

import numpy as np

from . import _capi as K
from .models import as_target
from .samplers import Ensemble, MALA, MHSampler, RobustAdaptiveMetropolis


_TRACE = os.environ.get("AMH_TRACE") is not None     # phase timings of `sample` on stderr (the library prints its own)

# ------------------------------------------------------------------ parallel
class MCMCSerial:
    pass


class MCMCThreads:
    pass


class MCMCDistributed:
    pass


@dataclass
class MCMCB200:
    """The ensemble type the Julia shim adds (`MCMCB200 <: AbstractMCMCEnsemble`): run all chains on
    B200s.  Under torchrun (one process per GPU) the chains are sharded in contiguous blocks
    across ranks; there is no per-step collective."""
    device: int | None = None
    gather: bool = True
    # ngpus = k: ALL chains in THIS process on k devices of the box through the library's multi-GPU job (amh_job_*):
    # the chains are sharded by the library, the target is broadcast once (NCCL over NVLink), every device writes its
    # column block of the output array directly; no torch.distributed involved.  This is what the Julia shim's
    # `MCMCB200(ngpus = 8)` lowers to.  `devices` optionally names the CUDA device indices.
    ngpus: int | None = None
    devices: tuple | None = None
    # "fp64" (default): every kernel bit-exact against the oracle.  "bf16x2": OPT-IN tensor-core arithmetic with a stated
    # tolerance -- MALA on the many-row logistic target (dim 128) runs its design-matrix contractions as split-bf16
    # tcgen05 GEMMs; anything else raises.
    dtype: str = "fp64"
    # this rank's chains as `streams` contiguous shards, each with its own context (stream, copy stream) on the same
    # GPU and driven by its own host thread: the host->device copy of one shard's initial parameters and the
    # device->host copy of its samples overlap the stepping kernels of the others.  Same results (global chain identity).
    streams: int = 1
    # RAM: a failed rank-1 downdate raises PosDefException after the run, like the reference (RAM :170); True keeps the
    # samples instead (the chain stopped adapting at that step, nothing else happened to it)
    ignore_failed_downdates: bool = False


# ---------------------------------------------------------------- chain types
class Transition:
    """Transition(params, lp, accepted)  (src/AdvancedMH.jl:61-65)"""
    __slots__ = ("params", "lp", "accepted")
    def __init__(self, params, lp, accepted):
        self.params, self.lp, self.accepted = params, lp, accepted


class Chains:
    """MCMCChains.Chains look-alike: value[iter, param..+lp, chain]
    (ext/AdvancedMHMCMCChainsExt.jl:24-38, 93-121)."""
    def __init__(self, value, names, start=1, thin=1, accepted=None, info=None):
        self.value = value
        self.names = list(names)
        self.start, self.thin = start, thin
        self.accepted = accepted
        self.info = info or {}
    def range(self):
        n = self.value.shape[0]
        return range(self.start, self.start + self.thin * n, self.thin)
    def __getitem__(self, name):
        return self.value[:, self.names.index(name), :]
    @property
    def parameters(self):
        return self.names[:-1]
    def array(self):
        """Array(chain): iterations x parameters, chains stacked, internals (lp) dropped"""
        v = self.value[:, :-1, :]
        return np.concatenate([v[:, :, c] for c in range(v.shape[2])], axis=0)


class StructArray(dict):
    """StructArrays look-alike (ext/AdvancedMHStructArraysExt.jl:12-27): one array per field"""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class TransitionVector:
    """Vector{Transition} view over the stored samples of ONE chain (or, for Ensemble, of the walkers)"""
    def __init__(self, value, accepted):
        self.value, self.accepted = value, accepted
    def __len__(self):
        return self.value.shape[0]
    def __getitem__(self, i):
        v = self.value[i]
        if v.shape[1] == 1:
            return Transition(v[:-1, 0].copy(), float(v[-1, 0]), bool(self.accepted[i, 0]))
        return [Transition(v[:-1, c].copy(), float(v[-1, c]), bool(self.accepted[i, c])) for c in range(v.shape[1])]


# --------------------------------------------------------------------- engine
_ENGINES: dict = {}


def default_engine(device: int | None = None) -> K.Engine:
    """The CUDA engine of this process (LOCAL_RANK selects the GPU under torchrun)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _ENGINES:
        _ENGINES[device] = K.Engine(device=device)
    return _ENGINES[device]


_STREAM_ENGINES: dict = {}
_POOL = None
_JOBS: dict = {}


def _job_for(eng: K.Engine, ngpus: int, devices=None) -> K.Job:
    """the multi-GPU job of this process for (library, device list): created once (communicator set-up is the expensive
    part), reused by every `sample` call"""
    key = (eng.path, eng.prefix, ngpus, None if devices is None else tuple(devices))
    if key not in _JOBS:
        _JOBS[key] = eng.job(ngpus, devices)
    return _JOBS[key]


def _stream_engines(eng: K.Engine, k: int):
    """`k` engines (contexts) on eng's device and library; eng itself is the first"""
    key = (eng.path, eng.prefix, eng.device)
    lst = _STREAM_ENGINES.setdefault(key, [eng])
    while len(lst) < k:
        lst.append(K.Engine(lib_path=eng.path, prefix=eng.prefix, device=eng.device))
    return lst[:k]


def _sample_block(eng, target, sampler, dim, n_blk, seeds, init, offset, N, discard_initial, thinning, num_warmup, out, acc):
    th = eng.target_of(target)
    sh = sampler.lower(eng, dim)
    run = None
    try:
        run = eng.run(th, sh, n_blk, seeds, init, chain_offset=offset)
        run.sample(N, discard_initial, thinning, num_warmup, store=True, store_accepted=True, summary=False,
                   chain_means=False, out=out, acc=acc)
        return run.launch_count()
    finally:
        if run is not None:
            run.close()
        sh.close()
        th.close()


def _dist_info():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), dist
    except Exception:
        pass
    return 0, 1, None


def _draw_seeds(rng, nchains: int, lo: int, hi: int):
    """seeds[lo:hi] of `seeds = rand(rng, UInt, nchains)` (AbstractMCMC's per-chain seeding) WITHOUT drawing
    the other ranks' seeds: chain c always gets raw 64-bit output number c of the generator, so the result does not
    depend on the sharding, and a rank's host work is O(local chains).  Needs a jumpable bit generator (numpy's default
    PCG64 family); any other generator falls back to drawing all of them.  The caller's rng ends up advanced by
    `nchains` draws either way."""
    bg = rng.bit_generator
    if isinstance(bg, (np.random.PCG64, np.random.PCG64DXSM)):
        # one raw 64-bit output per seed: `random_raw` gives the same values as integers(0, 2**64) without its range
        # handling (1.7x faster), and `advance` jumps over the other ranks' blocks
        st = bg.state
        if lo:
            bg.advance(lo)
        local = bg.random_raw(hi - lo) if hi > lo else np.empty(0, dtype=np.uint64)
        bg.state = st
        bg.advance(nchains)
        return local
    return rng.integers(0, 2 ** 64, size=nchains, dtype=np.uint64)[lo:hi]


class SamplerState:
    """What `callback(rng, model, sampler, sample, state, iteration)` receives as `state` (AbstractMCMC's callback
    signature; test/RobustAdaptiveMetropolis.jl:11-28 records it per saved sample): the sampler state of ALL local
    chains right after the saved step, chains on the last axis.  Fields follow the reference's state structs --
    Transition (src/AdvancedMH.jl:61-65): params, lp, accepted; GradientTransition (MALA.jl:14-19): + gradient;
    RobustAdaptiveMetropolisState (RAM :99-114): x, logprob, S, logalpha (`logα`), eta (`η`), iteration, isaccept."""
    def __init__(self, st, dim, ram):
        self.params = self.x = st["x"]
        self.lp = self.logprob = st["lp"]
        self.accepted = self.isaccept = st["accepted"].astype(bool)
        self.gradient = st.get("grad")
        self.naccept = st["naccept"]
        self.step = st["step"]
        self._S, self._dim = st.get("S"), dim
        if ram:
            self.logalpha, self.eta = st["logalpha"], st["eta"]
            self.iteration = st["step"] + 1                  # state.iteration starts at 1 (RAM :211)
            self.failed = st["failed"].astype(bool)

    def S(self, chain=0):
        """dense lower-triangular factor of one chain (RobustAdaptiveMetropolisState.S)"""
        d = self._dim
        out = np.zeros((d, d))
        out[np.tril_indices(d)] = self._S[:, chain]
        return out

    def S_diag(self):
        """eigvals(S) of every chain = the diagonal of the triangular factor (RAM :239-245): (dim, nchains)"""
        d = self._dim
        return self._S[[i * (i + 1) // 2 + i for i in range(d)], :]


def _sample_with_callback(run, sampler, rng, model, N, discard_initial, thinning, num_warmup, callback, dim, out, acc):
    """the AbstractMCMC schedule driven save point by save point, so that `callback` sees the real sampler state after
    every saved step (one state read-back per saved sample: the price of the reference's per-sample callback)"""
    is_ram, is_mala = isinstance(sampler, RobustAdaptiveMetropolis), isinstance(sampler, MALA)
    for i in range(N):
        k = discard_initial if i == 0 else thinning
        while k > 0:
            done = run.state_step()
            wu = done < num_warmup
            m = min(k, num_warmup - done) if wu else k
            run.steps(m, warmup=wu)
            k -= m
        st = run.state(grad=is_mala, S=is_ram)
        out[i, :dim, :] = st["x"]
        out[i, dim, :] = st["lp"]
        acc[i] = st["accepted"]
        callback(rng, model, sampler, out[i], SamplerState(st, dim, is_ram), i + 1)
    return out, acc


def shard_bounds(nunits: int, rank: int, world: int):
    """contiguous block partition of `nunits` chains (or ensembles) over `world` ranks"""
    base, rem = divmod(nunits, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class LocalParams:
    """`initial_params=LocalParams(block)`: the initial parameters of THIS rank's chains only, (dim, n_local) float64 in
    the device layout (chains fastest; n_local counts walkers for an Ensemble).  Under torchrun every rank then builds
    and uploads O(local chains) of host data instead of the whole (dim, nchains) matrix."""
    def __init__(self, block):
        self.block = block


def _initial_matrix(initial_params, sampler, dim, nchains, multi):
    """-> (dim, nchains_total) float64 or None; mirrors AbstractMCMC: a multi-chain call takes one entry per
    chain.  Fast path: a (dim, nchains_total) float64 array in the device layout is used as is."""
    if initial_params is None:
        return None
    nw = sampler.n_walkers if isinstance(sampler, Ensemble) else 1
    ip = initial_params
    if isinstance(ip, np.ndarray) and ip.ndim == 2 and ip.shape == (dim, nchains * nw) and (multi or nw > 1 or dim == 1):
        return ip if (ip.dtype == np.float64 and ip.flags.c_contiguous) else np.ascontiguousarray(ip, dtype=np.float64)
    if not multi:
        ip = [ip]
    if len(ip) != nchains:
        raise ValueError("initial_params must have one entry per chain")
    cols = []
    for entry in ip:
        e = np.asarray(entry, dtype=np.float64)
        if nw == 1:
            e = e.reshape(-1)
            if e.size != dim:
                raise ValueError(f"initial_params entry has length {e.size}, model dimension is {dim}")
            cols.append(e[:, None])
        else:
            e = e.reshape(nw, dim)
            cols.append(e.T)
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def sample(*args, **kw):
    """see `_sample`; `MCMCB200(dtype=...)` selects the arithmetic the samplers are lowered with"""
    par = next((a for a in args if isinstance(a, MCMCB200)), None)
    if par is not None and par.dtype != "fp64":
        with K.precision(par.dtype):
            return _sample(*args, **kw)
    return _sample(*args, **kw)


def _sample(*args, rng=None, seed=None, initial_params=None, discard_initial=None, thinning=1, num_warmup=0,
           chain_type=None, param_names=None, progress=False, callback=None, engine=None, store=True,
           summary=False, steps_per_launch=0, out=None, initial_state=None, save_state=False):
    """See module docstring.  Extra device-side keywords (never on the samplers): `engine`
    (a _capi.Engine; default = the CUDA library), `store=False` + `summary=True` for runs whose
    samples cannot be stored (SURVEY.md 7 hard part 7), `out=(values, accepted)` caller-owned (ideally
    pinned, Engine.pinned_empty) buffers of this rank's shard: (N, dim+1, n_local) float64, (N, n_local) uint8.
    `initial_state=` (AbstractMCMC keyword): a state dict of ALL chains as returned in `info["state"]` by a call with
    `save_state=True`; the run then continues the old one bit for bit -- its first sample is one `step` from that
    state, the step counter (hence the noise stream, RAM's `iteration` and the warm-up boundary) carries on."""
    _t = [time.perf_counter()] if _TRACE else None
    def lap(what):
        if _TRACE:
            now = time.perf_counter()
            print(f"[amh.py] sample {what:22s} +{(now - _t[-1]) * 1e3:7.3f} ms", file=sys.stderr)
            _t.append(now)
    args = list(args)
    if args and isinstance(args[0], np.random.Generator):
        rng = args.pop(0)
    if len(args) == 3:
        model, sampler, N = args
        parallel, nchains, multi = None, 1, False
    elif len(args) == 5:
        model, sampler, parallel, N, nchains = args
        multi = True
    else:
        raise TypeError("sample(model, sampler, N) or sample(model, sampler, parallel, N, nchains)")
    if not isinstance(sampler, MHSampler):
        raise TypeError("sampler must be an AdvancedMH sampler")
    target = as_target(model)
    dim = target.dim
    N, nchains = int(N), int(nchains)
    if discard_initial is None:
        discard_initial = num_warmup          # AbstractMCMC default (RAM docstring :41-44)
    if isinstance(sampler, RobustAdaptiveMetropolis) and not hasattr(target, "kind"):
        raise ValueError("RobustAdaptiveMetropolis needs a LogDensityProblems-style target")
    rank, world, dist = (0, 1, None)
    in_process_job = isinstance(parallel, MCMCB200) and parallel.ngpus is not None
    if isinstance(parallel, MCMCB200) and not in_process_job:
        rank, world, dist = _dist_info()
    if rng is None:
        if seed is None and world > 1:
            # every rank must key the SAME global seed vector: rank 0 draws the master seed, everybody gets it
            box = [int(np.random.SeedSequence().entropy) if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            seed = box[0]
        rng = np.random.default_rng(seed)
    nw = sampler.n_walkers if isinstance(sampler, Ensemble) else 1
    lo, hi = shard_bounds(nchains, rank, world)
    # seeds = rand(rng, UInt, nchains) with chain c <- draw number c, so results do not depend on the sharding;
    # only this rank's block [lo, hi) is materialised
    seeds = _draw_seeds(rng, nchains, lo, hi)
    if isinstance(initial_params, LocalParams):
        init = np.asarray(initial_params.block)
        if init.shape != (dim, (hi - lo) * nw):
            raise ValueError(f"LocalParams block must have shape {(dim, (hi - lo) * nw)}, got {init.shape}")
    else:
        init = _initial_matrix(initial_params, sampler, dim, nchains, multi)
        if init is not None:
            init = init[:, lo * nw:hi * nw]                    # this rank's column block (a view: no copy)
    eng = engine or default_engine(parallel.device if isinstance(parallel, MCMCB200) else None)
    if in_process_job:
        eng = _job_for(eng, int(parallel.ngpus), parallel.devices)     # same surface as an Engine, arrays job-wide

    lap("seeds + init + engine")
    n_local = (hi - lo) * nw
    nstreams = parallel.streams if isinstance(parallel, MCMCB200) else 1
    if (nstreams > 1 and not in_process_job and hi - lo >= nstreams and store and not summary and initial_state is None and not save_state
            and callback is None and getattr(target, "kind", None) != K.TARGET_USER):
        # shards of this rank's chains on separate streams of the GPU, one host thread each
        global _POOL
        from concurrent.futures import ThreadPoolExecutor
        if _POOL is None:
            _POOL = ThreadPoolExecutor(max_workers=16)
        engs = _stream_engines(eng, nstreams)
        if out is None:
            out = (np.empty((N, dim + 1, n_local)), np.empty((N, n_local), dtype=np.uint8))
        vals, accs = out
        futs = []
        for k, e in enumerate(engs):
            a, b = shard_bounds(hi - lo, k, nstreams)
            ca, cb = (lo + a) * nw, (lo + b) * nw                      # global chain (walker) range of the shard
            futs.append(_POOL.submit(_sample_block, e, target, sampler, dim, cb - ca, seeds[a:b],
                                     None if init is None else init[:, ca - lo * nw:cb - lo * nw], ca, N, discard_initial, thinning, num_warmup,
                                     vals[:, :, ca - lo * nw:cb - lo * nw], accs[:, ca - lo * nw:cb - lo * nw]))
        launches = sum(f.result() for f in futs)
        out, acc = vals, accs
        info = dict(summary=None, launches=launches, rank=rank, world=world, chains=(lo * nw, hi * nw), streams=nstreams)
        n_local = -1                                                   # done: skip the single-stream path below
    th = eng.target_of(target) if n_local >= 0 else None
    sh = sampler.lower(eng, dim) if n_local >= 0 else None
    run = None
    try:
        if n_local < 0:
            pass
        elif n_local > 0:
            sl = slice(lo * nw, hi * nw)
            if initial_state is not None:
                if "seeds" in initial_state:
                    allseeds = np.asarray(initial_state["seeds"], dtype=np.uint64)
                    if allseeds.shape != (nchains,):
                        raise ValueError("initial_state was saved for a different number of chains")
                    seeds = allseeds[lo:hi]
                init = np.asarray(initial_state["x"], dtype=np.float64)
                if init.shape != (dim, nchains * nw):
                    raise ValueError(f"initial_state['x'] must have shape {(dim, nchains * nw)}")
                init = init[:, sl]
            lap("target + sampler")
            run = eng.run(th, sh, n_local, seeds, init, chain_offset=lo * nw)
            lap("run_create")
            if initial_state is not None:
                run.set_state({k: (v if k == "step" or v is None else np.asarray(v)[..., sl])
                               for k, v in initial_state.items() if k != "seeds"})
                run.steps(1, warmup=run.state_step() < num_warmup)
            if callback is not None and store:
                vals = np.empty((N, dim + 1, n_local)) if out is None else out[0]
                accs = np.empty((N, n_local), dtype=np.uint8) if out is None else out[1]
                out, acc = _sample_with_callback(run, sampler, rng, model, N, discard_initial, thinning, num_warmup,
                                                 callback, dim, vals, accs)
                summ = None
            else:
                out, acc, summ = run.sample(N, discard_initial, thinning, num_warmup, store=store,
                                            store_accepted=store, summary=summary or not store, chain_means=False,
                                            out=None if out is None else out[0], acc=None if out is None else out[1])
            lap("run.sample")
            if isinstance(sampler, RobustAdaptiveMetropolis):
                nfail, first, _ = run.ram_failed()
                if nfail and not getattr(parallel, "ignore_failed_downdates", False):
                    raise K.PosDefException(first, nfail)        # lowrankdowndate throws (RAM :170)
            info = dict(summary=summ, launches=run.launch_count(), rank=rank, world=world,
                        chains=(lo * nw, hi * nw))
            if save_state:
                stt = run.state(grad=isinstance(sampler, MALA), S=isinstance(sampler, RobustAdaptiveMetropolis))
                stt["seeds"] = seeds
                info["state"] = stt
        else:
            out = np.empty((N, dim + 1, 0)) if store else None
            acc = np.empty((N, 0), dtype=np.uint8) if store else None
            info = dict(summary=None, launches=0, rank=rank, world=world, chains=(0, 0))
    finally:
        if run is not None:
            run.close()
        if sh is not None:
            sh.close()
        if th is not None:
            th.close()

    lap("close")
    if world > 1 and isinstance(parallel, MCMCB200) and parallel.gather and store:
        out, acc = _gather_samples(dist, out, acc, nchains * nw, world, nw)
    if world > 1 and isinstance(parallel, MCMCB200) and parallel.gather and save_state:
        # the resumable state of ALL chains on every rank (chains are the last axis of every per-chain array)
        parts = [None] * world
        dist.all_gather_object(parts, info.get("state"))
        parts = [p for p in parts if p is not None]
        info["state"] = {k: (parts[0][k] if k == "step" else
                             None if parts[0][k] is None else np.concatenate([p[k] for p in parts], axis=-1))
                         for k in parts[0]}

    if not store:
        return info
    if param_names is None and getattr(sampler, "names", None):
        param_names = sampler.names          # NamedTuple of proposals: the field names (src/AdvancedMH.jl:80-104)
    names = list(param_names) if param_names is not None else [f"param_{i + 1}" for i in range(dim)]
    if len(names) != dim:
        raise ValueError("param_names must have one entry per parameter")
    if chain_type is None:
        return TransitionVector(out, acc) if not multi else [TransitionVector(out[:, :, c * nw:(c + 1) * nw], acc[:, c * nw:(c + 1) * nw]) for c in range(out.shape[2] // nw)]
    if chain_type is Chains or chain_type == "Chains":
        return Chains(out, names + ["lp"], start=discard_initial + 1, thin=thinning, accepted=acc, info=info)
    if chain_type is StructArray or chain_type == "StructArray":
        sa = StructArray({nm: out[:, i, :].reshape(-1, order="F") if out.shape[2] > 1 else out[:, i, 0] for i, nm in enumerate(names)})
        sa["lp"] = out[:, dim, :].reshape(-1, order="F") if out.shape[2] > 1 else out[:, dim, 0]
        return sa
    if chain_type == "namedtuples":
        keys = names + ["lp"]
        return [dict(zip(keys, out[i, :, 0])) for i in range(out.shape[0])]
    raise ValueError("chain_type must be None, Chains, StructArray or 'namedtuples'")


def _gather_samples(dist, out, acc, ntotal, world, nw):
    """final gather of every rank's chains to all ranks (NCCL all_gather over NVLink when the
    tensors live on the GPUs; gloo on CPU in the tests)."""
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    N, dp1 = out.shape[0], out.shape[1]
    counts = [(shard_bounds(ntotal // nw, r, world)[1] - shard_bounds(ntotal // nw, r, world)[0]) * nw for r in range(world)]
    cmax = max(counts)
    buf = torch.zeros((N, dp1, cmax), dtype=torch.float64, device=dev)
    buf[:, :, :out.shape[2]] = torch.from_numpy(out).to(dev)
    abuf = torch.zeros((N, cmax), dtype=torch.uint8, device=dev)
    abuf[:, :acc.shape[1]] = torch.from_numpy(acc).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    accs = [torch.empty_like(abuf) for _ in range(world)]
    dist.all_gather(outs, buf)
    dist.all_gather(accs, abuf)
    full = np.concatenate([o[:, :, :c].cpu().numpy() for o, c in zip(outs, counts)], axis=2)
    facc = np.concatenate([a[:, :c].cpu().numpy() for a, c in zip(accs, counts)], axis=1)
    return full, facc
