#!/usr/bin/env python3
"""bench.py -- chain-steps/sec of the many-chain Metropolis-Hastings hot path.

Workload (BASELINE.json configs[1], "C2"): RWMH on a d=32 full-covariance MvNormal target,
65 536 chains PER GPU (weak scaling: chains are independent, no data-path collective), proposal
MvNormal(0, 2.38^2/d * Sigma) through its full Cholesky factor, fp64 throughout.

A bench "step" is ONE launch of the fused step kernel over all local chains; it advances every
chain by `--mcmc-steps-per-launch` MCMC steps (default 500 = 0.5 % of the 100 000-iteration C2 job; the
library's amh_run_steps call).  Between timed steps L2 is flushed (a 512 MB buffer is rewritten) because
the 17 MB chain state is L2 resident.  `e2e` is the same step through the public `sample()` call with
HOST buffers: initial parameters, seeds and target go host->device, two saved samples of every chain
come device->host, handle creation and destruction included.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this engine
  python bench.py --impl reference ...                           # the CPU arm (oracle port; Julia is not installed)

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chain-steps/sec (all chains) on d=32 MvNormal"
UNIT = "chain-steps/s"


def make_problem(amh, d, seed=32):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    lam = np.exp(np.linspace(0.0, np.log(100.0), d))
    Sigma = (Q * lam) @ Q.T
    Sigma = (Sigma + Sigma.T) / 2
    target = amh.MvNormalTarget(None, Sigma)
    sampler = amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sigma))
    return target, sampler, Sigma


def algorithmic_bytes_per_chain_step(d, T=8):
    # state read + state write, nothing else
    return 2 * (d + 1) * T


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU during the timed region (NVML)"""
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)
    def result(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_baseline(amh, d, spl, seconds=12.0, nchains=2048):
    """the oracle port timed on this box's host cores on a bounded sample of the same workload"""
    orc = amh.Engine(lib_path=os.path.join(ROOT, "oracle", "libamh_oracle.so"), prefix="amho_")
    cores = int(orc.lib.amho_get_threads())
    target, sampler, Sigma = make_problem(amh, d)
    seeds = np.random.default_rng(7).integers(0, 2 ** 64, size=nchains, dtype=np.uint64)
    run = orc.run(orc.target(target.kind, d, target.blob()), sampler.lower(orc, d), nchains, seeds)
    run.steps(5)
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < seconds:
        run.steps(spl)
        steps += spl
    dt = time.perf_counter() - t0
    run.close()
    return {"value": nchains * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nchains} chains x {steps} MCMC steps of the same d={d} workload, C++ oracle (restatement of "
                      f"mh-core.jl:92-117; Julia is not installed), std::thread over chains"}


def run_reference(args):
    """--impl reference: the reference's CPU path.  Julia cannot run here, so this is the oracle port
    with all host threads, on the same config/metric; each step is a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import amh_b200 as amh
    d, spl = args.dim, args.mcmc_steps_per_launch
    orc = amh.Engine(lib_path=os.path.join(ROOT, "oracle", "libamh_oracle.so"), prefix="amho_")
    cores = int(orc.lib.amho_get_threads())
    nchains = args.ref_chains
    target, sampler, _ = make_problem(amh, d)
    seeds = np.random.default_rng(7).integers(0, 2 ** 64, size=nchains, dtype=np.uint64)
    run = orc.run(orc.target(target.kind, d, target.blob()), sampler.lower(orc, d), nchains, seeds)
    for _ in range(args.warmup):
        run.steps(spl)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run.steps(spl)
    dt = time.perf_counter() - t0
    value = nchains * spl * args.steps / dt
    sample = (f"{nchains} chains x {spl} MCMC steps per bench step (bounded sample of the 65536-chain workload; "
              f"chain-steps/s is chain-count independent once every core is busy)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2: RWMH MvNormal d={d}, full-Cholesky proposal", "sampler": "RWMH",
                   "chains_timed": nchains, "mcmc_steps_per_launch": spl},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dim", type=int, default=32)
    ap.add_argument("--chains", type=int, default=65536, help="chains per GPU")
    ap.add_argument("--mcmc-steps-per-launch", type=int, default=500)
    ap.add_argument("--ref-chains", type=int, default=16384)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-streams", type=int, default=1,
                    help="MCMCB200(streams=...) of the e2e call: shards of a rank's chains on separate streams, so that "
                         "host<->device copies overlap the stepping kernels")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    # stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import amh_b200 as amh
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    d, n, spl = args.dim, args.chains, args.mcmc_steps_per_launch
    target, sampler, Sigma = make_problem(amh, d)
    blob = torch.from_numpy(target.blob()).to(dev)
    if world > 1:
        # one-time NCCL broadcast of the target's fixed data (north star); every rank then builds its shard
        dist.broadcast(blob, src=0)
    eng = amh.default_engine(local)
    th = eng.target(target.kind, d, blob.cpu().numpy())
    sh = sampler.lower(eng, d)
    # global chain identity: seeds are drawn for ALL chains, rank r owns [r*n, (r+1)*n)
    seeds_all = np.random.default_rng(20261017).integers(0, 2 ** 64, size=n * world, dtype=np.uint64)
    L = np.linalg.cholesky(Sigma)
    init = np.ascontiguousarray(L @ np.random.default_rng(100 + rank).normal(size=(d, n)))
    run = eng.run(th, sh, n, seeds_all[rank * n:(rank + 1) * n], init, chain_offset=rank * n)

    flush_buf = None if args.no_flush else torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def flush():
        if flush_buf is not None:
            flush_buf.add_(1)
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush()
        run.steps(spl, steps_per_launch=spl)
        run.sync()
    run.kernel_time_ms(reset=True)
    launches0 = run.launch_count()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        run.sync()

    clk = ClockSampler(local)
    barrier()
    clk.start()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush()
        run.steps(spl, steps_per_launch=spl)      # device time is taken by CUDA events on the launching stream
        run.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk.stop_flag = True
    ms, nl = run.kernel_time_ms(reset=True)
    launches = run.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    chain_steps = float(n) * spl * args.steps * world
    value = chain_steps / (ms_max * 1e-3)

    # ---- e2e: the public sample() call, host buffers in, host samples out --------------------
    e2e = None
    if args.e2e_steps > 0:
        init_all = np.concatenate([L @ np.random.default_rng(100 + r).normal(size=(d, n)) for r in range(world)], axis=1)
        hinit = eng.pinned_empty((d, n * world))
        hinit[...] = init_all
        pout = eng.pinned_empty((2, d + 1, n))          # this rank's shard of the two saved samples
        pacc = eng.pinned_empty((2, n), dtype=np.uint8)
        model = amh.DensityModel(target)
        amh.sample(model, sampler, amh.MCMCB200(device=local, gather=False, streams=args.e2e_streams), 2, n * world, initial_params=hinit,
                   thinning=spl, chain_type=amh.Chains, seed=99, out=(pout, pacc))          # warm-up call
        barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            ch = amh.sample(model, sampler, amh.MCMCB200(device=local, gather=False, streams=args.e2e_streams), 2, n * world,
                            initial_params=hinit, thinning=spl, chain_type=amh.Chains, seed=i, out=(pout, pacc))
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_val = float(n) * world * spl * args.e2e_steps / float(tt.item())
        h2d = 8 * d * n + 8 * n + target.blob().nbytes + 8 * (d * (d + 1) // 2)       # per rank: init, seeds, target, L
        d2h = 2 * (d + 1) * n * 8 + 2 * n                                             # per rank: 2 samples + accepted flags
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "call": f"sample(model, RWMH(MvNormal), MCMCB200(streams={args.e2e_streams}), N=2, nchains; thinning=spl, out=pinned) incl. handle "
                       "creation, H2D of initial_params/seeds/target from pinned host memory, D2H of 2 samples into pinned memory"}
        del hinit, ch, pout, pacc

    B = algorithmic_bytes_per_chain_step(d)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # per-rank achieved bandwidth of the step kernel (algorithmic bytes / mean launch duration)
    achieved = B * float(n) * spl / (ms / max(1, nl) * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_mh_step.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "kernel": "mh_step_tc16_kernel<32,28,true,true> (K1T16: DMMA mat-vecs, 16 chains per warp, one 28-warp CTA per SM)",
                "algorithmic_bytes_per_chain_step": B,
                "chain_steps_per_launch": n * spl,
                "note": "state (17 MB) is L2 resident and the kernel is bound by the shared FP64 datapath (DMMA + DFMA + the wide integer multiplies of Philox, 64 FMA/clk/SM); the HBM figure is the contractual denominator (SURVEY.md 8d)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(amh, d, spl)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C2: RWMH MvNormal d={d}, {n} chains per GPU, full-Cholesky proposal, fp64",
                       "chains_per_gpu": n, "mcmc_steps_per_launch": spl,
                       "l2": "state L2-resident by nature; L2 flushed (512 MB rewrite) between timed steps" if not args.no_flush else "no flush",
                       "parallelism": f"chains sharded x{world}, no per-step collective"},
            "clocks": clk.result(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "wall_s_timed_region": t_wall,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    run.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
