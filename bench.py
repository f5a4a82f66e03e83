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

The same line carries `configs.{c3,c4,c5}`: short runs of the other BASELINE configs through the C ABI (value, algorithmic
bytes, roofline fraction; C4 also the FP64-pipe fraction, C5 warm-up at 1 and 16 fused steps per launch from S = I and from
an adapted S0), every rank running its own copy (weak scaling: C5 at 8 GPUs = 262 144 chains), so that a driver run observes
all five configs.

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


def workload_config(d, n, spl, world, flush=True):
    """the `config` object: identical for this engine and for --impl reference (same workload, same sizes)"""
    return {"workload": f"C2: RWMH MvNormal d={d}, {n} chains per GPU, full-Cholesky proposal, fp64",
            "chains_per_gpu": n, "mcmc_steps_per_launch": spl,
            "l2": "state L2-resident by nature; L2 flushed (512 MB rewrite) between timed steps" if flush else "no flush",
            "parallelism": f"chains sharded x{world}, no per-step collective"}


def cpu_baseline(amh, d, spl, seconds=12.0, nchains=65536):
    """the oracle port timed on this box's host cores on a bounded sample of the same workload: the SAME 65 536 chains,
    fewer MCMC steps (about `seconds` of CPU work)"""
    orc = amh.Engine(lib_path=os.path.join(ROOT, "oracle", "libamh_oracle.so"), prefix="amho_")
    cores = int(orc.lib.amho_get_threads())
    target, sampler, Sigma = make_problem(amh, d)
    seeds = np.random.default_rng(7).integers(0, 2 ** 64, size=nchains, dtype=np.uint64)
    run = orc.run(orc.target(target.kind, d, target.blob()), sampler.lower(orc, d), nchains, seeds)
    run.steps(2)
    chunk = 20
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < seconds:
        run.steps(chunk)
        steps += chunk
    dt = time.perf_counter() - t0
    run.close()
    return {"value": nchains * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nchains} chains x {steps} MCMC steps of the same d={d} workload, C++ oracle (restatement of "
                      f"mh-core.jl:92-117; Julia is not installed), std::thread over chains"}


# ---------------------------------------------------------------- the other BASELINE configs, as sub-records
def _spd(d, seed, lo, hi):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    lam = np.exp(np.linspace(np.log(lo), np.log(hi), d))
    S = (Q * lam) @ Q.T
    return (S + S.T) / 2


def _timed(run, nsteps, warmup=False, spl=0, reps=3, warm=1):
    """mean device time (CUDA events on the launching stream, inside the library) of `nsteps` MCMC steps"""
    for _ in range(warm):
        run.steps(nsteps, warmup=warmup, steps_per_launch=spl)
    run.sync()
    run.kernel_time_ms(reset=True)
    for _ in range(reps):
        run.steps(nsteps, warmup=warmup, steps_per_launch=spl)
    run.sync()
    ms, _ = run.kernel_time_ms(reset=True)
    return ms / reps


# FP64-pipe cycles of one 16-chain warp-step of K1T16 at d = 32, by contract version (instruction mix from the ncu source
# page x per-instruction pipe costs of tools/ubench/issue_probe.cu): (DMMA, DFMA-class, IMAD.WIDE/HI)
PIPE_MIX = {1: (80, 398, 182), 2: (80, 377, 62)}        # v2 measured: profiles/r2_k1t16_v2_ncu_summary.txt
PIPE_COST = (16.2, 2.07, 4.1)

FP64_TFLOPS_PEAK = 148 * 64 * 2 * 1.965e9 / 1e12      # 64 FP64 FMA / clk / SM (DMMA or DFMA, one shared datapath) at 1965 MHz = 37.2


def _rec(chain_steps, ms, bytes_per, peak, **extra):
    v = chain_steps / (ms * 1e-3)
    out = {"value": v, "unit": UNIT, "ms": ms, "algorithmic_bytes_per_chain_step": bytes_per,
           "achieved_gbs": v * bytes_per / 1e9, "frac": v * bytes_per / 1e9 / peak}
    out.update(extra)
    return out


def bench_c3(amh, eng, peak, seed=2):
    d, nw, ne = 10, 4096, 64
    t = amh.RosenbrockTarget(d)
    s = amh.Ensemble(nw, amh.StretchProposal(amh.MvNormal(np.zeros(d), amh.I)))
    sd = np.random.default_rng(seed).integers(0, 2 ** 64, size=ne, dtype=np.uint64)
    run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), nw * ne, sd)
    ms = _timed(run, 128, spl=0)          # 0 = the library's choice: 64 sweeps per launch for this shape (2 x 0.8 GB of plan entries)
    st = run.state()
    out = _rec(nw * ne * 128, ms, 2 * (d + 1) * 8, peak, workload=f"C3: Ensemble(4096, StretchProposal) Rosenbrock d=10, {ne} ensembles per GPU, exact sequential sweep",
               kernel="stretch_plan_res_kernel + stretch_sweep_res_kernel (K2R: ensemble resident in the shared memory of a 2-CTA cluster, "
                      "in-place records, st.async + mbarrier hand-off; plan made ahead on a second stream)",
               accept_rate=float(st["naccept"].sum() / (nw * ne * st["step"])), sweeps_timed=128 * 3, sweeps_per_launch=64)
    run.close()
    return out


def bench_c4(amh, eng, peak, seed=3, n=16384, bf16_peak=1403.9):
    d, nrows = 128, 10000
    rng = np.random.default_rng(128)
    X = rng.normal(size=(nrows, d)) / np.sqrt(d)
    beta = rng.normal(size=d)
    y = (rng.random(nrows) < 1 / (1 + np.exp(-X @ beta))).astype(float)
    t = amh.LogisticRegressionTarget(X, y, tau=10.0)
    s2 = 3.3e-2
    s = amh.MALA(lambda g: amh.MvNormal((s2 / 2) * g, s2 * amh.I))
    sd = np.random.default_rng(seed).integers(0, 2 ** 64, size=n, dtype=np.uint64)
    run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, sd, np.zeros((d, n)))
    run.steps(60, steps_per_launch=4)                  # from zeros to the posterior mode region
    st0 = run.state()
    ms = _timed(run, 4, spl=2, reps=2, warm=1)
    st = run.state()
    tf = 4.0 * nrows * d * n * 4 / (ms * 1e-3) / 1e12  # two 10 000 x 128 mat-vecs per chain-step = 5.12 MFLOP
    out = _rec(n * 4, ms, 2 * (2 * d + 1) * 8, peak, workload=f"C4: MALA logistic regression d=128, 10 000 rows, {n} chains per GPU, analytic device gradient",
               kernel="mala_logistic_kernel<128> (K3L: TMA ring + two chained FP64 DMMA GEMMs)", bound="fp64 tensor",
               fp64_tflops=tf, fp64_pipe_frac=tf / FP64_TFLOPS_PEAK, fp64_tflops_peak=FP64_TFLOPS_PEAK,
               accept_rate_recent=float((st["naccept"].sum() - st0["naccept"].sum()) / (n * (st["step"] - st0["step"]))))
    run.close()
    # the same config on the OPT-IN split-bf16 tcgen05 path (K3T): not bit-exact, tolerance stated in DESIGN.md / the tests
    with amh.precision("bf16x2"):
        run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, sd, np.zeros((d, n)))
    run.steps(60)
    st0 = run.state()
    ms = _timed(run, 20, reps=2, warm=1)
    st = run.state()
    v = n * 20 / (ms * 1e-3)
    issued = 3 * 4.0 * nrows * d * v / 1e12            # three bf16 slice products per contraction
    out["tensor_bf16x2"] = {
        "value": v, "unit": UNIT, "ms": ms, "speedup_vs_fp64_path": v / out["value"],
        "kernel": "mala_tensor_kernel (K3T: tcgen05.mma kind::f16, TMEM accumulators, TMA-staged split-bf16 operands, fp32 link function)",
        "fp64_equivalent_tflops": 4.0 * nrows * d * v / 1e12, "bf16_tflops_issued": issued, "bf16_peak_tflops_sustained": bf16_peak,
        "tensor_frac": issued / bf16_peak, "bit_exact": False,
        "tolerance": "one step from the same state: |d lp| <= 2e-3 + 2e-6 |lp|, |d grad| <= 2e-3 max|grad|, identical fp64 candidates; "
                     "acceptance rate and posterior means agree with the fp64 path (tests/test_parity_baseline_gpu.py::test_mala_tensor_*)",
        "accept_rate_recent": float((st["naccept"].sum() - st0["naccept"].sum()) / (n * (st["step"] - st0["step"])))}
    run.close()
    return out


def bench_c5(amh, eng, peak, seed=4, n=32768):
    d = 64
    Sigma = _spd(d, 64, 1e-4, 1.0)
    t = amh.MvNormalTarget(None, Sigma)
    out = {"workload": f"C5: RobustAdaptiveMetropolis ill-conditioned Gaussian d=64 (eigenvalues 1e-4..1), {n} chains per GPU",
           "kernel": "ram_warp_kernel (K4W: warp per chain, factor in shared memory via bulk copies)", "bound": "hbm"}
    Bw, Bs = 2 * (d + 1) * 8 + d * (d + 1) * 8, 2 * (d + 1) * 8 + d * (d + 1) // 2 * 8
    sd = np.random.default_rng(seed).integers(0, 2 ** 64, size=n, dtype=np.uint64)
    for tag, S0 in (("from_identity", None), ("from_adapted_S0", (2.38 / np.sqrt(d)) * np.linalg.cholesky(Sigma))):
        # S = I is the reference default (RAM :198-199); on this target it does not get moving within any affordable
        # warm-up (acceptance 0: every step is a downdate), so the stationary mix of updates and downdates is timed from
        # the usual 2.38/sqrt(d) scaling of the target's factor as well.  Both are reported.
        s = amh.RobustAdaptiveMetropolis() if S0 is None else amh.RobustAdaptiveMetropolis(S=S0)
        run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, sd, np.zeros((d, n)))
        run.steps(128, warmup=True, steps_per_launch=16)
        st0 = run.state()
        rec = {}
        for spl in (1, 16):
            ms = _timed(run, 32, warmup=True, spl=spl)
            rec[f"warmup_{spl}_per_launch"] = _rec(n * 32, ms, Bw, peak)
        st1 = run.state()
        rec["accept_rate_recent"] = float((st1["naccept"].sum() - st0["naccept"].sum()) / (n * max(1, st1["step"] - st0["step"])))
        for spl in (1, 16):
            ms = _timed(run, 32, warmup=False, spl=spl)
            rec[f"sampling_{spl}_per_launch"] = _rec(n * 32, ms, Bs, peak)
        rec["failed_downdates"] = int(run.ram_failed()[0])
        out[tag] = rec
        run.close()
    # the headline figure of the sub-record: warm-up, one kernel per MCMC step (the genuinely HBM-bound case), adapted start
    h = out["from_adapted_S0"]["warmup_1_per_launch"]
    out.update(value=h["value"], unit=UNIT, frac=h["frac"], algorithmic_bytes_per_chain_step=Bw)
    return out


def bench_off_shape(amh, eng, peak, seed=7):
    """The same samplers OFF the hand-sized shapes of configs 2-5 ("The dimensions in between"): dimensions without
    an exact kernel, few chains, the reference's default StaticMH (issymmetric = false), RWMH / MALA on a logistic regression whose
    feature count is not 32 / 64 / 128.  Short runs; each entry: chain-steps/s and the fraction of the algorithmic HBM roofline."""
    out = {"workload": "off-shape cases: padded tensor-core MH kernels, small-run CTA shape, Hastings-term kernels, exact-dimension MALA, padded tiled logistic kernels"}
    sd = lambda n, k: np.random.default_rng(seed + k).integers(0, 2 ** 64, size=n, dtype=np.uint64)

    def one(tag, t, s, d, n, nsteps, spl, bytes_per, init=None):
        run = eng.run(eng.target(t.kind, d, t.blob()), s.lower(eng, d), n, sd(n, d), init)
        ms = _timed(run, nsteps, spl=spl)
        out[tag] = _rec(n * nsteps, ms, bytes_per, peak)
        run.close()

    for d, n in ((28, 65536), (48, 65536), (100, 65536), (32, 16384), (32, 4096)):
        Sg = _spd(d, 32, 1.0, 100.0)
        one(f"rwmh_mvnormal_d{d}_n{n}", amh.MvNormalTarget(None, Sg), amh.RWMH(amh.MvNormal(np.zeros(d), (2.38 ** 2 / d) * Sg)), d, n, 200, 100,
            2 * (d + 1) * 8)
    d = 32
    Sg = _spd(d, 32, 1.0, 100.0)
    one("staticmh_default_asymmetric_mvnormal_d32", amh.MvNormalTarget(None, Sg), amh.StaticMH(amh.MvNormal(np.zeros(d), 1.3 * Sg)), d, 65536, 200, 100,
        2 * (d + 1) * 8)
    d = 20
    Sg = _spd(d, 32, 0.5, 2.0)
    sg2 = (0.3 / d ** (1 / 3)) ** 2
    one("mala_mvnormal_d20", amh.MvNormalTarget(None, Sg), amh.MALA(lambda g: amh.MvNormal(0.5 * sg2 * g, sg2 * amh.I)), d, 65536, 100, 50,
        2 * (2 * d + 1) * 8, init=np.zeros((d, 65536)))
    d, nrows, n = 20, 2000, 16384
    rng = np.random.default_rng(128)
    X = rng.normal(size=(nrows, d)) / np.sqrt(d)
    y = (rng.random(nrows) < 1 / (1 + np.exp(-X @ rng.normal(size=d)))).astype(float)
    t = amh.LogisticRegressionTarget(X, y, tau=10.0)
    one("rwmh_logistic_d20_2000rows", t, amh.RWMH(amh.MvNormal(np.zeros(d), (0.05 ** 2) * amh.I)), d, n, 8, 4, 2 * (d + 1) * 8, init=np.zeros((d, n)))
    s2 = 0.002
    one("mala_logistic_d20_2000rows", t, amh.MALA(lambda g: amh.MvNormal(0.5 * s2 * g, s2 * amh.I)), d, n, 8, 4, 2 * (2 * d + 1) * 8, init=np.zeros((d, n)))
    out.update(value=out["rwmh_mvnormal_d28_n65536"]["value"], unit=UNIT, frac=out["rwmh_mvnormal_d28_n65536"]["frac"])
    return out


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
    nchains = args.ref_chains if args.ref_chains > 0 else args.chains
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
    sample = (f"{nchains} chains x {spl} MCMC steps per bench step: the whole per-GPU workload of this engine's arm, C++ oracle "
              f"(restatement of mh-core.jl:92-117; Julia is not installed), std::thread over chains")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(d, args.chains, spl, args.gpus, not args.no_flush), "chains_timed": nchains,
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
    ap.add_argument("--ref-chains", type=int, default=0, help="chains of the --impl reference run (0 = --chains: the same config)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3/C4/C5 sub-records")
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
    # global chain identity: chain c is keyed by draw number c of one generator; rank r materialises only its block
    # [r*n, (r+1)*n) (jumpable PCG64): per-rank host work is O(local chains)
    seeds = amh.sampling._draw_seeds(np.random.default_rng(20261017), n * world, rank * n, (rank + 1) * n)
    L = np.linalg.cholesky(Sigma)
    init = np.ascontiguousarray(L @ np.random.default_rng(100 + rank).normal(size=(d, n)))
    run = eng.run(th, sh, n, seeds, init, chain_offset=rank * n)

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
        hinit = eng.pinned_empty((d, n))                # this rank's block of the initial parameters, page-locked
        hinit[...] = init
        linit = amh.LocalParams(hinit)
        pout = eng.pinned_empty((2, d + 1, n))          # this rank's shard of the two saved samples
        pacc = eng.pinned_empty((2, n), dtype=np.uint8)
        model = amh.DensityModel(target)
        amh.sample(model, sampler, amh.MCMCB200(device=local, gather=False, streams=args.e2e_streams), 2, n * world, initial_params=linit,
                   thinning=spl, chain_type=amh.Chains, seed=99, out=(pout, pacc))          # warm-up call
        barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            ch = amh.sample(model, sampler, amh.MCMCB200(device=local, gather=False, streams=args.e2e_streams), 2, n * world,
                            initial_params=linit, thinning=spl, chain_type=amh.Chains, seed=i, out=(pout, pacc))
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
    # DRAM bytes of one launch: only quoted when the committed ncu --set full capture is of THIS launch shape
    # (same chains, same fused steps); anything else is not a measurement of the timed launch -> null
    cv = run.contract()
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_mh_step.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if (int(tj.get("mcmc_steps_in_captured_launch", -1)) == spl and int(tj.get("chains", 65536)) == n and int(tj.get("dim", 32)) == d
                    and int(tj.get("contract_version", 1)) == cv):
                traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("capture")
        except (OSError, ValueError, TypeError):
            traffic = None
    # what really bounds the kernel: the shared FP64 datapath.  Pipe cycles one 16-chain warp-step needs, from the
    # instruction mix of the kernel (ncu source page) x the per-instruction pipe costs measured with tools/ubench/issue_probe.cu
    # (profiles/r1_issue_probe_b200.txt: DMMA 16.2, DFMA-class 2.07, IMAD.WIDE/HI 4.1 cycles of a scheduler's FP64 pipe)
    mix = PIPE_MIX.get(cv) if d == 32 else None
    pipe_cycles_per_warp_step = None if mix is None else sum(m * c for m, c in zip(mix, PIPE_COST))
    fp64_pipe = None
    if pipe_cycles_per_warp_step is not None and n % 16 == 0:
        sm_hz = 1e6 * float(peaks.get("sm_max_mhz", 1965.0))
        us_step = 1e3 * (ms / max(1, nl)) / spl
        warp_steps_per_sched = (n / 16.0) / (148 * 4)
        floor_us = warp_steps_per_sched * pipe_cycles_per_warp_step / sm_hz * 1e6
        fp64_pipe = {"frac": floor_us / us_step, "floor_us_per_mcmc_step": floor_us, "measured_us_per_mcmc_step": us_step,
                     "pipe_cycles_per_16_chain_warp_step": pipe_cycles_per_warp_step,
                     "model": f"{mix[0]} DMMA x 16.2 + {mix[1]} DFMA-class x 2.07 + {mix[2]} IMAD.WIDE x 4.1 cycles (profiles/r1_issue_probe_b200.txt, DESIGN.md 5)"}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "kernel": "mh_step_tc16_kernel<32,28,true,true> (K1T16: DMMA mat-vecs, 16 chains per warp, one 28-warp CTA per SM)",
                "algorithmic_bytes_per_chain_step": B,
                "chain_steps_per_launch": n * spl,
                "fp64_pipe_frac": None if fp64_pipe is None else fp64_pipe["frac"], "fp64_pipe": fp64_pipe,
                "note": "state (17 MB) is L2 resident and the kernel is bound by the shared FP64 datapath (DMMA + DFMA + the wide integer multiplies of Philox, 64 FMA/clk/SM): fp64_pipe_frac is the honest utilisation figure; the HBM figure is the contractual denominator (SURVEY.md 8d)"}

    # ---- the other BASELINE configs (every rank runs its own copy: weak scaling; rank 0 reports the max-time aggregate) ----
    configs = None
    if not args.no_configs:
        configs = {}
        for name, fn in (("c3", bench_c3), ("c4", bench_c4), ("c5", bench_c5), ("off_shape", bench_off_shape)):
            barrier()
            rec = fn(amh, eng, peak, seed=10 * (rank + 1) + len(configs), **({"bf16_peak": float(peaks.get("bf16_tflops_sustained", 1403.9))} if name == "c4" else {}))
            if dist is not None:
                # whole-job value: every rank processed the same number of units; time = max over ranks
                tm = torch.tensor([rec["value"]], dtype=torch.float64, device=dev)
                dist.all_reduce(tm, op=dist.ReduceOp.MIN)
                rec["value_per_gpu_min"] = float(tm.item())
                rec["value"] = float(tm.item()) * world
                rec["n_gpus"] = world
                if "tensor_bf16x2" in rec:
                    tm = torch.tensor([rec["tensor_bf16x2"]["value"]], dtype=torch.float64, device=dev)
                    dist.all_reduce(tm, op=dist.ReduceOp.MIN)
                    rec["tensor_bf16x2"]["value"] = float(tm.item()) * world
            configs[name] = rec
        # config 2 in the reference's default output mode: EVERY step is a sample (thinning = 1), i.e. one launch per MCMC
        # step with the save epilogue (device ring), and in between: 16 steps per launch.  Per rank; wall clock around the
        # launches + sync beside the event-timed kernel durations shows what the launch cadence costs.
        barrier()
        c2s = {"workload": f"C2 with one launch per MCMC step / 16 steps per launch ({n} chains per GPU, d = {d})"}
        for k_per in (1, 16):
            nst = 256
            run.steps(nst, steps_per_launch=k_per)
            run.sync()
            run.kernel_time_ms(reset=True)
            t0 = time.perf_counter()
            run.steps(nst, steps_per_launch=k_per)
            run.sync()
            wall = (time.perf_counter() - t0) * 1e3
            kms, knl = run.kernel_time_ms(reset=True)
            c2s[f"steps_per_launch_{k_per}"] = {
                "value": float(n) * nst / (wall * 1e-3) * world, "unit": UNIT, "us_per_mcmc_step_wall": wall * 1e3 / nst,
                "us_per_mcmc_step_kernels": kms * 1e3 / nst, "launches": int(knl),
                "frac": B * float(n) * nst / (wall * 1e-3) / 1e9 / peak}
        c2s["note"] = ("not CUDA-graphed: the launches are already back to back on one stream (wall ~ sum of the kernel durations); what a 1-step "
                       "launch costs is the kernel's own prologue / epilogue (seeds, lp, counters, the state round trip through L2) and the "
                       "end-of-launch tail of a single-wave kernel, which a graph does not remove")
        configs["c2_launch_cadence"] = c2s

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(amh, d, spl)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "contract_version": cv,
            "config": workload_config(d, n, spl, world, not args.no_flush),
            "clocks": clk.result(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "configs": configs, "wall_s_timed_region": t_wall,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    run.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
