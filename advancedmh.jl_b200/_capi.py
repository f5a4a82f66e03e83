"""ctypes binding of the C ABI declared in include/amh.h.

`Engine()` loads the product library `libamh_b200.so` (hand-written sm_100a
kernels) that sits next to this file and fails loudly when it is missing or
when no CUDA device is present -- there is no CPU fallback.  The same binding
class can be pointed at another library exporting the same ABI under another
prefix; only the test-suite does that (with the CPU oracle, prefix ``amho_``).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "libamh_b200.so")

AMH_OK, AMH_ERR_INVALID, AMH_ERR_CUDA, AMH_ERR_UNSUPPORTED, AMH_ERR_STATE = 0, 1, 2, 3, 4

TARGET_IID_NORMAL, TARGET_MVNORMAL, TARGET_ROSENBROCK, TARGET_LOGISTIC = 1, 2, 3, 4
TARGET_GAUSS_PREC, TARGET_NIG_TOY, TARGET_NIG_TOY_LOG = 5, 6, 7
TARGET_USER = 100
SAMPLER_STATIC, SAMPLER_RW, SAMPLER_STRETCH, SAMPLER_MALA, SAMPLER_RAM, SAMPLER_MIXED = 1, 2, 3, 4, 5, 6
COV_SCALAR, COV_DIAG, COV_FULL, COV_COMPONENTS = 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)


class Component(C.Structure):
    """struct amh_component (include/amh.h)"""
    _fields_ = [("family", C.c_int32), ("rw", C.c_int32), ("symmetric", C.c_int32), ("reserved", C.c_int32),
                ("p0", C.c_double), ("p1", C.c_double), ("logc", C.c_double)]


class SamplerDesc(C.Structure):
    """struct amh_sampler_desc (include/amh.h)"""
    _fields_ = [
        ("kind", C.c_int32), ("dim", C.c_int32), ("symmetric", C.c_int32), ("cov_kind", C.c_int32),
        ("mean", _dp), ("scale", _dp),
        ("stretch_a", C.c_double), ("n_walkers", C.c_int64),
        ("mala_sigma2", C.c_double), ("mala_drift", C.c_double),
        ("ram_alpha", C.c_double), ("ram_gamma", C.c_double),
        ("ram_eig_lo", C.c_double), ("ram_eig_hi", C.c_double),
        ("ram_S0", _dp),
        ("components", C.POINTER(Component)),
        ("contract", C.c_int32), ("precision", C.c_int32),
    ]


class Summary(C.Structure):
    """struct amh_summary (include/amh.h)"""
    _fields_ = [
        ("n_saved", C.c_int64), ("n_steps", C.c_int64), ("accept_rate", C.c_double),
        ("mean", _dp), ("var", _dp), ("chain_mean", _dp),
    ]


class AMHError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class AMHArgumentError(AMHError, ValueError):
    """AMH_ERR_INVALID / AMH_ERR_UNSUPPORTED -- Julia's ArgumentError"""


class AMHStateError(AMHError):
    """AMH_ERR_STATE -- e.g. MALA without initial parameters (MALA.jl:37)"""


class PosDefException(ArithmeticError):
    """LinearAlgebra.PosDefException: RAM's rank-1 downdate left the positive-definite cone
    (`lowrankdowndate`, RobustAdaptiveMetropolis.jl:170).  The reference aborts `sample`; the device flags the chain
    (amh_run_ram_failed) and the host raises after the run.  `.chain` is the first flagged chain, `.count` how many."""
    def __init__(self, chain, count):
        super().__init__(f"matrix is not positive definite; rank-1 downdate failed for {count} chain(s), first: chain {chain}")
        self.chain, self.count = chain, count


# every symbol include/amh.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "version", "last_error", "contract_version", "ctx_create", "ctx_destroy", "ctx_sync",
    "target_create", "target_create_source", "target_destroy", "sampler_create", "sampler_destroy",
    "run_create", "run_destroy", "run_steps", "run_sync", "run_sample", "run_sample_ld",
    "run_get_state", "run_set_params", "run_set_state", "run_get_ram_adapt", "run_set_ram_adapt", "run_ram_failed", "run_contract", "run_dim", "run_nchains", "run_launch_count",
    "run_kernel_time_ms", "host_alloc", "host_free", "run_get_state_ld", "run_set_state_ld",
    "job_create", "job_destroy", "job_ngpus", "job_target_create", "job_target_create_source", "job_broadcast_mode",
    "job_broadcast_ms", "job_comm_init_ms", "job_sampler_create", "job_run_create", "job_run_destroy", "job_run_steps",
    "job_run_sync", "job_run_sample", "job_run_get_state", "job_run_set_state", "job_run_get_ram_adapt",
    "job_run_set_ram_adapt", "job_run_ram_failed", "job_run_contract", "job_run_shard", "job_run_launch_count", "job_run_kernel_time_ms",
]


# contract version new samplers are lowered with when the caller does not say (0 = the library's default, v2);
# tests pin v1 with `with contract(1): ...`
DEFAULT_CONTRACT = 0


class contract:
    """context manager: lower samplers under numerical-contract version `v` inside the block"""
    def __init__(self, v):
        self.v = int(v)
    def __enter__(self):
        global DEFAULT_CONTRACT
        self.old, DEFAULT_CONTRACT = DEFAULT_CONTRACT, self.v
        return self
    def __exit__(self, *exc):
        global DEFAULT_CONTRACT
        DEFAULT_CONTRACT = self.old
        return False


# arithmetic of the design-matrix contractions: 0 = fp64 (bit-exact, default), 1 = split-bf16 tensor-core GEMMs (opt-in,
# stated tolerance; MALA x logistic target x dim 128 only).  `with precision("bf16x2"): ...` or MCMCB200(dtype="bf16x2").
PRECISION_FP64, PRECISION_BF16X2 = 0, 1
DEFAULT_PRECISION = 0


class precision:
    """context manager: lower samplers with `amh_sampler_desc.precision` = p ("fp64" / "bf16x2" or 0 / 1) inside the block"""
    def __init__(self, p):
        self.p = {"fp64": 0, "f64": 0, "bf16x2": 1}.get(p, p)
        if self.p not in (0, 1):
            raise ValueError("precision must be 'fp64' or 'bf16x2'")
    def __enter__(self):
        global DEFAULT_PRECISION
        self.old, DEFAULT_PRECISION = DEFAULT_PRECISION, self.p
        return self
    def __exit__(self, *exc):
        global DEFAULT_PRECISION
        DEFAULT_PRECISION = self.old
        return False


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def sampler_desc(*, kind, dim, symmetric=False, cov_kind=COV_SCALAR, mean=None, scale=None,
                 stretch_a=2.0, n_walkers=0, mala_sigma2=0.0, mala_drift=0.0,
                 ram_alpha=0.234, ram_gamma=0.6, ram_eig_lo=0.0, ram_eig_hi=float("inf"), ram_S0=None,
                 components=None, contract=None, precision=None):
    """-> (amh_sampler_desc, objects that must stay alive while it is used).
    components: list of (family, p0, p1, logc[, rw, symmetric]) -- one univariate law per coordinate
    contract: version of the numerical contract (include/amh_contract.h); None = `DEFAULT_CONTRACT` (0 = library default)"""
    keep = []
    def ptr(a):
        if a is None:
            return None
        a = _as_f64(a).ravel()
        keep.append(a)
        return a.ctypes.data_as(_dp)
    d = SamplerDesc(kind, dim, int(bool(symmetric)), cov_kind, ptr(mean), ptr(scale), float(stretch_a),
                    int(n_walkers), float(mala_sigma2), float(mala_drift), float(ram_alpha), float(ram_gamma),
                    float(ram_eig_lo), float(ram_eig_hi), ptr(ram_S0), None,
                    int(DEFAULT_CONTRACT if contract is None else contract),
                    int(DEFAULT_PRECISION if precision is None else precision))
    if components is not None:
        if len(components) != dim:
            raise AMHArgumentError(AMH_ERR_INVALID, f"need one component per coordinate ({dim}), got {len(components)}")
        arr = (Component * dim)()
        for i, c in enumerate(components):
            fam, p0, p1, logc = c[:4]
            rw, sym = (c[4], c[5]) if len(c) >= 6 else (0, 0)
            arr[i] = Component(int(fam), int(bool(rw)), int(bool(sym)), 0, float(p0), float(p1), float(logc))
        keep.append(arr)
        d.components = C.cast(arr, C.POINTER(Component))
    return d, keep


def _point_at_bundled_nccl():
    """The library dlopen()s NCCL lazily (multi-GPU jobs with an NCCL broadcast only).  If the NCCL wheel PyTorch was built
    against is installed, make that the copy it loads (AMH_NCCL_LIB, unless the caller set it): the loader keys libraries by
    soname, so an older system libnccl.so.2 loaded first would be handed to a later `import torch` and break it."""
    if os.environ.get("AMH_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["AMH_NCCL_LIB"] = cand
                return
    except (ImportError, ValueError, AttributeError):
        pass


class Engine:
    """One library + one device context."""

    def __init__(self, lib_path: str | None = None, prefix: str = "amh_", device: int = 0):
        path = lib_path or PRODUCT_LIB
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "advancedmh.jl_b200 has no CPU fallback.")
        self.path = path
        self.prefix = prefix
        _point_at_bundled_nccl()
        self.lib = C.CDLL(path)
        self._bind()
        ctx = C.c_void_p()
        self._check(self._f("ctx_create")(C.c_int32(device), C.byref(ctx)))
        self.ctx = ctx
        self.device = device

    # -- plumbing ---------------------------------------------------------
    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _bind(self):
        f = self._f
        f("last_error").restype = C.c_char_p
        f("run_nchains").restype = C.c_int64
        f("run_launch_count").restype = C.c_int64
        for n in ("run_nchains", "run_launch_count", "run_dim", "run_contract", "run_sync", "run_destroy", "ctx_sync",
                  "ctx_destroy", "target_destroy", "sampler_destroy"):
            f(n).argtypes = [C.c_void_p]
        f("ctx_create").argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
        f("target_create").argtypes = [C.c_void_p, C.c_int32, C.c_int32, _dp, C.c_int64, C.POINTER(C.c_void_p)]
        f("target_create_source").argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, _dp, C.c_int64,
                                              C.POINTER(C.c_void_p)]
        f("sampler_create").argtypes = [C.c_void_p, C.POINTER(SamplerDesc), C.POINTER(C.c_void_p)]
        f("run_create").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, _u64p, _dp, C.c_int64,
                                    C.POINTER(C.c_void_p)]
        f("run_steps").argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
        f("run_sample").argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _dp, _u8p,
                                    C.POINTER(Summary)]
        f("run_sample_ld").argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int64, _u8p, C.c_int64,
                                       C.POINTER(Summary)]
        f("run_get_state").argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _u8p, _i64p, _i64p]
        f("run_set_params").argtypes = [C.c_void_p, _dp]
        f("run_set_state").argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _u8p, _i64p, C.c_int64]
        f("run_get_ram_adapt").argtypes = [C.c_void_p, _dp, _dp]
        f("run_set_ram_adapt").argtypes = [C.c_void_p, _dp, _dp, _u8p]
        f("run_ram_failed").argtypes = [C.c_void_p, _i64p, _i64p, _u8p]
        f("run_kernel_time_ms").argtypes = [C.c_void_p, C.c_int32, _dp, _i64p]
        f("host_alloc").argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        f("host_free").argtypes = [C.c_void_p]
        f("run_get_state_ld").argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _u8p, _i64p, _i64p]
        f("run_set_state_ld").argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _u8p, _i64p, C.c_int64]
        # multi-GPU job
        f("job_create").argtypes = [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]
        for n in ("job_destroy", "job_ngpus", "job_run_contract", "job_run_destroy", "job_run_sync", "job_run_launch_count", "job_broadcast_mode",
                  "job_broadcast_ms", "job_comm_init_ms"):
            f(n).argtypes = [C.c_void_p]
        f("job_broadcast_mode").restype = C.c_char_p
        f("job_broadcast_ms").restype = C.c_double
        f("job_comm_init_ms").restype = C.c_double
        f("job_run_launch_count").restype = C.c_int64
        f("job_target_create").argtypes = [C.c_void_p, C.c_int32, C.c_int32, _dp, C.c_int64]
        f("job_target_create_source").argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, _dp, C.c_int64]
        f("job_sampler_create").argtypes = [C.c_void_p, C.POINTER(SamplerDesc)]
        f("job_run_create").argtypes = [C.c_void_p, C.c_int64, _u64p, _dp, C.c_int64]
        f("job_run_steps").argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
        f("job_run_sample").argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _dp, _u8p, C.POINTER(Summary)]
        f("job_run_get_state").argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _u8p, _i64p, _i64p]
        f("job_run_set_state").argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _u8p, _i64p, C.c_int64]
        f("job_run_get_ram_adapt").argtypes = [C.c_void_p, _dp, _dp]
        f("job_run_set_ram_adapt").argtypes = [C.c_void_p, _dp, _dp, _u8p]
        f("job_run_ram_failed").argtypes = [C.c_void_p, _i64p, _i64p, _u8p]
        f("job_run_shard").argtypes = [C.c_void_p, C.c_int32, _i64p, _i64p, C.POINTER(C.c_int32)]
        f("job_run_kernel_time_ms").argtypes = [C.c_void_p, C.c_int32, _dp, _i64p]

    def _check(self, rc):
        if rc == AMH_OK:
            return
        msg = (self._f("last_error")() or b"").decode()
        if rc in (AMH_ERR_INVALID, AMH_ERR_UNSUPPORTED):
            raise AMHArgumentError(rc, msg)
        if rc == AMH_ERR_STATE:
            raise AMHStateError(rc, msg)
        raise AMHError(rc, msg)

    def version(self):
        a, b = C.c_int32(), C.c_int32()
        self._f("version")(C.byref(a), C.byref(b))
        return a.value, b.value

    def contract_version(self):
        return int(self._f("contract_version")())

    def sync(self):
        self._check(self._f("ctx_sync")(self.ctx))

    def close(self):
        if getattr(self, "ctx", None):
            self._f("ctx_destroy")(self.ctx)
            self.ctx = None

    def pinned_empty(self, shape, dtype=np.float64):
        """numpy array over page-locked host memory (amh_host_alloc): use it for `initial_params` and for the
        `out=` buffers of `sample` to get full-speed asynchronous host<->device copies.  Freed with the array."""
        import weakref
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self._f("host_alloc")(C.c_size_t(max(nbytes, 1)), C.byref(p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(buf, self._f("host_free"), p)
        return arr

    # -- objects ----------------------------------------------------------
    def target(self, kind: int, dim: int, blob) -> "TargetHandle":
        blob = _as_f64(blob).ravel()
        h = C.c_void_p()
        self._check(self._f("target_create")(self.ctx, kind, dim, blob.ctypes.data_as(_dp), blob.size, C.byref(h)))
        return TargetHandle(self, h, kind, dim)

    def target_source(self, dim: int, source: str, data=None, has_gradient: bool = False) -> "TargetHandle":
        """amh_target_create_source: a log-density given as source text (DensityModel(f), src/AdvancedMH.jl:52-54)"""
        data = np.zeros(0) if data is None else _as_f64(data).ravel()
        h = C.c_void_p()
        self._check(self._f("target_create_source")(self.ctx, dim, source.encode(), int(bool(has_gradient)),
                                                    data.ctypes.data_as(_dp) if data.size else None, data.size, C.byref(h)))
        return TargetHandle(self, h, TARGET_USER, dim)

    def target_of(self, t) -> "TargetHandle":
        """handle for a host-side target description (models.py): catalogue entry or source text"""
        if getattr(t, "kind", None) == TARGET_USER:
            return self.target_source(t.dim, t.source, t.data, t.has_gradient)
        return self.target(t.kind, t.dim, t.blob())

    def sampler(self, **kw) -> "SamplerHandle":
        """keywords: see `sampler_desc`"""
        d, keep = sampler_desc(**kw)
        h = C.c_void_p()
        self._check(self._f("sampler_create")(self.ctx, C.byref(d), C.byref(h)))
        return SamplerHandle(self, h, d.kind, d.dim, int(d.n_walkers))

    def job(self, ngpus: int, devices=None) -> "Job":
        """amh_job_create: `ngpus` devices of this box behind one handle, in this process (MCMCB200(ngpus=k))"""
        return Job(self, ngpus, devices)

    def run(self, target: "TargetHandle", sampler: "SamplerHandle", nchains: int, seeds, init=None,
            chain_offset: int = 0) -> "Run":
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64).ravel()
        init_p, init_ld = None, 0
        if init is not None:
            init = np.asarray(init)
            if init.shape != (target.dim, nchains):
                raise AMHArgumentError(AMH_ERR_INVALID, f"init must have shape (dim, nchains) = {(target.dim, nchains)}, got {init.shape}")
            # a column block of a larger C-contiguous float64 matrix is passed as is (row stride = init_ld)
            ok = (init.dtype == np.float64 and init.strides[1] == 8 and init.strides[0] % 8 == 0
                  and init.strides[0] >= 8 * nchains)
            if not ok:
                init = _as_f64(init)
            init_ld = init.strides[0] // 8
            init_p = C.cast(init.ctypes.data, _dp)
        h = C.c_void_p()
        self._check(self._f("run_create")(self.ctx, target.h, sampler.h, nchains, chain_offset,
                                          seeds.ctypes.data_as(_u64p), init_p, init_ld, C.byref(h)))
        return Run(self, h, target, sampler, nchains)


@dataclass
class TargetHandle:
    eng: Engine
    h: C.c_void_p
    kind: int
    dim: int

    def close(self):
        if self.h:
            self.eng._f("target_destroy")(self.h)
            self.h = None


@dataclass
class SamplerHandle:
    eng: Engine
    h: C.c_void_p
    kind: int
    dim: int
    n_walkers: int

    def close(self):
        if self.h:
            self.eng._f("sampler_destroy")(self.h)
            self.h = None


class Run:
    _p = "run_"            # entry-point family: amh_run_* (one device) or amh_job_run_* (JobRun)

    def __init__(self, eng, h, target, sampler, n):
        self.eng, self.h, self.target, self.sampler, self.n = eng, h, target, sampler, n
        self.dim = target.dim

    def _r(self, name):
        return self.eng._f(self._p + name)

    def steps(self, nsteps: int, warmup: bool = False, steps_per_launch: int = 0):
        self.eng._check(self._r("steps")(self.h, nsteps, int(warmup), steps_per_launch))

    def sync(self):
        self.eng._check(self._r("sync")(self.h))

    def sample(self, N, discard_initial=0, thinning=1, num_warmup=0, store=True, store_accepted=True,
               summary=True, chain_means=False, out=None, acc=None):
        """`out` / `acc`: optional caller buffers ((N, d+1, n) float64 / (N, n) uint8, C-contiguous), e.g. pinned
        ones from Engine.pinned_empty; otherwise pageable numpy arrays are allocated."""
        d, n = self.dim, self.n
        out_ld = acc_ld = n
        if out is not None:
            # C-contiguous, or a block of chains out[:, :, a:b] of a C-contiguous (N, d+1, n_total) array
            ok = (out.shape == (N, d + 1, n) and out.dtype == np.float64 and out.strides[2] == 8 and
                  out.strides[1] % 8 == 0 and out.strides[1] >= 8 * n and out.strides[0] == (d + 1) * out.strides[1])
            if not ok:
                raise AMHArgumentError(AMH_ERR_INVALID, f"out must be a float64 array of shape {(N, d + 1, n)}, C-contiguous or a block of chains of one")
            out_ld = out.strides[1] // 8
        elif store:
            out = np.empty((N, d + 1, n), dtype=np.float64)
        if acc is not None:
            ok = acc.shape == (N, n) and acc.dtype == np.uint8 and acc.strides[1] == 1 and acc.strides[0] >= n
            if not ok:
                raise AMHArgumentError(AMH_ERR_INVALID, f"acc must be a uint8 array of shape {(N, n)}, C-contiguous or a block of chains of one")
            acc_ld = acc.strides[0]
        elif store_accepted:
            acc = np.empty((N, n), dtype=np.uint8)
        summ = None
        s = None
        if summary:
            mean = np.zeros(d); var = np.zeros(d)
            cm = np.zeros((d, n)) if chain_means else None
            s = Summary(0, 0, 0.0, mean.ctypes.data_as(_dp), var.ctypes.data_as(_dp),
                        cm.ctypes.data_as(_dp) if cm is not None else None)
        self._sample_call(N, discard_initial, thinning, num_warmup,
                          C.cast(out.ctypes.data, _dp) if out is not None else None, out_ld,
                          C.cast(acc.ctypes.data, _u8p) if acc is not None else None, acc_ld,
                          C.byref(s) if s is not None else None)
        if s is not None:
            summ = dict(n_saved=s.n_saved, n_steps=s.n_steps, accept_rate=s.accept_rate, mean=mean, var=var,
                        chain_mean=cm)
        return out, acc, summ

    def _sample_call(self, N, discard_initial, thinning, num_warmup, out_p, out_ld, acc_p, acc_ld, summ_p):
        self.eng._check(self._r("sample_ld")(self.h, N, discard_initial, thinning, num_warmup, out_p, out_ld, acc_p, acc_ld, summ_p))

    def state(self, grad=False, S=False):
        d, n = self.dim, self.n
        x = np.empty((d, n)); lp = np.empty(n)
        g = np.empty((d, n)) if grad else None
        Sm = np.empty((d * (d + 1) // 2, n)) if S else None
        acc = np.empty(n, dtype=np.uint8); nacc = np.empty(n, dtype=np.int64)
        step = C.c_int64()
        self.eng._check(self._r("get_state")(
            self.h, x.ctypes.data_as(_dp), lp.ctypes.data_as(_dp),
            g.ctypes.data_as(_dp) if g is not None else None,
            Sm.ctypes.data_as(_dp) if Sm is not None else None,
            acc.ctypes.data_as(_u8p), nacc.ctypes.data_as(_i64p), C.byref(step)))
        st = dict(x=x, lp=lp, grad=g, S=Sm, accepted=acc, naccept=nacc, step=step.value)
        if self.sampler.kind == SAMPLER_RAM:
            # the report-only fields of RobustAdaptiveMetropolisState (RAM :107-113) and the failed-downdate flags
            st["logalpha"], st["eta"] = self.ram_adapt()
            st["failed"] = self.ram_failed()[2]
        return st

    def state_step(self):
        step = C.c_int64()
        self.eng._check(self._r("get_state")(self.h, None, None, None, None, None, None, C.byref(step)))
        return step.value

    def set_params(self, x):
        x = _as_f64(x)
        assert x.shape == (self.dim, self.n)
        self.eng._check(self._r("set_params")(self.h, x.ctypes.data_as(_dp)))

    def set_state(self, state):
        """resume from a dict returned by `state()` (AbstractMCMC's `initial_state`); missing / None entries keep
        the run's current values"""
        d, n = self.dim, self.n
        keep = []
        def arr(key, shape, dtype, ptr):
            a = state.get(key)
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dtype)
            if a.shape != shape:
                raise AMHArgumentError(AMH_ERR_INVALID, f"state[{key!r}] must have shape {shape}, got {a.shape}")
            keep.append(a)
            return a.ctypes.data_as(ptr)
        step = state.get("step")
        if self.sampler.kind == SAMPLER_RAM and any(state.get(k) is not None for k in ("logalpha", "eta", "failed")):
            self.eng._check(self._r("set_ram_adapt")(
                self.h, arr("logalpha", (n,), np.float64, _dp), arr("eta", (n,), np.float64, _dp),
                arr("failed", (n,), np.uint8, _u8p)))
        self.eng._check(self._r("set_state")(
            self.h, arr("x", (d, n), np.float64, _dp), arr("lp", (n,), np.float64, _dp),
            arr("grad", (d, n), np.float64, _dp), arr("S", (d * (d + 1) // 2, n), np.float64, _dp),
            arr("accepted", (n,), np.uint8, _u8p), arr("naccept", (n,), np.int64, _i64p),
            C.c_int64(-1 if step is None else int(step))))

    def ram_adapt(self):
        """(log-alpha, eta) of the last step of every chain -- RobustAdaptiveMetropolisState fields (RAM :107-110)"""
        la = np.empty(self.n); eta = np.empty(self.n)
        self.eng._check(self._r("get_ram_adapt")(self.h, la.ctypes.data_as(_dp), eta.ctypes.data_as(_dp)))
        return la, eta

    def ram_failed(self):
        """(count, first global chain or -1, flags[n]) of chains whose rank-1 downdate failed (amh_run_ram_failed)"""
        nf, first = C.c_int64(), C.c_int64()
        flags = np.empty(self.n, dtype=np.uint8)
        self.eng._check(self._r("ram_failed")(self.h, C.byref(nf), C.byref(first), flags.ctypes.data_as(_u8p)))
        return int(nf.value), int(first.value), flags

    def contract(self):
        """the numerical-contract version this run was created under"""
        return int(self._r("contract")(self.h))

    def launch_count(self):
        return int(self._r("launch_count")(self.h))

    def kernel_time_ms(self, reset=False):
        ms = C.c_double(); nl = C.c_int64()
        self.eng._check(self._r("kernel_time_ms")(self.h, int(reset), C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def close(self):
        if self.h:
            self._r("destroy")(self.h)
            self.h = None


class _JobPart:
    """what Job.target*/sampler return: the job owns the object, closing is a no-op"""
    def __init__(self, kind, dim, n_walkers=0):
        self.kind, self.dim, self.n_walkers = kind, dim, n_walkers
    def close(self):
        pass


class Job:
    """amh_job: `ngpus` devices of one box in ONE process (include/amh.h "multi-GPU job").  Same surface as Engine for the
    host mirrors -- target / target_source / target_of / sampler / run -- but every array is job-wide and the library
    shards the chains, broadcasts the target and writes each device's column block of the output in place."""

    def __init__(self, eng: "Engine", ngpus: int, devices=None):
        self.eng, self.ngpus = eng, int(ngpus)
        devs = None
        if devices is not None:
            if len(devices) != ngpus:
                raise AMHArgumentError(AMH_ERR_INVALID, "need one device index per GPU of the job")
            devs = (C.c_int32 * ngpus)(*[int(v) for v in devices])
        h = C.c_void_p()
        eng._check(eng._f("job_create")(C.c_int32(ngpus), devs, C.byref(h)))
        self.h = h
        self._target = self._sampler = None
        self._keep = None

    # Engine look-alikes used by samplers.lower / sampling
    _check = property(lambda self: self.eng._check)
    _f = property(lambda self: self.eng._f)
    pinned_empty = property(lambda self: self.eng.pinned_empty)

    def target(self, kind, dim, blob):
        blob = _as_f64(blob).ravel()
        self.eng._check(self.eng._f("job_target_create")(self.h, kind, dim, blob.ctypes.data_as(_dp), blob.size))
        self._target = _JobPart(kind, dim)
        return self._target

    def target_source(self, dim, source, data=None, has_gradient=False):
        data = np.zeros(0) if data is None else _as_f64(data).ravel()
        self.eng._check(self.eng._f("job_target_create_source")(self.h, dim, source.encode(), int(bool(has_gradient)),
                                                                data.ctypes.data_as(_dp) if data.size else None, data.size))
        self._target = _JobPart(TARGET_USER, dim)
        return self._target

    def target_of(self, t):
        if getattr(t, "kind", None) == TARGET_USER:
            return self.target_source(t.dim, t.source, t.data, t.has_gradient)
        return self.target(t.kind, t.dim, t.blob())

    def sampler(self, **kw):
        d, keep = sampler_desc(**kw)
        self.eng._check(self.eng._f("job_sampler_create")(self.h, C.byref(d)))
        self._sampler = _JobPart(d.kind, d.dim, int(d.n_walkers))
        return self._sampler

    def broadcast_info(self):
        """(mode, ms of the last target broadcast, ms of the one-time communicator set-up)"""
        return ((self.eng._f("job_broadcast_mode")(self.h) or b"").decode(), float(self.eng._f("job_broadcast_ms")(self.h)),
                float(self.eng._f("job_comm_init_ms")(self.h)))

    def run(self, target, sampler, nchains, seeds, init=None, chain_offset=0) -> "JobRun":
        """seeds / init are JOB-WIDE: one seed per chain (per ensemble for Ensemble), init (dim, nchains)"""
        if chain_offset:
            raise AMHArgumentError(AMH_ERR_INVALID, "a job holds all chains: chain_offset must be 0")
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64).ravel()
        init_p, init_ld = None, 0
        if init is not None:
            init = np.asarray(init)
            if init.shape != (target.dim, nchains):
                raise AMHArgumentError(AMH_ERR_INVALID, f"init must have shape (dim, nchains) = {(target.dim, nchains)}, got {init.shape}")
            ok = (init.dtype == np.float64 and init.strides[1] == 8 and init.strides[0] % 8 == 0 and init.strides[0] >= 8 * nchains)
            if not ok:
                init = _as_f64(init)
            init_ld = init.strides[0] // 8
            init_p = C.cast(init.ctypes.data, _dp)
        self.eng._check(self.eng._f("job_run_create")(self.h, nchains, seeds.ctypes.data_as(_u64p), init_p, init_ld))
        return JobRun(self, target, sampler, nchains)

    def shards(self):
        """[(lo, hi, device)] of the current run"""
        out = []
        for k in range(self.ngpus):
            lo, hi, dev = C.c_int64(), C.c_int64(), C.c_int32()
            self.eng._check(self.eng._f("job_run_shard")(self.h, k, C.byref(lo), C.byref(hi), C.byref(dev)))
            out.append((lo.value, hi.value, dev.value))
        return out

    def close(self):
        if self.h:
            self.eng._f("job_destroy")(self.h)
            self.h = None


class JobRun(Run):
    """the run of a Job: the Run surface over amh_job_run_*, all arrays job-wide"""
    _p = "job_run_"

    def __init__(self, job, target, sampler, n):
        super().__init__(job.eng, job.h, target, sampler, n)
        self.job = job

    def _sample_call(self, N, discard_initial, thinning, num_warmup, out_p, out_ld, acc_p, acc_ld, summ_p):
        if out_ld != self.n or acc_ld != self.n:
            raise AMHArgumentError(AMH_ERR_INVALID, "a job writes the whole C-contiguous (N, dim+1, nchains) array")
        self.eng._check(self._r("sample")(self.h, N, discard_initial, thinning, num_warmup, out_p, acc_p, summ_p))

    def close(self):
        if self.h:
            self._r("destroy")(self.h)      # releases the run, the job handle stays
            self.h = None
