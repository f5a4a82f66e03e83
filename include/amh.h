/* amh.h -- C ABI of libamh_b200.so, the B200-native many-chain
 * Metropolis-Hastings engine.  This is the drop-in boundary for the multi-chain
 * `sample(model, sampler, parallel, N, nchains; kw...)` path of AdvancedMH.jl
 * (SURVEY.md 8b).  A Julia shim `ccall`s exactly these entry points
 * (INTEGRATION.md shows the binding); tests drive them through Python ctypes.
 *
 * Conventions
 *   - every function returns an int32 status (AMH_OK == 0); on failure the
 *     text is available from amh_last_error() (thread-local, library owned).
 *     AMH_ERR_INVALID maps to Julia's ArgumentError / ErrorException.
 *   - no C++ exception and no callback crosses the boundary.
 *   - handles are opaque and owned by the library; caller buffers are plain
 *     pointers + sizes and are only read/written during the call.
 *   - one process (or one handle set) per GPU; a handle must not be used from
 *     two host threads at once.  Multi-GPU = one amh_ctx per device, chains
 *     sharded by (chain_offset, nchains_local); no per-step collective.
 *   - there is NO CPU fallback: without a CUDA device amh_ctx_create fails.
 *
 * The CPU oracle (oracle/amh_oracle.cpp) exports the same functions with the
 * prefix amho_ so that one harness drives both.
 */
#ifndef AMH_H
#define AMH_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMH_VERSION_MAJOR 0
#define AMH_VERSION_MINOR 1

/* ---- status codes ---- */
#define AMH_OK               0
#define AMH_ERR_INVALID      1   /* bad argument (Julia: ArgumentError) */
#define AMH_ERR_CUDA         2   /* CUDA runtime failure, or no device */
#define AMH_ERR_UNSUPPORTED  3   /* representable in the reference, not on the device */
#define AMH_ERR_STATE        4   /* call out of order (e.g. MALA without initial params) */

/* ---- device target catalogue ------------------------
 * Replaces the opaque closure of DensityModel(f) (src/AdvancedMH.jl:52-54) and
 * LogDensityProblems.logdensity[_and_gradient] (src/AdvancedMH.jl:74-77,
 * MALA.jl:100-105) by a fixed catalogue.  `blob` layouts (all float64):        */
#define AMH_TARGET_IID_NORMAL  1 /* theta=(mu,sigma); blob=[y_1..y_n]; dim==2.
                                    lp = sigma>=0 ? sum_i logpdf(Normal(mu,sigma),y_i) : -Inf
                                    (README.md:26-31, test/runtests.jl:23-31)            */
#define AMH_TARGET_MVNORMAL    2 /* blob=[c0, mu[d], U[d(d+1)/2]] ; U lower-triangular
                                    packed by rows, U'U = inv(Sigma);
                                    lp = c0 - 0.5*|U(x-mu)|^2
                                    (test/RobustAdaptiveMetropolis.jl:1-9 `Gaussian`)     */
#define AMH_TARGET_ROSENBROCK  3 /* blob=[a, b, s]; lp = -sum_{i<d-1}[b(x_{i+1}-x_i^2)^2+(a-x_i)^2]/s */
#define AMH_TARGET_LOGISTIC    4 /* blob=[tau, X[n*d] row-major, y[n]];
                                    lp = sum_i[y_i eta_i - log1pexp(eta_i)] - |beta|^2/(2 tau^2);
                                    the row sum runs as 8 interleaved partial sums + a fixed tree (contract)  */
#define AMH_TARGET_GAUSS_PREC  5 /* blob=A[d*d] row-major symmetric; lp = -x'Ax/2, grad = -Ax
                                    (test/runtests.jl:335-347 `TheNormalLogDensity`)      */
#define AMH_TARGET_NIG_TOY     6 /* theta=(s,m); blob=[alpha, beta, y_1..y_n]; dim==2;
                                    s>0 ? logpdf(InverseGamma(alpha,beta),s)+logpdf(Normal(0,sqrt s),m)
                                          + sum_i logpdf(Normal(m,sqrt s),y_i) : -Inf
                                    (test/emcee.jl:5-15)                                   */
#define AMH_TARGET_NIG_TOY_LOG 7 /* same in (log s, m) with the Jacobian term (test/emcee.jl:46-56) */
#define AMH_TARGET_USER      100 /* log-density given as source text: amh_target_create_source below        */

/* ---- sampler kinds: the reference constructors they stand for ---- */
#define AMH_SAMPLER_STATIC   1  /* MetropolisHastings(StaticProposal(..)), StaticMH   mh-core.jl:44-49, proposal.jl:70-85 */
#define AMH_SAMPLER_RW       2  /* MetropolisHastings(RandomWalkProposal(..)), RWMH   mh-core.jl:50-51, proposal.jl:41-64 */
#define AMH_SAMPLER_STRETCH  3  /* Ensemble(n_walkers, StretchProposal(p, a))         emcee.jl:1-102                      */
#define AMH_SAMPLER_MALA     4  /* MALA(g -> MvNormal(drift*g, sigma2*I))             MALA.jl:1-93                        */
#define AMH_SAMPLER_RAM      5  /* RobustAdaptiveMetropolis(alpha, gamma, S, lo, hi)  RobustAdaptiveMetropolis.jl:75-278  */
#define AMH_SAMPLER_MIXED    6  /* MetropolisHastings([StaticProposal(d1), RandomWalkProposal(d2), ...]): an ARRAY OF
                                   PROPOSALS, one univariate law per coordinate, each static or random-walk and with its
                                   own `issymmetric` (proposal.jl:132-150, 236-240; README.md:125-133)                  */

/* proposal covariance representation (Distributions' ScalMat / PDiagMat / PDMat) */
#define AMH_COV_SCALAR 1   /* scale[0]   = sigma                                      */
#define AMH_COV_DIAG   2   /* scale[dim] = sigma_i  (also: array of Normal(mu_i,sigma_i), proposal.jl:26-35) */
#define AMH_COV_FULL   3   /* scale[dim(dim+1)/2] = lower Cholesky factor L, packed by rows */
#define AMH_COV_COMPONENTS 4 /* components[dim]: an array of univariate distributions of any catalogue family
                                (proposal.jl:26-35), e.g. StaticProposal([Normal(0,1), InverseGamma(2,3)]) README.md:106;
                                mean / scale are unused */

/* one coordinate of an AMH_COV_COMPONENTS proposal; families and parametrisation: AMH_FAM_* in amh_contract.h */
typedef struct amh_component {
    int32_t family;          /* AMH_FAM_NORMAL, _INVGAMMA, _GAMMA, _UNIFORM, _EXPONENTIAL, _LOGNORMAL               */
    int32_t rw;              /* AMH_SAMPLER_MIXED only: 1 = RandomWalkProposal, 0 = StaticProposal                   */
    int32_t symmetric;       /* AMH_SAMPLER_MIXED only: the component's `issymmetric` (Hastings term literal 0)      */
    int32_t reserved;
    double  p0, p1;          /* distribution parameters                                                              */
    double  logc;            /* normalising constant of the log-density, computed by the caller                      */
} amh_component;

typedef struct amh_sampler_desc {
    int32_t kind;            /* AMH_SAMPLER_*                                                   */
    int32_t dim;
    int32_t symmetric;       /* the `issymmetric` type parameter (proposal.jl:1-21): 1 => the
                                Hastings term is the literal 0 (proposal.jl:195-196)            */
    int32_t cov_kind;        /* AMH_COV_*  (STATIC / RW; also the initial-draw law of STRETCH)  */
    const double* mean;      /* proposal mean[dim], NULL = zeros                                */
    const double* scale;     /* see AMH_COV_*                                                   */
    double  stretch_a;       /* StretchProposal.stretch_length, default 2.0 (emcee.jl:68)       */
    int64_t n_walkers;       /* Ensemble.n_walkers (emcee.jl:2)                                 */
    double  mala_sigma2;     /* MALA proposal g -> MvNormal(mala_drift*g, mala_sigma2*I)        */
    double  mala_drift;      /*   canonical: mala_drift = sigma2/2 (README.md:180)              */
    double  ram_alpha;       /* RAM.alpha  = 0.234 (RobustAdaptiveMetropolis.jl:78)             */
    double  ram_gamma;       /* RAM.gamma  = 0.6   (:80)                                        */
    double  ram_eig_lo;      /* RAM.eigenvalue_lower_bound = 0   (:84)                          */
    double  ram_eig_hi;      /* RAM.eigenvalue_upper_bound = Inf (:86)                          */
    const double* ram_S0;    /* RAM.S: dense dim x dim row-major (lower triangle used), NULL = I (:82,198-207) */
    const amh_component* components;  /* [dim], cov_kind == AMH_COV_COMPONENTS or kind == AMH_SAMPLER_MIXED, else NULL */
    int32_t contract;        /* version of the numerical contract the run's step noise follows (include/amh_contract.h):
                                0 = the library default (AMH_CONTRACT_VERSION, or the environment variable AMH_CONTRACT),
                                1 = v1, 2 = v2.  The oracle takes the same field, so parity is per version.           */
    int32_t precision;       /* AMH_PRECISION_FP64 (0, default): fp64 throughout, bit-exact against the oracle.
                                AMH_PRECISION_BF16X2 (1): OPT-IN tensor-core path with a stated tolerance -- MALA on the
                                logistic target with dim = 128 runs its two design-matrix contractions as split-bf16
                                tcgen05 GEMMs with fp32 accumulation (csrc/amh_launch_mala_tensor.cu); every other
                                sampler / target combination rejects it with AMH_ERR_UNSUPPORTED                      */
} amh_sampler_desc;
#define AMH_PRECISION_FP64   0
#define AMH_PRECISION_BF16X2 1

/* pooled and per-chain summaries accumulated on the device over SAVED samples */
typedef struct amh_summary {
    int64_t n_saved;         /* samples accumulated per chain                                   */
    int64_t n_steps;         /* stateful steps taken per chain                                  */
    double  accept_rate;     /* accepted moves / steps, pooled over local chains                */
    double* mean;            /* [dim]   pooled posterior mean  (caller buffer, may be NULL)     */
    double* var;             /* [dim]   pooled posterior variance (population, /n)              */
    double* chain_mean;      /* [dim][nchains_local] per-chain means (may be NULL)              */
} amh_summary;

typedef struct amh_ctx     amh_ctx;
typedef struct amh_target  amh_target;
typedef struct amh_sampler amh_sampler;
typedef struct amh_run     amh_run;
typedef struct amh_job     amh_job;

#ifndef AMH_RTC   /* (the NVRTC translation unit of user-supplied targets needs the constants and structs only) */
/* handshake, diagnostics */
int32_t     amh_version(int32_t* major, int32_t* minor);
const char* amh_last_error(void);
int32_t     amh_contract_version(void);

/* one context per GPU: device selection + the stream all work is ordered on */
int32_t amh_ctx_create(int32_t device, amh_ctx** out);
int32_t amh_ctx_destroy(amh_ctx* ctx);
int32_t amh_ctx_sync(amh_ctx* ctx);

/* model: stands for DensityModel / LogDensityModel of a catalogue target */
int32_t amh_target_create(amh_ctx* ctx, int32_t kind, int32_t dim,
                          const double* blob, int64_t nblob, amh_target** out);
int32_t amh_target_destroy(amh_target* t);

/* model given as SOURCE TEXT: the route from the catalogue to DensityModel(f)
 * (src/AdvancedMH.jl:52-54: an arbitrary log-density closure) and to LogDensityProblems objects with
 * `logdensity` / `logdensity_and_gradient` (src/AdvancedMH.jl:76, MALA.jl:100-105).  A Julia closure cannot cross a
 * C ABI into a kernel, so the caller states the function in the C++ subset described in include/amh_user_target.h:
 *
 *   AMH_TARGET double amh_user_logdensity(const double* x, int dim, const double* data, long long ndata);
 *   AMH_TARGET void   amh_user_logdensity_and_gradient(const double* x, int dim, const double* data, long long ndata,
 *                                                      double* lp, double* grad);      // iff has_gradient != 0
 *
 * The library compiles it with NVRTC for sm_100a (-fmad=false, as the numerical contract requires) together with
 * its own generic kernels; every sampler then runs on it: StaticMH / RWMH incl. arrays of proposals, Ensemble, RAM,
 * and MALA when has_gradient != 0 (otherwise MALA fails like the reference: "The gradient of the log density
 * function is not defined", MALA.jl:44-50).  `data` ([ndata] float64, may be NULL) is copied to the device and handed
 * to every call.  dim <= 128.  A compile error returns AMH_ERR_INVALID with the compiler log in amh_last_error(). */
int32_t amh_target_create_source(amh_ctx* ctx, int32_t dim, const char* source, int32_t has_gradient,
                                 const double* data, int64_t ndata, amh_target** out);

/* sampler: POD image of the unchanged AdvancedMH constructor */
int32_t amh_sampler_create(amh_ctx* ctx, const amh_sampler_desc* desc, amh_sampler** out);
int32_t amh_sampler_destroy(amh_sampler* s);

/* run = the state of `nchains_local` chains [chain_offset, chain_offset+nchains_local)
 * of a job with global seeds.  Performs the FIRST step of AbstractMCMC
 * (mh-core.jl:76-86, emcee.jl:29-34, MALA.jl:37-40, RAM :175-214):
 *   init == NULL : draw from the proposal (STATIC/RW/STRETCH) or randn (RAM);
 *                  MALA -> AMH_ERR_STATE "please specify initial parameters" (MALA.jl:37)
 *   init != NULL : [dim][nchains_local] float64, chains fastest; row i starts at init + i*init_ld
 *                  (init_ld = 0 means dense, init_ld = nchains_total lets a rank pass its column block
 *                  of the global initial_params matrix without repacking).
 * seeds: one uint64 per local chain (MH/MALA/RAM) or per local ENSEMBLE (STRETCH,
 * nchains_local = n_ensembles * n_walkers), i.e. rand(rng, UInt, nchains). */
int32_t amh_run_create(amh_ctx* ctx, amh_target* target, amh_sampler* sampler,
                       int64_t nchains_local, int64_t chain_offset,
                       const uint64_t* seeds, const double* init, int64_t init_ld, amh_run** out);
int32_t amh_run_destroy(amh_run* run);

/* `nsteps` stateful steps for every chain, asynchronously on the ctx stream
 * (AbstractMCMC.step / step_warmup: mh-core.jl:92-117, emcee.jl:14-24,
 * MALA.jl:54-93, RAM :216-278).  warmup != 0 selects step_warmup (RAM adapts).
 * steps_per_launch <= 0 lets the library choose how many steps one kernel
 * launch fuses; 1 gives exactly one kernel per MCMC step. */
int32_t amh_run_steps(amh_run* run, int64_t nsteps, int32_t warmup, int32_t steps_per_launch);
int32_t amh_run_sync(amh_run* run);

/* The AbstractMCMC.mcmcsample schedule:
 *   discard_initial steps, save, then (N-1) x { thinning steps, save };
 *   stateful step s (1-based) is a warm-up step iff s <= num_warmup.
 * out (may be NULL): [N][dim+1][nchains_local] float64 = params..., lp
 *   (ext/AdvancedMHMCMCChainsExt.jl:24-38,93-106 layout, chains fastest);
 * accepted_out (may be NULL): [N][nchains_local] uint8 Transition.accepted;
 * summary (may be NULL): filled with moments over the N saved samples. */
int32_t amh_run_sample(amh_run* run, int64_t N, int64_t discard_initial, int64_t thinning,
                       int64_t num_warmup, double* out, uint8_t* accepted_out,
                       amh_summary* summary);

/* the same with caller buffers that are COLUMN BLOCKS of larger arrays: row stride out_ld >= nchains_local doubles
 * ([N][dim+1][out_ld]) and acc_ld >= nchains_local bytes.  Lets several runs -- shards of one job on several streams of
 * a GPU, or on several GPUs of a process -- fill one [N][dim+1][nchains] array in place. */
int32_t amh_run_sample_ld(amh_run* run, int64_t N, int64_t discard_initial, int64_t thinning,
                          int64_t num_warmup, double* out, int64_t out_ld, uint8_t* accepted_out, int64_t acc_ld,
                          amh_summary* summary);

/* state introspection / resume (getparams/setparams!!: src/AdvancedMH.jl:146-157,
 * MALA.jl:23-35, RAM :116-121; StatesExtractor: test/RobustAdaptiveMetropolis.jl:11-28).
 * Any pointer may be NULL.  x:[dim][n] lp:[n] grad:[dim][n] (MALA)
 * S:[dim(dim+1)/2][n] packed rows (RAM) accepted:[n] naccept:[n] */
int32_t amh_run_get_state(amh_run* run, double* x, double* lp, double* grad, double* S,
                          uint8_t* accepted, int64_t* naccept, int64_t* step_counter);
/* the same into COLUMN BLOCKS of job-wide arrays: every 2-d array has row stride `ld` >= nchains_local (x, grad:
 * [dim][ld]; S: [dim(dim+1)/2][ld]); the 1-d arrays are written at the pointer given */
int32_t amh_run_get_state_ld(amh_run* run, int64_t ld, double* x, double* lp, double* grad, double* S,
                             uint8_t* accepted, int64_t* naccept, int64_t* step_counter);
/* setparams!!: replaces x and recomputes lp (and the gradient for MALA) on the device */
int32_t amh_run_set_params(amh_run* run, const double* x);
/* resume: installs a state read earlier with amh_run_get_state, so that a new run continues an old one bit for bit
 * (AbstractMCMC's `initial_state=` keyword; the state structs are Transition src/AdvancedMH.jl:61-65,
 * GradientTransition MALA.jl:14-19, RobustAdaptiveMetropolisState RAM :99-114 -- RAM's `iteration` is
 * step_counter + 1 and its eta / log-alpha fields are outputs only).  Same layouts as amh_run_get_state; a NULL
 * pointer keeps the run's current value, step_counter < 0 keeps the counter.  Nothing is recomputed except static
 * MH's cached proposal log-density of the state (a pure function of x). */
int32_t amh_run_set_state(amh_run* run, const double* x, const double* lp, const double* grad, const double* S,
                          const uint8_t* accepted, const int64_t* naccept, int64_t step_counter);
int32_t amh_run_set_state_ld(amh_run* run, int64_t ld, const double* x, const double* lp, const double* grad, const double* S,
                             const uint8_t* accepted, const int64_t* naccept, int64_t step_counter);
/* RAM: the state fields that only report the last step -- log acceptance ratio `log-alpha` and adaptation step size
 * `eta` (RAM :107-110), [nchains_local] each, either may be NULL (StatesExtractor, test/RobustAdaptiveMetropolis.jl:11-28) */
int32_t amh_run_get_ram_adapt(amh_run* run, double* logalpha, double* eta);
/* resume of a RAM run: installs the report-only fields read with amh_run_get_ram_adapt / amh_run_ram_failed, so that a
 * resumed run reports the same `log-alpha` / `eta` as the uninterrupted one (RobustAdaptiveMetropolisState RAM :99-114).
 * [nchains_local] each, any may be NULL. */
int32_t amh_run_set_ram_adapt(amh_run* run, const double* logalpha, const double* eta, const uint8_t* failed);
/* RAM: chains whose rank-1 downdate left the positive-definite cone.  In the reference `lowrankdowndate` throws
 * `PosDefException` (RAM :170, SURVEY.md A.4: (v_i / A_ii)^2 > 1) and the whole `sample` call aborts; a lock-step
 * kernel cannot throw, so the chain keeps its last good factor, stops nothing else, and raises a sticky per-chain flag.
 * nfailed: number of flagged local chains; first_chain: GLOBAL index of the first one or -1; failed: [nchains_local]
 * flags (any may be NULL).  The host layer turns nfailed > 0 into the reference's exception after the run. */
int32_t amh_run_ram_failed(amh_run* run, int64_t* nfailed, int64_t* first_chain, uint8_t* failed);

/* page-locked host memory for initial_params / sample buffers: host<->device copies of pinned buffers run at
 * full PCIe/C2C speed and asynchronously (Julia: unsafe_wrap(Array, Ptr{Float64}(p), dims)) */
int32_t amh_host_alloc(size_t bytes, void** out);
int32_t amh_host_free(void* p);

int32_t amh_run_contract(amh_run* run);      /* the contract version this run was created under (1 or 2) */
int32_t amh_run_dim(amh_run* run);
int64_t amh_run_nchains(amh_run* run);
/* number of kernel launches issued by this run so far (bench.py's gpu_launches) */
int64_t amh_run_launch_count(amh_run* run);
/* device time of the stepping kernels since the last reset, measured with CUDA
 * events on the ctx stream (milliseconds); reset != 0 zeroes the accumulator.
 * The first call switches event recording on for this run (off by default: two
 * event records per amh_run_steps call are not free at one launch per step). */
int32_t amh_run_kernel_time_ms(amh_run* run, int32_t reset, double* ms, int64_t* launches);

/* ---- multi-GPU job: N devices of one box behind ONE handle, in ONE process ---------------------------------------
 * The multi-chain call of the reference is `sample(model, sampler, MCMCThreads() | MCMCDistributed(), N, nchains)`
 * (src/AdvancedMH.jl:30 re-export; README.md:135-148; test/runtests.jl:96-110): one keyword turns on chain-level data
 * parallelism.  Here that keyword is `MCMCB200(ngpus = k)` and the job handle is what it lowers to: the caller passes
 * JOB-WIDE arrays (seeds[nchains], init[dim][nchains], out[N][dim+1][nchains], state arrays) and the library shards
 * the chains in contiguous blocks over the devices (an ensemble stays on one device), drives every device from its own
 * host thread, broadcasts the target's fixed data once (ncclBroadcast over NVLink when libnccl can be loaded, else peer
 * copies; no per-step collective), lets every device write its column block of `out` directly, and pools the
 * summaries.  Results are bit-identical to a one-device run of the same seeds (global chain identity).
 * A job owns at most one target, one sampler and one run at a time; creating a new one releases the old. */
int32_t amh_job_create(int32_t ngpus, const int32_t* devices /* [ngpus] or NULL = 0..ngpus-1 */, amh_job** out);
int32_t amh_job_destroy(amh_job* job);
int32_t amh_job_ngpus(amh_job* job);
int32_t amh_job_target_create(amh_job* job, int32_t kind, int32_t dim, const double* blob, int64_t nblob);
int32_t amh_job_target_create_source(amh_job* job, int32_t dim, const char* source, int32_t has_gradient,
                                     const double* data, int64_t ndata);
/* how the last target reached the devices ("nccl", "peer", "h2d", "source") and how long it took (host clock, ms) */
const char* amh_job_broadcast_mode(amh_job* job);
double      amh_job_broadcast_ms(amh_job* job);
double      amh_job_comm_init_ms(amh_job* job);     /* one-time cost of ncclCommInitAll at amh_job_create (0 without NCCL) */
int32_t amh_job_sampler_create(amh_job* job, const amh_sampler_desc* desc);
/* nchains counts walkers for an Ensemble; seeds: one per chain (per ensemble for STRETCH); init: [dim][init_ld] or NULL */
int32_t amh_job_run_create(amh_job* job, int64_t nchains, const uint64_t* seeds, const double* init, int64_t init_ld);
int32_t amh_job_run_destroy(amh_job* job);
int32_t amh_job_run_steps(amh_job* job, int64_t nsteps, int32_t warmup, int32_t steps_per_launch);
int32_t amh_job_run_sync(amh_job* job);
/* out: [N][dim+1][nchains], accepted_out: [N][nchains], summary pooled over all devices (any may be NULL) */
int32_t amh_job_run_sample(amh_job* job, int64_t N, int64_t discard_initial, int64_t thinning, int64_t num_warmup,
                           double* out, uint8_t* accepted_out, amh_summary* summary);
int32_t amh_job_run_get_state(amh_job* job, double* x, double* lp, double* grad, double* S, uint8_t* accepted,
                              int64_t* naccept, int64_t* step_counter);
int32_t amh_job_run_set_state(amh_job* job, const double* x, const double* lp, const double* grad, const double* S,
                              const uint8_t* accepted, const int64_t* naccept, int64_t step_counter);
int32_t amh_job_run_get_ram_adapt(amh_job* job, double* logalpha, double* eta);
int32_t amh_job_run_set_ram_adapt(amh_job* job, const double* logalpha, const double* eta, const uint8_t* failed);
int32_t amh_job_run_ram_failed(amh_job* job, int64_t* nfailed, int64_t* first_chain, uint8_t* failed);
/* which chains device number k of the job holds: [lo, hi) and the CUDA device index */
int32_t amh_job_run_shard(amh_job* job, int32_t k, int64_t* lo, int64_t* hi, int32_t* device);
int32_t amh_job_run_contract(amh_job* job);
int64_t amh_job_run_launch_count(amh_job* job);
/* device time of the stepping kernels: max over the devices (they run concurrently), launches summed */
int32_t amh_job_run_kernel_time_ms(amh_job* job, int32_t reset, double* ms, int64_t* launches);
#endif /* AMH_RTC */

#ifdef __cplusplus
}
#endif
#endif /* AMH_H */
