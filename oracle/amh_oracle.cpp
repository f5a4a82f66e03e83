/* amh_oracle.cpp -- CPU ORACLE for the many-chain MH hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is a scalar, one-chain-at-
 * a-time restatement of the reference algorithms, used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * to check (and time beside) the CUDA engine.  Nothing under
 * advancedmh.jl_b200/ links, imports or calls it.
 *
 * What it restates (all paths relative to /root/reference):
 *   src/mh-core.jl:65-117        transition / first step / MH step
 *   src/proposal.jl:24-35,41-85,190-196   rand, logpdf, propose, q, logratio
 *   src/emcee.jl:6-8,29-58,70-102         ensemble init, sequential sweep, stretch move
 *   src/MALA.jl:37-40,54-93               gradient transition and MALA step
 *   src/RobustAdaptiveMetropolis.jl:123-173,175-278   RAM inner step, adaptation, steps
 *   src/AdvancedMH.jl:61-77               Transition / logdensity caching
 * plus the AbstractMCMC.mcmcsample schedule (not vendored) and
 * the LinearAlgebra rank-1 Cholesky up/down-date.
 *
 * Randomness and transcendental functions follow include/amh_contract.h (the
 * reference's Xoshiro/ziggurat stream cannot be consumed in lock-step and is
 * not stable across Julia versions).  PARITY PINNING: Julia is not installed
 * here or on the GPU box and the reference's tests hold no bit-level golden
 * vectors, so the oracle is pinned (tests/test_oracle_kat.py) against the
 * reference's own statistical known answers,
 * against Random123's published Philox4x32-10 vectors and against mpmath for
 * the contract math; bit-level agreement with a Julia run is UNPINNED.
 *
 * Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off, never -ffast-math).
 */
#include <dlfcn.h>
#include <fcntl.h>
#include <spawn.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <thread>
#include <algorithm>
#include <limits>

#include "../include/amh_contract.h"
#include "../include/amh.h"

extern char** environ;

namespace {

thread_local std::string g_err;
int g_threads = 0;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

inline int64_t tri(int64_t i, int64_t j) { return i * (i + 1) / 2 + j; }

/* ------------------------------------------------------------------ target */
struct Target {
    int kind = 0, dim = 0;
    std::vector<double> blob;
    int64_t ndata = 0;       /* IID_NORMAL / NIG / LOGISTIC rows */
    double inv2tau2 = 0, invtau2 = 0;
    /* AMH_TARGET_USER: the user's source text compiled by g++ with the contract flags (amho_target_create_source) */
    typedef double (*user_lp_fn)(const double*, int, const double*, long long);
    typedef void (*user_lpg_fn)(const double*, int, const double*, long long, double*, double*);
    void* user_lib = nullptr;
    user_lp_fn user_lp = nullptr;
    user_lpg_fn user_lpg = nullptr;
    std::string user_so;
    ~Target() {
        if (user_lib) dlclose(user_lib);
        if (!user_so.empty()) {
            std::remove(user_so.c_str());
            rmdir(user_so.substr(0, user_so.rfind('/')).c_str());     /* its private mkdtemp directory */
        }
    }

    /* Normal(mu, sigma) log-density of y: Distributions' normlogpdf
     * -(z^2 + log 2pi)/2 - log sigma  (SURVEY.md A.2). */
    static double normlogpdf(double mu, double sigma, double lsigma, double y) {
        const double z = (y - mu) / sigma;
        const double t = z * z + AMH_LOG_2PI;
        return -0.5 * t - lsigma;
    }

    double logp(const double* x) const {
        const int d = dim;
        switch (kind) {
        case AMH_TARGET_USER: return user_lp(x, d, blob.data(), (long long)ndata);
        case AMH_TARGET_IID_NORMAL: {
            /* density(theta) = insupport(theta) ? sum(logpdf.(Normal(theta1,theta2), data)) : -Inf
             * test/runtests.jl:26-28, README.md:29-31 */
            const double mu = x[0], sigma = x[1];
            if (!(sigma >= 0.0)) return -INFINITY;
            const double ls = amh::log_(sigma);
            double acc = 0.0;
            for (int64_t i = 0; i < ndata; ++i) acc = acc + normlogpdf(mu, sigma, ls, blob[i]);
            return acc;
        }
        case AMH_TARGET_MVNORMAL: {
            const double c0 = blob[0];
            const double* mu = &blob[1];
            const double* U = &blob[1 + d];
            double q = 0.0;
            for (int i = 0; i < d; ++i) {
                double w = U[tri(i, 0)] * (x[0] - mu[0]);
                for (int j = 1; j <= i; ++j) w = fma(U[tri(i, j)], x[j] - mu[j], w);
                q = (i == 0) ? w * w : fma(w, w, q);
            }
            return fma(-0.5, q, c0);
        }
        case AMH_TARGET_ROSENBROCK: {
            const double a = blob[0], b = blob[1], s = blob[2];
            double acc = 0.0;
            for (int i = 0; i + 1 < d; ++i) {
                const double t1 = fma(-x[i], x[i], x[i + 1]);
                const double t2 = a - x[i];
                acc = acc + fma(b * t1, t1, t2 * t2);
            }
            return -(acc / s);
        }
        case AMH_TARGET_GAUSS_PREC: {
            /* -x' * A * x / 2  (test/runtests.jl:341) */
            const double* A = blob.data();
            double q = 0.0;
            for (int j = 0; j < d; ++j) {
                double t = x[0] * A[j];
                for (int i = 1; i < d; ++i) t = fma(x[i], A[(int64_t)i * d + j], t);
                q = (j == 0) ? t * x[0] : fma(t, x[j], q);
            }
            return -0.5 * q;
        }
        case AMH_TARGET_NIG_TOY:
        case AMH_TARGET_NIG_TOY_LOG: {
            /* test/emcee.jl:5-15 (untransformed) and :46-56 (log s, + Jacobian) */
            const double alpha = blob[0], beta = blob[1], cig = blob[2];
            const double* y = &blob[3];
            double s, logs = 0.0;
            if (kind == AMH_TARGET_NIG_TOY) {
                s = x[0];
                if (!(s > 0.0)) return -INFINITY;
            } else {
                logs = x[0];
                s = amh::exp_(logs);
            }
            const double m = x[1];
            const double ls = (kind == AMH_TARGET_NIG_TOY) ? amh::log_(s) : amh::log_(s);
            const double sq = sqrt(s);
            const double lsq = amh::log_(sq);
            double acc = (cig - (alpha + 1.0) * ls) - beta / s;
            acc = acc + normlogpdf(0.0, sq, lsq, m);
            for (int64_t i = 0; i < ndata; ++i) acc = acc + normlogpdf(m, sq, lsq, y[i]);
            if (kind == AMH_TARGET_NIG_TOY_LOG) acc = acc + logs;
            return acc;
        }
        case AMH_TARGET_LOGISTIC: {
            double lp; logp_grad(x, lp, nullptr); return lp;
        }
        }
        return NAN;
    }

    bool has_grad() const {
        if (kind == AMH_TARGET_USER) return user_lpg != nullptr;
        return kind == AMH_TARGET_MVNORMAL || kind == AMH_TARGET_GAUSS_PREC ||
               kind == AMH_TARGET_IID_NORMAL || kind == AMH_TARGET_LOGISTIC ||
               kind == AMH_TARGET_ROSENBROCK;
    }

    /* LogDensityProblems.logdensity_and_gradient (MALA.jl:100-105) */
    void logp_grad(const double* x, double& lp, double* g) const {
        const int d = dim;
        switch (kind) {
        case AMH_TARGET_USER: user_lpg(x, d, blob.data(), (long long)ndata, &lp, g); return;
        case AMH_TARGET_GAUSS_PREC: {
            /* (-x'Ax/2, -A x)  test/runtests.jl:343-345 */
            lp = logp(x);
            if (g) {
                const double* A = blob.data();
                for (int i = 0; i < d; ++i) {
                    double t = A[(int64_t)i * d] * x[0];
                    for (int j = 1; j < d; ++j) t = fma(A[(int64_t)i * d + j], x[j], t);
                    g[i] = -t;
                }
            }
            return;
        }
        case AMH_TARGET_MVNORMAL: {
            /* grad = -U'(U(x-mu)) */
            const double c0 = blob[0];
            const double* mu = &blob[1];
            const double* U = &blob[1 + d];
            std::vector<double> w(d);
            double q = 0.0;
            for (int i = 0; i < d; ++i) {
                double t = U[tri(i, 0)] * (x[0] - mu[0]);
                for (int j = 1; j <= i; ++j) t = fma(U[tri(i, j)], x[j] - mu[j], t);
                w[i] = t;
                q = (i == 0) ? t * t : fma(t, t, q);
            }
            lp = fma(-0.5, q, c0);
            if (g) {
                for (int j = 0; j < d; ++j) {
                    double t = U[tri(j, j)] * w[j];
                    for (int i = j + 1; i < d; ++i) t = fma(U[tri(i, j)], w[i], t);
                    g[j] = -t;
                }
            }
            return;
        }
        case AMH_TARGET_IID_NORMAL: {
            lp = logp(x);
            if (g) {
                const double mu = x[0], sigma = x[1];
                double s1 = 0.0, s2 = 0.0;
                for (int64_t i = 0; i < ndata; ++i) {
                    const double r = blob[i] - mu;
                    s1 = s1 + r;
                    s2 = fma(r, r, s2);
                }
                const double v = sigma * sigma;
                g[0] = s1 / v;
                g[1] = s2 / (v * sigma) - (double)ndata / sigma;
            }
            return;
        }
        case AMH_TARGET_ROSENBROCK: {
            lp = logp(x);
            if (g) {
                const double a = blob[0], b = blob[1], s = blob[2];
                for (int i = 0; i < d; ++i) g[i] = 0.0;
                for (int i = 0; i + 1 < d; ++i) {
                    const double t1 = fma(-x[i], x[i], x[i + 1]);
                    const double t2 = a - x[i];
                    /* d/dx_i: b*2*t1*(-2 x_i) - 2 t2 ; d/dx_{i+1}: 2 b t1 */
                    g[i] = g[i] + (-4.0 * b * t1 * x[i] - 2.0 * t2);
                    g[i + 1] = g[i + 1] + 2.0 * b * t1;
                }
                for (int i = 0; i < d; ++i) g[i] = -(g[i] / s);
            }
            return;
        }
        case AMH_TARGET_LOGISTIC: {
            /* blob = [tau, X[n*d], y[n]] */
            const double* X = &blob[1];
            const double* y = &blob[1 + ndata * d];
            /* log-likelihood: 8 interleaved partial sums (rows i mod 8, each in ascending order) combined by a fixed
             * tree -- the order of the tiled device kernel (8-row DMMA tiles + warp-shuffle butterfly); the
             * reference's `sum` over rows has no defined order of its own (pairwise in Julia's Base). */
            double llp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            std::vector<double> gg;
            if (g) gg.assign(d, 0.0);
            for (int64_t i = 0; i < ndata; ++i) {
                const double* xi = X + i * d;
                double eta = xi[0] * x[0];
                for (int j = 1; j < d; ++j) eta = fma(xi[j], x[j], eta);
                llp[i & 7] = llp[i & 7] + (y[i] * eta - amh::log1pexp(eta));
                if (g) {
                    const double r = y[i] - amh::sigmoid(eta);
                    for (int j = 0; j < d; ++j) gg[j] = fma(xi[j], r, gg[j]);
                }
            }
            const double ll = ((llp[0] + llp[1]) + (llp[2] + llp[3])) + ((llp[4] + llp[5]) + (llp[6] + llp[7]));
            double q = x[0] * x[0];
            for (int j = 1; j < d; ++j) q = fma(x[j], x[j], q);
            lp = ll - q * inv2tau2;
            if (g) for (int j = 0; j < d; ++j) g[j] = gg[j] - x[j] * invtau2;
            return;
        }
        }
        lp = NAN;
    }
};

/* ----------------------------------------------------------------- sampler */
struct Sampler {
    amh_sampler_desc d{};
    std::vector<double> mean, scale, S0;
    bool has_mean = false;
    double mala_sigma = 0;
    std::vector<amh_component> comps;      /* AMH_COV_COMPONENTS / AMH_SAMPLER_MIXED */
    bool by_components() const { return !comps.empty(); }

    /* v = map(rand, p.proposal) for an array of univariate laws (proposal.jl:26-28); `blk0` = first sub-stream block */
    void draw_components(const double* z, uint64_t seed, uint64_t blk0, double* v) const {
        for (int i = 0; i < d.dim; ++i)
            v[i] = amh::family_draw(comps[i].family, comps[i].p0, comps[i].p1, z[i], seed, blk0 + (uint64_t)i);
    }
    /* logpdf(p, a) = mapreduce(logpdf, +, zip(p.proposal, a)): left-to-right sum (proposal.jl:32-35) */
    double logq_components(const double* a) const {
        double acc = 0.0;
        for (int i = 0; i < d.dim; ++i) {
            const double l = amh::family_logpdf(comps[i].family, comps[i].p0, comps[i].p1, comps[i].logc, a[i]);
            acc = (i == 0) ? l : acc + l;
        }
        return acc;
    }

    /* v = rand(rng, proposal) given the step's standard normals z
     * (proposal.jl:25-28; Distributions: mu + unwhiten(Sigma, z)) */
    void draw(const double* z, double* v) const {
        const int n = d.dim;
        for (int i = 0; i < n; ++i) {
            double t;
            if (d.cov_kind == AMH_COV_FULL) {
                t = scale[tri(i, 0)] * z[0];
                for (int j = 1; j <= i; ++j) t = fma(scale[tri(i, j)], z[j], t);
            } else if (d.cov_kind == AMH_COV_DIAG) {
                t = scale[i] * z[i];
            } else {
                t = scale[0] * z[i];
            }
            v[i] = has_mean ? t + mean[i] : t;
        }
    }

    /* log-density of the proposal distribution at `a`, up to its (constant)
     * normaliser: -1/2 |L^{-1}(a - mean)|^2  (proposal.jl:31-35, A.2) */
    double logq(const double* a) const {
        const int n = d.dim;
        std::vector<double> w(n);
        double q = 0.0;
        for (int i = 0; i < n; ++i) {
            double s = has_mean ? a[i] - mean[i] : a[i];
            double wi;
            if (d.cov_kind == AMH_COV_FULL) {
                for (int j = 0; j < i; ++j) s = fma(-scale[tri(i, j)], w[j], s);
                wi = s / scale[tri(i, i)];
            } else if (d.cov_kind == AMH_COV_DIAG) {
                wi = s / scale[i];
            } else {
                wi = s / scale[0];
            }
            w[i] = wi;
            q = (i == 0) ? wi * wi : fma(wi, wi, q);
        }
        return -0.5 * q;
    }
};

/* --------------------------------------------------------------------- run */
struct Run {
    Target* t = nullptr;
    Sampler* s = nullptr;
    int64_t n = 0, off = 0;
    int dim = 0;
    std::vector<uint64_t> seeds;
    std::vector<double> X, lp, lq, G, S;
    std::vector<double> logalpha, eta;
    std::vector<uint8_t> acc, failed;
    std::vector<int64_t> nacc;
    int64_t step = 0;          /* stateful steps taken so far */
    int cv = AMH_CONTRACT_VERSION;   /* contract version of the step noise (amh_sampler_desc.contract) */
    /* saved-sample moments */
    std::vector<double> sum, sumsq;
    int64_t nsaved = 0;
};

/* the d standard normals of step `step`: stream 0 = the chain's step noise under contract `cv`; stream 1 = the initial
 * draws of an ensemble's walkers (always the v1 layout: one block per pair) */
void normals(int cv, uint64_t seed, uint64_t step, uint32_t stream, int d, double* z) {
    if (stream == 0) {
        amh::step_normals_cv(cv, seed, step * amh::blocks_per_step_cv(cv, d), d, z);
        return;
    }
    const uint64_t B = (uint64_t)((d + 1) / 2);
    const uint64_t b0 = step * B;
    for (int j = 0; 2 * j < d; ++j) {
        const amh::Block b = amh::stream_block(seed, b0 + j, stream);
        double z0, z1;
        amh::normal_pair(b, z0, z1);
        z[2 * j] = z0;
        if (2 * j + 1 < d) z[2 * j + 1] = z1;
    }
}

double step_exponential(int cv, uint64_t seed, uint64_t step, int d) {
    return amh::step_exponential_cv(cv, seed, step * amh::blocks_per_step_cv(cv, d), d);
}

template <class F>
void parallel_for(int64_t n, F f) {
    int T = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if ((int64_t)T > n) T = (int)std::max<int64_t>(1, n);
    if (T == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    const int64_t chunk = (n + T - 1) / T;
    for (int k = 0; k < T; ++k) {
        const int64_t a = k * chunk, b = std::min<int64_t>(n, a + chunk);
        if (a >= b) break;
        th.emplace_back([=] { f(a, b); });
    }
    for (auto& x : th) x.join();
}

/* ---- MH step for chains [a,b): mh-core.jl:92-117 ---- */
void mh_steps(Run& r, int64_t a, int64_t b, int64_t nsteps) {
    const int d = r.dim;
    const Sampler& sp = *r.s;
    const bool is_rw = sp.d.kind == AMH_SAMPLER_RW;
    const bool sym = sp.d.symmetric != 0;
    std::vector<double> x(d), z(d), v(d), c(d), t1(d), t2(d);
    for (int64_t ch = a; ch < b; ++ch) {
        const uint64_t seed = r.seeds[ch];
        for (int i = 0; i < d; ++i) x[i] = r.X[(int64_t)i * r.n + ch];
        double lp = r.lp[ch], lq = r.lq[ch];
        int64_t nacc = r.nacc[ch];
        uint8_t accepted = r.acc[ch];
        for (int64_t s = 0; s < nsteps; ++s) {
            const uint64_t k = (uint64_t)(r.step + s + 1);
            normals(r.cv, seed, k, 0, d, z.data());
            sp.draw(z.data(), v.data());
            /* candidate = t + rand(rng, proposal) (proposal.jl:49-56) | rand(rng, proposal) (:70-77) */
            for (int i = 0; i < d; ++i) c[i] = is_rw ? x[i] + v[i] : v[i];
            const double lp_c = r.t->logp(c.data());
            /* logratio_proposal_density (mh-core.jl:119-123 -> proposal.jl:190-196) */
            double logratio = 0.0, lq_c = 0.0;
            if (!sym) {
                if (is_rw) {
                    /* q(p,state,cand) - q(p,cand,state) = logpdf(p, x - c) - logpdf(p, c - x);
                     * for a zero-mean Gaussian the two arguments are exact negatives and the
                     * difference is exactly 0.0, so it is skipped. */
                    if (sp.has_mean) {
                        for (int i = 0; i < d; ++i) { t1[i] = x[i] - c[i]; t2[i] = c[i] - x[i]; }
                        logratio = sp.logq(t1.data()) - sp.logq(t2.data());
                    }
                } else {
                    /* static: logpdf(p, state) - logpdf(p, cand); logpdf(p, state) was computed
                     * when `state` was itself the candidate (same value, cached in lq) */
                    lq_c = sp.logq(c.data());
                    logratio = lq - lq_c;
                }
            }
            const double loga = (lp_c - lp) + logratio;
            const double e = step_exponential(r.cv, seed, k, d);
            if (-e < loga) {                       /* mh-core.jl:108 (strict) */
                for (int i = 0; i < d; ++i) x[i] = c[i];
                lp = lp_c; lq = lq_c; accepted = 1; ++nacc;
            } else {
                accepted = 0;
            }
        }
        for (int i = 0; i < d; ++i) r.X[(int64_t)i * r.n + ch] = x[i];
        r.lp[ch] = lp; r.lq[ch] = lq; r.nacc[ch] = nacc; r.acc[ch] = accepted;
    }
}

/* ---- MH step with an array of univariate laws (one Proposal over an array: proposal.jl:26-35,41-85) or an array of
 * Proposals (AMH_SAMPLER_MIXED: proposal.jl:132-150 propose, :236-240 logratio = left-to-right sum of the per-component
 * ratios, symmetric components contributing the literal 0 of :195-196) ---- */
void mh_component_steps(Run& r, int64_t a, int64_t b, int64_t nsteps) {
    const int d = r.dim;
    const Sampler& sp = *r.s;
    const bool mixed = sp.d.kind == AMH_SAMPLER_MIXED;
    const bool is_rw = sp.d.kind == AMH_SAMPLER_RW;
    const bool sym = sp.d.symmetric != 0;
    std::vector<double> x(d), z(d), v(d), c(d), t1(d), t2(d);
    for (int64_t ch = a; ch < b; ++ch) {
        const uint64_t seed = r.seeds[ch];
        for (int i = 0; i < d; ++i) x[i] = r.X[(int64_t)i * r.n + ch];
        double lp = r.lp[ch];
        int64_t nacc = r.nacc[ch];
        uint8_t accepted = r.acc[ch];
        for (int64_t s = 0; s < nsteps; ++s) {
            const uint64_t k = (uint64_t)(r.step + s + 1);
            normals(r.cv, seed, k, 0, d, z.data());
            sp.draw_components(z.data(), seed, k * (uint64_t)d, v.data());
            for (int i = 0; i < d; ++i) {
                const bool rw_i = mixed ? sp.comps[i].rw != 0 : is_rw;
                c[i] = rw_i ? x[i] + v[i] : v[i];
            }
            const double lp_c = r.t->logp(c.data());
            double logratio = 0.0;
            if (mixed) {
                for (int i = 0; i < d; ++i) {
                    const amh_component& q = sp.comps[i];
                    double lr = 0.0;
                    if (!q.symmetric) {
                        if (q.rw)
                            lr = amh::family_logpdf(q.family, q.p0, q.p1, q.logc, x[i] - c[i]) -
                                 amh::family_logpdf(q.family, q.p0, q.p1, q.logc, c[i] - x[i]);
                        else
                            lr = amh::family_logpdf(q.family, q.p0, q.p1, q.logc, x[i]) -
                                 amh::family_logpdf(q.family, q.p0, q.p1, q.logc, c[i]);
                    }
                    logratio = (i == 0) ? lr : logratio + lr;
                }
            } else if (!sym) {
                if (is_rw) {
                    for (int i = 0; i < d; ++i) { t1[i] = x[i] - c[i]; t2[i] = c[i] - x[i]; }
                    logratio = sp.logq_components(t1.data()) - sp.logq_components(t2.data());
                } else {
                    logratio = sp.logq_components(x.data()) - sp.logq_components(c.data());
                }
            }
            const double loga = (lp_c - lp) + logratio;
            const double e = step_exponential(r.cv, seed, k, d);
            if (-e < loga) {                       /* mh-core.jl:108 (strict; NaN rejects) */
                for (int i = 0; i < d; ++i) x[i] = c[i];
                lp = lp_c; accepted = 1; ++nacc;
            } else {
                accepted = 0;
            }
        }
        for (int i = 0; i < d; ++i) r.X[(int64_t)i * r.n + ch] = x[i];
        r.lp[ch] = lp; r.nacc[ch] = nacc; r.acc[ch] = accepted;
    }
}

/* ---- MALA step: MALA.jl:54-93 ---- */
void mala_steps(Run& r, int64_t a, int64_t b, int64_t nsteps) {
    const int d = r.dim;
    const Sampler& sp = *r.s;
    const double sigma = sp.mala_sigma, sigma2 = sp.d.mala_sigma2, drift = sp.d.mala_drift;
    std::vector<double> x(d), g(d), z(d), c(d), gc(d);
    for (int64_t ch = a; ch < b; ++ch) {
        const uint64_t seed = r.seeds[ch];
        for (int i = 0; i < d; ++i) { x[i] = r.X[(int64_t)i * r.n + ch]; g[i] = r.G[(int64_t)i * r.n + ch]; }
        double lp = r.lp[ch];
        int64_t nacc = r.nacc[ch];
        uint8_t accepted = r.acc[ch];
        for (int64_t s = 0; s < nsteps; ++s) {
            const uint64_t k = (uint64_t)(r.step + s + 1);
            normals(r.cv, seed, k, 0, d, z.data());
            /* state + rand(MvNormal(drift*grad, sigma2*I)) (MALA.jl:70 -> proposal.jl:49-56) */
            for (int i = 0; i < d; ++i) c[i] = x[i] + (sigma * z[i] + drift * g[i]);
            double lp_c;
            r.t->logp_grad(c.data(), lp_c, gc.data());
            /* q(prop(grad_c), state, cand) - q(prop(grad), cand, state)  (MALA.jl:78-80) */
            double A = 0.0, B = 0.0;
            for (int i = 0; i < d; ++i) {
                const double da = (x[i] - c[i]) - drift * gc[i];
                const double db = (c[i] - x[i]) - drift * g[i];
                A = (i == 0) ? da * da : fma(da, da, A);
                B = (i == 0) ? db * db : fma(db, db, B);
            }
            const double logratio = (-0.5 * (A / sigma2)) - (-0.5 * (B / sigma2));
            const double loga = (lp_c - lp) + logratio;
            const double e = step_exponential(r.cv, seed, k, d);
            if (-e < loga) {
                x = c; g = gc; lp = lp_c; accepted = 1; ++nacc;
            } else {
                accepted = 0;
            }
        }
        for (int i = 0; i < d; ++i) { r.X[(int64_t)i * r.n + ch] = x[i]; r.G[(int64_t)i * r.n + ch] = g[i]; }
        r.lp[ch] = lp; r.nacc[ch] = nacc; r.acc[ch] = accepted;
    }
}

/* ---- RAM: RobustAdaptiveMetropolis.jl:123-173, 216-278; A.4 for the Givens sweeps ---- */
void ram_steps(Run& r, int64_t a, int64_t b, int64_t nsteps, bool warmup) {
    const int d = r.dim;
    const Sampler& sp = *r.s;
    const int64_t nt = (int64_t)d * (d + 1) / 2;
    const double alpha = sp.d.ram_alpha, gamma = sp.d.ram_gamma;
    const double lo = sp.d.ram_eig_lo, hi = sp.d.ram_eig_hi;
    const bool check = !(lo == 0.0 && hi == INFINITY);
    std::vector<double> x(d), U(d), su(d), xn(d), v(d), S(nt), Sn(nt);
    for (int64_t ch = a; ch < b; ++ch) {
        const uint64_t seed = r.seeds[ch];
        for (int i = 0; i < d; ++i) x[i] = r.X[(int64_t)i * r.n + ch];
        for (int64_t q = 0; q < nt; ++q) S[q] = r.S[q * r.n + ch];
        double lp = r.lp[ch], logalpha = r.logalpha[ch], eta = r.eta[ch];
        int64_t nacc = r.nacc[ch];
        uint8_t accepted = r.acc[ch], failed = r.failed[ch];
        for (int64_t s = 0; s < nsteps; ++s) {
            const uint64_t k = (uint64_t)(r.step + s + 1);
            const int64_t iteration = r.step + s + 1;     /* state.iteration (starts at 1, RAM :211) */
            normals(r.cv, seed, k, 0, d, U.data());
            /* x_new = muladd(S, U, x)  (RAM :136) */
            for (int i = 0; i < d; ++i) {
                double t = S[tri(i, 0)] * U[0];
                for (int j = 1; j <= i; ++j) t = fma(S[tri(i, j)], U[j], t);
                su[i] = t;
                xn[i] = t + x[i];
            }
            const double lp_new = r.t->logp(xn.data());
            const double dl = lp_new - lp;
            logalpha = (dl != dl) ? dl : (dl < 0.0 ? dl : 0.0);    /* min(lp_new - lp, 0)  (:147) */
            const double e = step_exponential(r.cv, seed, k, d);
            const bool isaccept = e > -logalpha;                     /* (:148) */
            if (warmup) {
                /* ram_adapt (:153-173) */
                const double dalpha = amh::exp_(logalpha) - alpha;
                eta = amh::exp_(-gamma * amh::log_((double)iteration));   /* iteration^(-gamma) */
                if (dalpha == dalpha) {
                    const double cc = sqrt(eta * fabs(dalpha));
                    double nu = U[0] * U[0];
                    for (int i = 1; i < d; ++i) nu = fma(U[i], U[i], nu);
                    nu = sqrt(nu);
                    for (int i = 0; i < d; ++i) v[i] = (cc * su[i]) / nu;
                    Sn = S;
                    bool ok = true;
                    if (dalpha > 0.0) {
                        /* lowrankupdate */
                        for (int i = 0; i < d; ++i) {
                            const double f = Sn[tri(i, i)], g = v[i];
                            const double rr = sqrt(fma(f, f, g * g));
                            const double c = f / rr, sn = g / rr;
                            Sn[tri(i, i)] = rr;
                            for (int j = i + 1; j < d; ++j) {
                                const double Aji = Sn[tri(j, i)];
                                Sn[tri(j, i)] = c * Aji + sn * v[j];
                                v[j] = c * v[j] - sn * Aji;
                            }
                        }
                    } else {
                        /* lowrankdowndate; s^2 > 1 is the reference's PosDefException */
                        for (int i = 0; i < d && ok; ++i) {
                            const double Aii = Sn[tri(i, i)];
                            const double sn = v[i] / Aii;
                            const double s2 = sn * sn;
                            if (s2 > 1.0) { ok = false; break; }
                            const double c = sqrt(1.0 - s2);
                            Sn[tri(i, i)] = c * Aii;
                            for (int j = i + 1; j < d; ++j) {
                                const double Aji = (Sn[tri(j, i)] - sn * v[j]) / c;
                                Sn[tri(j, i)] = Aji;
                                v[j] = -sn * Aji + c * v[j];
                            }
                        }
                        if (!ok) failed = 1;
                    }
                    /* valid_eigenvalues (:239-245): diagonal of the triangular factor */
                    if (ok && check) {
                        for (int i = 0; i < d; ++i) {
                            const double ev = Sn[tri(i, i)];
                            if (!(lo <= ev && ev <= hi)) { ok = false; break; }
                        }
                    }
                    if (ok) S = Sn;
                } else {
                    failed = 1;
                }
            }
            if (isaccept) { x = xn; lp = lp_new; ++nacc; }
            accepted = isaccept ? 1 : 0;
        }
        for (int i = 0; i < d; ++i) r.X[(int64_t)i * r.n + ch] = x[i];
        for (int64_t q = 0; q < nt; ++q) r.S[q * r.n + ch] = S[q];
        r.lp[ch] = lp; r.logalpha[ch] = logalpha; r.eta[ch] = eta;
        r.nacc[ch] = nacc; r.acc[ch] = accepted; r.failed[ch] = failed;
    }
}

/* ---- stretch-move sweeps for ensembles [a,b): emcee.jl:39-58, 70-102 ---- */
void stretch_steps(Run& r, int64_t ea, int64_t eb, int64_t nsteps) {
    const int d = r.dim;
    const Sampler& sp = *r.s;
    const int64_t nw = sp.d.n_walkers;
    const double aa = sp.d.stretch_a;
    std::vector<double> oldx((size_t)nw * d), newx((size_t)nw * d), oldlp(nw), newlp(nw), y(d);
    std::vector<uint8_t> newacc(nw);
    for (int64_t en = ea; en < eb; ++en) {
        const uint64_t seed = r.seeds[en];
        const int64_t base = en * nw;
        for (int64_t w = 0; w < nw; ++w) {
            for (int i = 0; i < d; ++i) oldx[w * d + i] = r.X[(int64_t)i * r.n + base + w];
            oldlp[w] = r.lp[base + w];
        }
        for (int64_t s = 0; s < nsteps; ++s) {
            const uint64_t k = (uint64_t)(r.step + s + 1);
            for (int64_t i = 0; i < nw; ++i) {
                const uint64_t blk = (k * (uint64_t)nw + (uint64_t)i) * 2;
                const amh::Block b0 = amh::stream_block(seed, blk, 0);
                const amh::Block b1 = amh::stream_block(seed, blk + 1, 0);
                /* idx = mod1(i + rand(1:(n-1)), n)  (emcee.jl:52) */
                const int64_t rr = (int64_t)amh::bounded(b0.v[0], b0.v[1], (uint64_t)(nw - 1));
                const int64_t idx = (i + rr + 1) % nw;
                /* other = idx < i ? new_walkers[idx] : walkers[idx]  (emcee.jl:53) */
                const double* other = (idx < i) ? &newx[idx * d] : &oldx[idx * d];
                const double* walker = &oldx[i * d];
                /* move (emcee.jl:70-102) */
                const double u = amh::u01(b0.v[2], b0.v[3]);
                const double t = (aa - 1.0) * u + 1.0;
                const double z = (t * t) / aa;
                const double alphamult = (double)(d - 1) * amh::log_(z);
                for (int j = 0; j < d; ++j) y[j] = other[j] + z * (walker[j] - other[j]);
                const double lpy = r.t->logp(y.data());
                const double alpha = (alphamult + lpy) - oldlp[i];
                const double e = amh::exponential(b1.v[0], b1.v[1]);
                if (-e <= alpha) {                 /* emcee.jl:93 (non-strict) */
                    for (int j = 0; j < d; ++j) newx[i * d + j] = y[j];
                    newlp[i] = lpy; newacc[i] = 1; r.nacc[base + i] += 1;
                } else {
                    for (int j = 0; j < d; ++j) newx[i * d + j] = walker[j];
                    newlp[i] = oldlp[i]; newacc[i] = 0;
                }
            }
            oldx.swap(newx); oldlp.swap(newlp);
        }
        for (int64_t w = 0; w < nw; ++w) {
            for (int i = 0; i < d; ++i) r.X[(int64_t)i * r.n + base + w] = oldx[w * d + i];
            r.lp[base + w] = oldlp[w];
            if (nsteps > 0) r.acc[base + w] = newacc[w];
        }
    }
}

int do_steps(Run& r, int64_t nsteps, bool warmup) {
    if (nsteps <= 0) return AMH_OK;
    switch (r.s->d.kind) {
    case AMH_SAMPLER_STATIC:
    case AMH_SAMPLER_RW:
    case AMH_SAMPLER_MIXED:
        if (r.s->by_components())
            parallel_for(r.n, [&](int64_t a, int64_t b) { mh_component_steps(r, a, b, nsteps); });
        else
            parallel_for(r.n, [&](int64_t a, int64_t b) { mh_steps(r, a, b, nsteps); });
        break;
    case AMH_SAMPLER_MALA:
        parallel_for(r.n, [&](int64_t a, int64_t b) { mala_steps(r, a, b, nsteps); });
        break;
    case AMH_SAMPLER_RAM:
        parallel_for(r.n, [&](int64_t a, int64_t b) { ram_steps(r, a, b, nsteps, warmup); });
        break;
    case AMH_SAMPLER_STRETCH:
        parallel_for(r.n / r.s->d.n_walkers, [&](int64_t a, int64_t b) { stretch_steps(r, a, b, nsteps); });
        break;
    default:
        return fail(AMH_ERR_INVALID, "unknown sampler kind");
    }
    r.step += nsteps;
    return AMH_OK;
}

/* Welford running mean (r.sum) and M2 (r.sumsq) per (coordinate, chain) over the saved samples */
void accumulate(Run& r) {
    const int64_t n = r.n;
    const double inv_n = 1.0 / (double)(r.nsaved + 1);
    for (int i = 0; i < r.dim; ++i)
        for (int64_t c = 0; c < n; ++c) {
            const int64_t o = (int64_t)i * n + c;
            const double v = r.X[o];
            const double m = r.sum[o];
            const double dl = v - m;
            const double m1 = fma(dl, inv_n, m);
            r.sum[o] = m1;
            r.sumsq[o] = fma(dl, v - m1, r.sumsq[o]);
        }
    r.nsaved += 1;
}

}  // namespace

/* =============================== C ABI (amho_) =============================== */
extern "C" {

int32_t amho_version(int32_t* major, int32_t* minor) {
    if (major) *major = AMH_VERSION_MAJOR;
    if (minor) *minor = AMH_VERSION_MINOR;
    return AMH_OK;
}
const char* amho_last_error(void) { return g_err.c_str(); }
int32_t amho_contract_version(void) { return AMH_CONTRACT_VERSION; }
int32_t amho_set_threads(int32_t n) { g_threads = n; return AMH_OK; }
int32_t amho_get_threads(void) {
    return g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
}

int32_t amho_ctx_create(int32_t, amh_ctx** out) { if (out) *out = (amh_ctx*)(uintptr_t)1; return AMH_OK; }
int32_t amho_ctx_destroy(amh_ctx*) { return AMH_OK; }
int32_t amho_ctx_sync(amh_ctx*) { return AMH_OK; }

int32_t amho_target_create(amh_ctx*, int32_t kind, int32_t dim, const double* blob, int64_t nblob,
                           amh_target** out) {
    if (!out) return fail(AMH_ERR_INVALID, "out is NULL");
    if (dim < 1) return fail(AMH_ERR_INVALID, "dim must be >= 1");
    if (nblob < 0 || (nblob > 0 && !blob)) return fail(AMH_ERR_INVALID, "blob is NULL");
    Target* t = new Target();
    t->kind = kind; t->dim = dim;
    t->blob.assign(blob, blob + nblob);
    const int64_t d = dim;
    bool ok = true;
    switch (kind) {
    case AMH_TARGET_IID_NORMAL: ok = (dim == 2 && nblob >= 1); t->ndata = nblob; break;
    case AMH_TARGET_MVNORMAL: ok = (nblob == 1 + d + d * (d + 1) / 2); break;
    case AMH_TARGET_ROSENBROCK: ok = (nblob == 3 && dim >= 2); break;
    case AMH_TARGET_GAUSS_PREC: ok = (nblob == d * d); break;
    case AMH_TARGET_NIG_TOY:
    case AMH_TARGET_NIG_TOY_LOG: ok = (dim == 2 && nblob >= 3); t->ndata = nblob - 3; break;
    case AMH_TARGET_LOGISTIC:
        ok = (nblob >= 1 + d + 1) && ((nblob - 1) % (d + 1) == 0);
        if (ok) {
            t->ndata = (nblob - 1) / (d + 1);
            const double tau = blob[0];
            t->inv2tau2 = 1.0 / (2.0 * tau * tau);
            t->invtau2 = 1.0 / (tau * tau);
        }
        break;
    default: ok = false;
    }
    if (!ok) { delete t; return fail(AMH_ERR_INVALID, "target kind/dim/blob size mismatch"); }
    *out = (amh_target*)t;
    return AMH_OK;
}
/* The oracle's side of amh_target_create_source: the SAME source text, compiled for the host with the contract flags
 * (g++ -O2 -ffp-contract=off, include/amh_user_target.h in front) into a scratch shared object and dlopen'ed. */
int32_t amho_target_create_source(amh_ctx*, int32_t dim, const char* source, int32_t has_gradient,
                                  const double* data, int64_t ndata, amh_target** out) {
    if (!out) return fail(AMH_ERR_INVALID, "out is NULL");
    if (!source || !*source) return fail(AMH_ERR_INVALID, "source is NULL or empty");
    if (dim < 1) return fail(AMH_ERR_INVALID, "dim must be >= 1");
    if (dim > 128) return fail(AMH_ERR_UNSUPPORTED, "user-supplied targets support dim <= 128");
    if (ndata < 0 || (ndata > 0 && !data)) return fail(AMH_ERR_INVALID, "data is NULL");
    Dl_info info;
    std::string incdir = "../include";
    if (dladdr((void*)&amho_target_create_source, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        const size_t k = p.rfind('/');
        incdir = (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/../include";
    }
    /* a PRIVATE scratch directory (mkdtemp: mode 0700, unpredictable name) holds source, log and shared object, and the
     * compiler is spawned without a shell, so neither a symlink race in /tmp nor AMHO_CXX can redirect what is loaded */
    char tmpl[] = "/tmp/amho_user_XXXXXX";
    if (!mkdtemp(tmpl)) return fail(AMH_ERR_STATE, "mkdtemp failed");
    const std::string dir(tmpl), cpp = dir + "/target.cpp", so = dir + "/target.so", log = dir + "/build.log";
    auto cleanup = [&](bool keep_so) {
        std::remove(cpp.c_str()); std::remove(log.c_str());
        if (!keep_so) { std::remove(so.c_str()); rmdir(dir.c_str()); }
    };
    {
        FILE* f = fopen(cpp.c_str(), "wx");
        if (!f) { cleanup(false); return fail(AMH_ERR_STATE, "cannot write the scratch source file"); }
        fputs("#include \"amh_user_target.h\"\n#line 1 \"amh_user_target.cu\"\n", f);
        fputs(source, f);
        fputs("\n", f);
        fclose(f);
    }
    const char* cxx_env = getenv("AMHO_CXX");
    const std::string cxx = cxx_env ? cxx_env : "g++", inc = "-I" + incdir;
    int rc = -1;
    {
        const int lfd = open(log.c_str(), O_WRONLY | O_CREAT | O_EXCL, 0600);
        posix_spawn_file_actions_t fa;
        posix_spawn_file_actions_init(&fa);
        if (lfd >= 0) { posix_spawn_file_actions_adddup2(&fa, lfd, 1); posix_spawn_file_actions_adddup2(&fa, lfd, 2); }
        const char* argv[] = {cxx.c_str(), "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-mfma",
                              inc.c_str(), "-o", so.c_str(), cpp.c_str(), nullptr};
        pid_t pid = 0;
        if (posix_spawnp(&pid, cxx.c_str(), &fa, nullptr, (char* const*)argv, environ) == 0) {
            int status = 0;
            if (waitpid(pid, &status, 0) == pid && WIFEXITED(status)) rc = WEXITSTATUS(status);
        }
        posix_spawn_file_actions_destroy(&fa);
        if (lfd >= 0) close(lfd);
    }
    std::string logtxt;
    if (FILE* f = fopen(log.c_str(), "r")) {
        char buf[4096];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), f)) > 0) logtxt.append(buf, n);
        fclose(f);
    }
    cleanup(rc == 0);
    if (rc != 0) return fail(AMH_ERR_INVALID, "the target source does not compile:\n" + logtxt);
    Target* t = new Target();
    t->kind = AMH_TARGET_USER; t->dim = dim; t->ndata = ndata; t->user_so = so;
    if (ndata > 0) t->blob.assign(data, data + ndata);
    else t->blob.assign(1, 0.0);
    t->user_lib = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!t->user_lib) { const std::string e = dlerror(); delete t; return fail(AMH_ERR_INVALID, "dlopen of the compiled target failed: " + e); }
    t->user_lp = (Target::user_lp_fn)dlsym(t->user_lib, "amh_user_logdensity");
    if (!t->user_lp) { delete t; return fail(AMH_ERR_INVALID, "the target source does not define amh_user_logdensity"); }
    if (has_gradient) {
        t->user_lpg = (Target::user_lpg_fn)dlsym(t->user_lib, "amh_user_logdensity_and_gradient");
        if (!t->user_lpg) { delete t; return fail(AMH_ERR_INVALID, "the target source does not define amh_user_logdensity_and_gradient"); }
    }
    *out = (amh_target*)t;
    return AMH_OK;
}
int32_t amho_target_destroy(amh_target* t) { delete (Target*)t; return AMH_OK; }

/* parameter validation of one univariate law (Distributions.jl constructors throw DomainError) */
static const char* check_component(const amh_component& q) {
    switch (q.family) {
    case AMH_FAM_NORMAL:
    case AMH_FAM_LOGNORMAL:   return (q.p1 > 0.0) ? nullptr : "Normal / LogNormal need sigma > 0";
    case AMH_FAM_INVGAMMA:
    case AMH_FAM_GAMMA:       return (q.p0 > 0.0 && q.p1 > 0.0) ? nullptr : "Gamma / InverseGamma need shape > 0 and scale > 0";
    case AMH_FAM_UNIFORM:     return (q.p1 > q.p0) ? nullptr : "Uniform needs a < b";
    case AMH_FAM_EXPONENTIAL: return (q.p0 > 0.0) ? nullptr : "Exponential needs scale > 0";
    }
    return "unknown distribution family";
}

int32_t amho_sampler_create(amh_ctx*, const amh_sampler_desc* desc, amh_sampler** out) {
    if (!desc || !out) return fail(AMH_ERR_INVALID, "desc/out is NULL");
    const int d = desc->dim;
    if (d < 1) return fail(AMH_ERR_INVALID, "dim must be >= 1");
    Sampler* s = new Sampler();
    s->d = *desc;
    const int64_t nt = (int64_t)d * (d + 1) / 2;
    auto bad = [&](const char* m) { delete s; return fail(AMH_ERR_INVALID, m); };
    switch (desc->kind) {
    case AMH_SAMPLER_STATIC:
    case AMH_SAMPLER_RW:
    case AMH_SAMPLER_MIXED:
    case AMH_SAMPLER_STRETCH: {
        if (desc->kind == AMH_SAMPLER_STRETCH) {
            if (desc->n_walkers < 2) return bad("Ensemble needs n_walkers >= 2");
            if (!(desc->stretch_a > 1.0)) return bad("stretch_length must be > 1");
        }
        if (desc->kind == AMH_SAMPLER_MIXED || desc->cov_kind == AMH_COV_COMPONENTS) {
            if (!desc->components) return bad("components is NULL");
            s->comps.assign(desc->components, desc->components + d);
            for (const amh_component& q : s->comps) {
                const char* m = check_component(q);
                if (m) return bad(m);
            }
            break;
        }
        const bool need_cov = desc->kind != AMH_SAMPLER_STRETCH || desc->scale != nullptr;
        if (need_cov) {
            if (!desc->scale) return bad("proposal scale is NULL");
            const int64_t ns = desc->cov_kind == AMH_COV_FULL ? nt : desc->cov_kind == AMH_COV_DIAG ? d
                               : desc->cov_kind == AMH_COV_SCALAR ? 1 : -1;
            if (ns < 0) return bad("unknown cov_kind");
            s->scale.assign(desc->scale, desc->scale + ns);
            for (int i = 0; i < d; ++i) {
                const double dg = desc->cov_kind == AMH_COV_FULL ? s->scale[tri(i, i)]
                                  : desc->cov_kind == AMH_COV_DIAG ? s->scale[i] : s->scale[0];
                if (!(dg > 0.0)) return bad("proposal scale must have a positive diagonal");
            }
        }
        if (desc->mean) { s->mean.assign(desc->mean, desc->mean + d); s->has_mean = true; }
        break;
    }
    case AMH_SAMPLER_MALA:
        if (!(desc->mala_sigma2 > 0.0)) return bad("MALA sigma2 must be > 0");
        s->mala_sigma = sqrt(desc->mala_sigma2);
        break;
    case AMH_SAMPLER_RAM:
        if (desc->ram_S0) {
            s->S0.resize(nt);
            for (int i = 0; i < d; ++i)
                for (int j = 0; j <= i; ++j) s->S0[tri(i, j)] = desc->ram_S0[(int64_t)i * d + j];
        }
        break;
    default:
        return bad("unknown sampler kind");
    }
    s->d.mean = nullptr; s->d.scale = nullptr; s->d.ram_S0 = nullptr; s->d.components = nullptr;
    if (s->d.contract == 0) {
        const char* ev = getenv("AMH_CONTRACT");
        s->d.contract = ev ? atoi(ev) : AMH_CONTRACT_VERSION;
    }
    if (s->d.contract != AMH_CONTRACT_V1 && s->d.contract != AMH_CONTRACT_V2) { delete s; return fail(AMH_ERR_INVALID, "unknown contract version"); }
    *out = (amh_sampler*)s;
    return AMH_OK;
}
int32_t amho_sampler_destroy(amh_sampler* s) { delete (Sampler*)s; return AMH_OK; }

int32_t amho_run_create(amh_ctx*, amh_target* target, amh_sampler* sampler, int64_t n, int64_t off,
                        const uint64_t* seeds, const double* init, int64_t init_ld, amh_run** out) {
    if (!target || !sampler || !out || !seeds) return fail(AMH_ERR_INVALID, "NULL argument");
    Target* t = (Target*)target;
    Sampler* s = (Sampler*)sampler;
    if (t->dim != s->d.dim) return fail(AMH_ERR_INVALID, "target and sampler dimensions differ");
    if (n < 1) return fail(AMH_ERR_INVALID, "nchains_local must be >= 1");
    if (init_ld == 0) init_ld = n;
    if (init && init_ld < n) return fail(AMH_ERR_INVALID, "init_ld must be >= nchains_local");
    const int kind = s->d.kind;
    const int d = t->dim;
    int64_t nseeds = n;
    if (kind == AMH_SAMPLER_STRETCH) {
        if (n % s->d.n_walkers) return fail(AMH_ERR_INVALID, "nchains_local must be a multiple of n_walkers");
        nseeds = n / s->d.n_walkers;
    }
    if (kind == AMH_SAMPLER_MALA) {
        /* propose(::MALA) = error("please specify initial parameters")  (MALA.jl:37) */
        if (!init) return fail(AMH_ERR_STATE, "please specify initial parameters");
        if (!t->has_grad()) return fail(AMH_ERR_INVALID,
            "The gradient of the log density function is not defined");   /* MALA.jl:49-51 */
    }
    if (kind == AMH_SAMPLER_STRETCH && !init && s->scale.empty() && !s->by_components())
        return fail(AMH_ERR_INVALID, "stretch move without init needs an initial-draw proposal");
    Run* r = new Run();
    r->t = t; r->s = s; r->n = n; r->off = off; r->dim = d;
    r->cv = s->d.contract;
    r->seeds.assign(seeds, seeds + nseeds);
    r->X.assign((size_t)d * n, 0.0);
    r->lp.assign(n, 0.0); r->lq.assign(n, 0.0);
    r->acc.assign(n, 0); r->failed.assign(n, 0); r->nacc.assign(n, 0);
    r->sum.assign((size_t)d * n, 0.0); r->sumsq.assign((size_t)d * n, 0.0);
    if (kind == AMH_SAMPLER_MALA) r->G.assign((size_t)d * n, 0.0);
    const int64_t nt = (int64_t)d * (d + 1) / 2;
    if (kind == AMH_SAMPLER_RAM) {
        r->S.assign((size_t)nt * n, 0.0);
        r->logalpha.assign(n, 0.0); r->eta.assign(n, 0.0);
    }
    /* ---- first step (mh-core.jl:76-86; emcee.jl:29-34; RAM :175-214) ---- */
    parallel_for(n, [&](int64_t a, int64_t b) {
        std::vector<double> x(d), z(d), g(d);
        for (int64_t ch = a; ch < b; ++ch) {
            if (init) {
                for (int i = 0; i < d; ++i) x[i] = init[(int64_t)i * init_ld + ch];
            } else if (kind == AMH_SAMPLER_RAM) {
                normals(r->cv, r->seeds[ch], 0, 0, d, x.data());     /* randn(rng, T, d)  (:193) */
            } else if (kind == AMH_SAMPLER_STRETCH) {
                /* n_walkers draws from the inner proposal (emcee.jl:29-34): stream 1 of the ensemble */
                const int64_t en = ch / s->d.n_walkers, w = ch % s->d.n_walkers;
                normals(r->cv, r->seeds[en], (uint64_t)w, 1, d, z.data());
                if (s->by_components()) s->draw_components(z.data(), r->seeds[en], (uint64_t)w * (uint64_t)d, x.data());
                else s->draw(z.data(), x.data());
            } else {
                normals(r->cv, r->seeds[ch], 0, 0, d, z.data());
                if (s->by_components()) s->draw_components(z.data(), r->seeds[ch], 0ull, x.data());   /* proposal.jl:132-140 */
                else s->draw(z.data(), x.data());                    /* propose(rng, sampler, model) */
            }
            for (int i = 0; i < d; ++i) r->X[(int64_t)i * n + ch] = x[i];
            if (kind == AMH_SAMPLER_MALA) {
                double lp;
                t->logp_grad(x.data(), lp, g.data());                /* transition(::MALA,...) MALA.jl:38-40 */
                r->lp[ch] = lp;
                for (int i = 0; i < d; ++i) r->G[(int64_t)i * n + ch] = g[i];
            } else {
                r->lp[ch] = t->logp(x.data());
            }
            if (kind == AMH_SAMPLER_STATIC && !s->d.symmetric && !s->by_components()) r->lq[ch] = s->logq(x.data());
            if (kind == AMH_SAMPLER_RAM) {
                for (int i = 0; i < d; ++i)
                    for (int j = 0; j <= i; ++j)
                        r->S[tri(i, j) * n + ch] = s->S0.empty() ? (i == j ? 1.0 : 0.0) : s->S0[tri(i, j)];
                r->acc[ch] = 1;                                      /* Transition(x, lp, true) (:213) */
            }
        }
    });
    *out = (amh_run*)r;
    return AMH_OK;
}
int32_t amho_run_destroy(amh_run* r) { delete (Run*)r; return AMH_OK; }

int32_t amho_run_steps(amh_run* run, int64_t nsteps, int32_t warmup, int32_t) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    if (nsteps < 0) return fail(AMH_ERR_INVALID, "nsteps must be >= 0");
    return do_steps(*(Run*)run, nsteps, warmup != 0);
}
int32_t amho_run_sync(amh_run*) { return AMH_OK; }

int32_t amho_run_sample_ld(amh_run* run, int64_t N, int64_t discard_initial, int64_t thinning,
                           int64_t num_warmup, double* out, int64_t out_ld, uint8_t* accepted_out, int64_t acc_ld,
                           amh_summary* summary);
int32_t amho_run_sample(amh_run* run, int64_t N, int64_t discard_initial, int64_t thinning,
                        int64_t num_warmup, double* out, uint8_t* accepted_out, amh_summary* summary) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    return amho_run_sample_ld(run, N, discard_initial, thinning, num_warmup, out, ((Run*)run)->n, accepted_out, ((Run*)run)->n, summary);
}
int32_t amho_run_sample_ld(amh_run* run, int64_t N, int64_t discard_initial, int64_t thinning,
                           int64_t num_warmup, double* out, int64_t out_ld, uint8_t* accepted_out, int64_t acc_ld,
                           amh_summary* summary) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    if (N < 1 || thinning < 1 || discard_initial < 0 || num_warmup < 0)
        return fail(AMH_ERR_INVALID, "need N >= 1, thinning >= 1, discard_initial >= 0, num_warmup >= 0");
    if ((out && out_ld < ((Run*)run)->n) || (accepted_out && acc_ld < ((Run*)run)->n))
        return fail(AMH_ERR_INVALID, "out_ld / acc_ld must be >= nchains_local");
    Run& r = *(Run*)run;
    const int64_t n = r.n;
    const int d = r.dim;
    std::fill(r.sum.begin(), r.sum.end(), 0.0);
    std::fill(r.sumsq.begin(), r.sumsq.end(), 0.0);
    r.nsaved = 0;
    auto advance = [&](int64_t k) -> int {
        /* stateful step s (1-based, cumulative) is step_warmup iff s <= num_warmup */
        while (k > 0) {
            const bool wu = r.step < num_warmup;
            int64_t m = k;
            if (wu) m = std::min<int64_t>(k, num_warmup - r.step);
            const int rc = do_steps(r, m, wu);
            if (rc) return rc;
            k -= m;
        }
        return AMH_OK;
    };
    for (int64_t i = 0; i < N; ++i) {
        const int rc = advance(i == 0 ? discard_initial : thinning);
        if (rc) return rc;
        if (out) {
            double* o = out + (size_t)i * (d + 1) * out_ld;
            for (int j = 0; j < d; ++j) std::memcpy(o + (size_t)j * out_ld, r.X.data() + (size_t)j * n, sizeof(double) * n);
            std::memcpy(o + (size_t)d * out_ld, r.lp.data(), sizeof(double) * n);
        }
        if (accepted_out) std::memcpy(accepted_out + (size_t)i * acc_ld, r.acc.data(), n);
        accumulate(r);
    }
    if (summary) {
        summary->n_saved = r.nsaved;
        summary->n_steps = r.step;
        double na = 0;
        for (int64_t c = 0; c < n; ++c) na += (double)r.nacc[c];
        summary->accept_rate = r.step > 0 ? na / ((double)n * (double)r.step) : 0.0;
        for (int i = 0; i < d; ++i) {
            double s1 = 0;
            for (int64_t c = 0; c < n; ++c) {
                s1 += r.sum[(int64_t)i * n + c];
                if (summary->chain_mean) summary->chain_mean[(int64_t)i * n + c] = r.sum[(int64_t)i * n + c];
            }
            const double m = s1 / (double)n;
            double m2 = 0, dev = 0;
            for (int64_t c = 0; c < n; ++c) {
                const double dl = r.sum[(int64_t)i * n + c] - m;
                m2 += r.sumsq[(int64_t)i * n + c];
                dev = fma(dl, dl, dev);
            }
            if (summary->mean) summary->mean[i] = m;
            if (summary->var) summary->var[i] = (m2 + (double)r.nsaved * dev) / ((double)n * (double)r.nsaved);
        }
    }
    return AMH_OK;
}

static void copy_rows(double* dst, int64_t dld, const double* src, int64_t sld, int64_t rows, int64_t n) {
    for (int64_t i = 0; i < rows; ++i) std::memcpy(dst + i * dld, src + i * sld, sizeof(double) * n);
}
int32_t amho_run_get_state_ld(amh_run* run, int64_t ld, double* x, double* lp, double* grad, double* S,
                              uint8_t* accepted, int64_t* naccept, int64_t* step_counter);
int32_t amho_run_get_state(amh_run* run, double* x, double* lp, double* grad, double* S,
                           uint8_t* accepted, int64_t* naccept, int64_t* step_counter) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    return amho_run_get_state_ld(run, ((Run*)run)->n, x, lp, grad, S, accepted, naccept, step_counter);
}
int32_t amho_run_get_state_ld(amh_run* run, int64_t ld, double* x, double* lp, double* grad, double* S,
                              uint8_t* accepted, int64_t* naccept, int64_t* step_counter) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    Run& r = *(Run*)run;
    if (ld < r.n) return fail(AMH_ERR_INVALID, "ld must be >= nchains_local");
    if (x) copy_rows(x, ld, r.X.data(), r.n, r.dim, r.n);
    if (lp) std::memcpy(lp, r.lp.data(), sizeof(double) * r.lp.size());
    if (grad) {
        if (r.G.empty()) return fail(AMH_ERR_INVALID, "sampler keeps no gradient");
        copy_rows(grad, ld, r.G.data(), r.n, r.dim, r.n);
    }
    if (S) {
        if (r.S.empty()) return fail(AMH_ERR_INVALID, "sampler keeps no Cholesky factor");
        copy_rows(S, ld, r.S.data(), r.n, (int64_t)r.dim * (r.dim + 1) / 2, r.n);
    }
    if (accepted) std::memcpy(accepted, r.acc.data(), r.acc.size());
    if (naccept) std::memcpy(naccept, r.nacc.data(), sizeof(int64_t) * r.nacc.size());
    if (step_counter) *step_counter = r.step;
    return AMH_OK;
}

int32_t amho_run_set_state_ld(amh_run* run, int64_t ld, const double* x, const double* lp, const double* grad, const double* S,
                              const uint8_t* accepted, const int64_t* naccept, int64_t step_counter);
int32_t amho_run_set_state(amh_run* run, const double* x, const double* lp, const double* grad, const double* S,
                           const uint8_t* accepted, const int64_t* naccept, int64_t step_counter) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    return amho_run_set_state_ld(run, ((Run*)run)->n, x, lp, grad, S, accepted, naccept, step_counter);
}
int32_t amho_run_set_state_ld(amh_run* run, int64_t ld, const double* x, const double* lp, const double* grad, const double* S,
                              const uint8_t* accepted, const int64_t* naccept, int64_t step_counter) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    Run& r = *(Run*)run;
    const int d = r.dim;
    const int64_t n = r.n;
    if (ld < n) return fail(AMH_ERR_INVALID, "ld must be >= nchains_local");
    if (grad && r.G.empty()) return fail(AMH_ERR_INVALID, "sampler keeps no gradient");
    if (S && r.S.empty()) return fail(AMH_ERR_INVALID, "sampler keeps no Cholesky factor");
    if (x) copy_rows(r.X.data(), n, x, ld, d, n);
    if (lp) std::memcpy(r.lp.data(), lp, sizeof(double) * r.lp.size());
    if (grad) copy_rows(r.G.data(), n, grad, ld, d, n);
    if (S) copy_rows(r.S.data(), n, S, ld, (int64_t)d * (d + 1) / 2, n);
    if (accepted) std::memcpy(r.acc.data(), accepted, r.acc.size());
    if (naccept) std::memcpy(r.nacc.data(), naccept, sizeof(int64_t) * r.nacc.size());
    if (step_counter >= 0) r.step = step_counter;
    if (x && r.s->d.kind == AMH_SAMPLER_STATIC && !r.s->d.symmetric && !r.s->by_components()) {
        std::vector<double> xx(d);
        for (int64_t ch = 0; ch < n; ++ch) {
            for (int i = 0; i < d; ++i) xx[i] = r.X[(int64_t)i * n + ch];
            r.lq[ch] = r.s->logq(xx.data());
        }
    }
    return AMH_OK;
}

int32_t amho_run_get_ram_adapt(amh_run* run, double* logalpha, double* eta) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    Run& r = *(Run*)run;
    if (r.s->d.kind != AMH_SAMPLER_RAM) return fail(AMH_ERR_INVALID, "not a RobustAdaptiveMetropolis run");
    if (logalpha) std::memcpy(logalpha, r.logalpha.data(), sizeof(double) * r.logalpha.size());
    if (eta) std::memcpy(eta, r.eta.data(), sizeof(double) * r.eta.size());
    return AMH_OK;
}

int32_t amho_run_set_ram_adapt(amh_run* run, const double* logalpha, const double* eta, const uint8_t* failed) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    Run& r = *(Run*)run;
    if (r.s->d.kind != AMH_SAMPLER_RAM) return fail(AMH_ERR_INVALID, "not a RobustAdaptiveMetropolis run");
    if (logalpha) std::memcpy(r.logalpha.data(), logalpha, sizeof(double) * r.logalpha.size());
    if (eta) std::memcpy(r.eta.data(), eta, sizeof(double) * r.eta.size());
    if (failed) std::memcpy(r.failed.data(), failed, r.failed.size());
    return AMH_OK;
}

int32_t amho_run_ram_failed(amh_run* run, int64_t* nfailed, int64_t* first_chain, uint8_t* failed) {
    if (!run) return fail(AMH_ERR_INVALID, "run is NULL");
    Run& r = *(Run*)run;
    if (r.s->d.kind != AMH_SAMPLER_RAM) return fail(AMH_ERR_INVALID, "not a RobustAdaptiveMetropolis run");
    int64_t nf = 0, first = -1;
    for (int64_t c = 0; c < r.n; ++c)
        if (r.failed[c]) { if (first < 0) first = c; ++nf; }
    if (nfailed) *nfailed = nf;
    if (first_chain) *first_chain = first < 0 ? -1 : r.off + first;
    if (failed) std::memcpy(failed, r.failed.data(), r.failed.size());
    return AMH_OK;
}

int32_t amho_run_set_params(amh_run* run, const double* x) {
    if (!run || !x) return fail(AMH_ERR_INVALID, "NULL argument");
    Run& r = *(Run*)run;
    const int d = r.dim;
    const int64_t n = r.n;
    std::memcpy(r.X.data(), x, sizeof(double) * r.X.size());
    const int kind = r.s->d.kind;
    /* setparams!! recomputes lp (src/AdvancedMH.jl:151-157) and the gradient (MALA.jl:27-35);
     * RAM's setparams!! keeps logprob (RAM :117-121) */
    if (kind == AMH_SAMPLER_RAM) return AMH_OK;
    std::vector<double> xx(d), g(d);
    for (int64_t ch = 0; ch < n; ++ch) {
        for (int i = 0; i < d; ++i) xx[i] = r.X[(int64_t)i * n + ch];
        if (kind == AMH_SAMPLER_MALA) {
            double lp;
            r.t->logp_grad(xx.data(), lp, g.data());
            r.lp[ch] = lp;
            for (int i = 0; i < d; ++i) r.G[(int64_t)i * n + ch] = g[i];
        } else {
            r.lp[ch] = r.t->logp(xx.data());
        }
        if (kind == AMH_SAMPLER_STATIC && !r.s->d.symmetric && !r.s->by_components()) r.lq[ch] = r.s->logq(xx.data());
    }
    return AMH_OK;
}

int32_t amho_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(AMH_ERR_INVALID, "out is NULL");
    *out = bytes ? std::malloc(bytes) : nullptr;
    return AMH_OK;
}
int32_t amho_host_free(void* p) { std::free(p); return AMH_OK; }

int32_t amho_run_dim(amh_run* run) { return run ? ((Run*)run)->dim : -1; }
int64_t amho_run_nchains(amh_run* run) { return run ? ((Run*)run)->n : -1; }
int64_t amho_run_launch_count(amh_run*) { return 0; }
int32_t amho_run_kernel_time_ms(amh_run*, int32_t, double* ms, int64_t* launches) {
    if (ms) *ms = 0;
    if (launches) *launches = 0;
    return AMH_OK;
}

/* ---- raw contract probes for tests (vectorised over n inputs) ---- */
void amho_probe_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                       uint32_t* out4) {
    const amh::Block b = amh::philox4x32_10(c0, c1, c2, c3, k0, k1);
    for (int i = 0; i < 4; ++i) out4[i] = b.v[i];
}
void amho_probe_log(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = amh::log_(x[i]); }
void amho_probe_exp(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = amh::exp_(x[i]); }
void amho_probe_log1pexp(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = amh::log1pexp(x[i]); }
void amho_probe_sigmoid(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = amh::sigmoid(x[i]); }
void amho_probe_family_logpdf(int32_t fam, double p0, double p1, double logc, const double* x, double* y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) y[i] = amh::family_logpdf(fam, p0, p1, logc, x[i]);
}
void amho_probe_u01(const uint64_t* w, double* y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) y[i] = amh::u01((uint32_t)w[i], (uint32_t)(w[i] >> 32));
}
void amho_probe_exponential(const uint64_t* w, double* y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) y[i] = amh::exponential((uint32_t)w[i], (uint32_t)(w[i] >> 32));
}
void amho_probe_normal_pair(const uint64_t* w0, const uint64_t* w1, double* z0, double* z1, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        amh::Block b;
        b.v[0] = (uint32_t)w0[i]; b.v[1] = (uint32_t)(w0[i] >> 32);
        b.v[2] = (uint32_t)w1[i]; b.v[3] = (uint32_t)(w1[i] >> 32);
        amh::normal_pair(b, z0[i], z1[i]);
    }
}
/* the d standard normals and the exponential of step `step` of the chain seeded `seed` */
void amho_probe_step_noise(uint64_t seed, uint64_t step, int32_t d, double* z, double* e) {
    normals(AMH_CONTRACT_V1, seed, step, 0, d, z);
    *e = step_exponential(AMH_CONTRACT_V1, seed, step, d);
}
void amho_probe_step_noise_cv(int32_t cv, uint64_t seed, uint64_t step, int32_t d, double* z, double* e) {
    normals(cv, seed, step, 0, d, z);
    *e = step_exponential(cv, seed, step, d);
}
void amho_probe_philox7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4) {
    const amh::Block b = amh::philox4x32_7(c0, c1, c2, c3, k0, k1);
    for (int i = 0; i < 4; ++i) out4[i] = b.v[i];
}
int32_t amho_run_contract(amh_run* run) { return run ? ((Run*)run)->cv : -1; }
void amho_probe_normal_pair32(const uint32_t* wr, const uint32_t* wa, double* z0, double* z1, int64_t n) {
    for (int64_t i = 0; i < n; ++i) amh::normal_pair32(wr[i], wa[i], z0[i], z1[i]);
}
double amho_probe_target_logp(amh_target* t, const double* x) { return ((Target*)t)->logp(x); }
void amho_probe_target_grad(amh_target* t, const double* x, double* lp, double* g) {
    ((Target*)t)->logp_grad(x, *lp, g);
}

}  /* extern "C" */

/* ---- the multi-device job layer (amh_job_*): the product's sharding logic (advancedmh.jl_b200/csrc/amh_job_impl.h)
 * instantiated over the oracle's own entry points, so that the CPU tests exercise the same code path the GPU job runs:
 * "devices" are just shards here. ---- */
#include "../advancedmh.jl_b200/csrc/amh_job_impl.h"

namespace {
struct OracleBackend {
    struct Shared {};
    static std::string last_error() { return amho_last_error(); }
    static int fail(int code, const std::string& m) { return ::fail(code, m); }
    static int ctx_create(int dev, amh_ctx** out) { return amho_ctx_create(dev, out); }
    static int ctx_destroy(amh_ctx* c) { return amho_ctx_destroy(c); }
    static int target_destroy(amh_target* t) { return amho_target_destroy(t); }
    static int target_create_source(amh_ctx* c, int32_t dim, const char* src, int32_t g, const double* data, int64_t nd, amh_target** out) {
        return amho_target_create_source(c, dim, src, g, data, nd, out);
    }
    static int sampler_create(amh_ctx* c, const amh_sampler_desc* d, amh_sampler** out) { return amho_sampler_create(c, d, out); }
    static int sampler_destroy(amh_sampler* s) { return amho_sampler_destroy(s); }
    static int run_create(amh_ctx* c, amh_target* t, amh_sampler* s, int64_t n, int64_t off, const uint64_t* seeds, const double* init,
                          int64_t ld, amh_run** out) { return amho_run_create(c, t, s, n, off, seeds, init, ld, out); }
    static int run_destroy(amh_run* r) { return amho_run_destroy(r); }
    static int run_steps(amh_run* r, int64_t n, int32_t w, int32_t spl) { return amho_run_steps(r, n, w, spl); }
    static int run_sync(amh_run* r) { return amho_run_sync(r); }
    static int run_sample_ld(amh_run* r, int64_t N, int64_t di, int64_t th, int64_t nw, double* out, int64_t old, uint8_t* acc, int64_t ald,
                             amh_summary* s) { return amho_run_sample_ld(r, N, di, th, nw, out, old, acc, ald, s); }
    static int run_get_state_ld(amh_run* r, int64_t ld, double* x, double* lp, double* g, double* S, uint8_t* a, int64_t* na, int64_t* st) {
        return amho_run_get_state_ld(r, ld, x, lp, g, S, a, na, st);
    }
    static int run_set_state_ld(amh_run* r, int64_t ld, const double* x, const double* lp, const double* g, const double* S, const uint8_t* a,
                                const int64_t* na, int64_t st) { return amho_run_set_state_ld(r, ld, x, lp, g, S, a, na, st); }
    static int run_get_ram_adapt(amh_run* r, double* la, double* eta) { return amho_run_get_ram_adapt(r, la, eta); }
    static int run_set_ram_adapt(amh_run* r, const double* la, const double* eta, const uint8_t* f) { return amho_run_set_ram_adapt(r, la, eta, f); }
    static int run_ram_failed(amh_run* r, int64_t* nf, int64_t* first, uint8_t* f) { return amho_run_ram_failed(r, nf, first, f); }
    static int64_t run_launch_count(amh_run* r) { return amho_run_launch_count(r); }
    static int run_contract(amh_run* r) { return amho_run_contract(r); }
    static int run_kernel_time_ms(amh_run* r, int32_t reset, double* ms, int64_t* l) { return amho_run_kernel_time_ms(r, reset, ms, l); }
    static int shared_init(amhjob::Job<OracleBackend>&) { return AMH_OK; }
    static void shared_destroy(amhjob::Job<OracleBackend>&) {}
    static int target_broadcast(amhjob::Job<OracleBackend>& j, int32_t kind, int32_t dim, const double* blob, int64_t nblob) {
        j.bcast_mode = "copy";
        return j.each([&](int k) { return (int)amho_target_create(j.ctx[k], kind, dim, blob, nblob, &j.target[k]); });
    }
};
}  // namespace

AMH_DEFINE_JOB_ABI(amho_, OracleBackend)
extern "C" double amho_job_comm_init_ms(amh_job*) { return 0.0; }
