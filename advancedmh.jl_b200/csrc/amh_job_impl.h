/* amh_job_impl.h -- the multi-GPU data plane behind the C ABI (include/amh.h, amh_job_*): ONE process, N device
 * contexts, one host worker thread per device.  This is what `MCMCThreads()` / `MCMCDistributed()` are to the
 * reference (src/AdvancedMH.jl:30; README.md:135-148; test/runtests.jl:96-110): the caller passes the job-wide
 * arrays -- seeds, initial parameters, the [N][dim+1][nchains] output -- and the library shards them.
 *
 *   - chains are split in contiguous blocks over the devices (an emcee ensemble is indivisible), chain identity
 *     (seed, index) is global, so results do not depend on the number of devices;
 *   - the target's fixed data goes host -> device 0 once and is BROADCAST to the other devices (backend hook:
 *     ncclBroadcast over NVLink, peer copies, or N host copies);
 *   - there is no per-step collective; every device steps its block on its own stream;
 *   - samples leave each device straight into its column block of the caller's array (amh_run_sample_ld) -- no
 *     gather, no staging copy -- and the summaries are pooled on the host (KBs).
 *
 * The logic is generic over a Backend (the per-device entry points), so that the test oracle can instantiate the
 * same code under its own symbol prefix and the sharding logic is testable without a GPU. */
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/amh.h"

namespace amhjob {

inline void shard_bounds(long long nunits, int k, int world, long long& lo, long long& hi) {
    const long long base = nunits / world, rem = nunits % world;
    lo = (long long)k * base + std::min<long long>(k, rem);
    hi = lo + base + (k < rem ? 1 : 0);
}

/* one persistent host thread per device: a phase of a job call (create, sample, read state ...) is posted to all of
 * them at once and joined; no thread is created per call (8 thread creations per phase were ~1 ms of a 10 ms call).
 *
 * Waiting is spin-then-sleep on both sides: a worker that has finished a phase polls for the next one for `spin_us`
 * (AMH_JOB_SPIN_US, default 4000; 0 = sleep at once) before it sleeps on the condition variable, and the caller polls
 * for the join (the devices have equal shares and finish within microseconds of each other).  The phases of one
 * `sample` call, and the calls of a loop, follow each other within a millisecond, so inside a burst no thread is ever
 * woken: 6.62 against 6.74-6.9 ms per 2-GPU call with sleeping workers (profiles/r2_job_fanout_2gpu.txt), and a
 * sleeping worker's wake-up can take a scheduler time slice (~4 ms) when every core is busy -- seen when the host
 * program's BLAS threads were still spinning after a matrix product (the same file; what looked like "one device starts
 * 4.7 ms late" in tools/job_check.py was that).  Workers fall asleep a few milliseconds after the last call. */
class Workers {
public:
    explicit Workers(int n) : n_(n) {
        if (const char* ev = std::getenv("AMH_JOB_SPIN_US")) spin_us_ = std::max(0l, std::atol(ev));
        for (int k = 1; k < n; ++k) th_.emplace_back([this, k] { loop(k); });      /* worker 0 is the calling thread */
    }
    ~Workers() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; gen_.fetch_add(1, std::memory_order_release); }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void run(const std::function<void(int)>& f) {
        if (n_ == 1) { f(0); return; }
        {
            std::lock_guard<std::mutex> g(m_);             /* the lock orders the post against a worker that is about to sleep */
            task_ = &f;
            done_.store(0, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        f(0);
        /* join: the other devices finish within microseconds of this one (equal shares), so poll first */
        const auto t0 = std::chrono::steady_clock::now();
        unsigned spins = 0;
        while (done_.load(std::memory_order_acquire) != n_ - 1) {
            relax();
            if ((++spins & 1023u) == 0 &&
                std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() >= spin_us_) {
                std::unique_lock<std::mutex> lk(m_);
                joiner_sleeps_.store(true, std::memory_order_release);
                cvd_.wait(lk, [&] { return done_.load(std::memory_order_acquire) == n_ - 1; });
                joiner_sleeps_.store(false, std::memory_order_release);
            }
        }
        task_ = nullptr;
    }
private:
    static void relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    void loop(int k) {
        unsigned long long seen = 0;
        for (;;) {
            /* poll for the next phase, then sleep */
            const auto t0 = std::chrono::steady_clock::now();
            unsigned spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen) {
                relax();
                if ((++spins & 1023u) == 0 &&
                    std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() >= spin_us_) {
                    std::unique_lock<std::mutex> lk(m_);
                    sleepers_.fetch_add(1, std::memory_order_acq_rel);
                    cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                    sleepers_.fetch_sub(1, std::memory_order_acq_rel);
                    break;
                }
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_) return;
            const std::function<void(int)>* f = task_;
            (*f)(k);
            done_.fetch_add(1, std::memory_order_release);
            /* pass through the mutex: either this comes before the joiner's critical section (then its predicate sees
             * done_) or after it has entered wait() (then the flag is visible here and it is woken) */
            { std::lock_guard<std::mutex> g(m_); }
            if (joiner_sleeps_.load(std::memory_order_acquire)) cvd_.notify_one();
        }
    }
    int n_;
    long spin_us_ = 4000;
    std::atomic<int> done_{0};
    std::atomic<int> sleepers_{0};
    std::atomic<unsigned long long> gen_{0};
    std::atomic<bool> stop_{false};
    const std::function<void(int)>* task_ = nullptr;
    std::atomic<bool> joiner_sleeps_{false};
    std::mutex m_;
    std::condition_variable cv_, cvd_;
    std::vector<std::thread> th_;
};

template <class B>
struct Job {
    int ngpus = 0;
    std::vector<int> devices;
    std::vector<amh_ctx*> ctx;
    std::vector<amh_target*> target;
    std::vector<amh_sampler*> sampler;
    std::vector<amh_run*> run;
    std::vector<long long> lo, hi;          /* chain (walker) range of every device: [lo, hi) */
    long long nchains = 0, nw = 1;
    int dim = 0, sampler_kind = 0;
    typename B::Shared shared;              /* backend state: communicators, peer-access flags */
    double bcast_ms = 0;
    std::string bcast_mode = "none";
    std::unique_ptr<Workers> workers;

    /* f(k) -> status on one worker thread per device; the first failure wins and its message becomes the caller's */
    template <class F>
    int each(F f, bool only_with_run = false) {
        std::vector<int> rc(ngpus, AMH_OK);
        std::vector<std::string> msg(ngpus);
        static const bool trace = std::getenv("AMH_TRACE") != nullptr;        /* absolute times of the fan-out, per worker */
        const auto t0 = std::chrono::steady_clock::now();
        auto ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
        std::vector<double> ta(ngpus, 0.0), tb(ngpus, 0.0);
        const std::function<void(int)> body = [&](int k) {
            if (only_with_run && !run[k]) return;
            ta[k] = ms();
            rc[k] = f(k);
            if (rc[k] != AMH_OK) msg[k] = B::last_error();
            tb[k] = ms();
        };
        if (!workers) workers.reset(new Workers(ngpus));
        workers->run(body);
        if (trace) {
            std::string line = "[amh] job fan-out: joined at " + std::to_string(ms()) + " ms; worker start/end";
            for (int k = 0; k < ngpus; ++k) line += "  [" + std::to_string(k) + "] " + std::to_string(ta[k]) + " / " + std::to_string(tb[k]);
            std::fprintf(stderr, "%s\n", line.c_str());
        }
        for (int k = 0; k < ngpus; ++k)
            if (rc[k] != AMH_OK) return B::fail(rc[k], "device " + std::to_string(devices[k]) + ": " + msg[k]);
        return AMH_OK;
    }

    void drop_run() {
        for (int k = 0; k < ngpus; ++k)
            if (run[k]) { B::run_destroy(run[k]); run[k] = nullptr; }
        nchains = 0;
    }
    void drop_target() {
        for (int k = 0; k < ngpus; ++k)
            if (target[k]) { B::target_destroy(target[k]); target[k] = nullptr; }
    }
    void drop_sampler() {
        for (int k = 0; k < ngpus; ++k)
            if (sampler[k]) { B::sampler_destroy(sampler[k]); sampler[k] = nullptr; }
    }
};

template <class B>
int job_create(int32_t ngpus, const int32_t* devices, Job<B>** out) {
    if (!out) return B::fail(AMH_ERR_INVALID, "out is NULL");
    if (ngpus < 1) return B::fail(AMH_ERR_INVALID, "ngpus must be >= 1");
    Job<B>* j = new Job<B>();
    j->ngpus = ngpus;
    for (int k = 0; k < ngpus; ++k) j->devices.push_back(devices ? devices[k] : k);
    j->ctx.assign(ngpus, nullptr);
    j->target.assign(ngpus, nullptr);
    j->sampler.assign(ngpus, nullptr);
    j->run.assign(ngpus, nullptr);
    j->lo.assign(ngpus, 0);
    j->hi.assign(ngpus, 0);
    for (int k = 0; k < ngpus; ++k) {
        const int rc = B::ctx_create(j->devices[k], &j->ctx[k]);
        if (rc) {
            const std::string m = B::last_error();
            for (int q = 0; q < k; ++q) B::ctx_destroy(j->ctx[q]);
            delete j;
            return B::fail(rc, m);
        }
    }
    const int rc = B::shared_init(*j);
    if (rc) {
        const std::string m = B::last_error();
        for (int k = 0; k < ngpus; ++k) B::ctx_destroy(j->ctx[k]);
        delete j;
        return B::fail(rc, m);
    }
    *out = j;
    return AMH_OK;
}

template <class B>
int job_destroy(Job<B>* j) {
    if (!j) return AMH_OK;
    j->drop_run();
    j->drop_sampler();
    j->drop_target();
    B::shared_destroy(*j);
    for (int k = 0; k < j->ngpus; ++k) B::ctx_destroy(j->ctx[k]);
    delete j;
    return AMH_OK;
}

template <class B>
int job_target_create(Job<B>* j, int32_t kind, int32_t dim, const double* blob, int64_t nblob) {
    if (!j) return B::fail(AMH_ERR_INVALID, "job is NULL");
    j->drop_run();
    j->drop_target();
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = B::target_broadcast(*j, kind, dim, blob, nblob);      /* fills j->target[*], j->bcast_mode */
    j->bcast_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc) { const std::string m = B::last_error(); j->drop_target(); return B::fail(rc, m); }
    j->dim = dim;
    return AMH_OK;
}

template <class B>
int job_target_create_source(Job<B>* j, int32_t dim, const char* source, int32_t has_gradient, const double* data, int64_t ndata) {
    if (!j) return B::fail(AMH_ERR_INVALID, "job is NULL");
    j->drop_run();
    j->drop_target();
    /* every device compiles the text for itself, in parallel (the run-time compiler is thread safe) */
    const int rc = j->each([&](int k) { return B::target_create_source(j->ctx[k], dim, source, has_gradient, data, ndata, &j->target[k]); });
    if (rc) { const std::string m = B::last_error(); j->drop_target(); return B::fail(rc, m); }
    j->dim = dim;
    j->bcast_mode = "source";
    return AMH_OK;
}

template <class B>
int job_sampler_create(Job<B>* j, const amh_sampler_desc* desc) {
    if (!j || !desc) return B::fail(AMH_ERR_INVALID, "job/desc is NULL");
    j->drop_run();
    j->drop_sampler();
    const int rc = j->each([&](int k) { return B::sampler_create(j->ctx[k], desc, &j->sampler[k]); });
    if (rc) { const std::string m = B::last_error(); j->drop_sampler(); return B::fail(rc, m); }
    j->sampler_kind = desc->kind;
    j->nw = desc->kind == AMH_SAMPLER_STRETCH ? std::max<long long>(1, desc->n_walkers) : 1;
    return AMH_OK;
}

template <class B>
int job_run_create(Job<B>* j, int64_t nchains, const uint64_t* seeds, const double* init, int64_t init_ld) {
    if (!j || !seeds) return B::fail(AMH_ERR_INVALID, "job/seeds is NULL");
    if (!j->target[0] || !j->sampler[0]) return B::fail(AMH_ERR_STATE, "create the job's target and sampler first");
    if (nchains < 1) return B::fail(AMH_ERR_INVALID, "nchains must be >= 1");
    if (nchains % j->nw) return B::fail(AMH_ERR_INVALID, "nchains must be a multiple of n_walkers");
    if (init_ld == 0) init_ld = nchains;
    if (init && init_ld < nchains) return B::fail(AMH_ERR_INVALID, "init_ld must be >= nchains");
    j->drop_run();
    const long long nunits = nchains / j->nw;
    for (int k = 0; k < j->ngpus; ++k) {
        long long a, b;
        shard_bounds(nunits, k, j->ngpus, a, b);
        j->lo[k] = a * j->nw;
        j->hi[k] = b * j->nw;
    }
    const int rc = j->each([&](int k) {
        const long long n = j->hi[k] - j->lo[k];
        if (n == 0) return (int)AMH_OK;                    /* more devices than chains: this one idles */
        return (int)B::run_create(j->ctx[k], j->target[k], j->sampler[k], n, j->lo[k], seeds + j->lo[k] / j->nw,
                                  init ? init + j->lo[k] : nullptr, init_ld, &j->run[k]);
    });
    if (rc) { const std::string m = B::last_error(); j->drop_run(); return B::fail(rc, m); }
    j->nchains = nchains;
    return AMH_OK;
}

template <class B>
int job_need_run(Job<B>* j) {
    if (!j) return B::fail(AMH_ERR_INVALID, "job is NULL");
    if (j->nchains == 0) return B::fail(AMH_ERR_STATE, "the job has no run (amh_job_run_create)");
    return AMH_OK;
}

template <class B>
int job_run_steps(Job<B>* j, int64_t nsteps, int32_t warmup, int32_t spl) {
    if (int rc = job_need_run(j)) return rc;
    /* asynchronous on every device's stream: one thread can enqueue for all of them */
    for (int k = 0; k < j->ngpus; ++k)
        if (j->run[k])
            if (int rc = B::run_steps(j->run[k], nsteps, warmup, spl)) return rc;
    return AMH_OK;
}

template <class B>
int job_run_sync(Job<B>* j) {
    if (int rc = job_need_run(j)) return rc;
    for (int k = 0; k < j->ngpus; ++k)
        if (j->run[k])
            if (int rc = B::run_sync(j->run[k])) return rc;
    return AMH_OK;
}

template <class B>
int job_run_sample(Job<B>* j, int64_t N, int64_t discard_initial, int64_t thinning, int64_t num_warmup, double* out,
                   uint8_t* accepted_out, amh_summary* summary) {
    if (int rc = job_need_run(j)) return rc;
    const int d = j->dim;
    const long long nall = j->nchains;
    std::vector<amh_summary> sm(j->ngpus);
    std::vector<std::vector<double>> mean(j->ngpus), var(j->ngpus), cmean(j->ngpus);
    if (summary)
        for (int k = 0; k < j->ngpus; ++k) {
            std::memset(&sm[k], 0, sizeof(amh_summary));
            mean[k].assign(d, 0.0); var[k].assign(d, 0.0);
            sm[k].mean = mean[k].data(); sm[k].var = var[k].data();
            if (summary->chain_mean) { cmean[k].assign((size_t)d * (j->hi[k] - j->lo[k]), 0.0); sm[k].chain_mean = cmean[k].data(); }
        }
    /* every device writes its COLUMN BLOCK of the caller's [N][dim+1][nchains] array in place */
    const int rc = j->each([&](int k) {
        return (int)B::run_sample_ld(j->run[k], N, discard_initial, thinning, num_warmup, out ? out + j->lo[k] : nullptr, nall,
                                     accepted_out ? accepted_out + j->lo[k] : nullptr, nall, summary ? &sm[k] : nullptr);
    }, true);
    if (rc) return rc;
    if (summary) {
        /* pooled moments of the union: weights = chains per device (every chain holds n_saved samples) */
        double wsum = 0, acc = 0;
        summary->n_saved = 0; summary->n_steps = 0;
        for (int k = 0; k < j->ngpus; ++k) {
            if (!j->run[k]) continue;
            const double w = (double)(j->hi[k] - j->lo[k]);
            wsum += w;
            acc += w * sm[k].accept_rate;
            summary->n_saved = sm[k].n_saved; summary->n_steps = sm[k].n_steps;
        }
        summary->accept_rate = acc / wsum;
        for (int i = 0; i < d; ++i) {
            double m = 0;
            for (int k = 0; k < j->ngpus; ++k)
                if (j->run[k]) m += (double)(j->hi[k] - j->lo[k]) * mean[k][i];
            m /= wsum;
            double v = 0;
            for (int k = 0; k < j->ngpus; ++k)
                if (j->run[k]) {
                    const double dl = mean[k][i] - m;
                    v += (double)(j->hi[k] - j->lo[k]) * (var[k][i] + dl * dl);
                }
            if (summary->mean) summary->mean[i] = m;
            if (summary->var) summary->var[i] = v / wsum;
            if (summary->chain_mean)
                for (int k = 0; k < j->ngpus; ++k) {
                    const long long n = j->hi[k] - j->lo[k];
                    if (n) std::memcpy(summary->chain_mean + (size_t)i * nall + j->lo[k], cmean[k].data() + (size_t)i * n, sizeof(double) * n);
                }
        }
    }
    return AMH_OK;
}

template <class B>
int job_run_get_state(Job<B>* j, double* x, double* lp, double* grad, double* S, uint8_t* accepted, int64_t* naccept,
                      int64_t* step_counter) {
    if (int rc = job_need_run(j)) return rc;
    std::vector<int64_t> steps(j->ngpus, 0);
    const int rc = j->each([&](int k) {
        const long long o = j->lo[k];
        return (int)B::run_get_state_ld(j->run[k], j->nchains, x ? x + o : nullptr, lp ? lp + o : nullptr, grad ? grad + o : nullptr,
                                        S ? S + o : nullptr, accepted ? accepted + o : nullptr, naccept ? naccept + o : nullptr, &steps[k]);
    }, true);
    if (rc) return rc;
    if (step_counter)
        for (int k = 0; k < j->ngpus; ++k)
            if (j->run[k]) { *step_counter = steps[k]; break; }
    return AMH_OK;
}

template <class B>
int job_run_set_state(Job<B>* j, const double* x, const double* lp, const double* grad, const double* S, const uint8_t* accepted,
                      const int64_t* naccept, int64_t step_counter) {
    if (int rc = job_need_run(j)) return rc;
    return j->each([&](int k) {
        const long long o = j->lo[k];
        return (int)B::run_set_state_ld(j->run[k], j->nchains, x ? x + o : nullptr, lp ? lp + o : nullptr, grad ? grad + o : nullptr,
                                        S ? S + o : nullptr, accepted ? accepted + o : nullptr, naccept ? naccept + o : nullptr, step_counter);
    }, true);
}

template <class B>
int job_run_get_ram_adapt(Job<B>* j, double* logalpha, double* eta) {
    if (int rc = job_need_run(j)) return rc;
    return j->each([&](int k) {
        return (int)B::run_get_ram_adapt(j->run[k], logalpha ? logalpha + j->lo[k] : nullptr, eta ? eta + j->lo[k] : nullptr);
    }, true);
}

template <class B>
int job_run_set_ram_adapt(Job<B>* j, const double* logalpha, const double* eta, const uint8_t* failed) {
    if (int rc = job_need_run(j)) return rc;
    return j->each([&](int k) {
        const long long o = j->lo[k];
        return (int)B::run_set_ram_adapt(j->run[k], logalpha ? logalpha + o : nullptr, eta ? eta + o : nullptr, failed ? failed + o : nullptr);
    }, true);
}

template <class B>
int job_run_ram_failed(Job<B>* j, int64_t* nfailed, int64_t* first_chain, uint8_t* failed) {
    if (int rc = job_need_run(j)) return rc;
    std::vector<int64_t> nf(j->ngpus, 0), first(j->ngpus, -1);
    const int rc = j->each([&](int k) {
        return (int)B::run_ram_failed(j->run[k], &nf[k], &first[k], failed ? failed + j->lo[k] : nullptr);
    }, true);
    if (rc) return rc;
    int64_t tot = 0, f = -1;
    for (int k = 0; k < j->ngpus; ++k) {
        tot += nf[k];
        if (f < 0 && first[k] >= 0) f = first[k];          /* run-level indices are already global (chain_offset) */
    }
    if (nfailed) *nfailed = tot;
    if (first_chain) *first_chain = f;
    return AMH_OK;
}

template <class B>
int job_run_shard(Job<B>* j, int32_t k, int64_t* lo, int64_t* hi, int32_t* device) {
    if (!j) return B::fail(AMH_ERR_INVALID, "job is NULL");
    if (k < 0 || k >= j->ngpus) return B::fail(AMH_ERR_INVALID, "shard index out of range");
    if (lo) *lo = j->lo[k];
    if (hi) *hi = j->hi[k];
    if (device) *device = j->devices[k];
    return AMH_OK;
}

template <class B>
int64_t job_run_launch_count(Job<B>* j) {
    if (!j) return -1;
    int64_t tot = 0;
    for (int k = 0; k < j->ngpus; ++k)
        if (j->run[k]) tot += B::run_launch_count(j->run[k]);
    return tot;
}

/* device time of the stepping kernels: the MAX over the devices (they run concurrently), launches summed */
template <class B>
int job_run_kernel_time_ms(Job<B>* j, int32_t reset, double* ms, int64_t* launches) {
    if (int rc = job_need_run(j)) return rc;
    double mx = 0;
    int64_t nl = 0;
    for (int k = 0; k < j->ngpus; ++k) {
        if (!j->run[k]) continue;
        double m = 0;
        int64_t l = 0;
        if (int rc = B::run_kernel_time_ms(j->run[k], reset, &m, &l)) return rc;
        mx = std::max(mx, m);
        nl += l;
    }
    if (ms) *ms = mx;
    if (launches) *launches = nl;
    return AMH_OK;
}

}  // namespace amhjob

/* the extern "C" entry points of include/amh.h for backend B under the symbol prefix P (amh_ for the product) */
#define AMH_JOB_CAT_(a, b) a##b
#define AMH_JOB_CAT(a, b) AMH_JOB_CAT_(a, b)
#define AMH_DEFINE_JOB_ABI(P, B)                                                                                                   \
    extern "C" {                                                                                                                   \
    int32_t AMH_JOB_CAT(P, job_create)(int32_t ngpus, const int32_t* devices, amh_job** out) {                                     \
        return amhjob::job_create<B>(ngpus, devices, (amhjob::Job<B>**)out);                                                       \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_destroy)(amh_job* j) { return amhjob::job_destroy<B>((amhjob::Job<B>*)j); }                         \
    int32_t AMH_JOB_CAT(P, job_ngpus)(amh_job* j) { return j ? ((amhjob::Job<B>*)j)->ngpus : -1; }                                 \
    int32_t AMH_JOB_CAT(P, job_target_create)(amh_job* j, int32_t kind, int32_t dim, const double* blob, int64_t nblob) {          \
        return amhjob::job_target_create<B>((amhjob::Job<B>*)j, kind, dim, blob, nblob);                                           \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_target_create_source)(amh_job* j, int32_t dim, const char* source, int32_t has_gradient,            \
                                                     const double* data, int64_t ndata) {                                         \
        return amhjob::job_target_create_source<B>((amhjob::Job<B>*)j, dim, source, has_gradient, data, ndata);                    \
    }                                                                                                                              \
    const char* AMH_JOB_CAT(P, job_broadcast_mode)(amh_job* j) { return j ? ((amhjob::Job<B>*)j)->bcast_mode.c_str() : ""; }       \
    double AMH_JOB_CAT(P, job_broadcast_ms)(amh_job* j) { return j ? ((amhjob::Job<B>*)j)->bcast_ms : -1.0; }                      \
    int32_t AMH_JOB_CAT(P, job_sampler_create)(amh_job* j, const amh_sampler_desc* desc) {                                         \
        return amhjob::job_sampler_create<B>((amhjob::Job<B>*)j, desc);                                                            \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_create)(amh_job* j, int64_t nchains, const uint64_t* seeds, const double* init, int64_t ld) {   \
        return amhjob::job_run_create<B>((amhjob::Job<B>*)j, nchains, seeds, init, ld);                                            \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_destroy)(amh_job* j) { if (j) ((amhjob::Job<B>*)j)->drop_run(); return AMH_OK; }                \
    int32_t AMH_JOB_CAT(P, job_run_steps)(amh_job* j, int64_t nsteps, int32_t warmup, int32_t spl) {                               \
        return amhjob::job_run_steps<B>((amhjob::Job<B>*)j, nsteps, warmup, spl);                                                  \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_sync)(amh_job* j) { return amhjob::job_run_sync<B>((amhjob::Job<B>*)j); }                       \
    int32_t AMH_JOB_CAT(P, job_run_sample)(amh_job* j, int64_t N, int64_t discard_initial, int64_t thinning, int64_t num_warmup,   \
                                           double* out, uint8_t* accepted_out, amh_summary* summary) {                             \
        return amhjob::job_run_sample<B>((amhjob::Job<B>*)j, N, discard_initial, thinning, num_warmup, out, accepted_out, summary); \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_get_state)(amh_job* j, double* x, double* lp, double* grad, double* S, uint8_t* accepted,       \
                                              int64_t* naccept, int64_t* step_counter) {                                          \
        return amhjob::job_run_get_state<B>((amhjob::Job<B>*)j, x, lp, grad, S, accepted, naccept, step_counter);                  \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_set_state)(amh_job* j, const double* x, const double* lp, const double* grad, const double* S,  \
                                              const uint8_t* accepted, const int64_t* naccept, int64_t step_counter) {             \
        return amhjob::job_run_set_state<B>((amhjob::Job<B>*)j, x, lp, grad, S, accepted, naccept, step_counter);                  \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_get_ram_adapt)(amh_job* j, double* logalpha, double* eta) {                                     \
        return amhjob::job_run_get_ram_adapt<B>((amhjob::Job<B>*)j, logalpha, eta);                                                \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_set_ram_adapt)(amh_job* j, const double* logalpha, const double* eta, const uint8_t* failed) {  \
        return amhjob::job_run_set_ram_adapt<B>((amhjob::Job<B>*)j, logalpha, eta, failed);                                        \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_ram_failed)(amh_job* j, int64_t* nfailed, int64_t* first_chain, uint8_t* failed) {              \
        return amhjob::job_run_ram_failed<B>((amhjob::Job<B>*)j, nfailed, first_chain, failed);                                    \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_shard)(amh_job* j, int32_t k, int64_t* lo, int64_t* hi, int32_t* device) {                      \
        return amhjob::job_run_shard<B>((amhjob::Job<B>*)j, k, lo, hi, device);                                                    \
    }                                                                                                                              \
    int32_t AMH_JOB_CAT(P, job_run_contract)(amh_job* j) {                                                                         \
        amhjob::Job<B>* q = (amhjob::Job<B>*)j;                                                                                    \
        if (!q) return -1;                                                                                                         \
        for (int k = 0; k < q->ngpus; ++k)                                                                                         \
            if (q->run[k]) return B::run_contract(q->run[k]);                                                                      \
        return -1;                                                                                                                 \
    }                                                                                                                              \
    int64_t AMH_JOB_CAT(P, job_run_launch_count)(amh_job* j) { return amhjob::job_run_launch_count<B>((amhjob::Job<B>*)j); }       \
    int32_t AMH_JOB_CAT(P, job_run_kernel_time_ms)(amh_job* j, int32_t reset, double* ms, int64_t* launches) {                     \
        return amhjob::job_run_kernel_time_ms<B>((amhjob::Job<B>*)j, reset, ms, launches);                                         \
    }                                                                                                                              \
    }
