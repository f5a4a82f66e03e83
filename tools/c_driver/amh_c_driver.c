/* amh_c_driver.c -- plain-C caller of the C ABI (include/amh.h): the calls a Julia `ccall` shim makes, without
 * Python.  Used for ncu captures (no interpreter start-up under the profiler) and as the C example of
 * INTEGRATION.md.  Workload: BASELINE config 2 (RWMH, d-dim full-covariance MvNormal, n chains).
 *
 *   amh_c_driver [d=32] [nchains=65536] [launches=5] [mcmc_steps_per_launch=20] [warmup_launches=3] [events=1]
 *   events=0 leaves the library's CUDA-event timing off and reports host wall time around the synchronised loop
 *   [sampler=rw|ram|ramwarm]  ram = RobustAdaptiveMetropolis sampling steps, ramwarm = warm-up (adapting) steps
 *
 * Build: gcc -O2 -o amh_c_driver amh_c_driver.c -I../../include -L../../advancedmh.jl_b200 -lamh_b200 -lm \
 *            -Wl,-rpath,'$ORIGIN/../../advancedmh.jl_b200'
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "amh.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != AMH_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, amh_last_error()); return 1; } } while (0)

static unsigned long long s_rng = 88172645463325252ull;
static double urand(void) { s_rng ^= s_rng << 13; s_rng ^= s_rng >> 7; s_rng ^= s_rng << 17; return (double)(s_rng >> 11) / 9007199254740992.0; }
static double nrand(void) { return sqrt(-2.0 * log(urand() + 1e-300)) * cos(6.283185307179586 * urand()); }

/* lower Cholesky factor of a (row-major d x d), in place into l */
static int cholesky(const double* a, double* l, int d) {
    memset(l, 0, sizeof(double) * d * d);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = a[i * d + j];
            for (int k = 0; k < j; ++k) s -= l[i * d + k] * l[j * d + k];
            if (i == j) { if (s <= 0) return 1; l[i * d + i] = sqrt(s); }
            else l[i * d + j] = s / l[j * d + j];
        }
    return 0;
}

int main(int argc, char** argv) {
    const int d = argc > 1 ? atoi(argv[1]) : 32;
    const long long n = argc > 2 ? atoll(argv[2]) : 65536;
    const int launches = argc > 3 ? atoi(argv[3]) : 5;
    const int spl = argc > 4 ? atoi(argv[4]) : 20;
    const int warm = argc > 5 ? atoi(argv[5]) : 3;
    const int events = argc > 6 ? atoi(argv[6]) : 1;
    const char* smp = argc > 7 ? argv[7] : "rw";
    const int is_ram = strncmp(smp, "ram", 3) == 0, ram_warm = strcmp(smp, "ramwarm") == 0;
    /* Sigma = G G' / d + diag(1..) : some SPD matrix with a spread spectrum */
    double* G = malloc(sizeof(double) * d * d), *S = malloc(sizeof(double) * d * d), *C = malloc(sizeof(double) * d * d);
    double* Ci = malloc(sizeof(double) * d * d), *L = malloc(sizeof(double) * d * d);
    for (int i = 0; i < d * d; ++i) G[i] = nrand();
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double s = 0;
            for (int k = 0; k < d; ++k) s += G[i * d + k] * G[j * d + k];
            S[i * d + j] = s / d * 10.0 + (i == j ? 1.0 + i : 0.0);
        }
    if (cholesky(S, C, d)) { fprintf(stderr, "not SPD\n"); return 1; }
    /* U = inv(C) (lower), logdet */
    memset(Ci, 0, sizeof(double) * d * d);
    double logdet = 0;
    for (int c = 0; c < d; ++c) {
        for (int i = c; i < d; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s -= C[i * d + k] * Ci[k * d + c];
            Ci[i * d + c] = s / C[i * d + i];
        }
        logdet += 2.0 * log(C[c * d + c]);
    }
    const int nt = d * (d + 1) / 2;
    double* blob = malloc(sizeof(double) * (1 + d + nt));
    blob[0] = -0.5 * (d * log(2 * 3.141592653589793) + logdet);
    for (int i = 0; i < d; ++i) blob[1 + i] = 0.0;
    for (int i = 0, q = 0; i < d; ++i) for (int j = 0; j <= i; ++j) blob[1 + d + q++] = Ci[i * d + j];
    /* proposal: MvNormal(0, 2.38^2/d Sigma) -> factor sqrt(2.38^2/d) C, packed by rows */
    double* scale = malloc(sizeof(double) * nt);
    const double f = 2.38 / sqrt((double)d);
    for (int i = 0, q = 0; i < d; ++i) for (int j = 0; j <= i; ++j) scale[q++] = f * C[i * d + j];
    (void)L;

    amh_ctx* ctx; amh_target* tg; amh_sampler* sp; amh_run* run;
    CHECK(amh_ctx_create(0, &ctx));
    CHECK(amh_target_create(ctx, AMH_TARGET_MVNORMAL, d, blob, 1 + d + nt, &tg));
    amh_sampler_desc desc; memset(&desc, 0, sizeof(desc));
    desc.kind = AMH_SAMPLER_RW; desc.dim = d; desc.symmetric = 0; desc.cov_kind = AMH_COV_FULL; desc.scale = scale;
    if (is_ram) {
        desc.kind = AMH_SAMPLER_RAM; desc.scale = NULL; desc.ram_alpha = 0.234; desc.ram_gamma = 0.6;
        desc.ram_eig_lo = 0.0; desc.ram_eig_hi = INFINITY;
    }
    CHECK(amh_sampler_create(ctx, &desc, &sp));
    unsigned long long* seeds = malloc(sizeof(unsigned long long) * n);
    for (long long i = 0; i < n; ++i) { urand(); seeds[i] = s_rng; }
    CHECK(amh_run_create(ctx, tg, sp, n, 0, (const uint64_t*)seeds, NULL, 0, &run));
    for (int i = 0; i < warm; ++i) CHECK(amh_run_steps(run, spl, ram_warm, spl));
    CHECK(amh_run_sync(run));
    double ms = 0; int64_t nl = launches;
    if (events) CHECK(amh_run_kernel_time_ms(run, 1, &ms, &nl));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < launches; ++i) CHECK(amh_run_steps(run, spl, ram_warm, spl));
    CHECK(amh_run_sync(run));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double wall_ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
    if (events) CHECK(amh_run_kernel_time_ms(run, 1, &ms, &nl));
    else { ms = wall_ms; nl = launches; }
    printf("wall %.3f ms for the synchronised loop (%s)\n", wall_ms, events ? "CUDA-event timing on" : "CUDA-event timing off");
    int64_t* nacc = malloc(sizeof(int64_t) * n); int64_t steps;
    CHECK(amh_run_get_state(run, NULL, NULL, NULL, NULL, NULL, nacc, &steps));
    double acc = 0; for (long long i = 0; i < n; ++i) acc += (double)nacc[i];
    printf("d=%d chains=%lld launches=%lld steps/launch=%d  %.3f ms  %.4g chain-steps/s  %.2f us/step  accept=%.3f\n",
           d, n, (long long)nl, spl, ms, (double)n * spl * launches / (ms * 1e-3), 1e3 * ms / (spl * launches), acc / ((double)n * steps));
    amh_run_destroy(run); amh_sampler_destroy(sp); amh_target_destroy(tg); amh_ctx_destroy(ctx);
    return 0;
}
