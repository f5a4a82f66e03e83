/* amh_ram_common.cuh -- argument block and small helpers of the warp-per-chain RAM kernel K4W (amh_launch_ram_warp.cu).
 * (A streaming variant, K4S -- factor read from global memory a few columns ahead of the recurrence, double-buffered in HBM,
 * 24 resident chains per SM instead of 12 -- was built and measured in round 2: bit-exact, but 4.3-4.9e7 chain-steps/s
 * against K4W's 7.1e7 on C5, profiles/r2_k4s_stream_ab.txt, so it was removed.  S2 / sflag below are what is left of its
 * interface: NULL for K4W.) */
#pragma once
#include "amh_params.cuh"
#include "amh_fastmath.cuh"

namespace amhh {
using namespace amhd;

struct RamWArgs {
    ChainState st;
    SaveArgs sv;
    int d;
    int nsteps;
    int warmup;
    unsigned long long step0;
    double* S;                 /* [chain][nt] column-packed lower factor */
    double* S2;                /* K4S: second factor buffer, same layout (the non-mutating lowrankupdate, RAM :167,:170) */
    unsigned char* sflag;      /* K4S: per chain, which buffer holds the current factor */
    unsigned char* failed;
    double* logalpha;
    double* eta;
    double alpha, gamma, lo, hi;
    int check;
    const double* Utc;         /* target factor, column-packed [nt] */
    int force_redo;            /* test switch: treat every speculative sweep as out of range (exercises the redo path) */
    const double* mu;          /* [d] */
    double c0;
};

__host__ __device__ __forceinline__ int colstart(int i, int d) { return i * d - (i * (i - 1)) / 2; }

/* out-of-line IEEE operators for the operands the branch-free sequences of amh_fastmath.cuh do not cover */
static __device__ __noinline__ double div_slow(double a, double b) { return a / b; }


/* range bookkeeping of the speculative sweeps: instead of testing every operand against the exponent range of the
 * straight-line sequences, the loops keep a running minimum / maximum of the operands' HIGH WORDS (for doubles of one
 * sign the integer order of the high words is the order of the values; a negative value gives a negative word, NaN and
 * Inf large positive ones) -- two integer min/max per operand, off the fp64 pipe and off the critical path */
constexpr int kHiLo = 0x2B800000;          /* high word of 2^-327 ~ 3.7e-99  */
constexpr int kHiHi = 0x54B00000;          /* high word of 2^332  ~ 8.7e99   */
__device__ __forceinline__ int hi_abs(double x) { return __double2hiint(x) & 0x7fffffff; }


}  // namespace amhh
