/* amh_user_target.h -- what the source text of a user-supplied target may rely on
 * (amh_target_create_source in amh.h).
 *
 * The text stands for the closure of DensityModel(f) (/root/reference/src/AdvancedMH.jl:52-54,74) or for a
 * LogDensityProblems object (src/AdvancedMH.jl:76; logdensity_and_gradient: MALA.jl:100-105).  It is compiled twice
 * from the SAME characters:
 *   - by NVRTC inside libamh_b200.so, for sm_100a, with --fmad=false      (AMH_TARGET = __device__)
 *   - by g++ -ffp-contract=off inside the CPU oracle (test infrastructure) (AMH_TARGET = extern "C")
 * and the two results agree bit for bit as long as the text keeps to the numerical contract (amh_contract.h):
 *   + - * /, sqrt(), fma(), fabs(), comparisons, integer and bit operations: IEEE-754, correctly rounded on both sides;
 *   amh::log_(x), amh::exp_(x), amh::log1pexp(x), amh::sigmoid(x): the contract's own transcendental functions;
 *   NO log()/exp()/pow()/sin()/... of the C library (libm and libdevice round differently), no static state.
 * Every fused multiply-add must be written as fma(); a*b+c is two roundings on both sides.
 *
 * Entry points the text defines (global namespace):
 *
 *   AMH_TARGET double amh_user_logdensity(const double* x, int dim, const double* data, long long ndata);
 *       log-density at x[0..dim-1]; -INFINITY outside the support (the move is rejected, mh-core.jl:108);
 *       `data` is the array handed to amh_target_create_source (read-only, device global memory on the GPU).
 *
 *   AMH_TARGET void amh_user_logdensity_and_gradient(const double* x, int dim, const double* data, long long ndata,
 *                                                    double* lp, double* grad);
 *       only when has_gradient != 0 (needed by MALA): *lp and grad[0..dim-1].
 */
#ifndef AMH_USER_TARGET_H
#define AMH_USER_TARGET_H
#include "amh_contract.h"
#ifndef AMH_TARGET
#if defined(__CUDACC__)
#define AMH_TARGET __device__
#else
#define AMH_TARGET extern "C"
#endif
#endif
#endif /* AMH_USER_TARGET_H */
