/* amh_launch_mh_dims.cu -- more exact-dimension instantiations of the per-thread MH step kernel K1 (mh-core.jl:92-117) for
 * the targets the tensor-core kernels do not cover.
 *
 * amh_launch_mh.cu holds d = 1..6, 8, 10, 12, 16, 20, 24, 32; every other dimension ran the generic kernel (run-time
 * dimension, vectors in local memory): RWMH x GaussianPrecision on 65 536 chains d = 7 -> 6.0e9 chain-steps/s next to
 * 2.6e10 at d = 8, d = 14 -> 2.0e9 (16: 1.0e10), d = 28 -> 6.0e8 (32: 2.6e9); x Rosenbrock 2-3 x
 * (profiles/r2_dim_cliffs.txt).  MvNormal targets of these dimensions run the padded tensor-core kernels
 * (amh_launch_mh_tcp.cu) and are not instantiated here.  A second translation unit so that the halves compile in parallel. */
#define AMH_MH_EXTRA_TU
#include "amh_launch_mh.cu"

namespace amhh {

template <class T>
static int more_dims(amh_run& r, int nsteps, const SaveArgs& sv, bool& taken) {
    taken = true;
    switch (r.dim) {
    case 7: return launch_mh_t<7, T>(r, nsteps, sv);
    case 9: return launch_mh_t<9, T>(r, nsteps, sv);
    case 11: return launch_mh_t<11, T>(r, nsteps, sv);
    case 14: return launch_mh_t<14, T>(r, nsteps, sv);
    case 18: return launch_mh_t<18, T>(r, nsteps, sv);
    case 28: return launch_mh_t<28, T>(r, nsteps, sv);
    }
    taken = false;
    return AMH_OK;
}

/* MvNormal targets: only where the per-thread kernel beats the padded tensor-core one -- d = 9 and 11 would be padded to 16
 * (3 x / 2 x the mat-vec work: 1.5e10 chain-steps/s against 2.2e10 for the per-thread kernel at d = 10), d = 17 ... 19 to 24
 * (8.9e9 against 1.0-1.2e10; with few chains those three stay on the tensor-core kernels' 4-warp CTAs, which spread better) */
static int mvnormal_dims(amh_run& r, int nsteps, const SaveArgs& sv, bool& taken) {
    taken = true;
    switch (r.dim) {
    case 9: return launch_mh_t<9, TMvNormal>(r, nsteps, sv);
    case 11: return launch_mh_t<11, TMvNormal>(r, nsteps, sv);
    case 17: return launch_mh_t<17, TMvNormal>(r, nsteps, sv);
    case 18: return launch_mh_t<18, TMvNormal>(r, nsteps, sv);
    case 19: return launch_mh_t<19, TMvNormal>(r, nsteps, sv);
    }
    taken = false;
    return AMH_OK;
}

int launch_mh_more_dims(amh_run& r, int nsteps, const SaveArgs& sv, bool& taken) {
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return mvnormal_dims(r, nsteps, sv, taken);
    case AMH_TARGET_GAUSS_PREC: return more_dims<TGaussPrec>(r, nsteps, sv, taken);
    case AMH_TARGET_ROSENBROCK: return more_dims<TRosenbrock>(r, nsteps, sv, taken);
    }
    taken = false;
    return AMH_OK;
}

}  // namespace amhh
