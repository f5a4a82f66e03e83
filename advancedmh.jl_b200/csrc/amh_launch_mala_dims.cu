/* amh_launch_mala_dims.cu -- more exact-dimension instantiations of the MALA step kernel (MALA.jl:54-93).
 *
 * amh_launch_mala.cu holds d = 2, 3, 4, 5, 8, 10, 16; every other dimension ran the generic kernel (run-time dimension,
 * vectors in local memory), measured 3-65 x slower on 65 536 chains of a MvNormal target (profiles/r2_dim_cliffs.txt):
 * d = 6 -> 1.1e10 chain-steps/s next to 3.2e10 at d = 5, d = 12 -> 1.8e9 (10: 1.9e10), d = 32 -> 1.7e8 (16: 1.1e10).
 * This translation unit adds the dimensions below for the catalogue targets that take any dimension; it exists only so
 * that the two halves compile in parallel (the kernel is the same template: the source file is included). */
#define AMH_MALA_EXTRA_TU
#include "amh_launch_mala.cu"

namespace amhh {

template <class T>
static int more_dims(amh_run& r, int nsteps, const SaveArgs& sv, bool& taken) {
    taken = true;
    switch (r.dim) {
    case 6: return launch_mala_t<6, T>(r, nsteps, sv);
    case 7: return launch_mala_t<7, T>(r, nsteps, sv);
    case 9: return launch_mala_t<9, T>(r, nsteps, sv);
    case 12: return launch_mala_t<12, T>(r, nsteps, sv);
    case 14: return launch_mala_t<14, T>(r, nsteps, sv);
    case 20: return launch_mala_t<20, T>(r, nsteps, sv);
    case 24: return launch_mala_t<24, T>(r, nsteps, sv);
    case 32: return launch_mala_t<32, T>(r, nsteps, sv);
    }
    taken = false;
    return AMH_OK;
}

int launch_mala_more_dims(amh_run& r, int nsteps, const SaveArgs& sv, bool& taken) {
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return more_dims<TMvNormal>(r, nsteps, sv, taken);
    case AMH_TARGET_GAUSS_PREC: return more_dims<TGaussPrec>(r, nsteps, sv, taken);
    case AMH_TARGET_ROSENBROCK: return more_dims<TRosenbrock>(r, nsteps, sv, taken);
    }
    taken = false;
    return AMH_OK;
}

}  // namespace amhh
