/* amh_launch_stretch_dims.cu -- more exact-dimension instantiations of the stretch-move kernels (emcee.jl:39-102).
 *
 * The sweep kernels keep a walker's record in registers, so they are templates on the dimension; a dimension without an
 * instantiation runs the generic (run-time dimension, local-memory) variant, which is 4-8 x slower: Rosenbrock,
 * 64 x 4 096 walkers, d = 10 -> 1.13e10 moves/s, d = 12 (generic) -> 3.0e9, d = 20 (generic, 64 x 2 048) -> 1.36e9
 * (profiles/r2_c3_shape_sweep.txt).  amh_launch_stretch.cu holds d = 2, 3, 4, 5, 8, 10, 16; this translation unit adds
 * the dimensions below (6, 7, 9, 12, 14, 20, 24, 32) for the three catalogue targets that take any dimension.  It is a second translation unit only so
 * that the two halves compile in parallel (the kernels are the same templates: the source file is included). */
#define AMH_STRETCH_EXTRA_TU
#include "amh_launch_stretch.cu"

namespace amhh {

template <class T>
static int more_dims(amh_run& r, int nsteps, const SaveArgs& sv, bool& taken) {
    taken = true;
    switch (r.dim) {
    case 6: return launch_stretch_t<6, T>(r, nsteps, sv);
    case 7: return launch_stretch_t<7, T>(r, nsteps, sv);
    case 9: return launch_stretch_t<9, T>(r, nsteps, sv);
    case 12: return launch_stretch_t<12, T>(r, nsteps, sv);
    case 14: return launch_stretch_t<14, T>(r, nsteps, sv);
    case 20: return launch_stretch_t<20, T>(r, nsteps, sv);
    case 24: return launch_stretch_t<24, T>(r, nsteps, sv);
    case 32: return launch_stretch_t<32, T>(r, nsteps, sv);
    }
    taken = false;
    return AMH_OK;
}

int launch_stretch_more_dims(amh_run& r, int nsteps, const SaveArgs& sv, bool& taken) {
    switch (r.target->kind) {
    case AMH_TARGET_MVNORMAL: return more_dims<TMvNormal>(r, nsteps, sv, taken);
    case AMH_TARGET_GAUSS_PREC: return more_dims<TGaussPrec>(r, nsteps, sv, taken);
    case AMH_TARGET_ROSENBROCK: return more_dims<TRosenbrock>(r, nsteps, sv, taken);
    }
    taken = false;
    return AMH_OK;
}

}  // namespace amhh
