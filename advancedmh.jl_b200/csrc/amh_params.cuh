/* amh_params.cuh -- host-side packing of sampler / target descriptions into the
 * kernel parameter structs (constant-bank arrays when the dimension bucket is a
 * compile-time constant, device pointers on the generic path). */
#pragma once
#include <cstring>
#include "amh_host.h"
#include "amh_kernels.cuh"

namespace amhh {

template <int DMAX>
amhd::PropP<DMAX> make_prop(const amh_sampler& s) {
    amhd::PropP<DMAX> p;
    std::memset(&p, 0, sizeof(p));
    p.cov_kind = s.d.cov_kind;
    p.has_mean = s.has_mean ? 1 : 0;
    if constexpr (DMAX > 0) {
        for (size_t i = 0; i < s.mean.size() && i < (size_t)DMAX; ++i) p.mean.v[i] = s.mean[i];
        for (size_t i = 0; i < s.scale.size() && i < (size_t)(DMAX * (DMAX + 1) / 2); ++i) p.scale.v[i] = s.scale[i];
    } else {
        p.mean.p = s.dmean;
        p.scale.p = s.dscale;
    }
    return p;
}

template <class T, int DMAX>
typename T::template Params<DMAX> make_tp(const amh_target& t) {
    typename T::template Params<DMAX> p;
    std::memset(&p, 0, sizeof(p));
    const int d = t.dim;
    if constexpr (T::kind == AMH_TARGET_MVNORMAL) {
        p.c0 = t.blob[0];
        if constexpr (DMAX > 0) {
            for (int i = 0; i < d; ++i) p.mu.v[i] = t.blob[1 + i];
            for (int i = 0; i < d * (d + 1) / 2; ++i) p.U.v[i] = t.blob[1 + d + i];
        } else {
            p.mu.p = t.dblob + 1;
            p.U.p = t.dblob + 1 + d;
        }
    } else if constexpr (T::kind == AMH_TARGET_GAUSS_PREC) {
        if constexpr (DMAX > 0) {
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) p.A.v[i * DMAX + j] = t.blob[(size_t)i * d + j];
        } else {
            p.A.p = t.dblob;
        }
    } else if constexpr (T::kind == AMH_TARGET_ROSENBROCK) {
        p.a = t.blob[0]; p.b = t.blob[1]; p.s = t.blob[2];
    } else if constexpr (T::kind == AMH_TARGET_IID_NORMAL) {
        p.y = t.dblob; p.n = t.ndata;
    } else if constexpr (T::kind == AMH_TARGET_NIG_TOY) {
        p.alpha = t.blob[0]; p.beta = t.blob[1]; p.cig = t.blob[2];
        p.y = t.dblob + 3; p.n = t.ndata;
        p.logspace = (t.kind == AMH_TARGET_NIG_TOY_LOG) ? 1 : 0;
    } else if constexpr (T::kind == AMH_TARGET_LOGISTIC) {
        p.X = t.dblob + 1;
        p.y = t.dblob + 1 + t.ndata * d;
        p.n = t.ndata;
        p.inv2tau2 = t.inv2tau2; p.invtau2 = t.invtau2;
    } else if constexpr (T::kind == AMH_TARGET_USER) {
        p.data = t.dblob; p.n = t.ndata;
    }
    return p;
}

}  // namespace amhh
