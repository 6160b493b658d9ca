// beta / gamma as the GDN kernels see them: either the effective values, or the raw parameters of
// NonNegativeParametrizer (effective = max(p, bound)^2 - pedestal), re-parametrised on the fly while they are staged.
// Fusing the re-parametrisation saves two launches per GDN call each way (kernel K4).
#pragma once
#include "hd_math.cuh"

namespace mmnc {

struct GdnParams {
    const float *beta;
    const float *gamma;
    float beta_bound, gamma_bound, pedestal;
    int raw;

    __device__ __forceinline__ float b(int i) const {
        float v = beta[i];
        if (raw) { v = max_nan(v, beta_bound); v = v * v - pedestal; }
        return v;
    }
    __device__ __forceinline__ float g(int64_t idx) const {
        float v = gamma[idx];
        if (raw) { v = max_nan(v, gamma_bound); v = v * v - pedestal; }
        return v;
    }
    // the same two maps applied to a value that has already been loaded (staging loops issue all their loads first)
    __device__ __forceinline__ float b_of(float v) const {
        if (raw) { v = max_nan(v, beta_bound); v = v * v - pedestal; }
        return v;
    }
    __device__ __forceinline__ float g_of(float v) const {
        if (raw) { v = max_nan(v, gamma_bound); v = v * v - pedestal; }
        return v;
    }
    // chain rule back to the raw parameter, with LowerBound's custom gradient
    __device__ __forceinline__ float db(int i, float d_eff) const {
        if (!raw) return d_eff;
        const float p = beta[i];
        return lower_bound_grad(p, beta_bound, 2.f * max_nan(p, beta_bound) * d_eff);
    }
    __device__ __forceinline__ float dg(int64_t idx, float d_eff) const {
        if (!raw) return d_eff;
        const float p = gamma[idx];
        return lower_bound_grad(p, gamma_bound, 2.f * max_nan(p, gamma_bound) * d_eff);
    }
};

inline GdnParams gdn_effective(const float *beta, const float *gamma) {
    GdnParams p;
    p.beta = beta; p.gamma = gamma; p.beta_bound = p.gamma_bound = p.pedestal = 0.f; p.raw = 0;
    return p;
}

}  // namespace mmnc
