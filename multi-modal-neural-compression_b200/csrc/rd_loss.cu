// K3 epilogue: distortion terms, the multi-task RD-loss epilogue and the GDN re-parametrisation
// (NonNegativeParametrizer of row a8).
//
// Reference formulas: /root/reference/src/models/multi_task_compressor.py:223-293, 302-357, 437;
// mixed_latent.py:70-118; shared_latent.py:118-147; /root/reference/src/loss_balancing.py:31-54.
#include "common.cuh"
#include "hd_math.cuh"

namespace mmnc {

constexpr int D_THREADS = 256;

// ---- distortion: out += scale * sum d(a, b); float4 streaming, grid-stride, one atomic per block
template <int kKind>
__device__ __forceinline__ float dist_term(float a, float b) {
    const float d = a - b;
    return kKind == 0 ? d * d : fabsf(d);
}

template <int kKind>
__global__ void __launch_bounds__(D_THREADS)
distortion_forward_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t n, float scale,
                          float *__restrict__ out) {
    __shared__ float red[32];
    float acc = 0.f;
    const int64_t n4 = n >> 2;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (aligned) {
        const float4 *a4 = reinterpret_cast<const float4 *>(a);
        const float4 *b4 = reinterpret_cast<const float4 *>(b);
        for (int64_t i = tid; i < n4; i += stride) {
            const float4 u = a4[i], v = b4[i];
            acc += dist_term<kKind>(u.x, v.x) + dist_term<kKind>(u.y, v.y) + dist_term<kKind>(u.z, v.z) +
                   dist_term<kKind>(u.w, v.w);
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) acc += dist_term<kKind>(a[i], b[i]);
    } else {
        for (int64_t i = tid; i < n; i += stride) acc += dist_term<kKind>(a[i], b[i]);
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, tot * scale);
}

template <int kKind>
__global__ void __launch_bounds__(D_THREADS)
distortion_backward_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t n, float scale,
                           const float *__restrict__ g_scalar, float *__restrict__ g_a) {
    const float g = g_scalar[0] * scale;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    auto term = [&](float u, float v) { const float d = u - v; return kKind == 0 ? 2.f * g * d : g * sign_t(d); };
    const bool aligned =
        ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(g_a)) & 15) == 0;
    if (aligned) {  // float4 streaming like the forward
        const int64_t n4 = n >> 2;
        const float4 *a4 = reinterpret_cast<const float4 *>(a);
        const float4 *b4 = reinterpret_cast<const float4 *>(b);
        float4 *o4 = reinterpret_cast<float4 *>(g_a);
        for (int64_t i = tid; i < n4; i += stride) {
            const float4 u = a4[i], v = b4[i];
            o4[i] = make_float4(term(u.x, v.x), term(u.y, v.y), term(u.z, v.z), term(u.w, v.w));
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) g_a[i] = term(a[i], b[i]);
    } else {
        for (int64_t i = tid; i < n; i += stride) g_a[i] = term(a[i], b[i]);
    }
}

// ---- RD epilogue: one block.  All inputs are tiny (M, N <= a few hundred; T, groups <= 8).
constexpr int RD_THREADS = 256;
constexpr int RD_MAX_GROUPS = 16;
constexpr float LN2 = 0.69314718055994530942f;

__global__ void __launch_bounds__(RD_THREADS)
rd_epilogue_kernel(const float *__restrict__ lnsum_y, int64_t M, const float *__restrict__ lnsum_z, int64_t N,
                   const int32_t *__restrict__ group_of_channel, int n_groups,
                   const float *__restrict__ group_inv_pixels, const float *__restrict__ group_weight,
                   float z_inv_pixels, float z_weight, const float *__restrict__ task_losses, int T,
                   const float *__restrict__ log_vars, float lmbda, float *__restrict__ scalars,
                   float *__restrict__ g_lnsum_y, float *__restrict__ g_lnsum_z, float *__restrict__ g_task_losses,
                   float *__restrict__ g_log_vars) {
    __shared__ float gsum[RD_MAX_GROUPS + 1];
    __shared__ float red[32];
    // group sums of ln(lik): deterministic order (thread-strided partials, then the block tree)
    for (int g = 0; g <= n_groups; ++g) {
        float acc = 0.f;
        if (g < n_groups) {
            for (int64_t c = threadIdx.x; c < M; c += blockDim.x)
                if (group_of_channel[c] == g) acc += lnsum_y[c];
        } else {
            for (int64_t c = threadIdx.x; c < N; c += blockDim.x) acc += lnsum_z[c];
        }
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) gsum[g] = tot;
        __syncthreads();
    }
    // gradients of the loss w.r.t. the per-channel sums: bpp(G) = sum / (-ln 2) * inv_pixels, weight w_G
    for (int64_t c = threadIdx.x; c < M; c += blockDim.x) {
        const int g = group_of_channel[c];
        g_lnsum_y[c] = (g >= 0 && g < n_groups) ? (-group_weight[g] * group_inv_pixels[g] / LN2) : 0.f;
    }
    for (int64_t c = threadIdx.x; c < N; c += blockDim.x) g_lnsum_z[c] = -z_weight * z_inv_pixels / LN2;
    if (threadIdx.x == 0) {
        float comp = 0.f;
        for (int g = 0; g < n_groups; ++g) {
            const float bpp = gsum[g] / (-LN2) * group_inv_pixels[g];
            scalars[4 + g] = bpp;
            comp += group_weight[g] * bpp;
        }
        const float zbpp = gsum[n_groups] / (-LN2) * z_inv_pixels;
        comp += z_weight * zbpp;
        float rec = 0.f;
        for (int t = 0; t < T; ++t) {
            const float L = task_losses[t];
            float w, gL, gs;
            if (log_vars != nullptr) {
                const float s = log_vars[t], e = expf(-s);
                const float mask = (L != 0.f) ? 1.f : 0.f;
                w = (e * L + s) * mask;
                gL = e * mask;
                gs = (1.f - e * L) * mask;
            } else {
                w = L; gL = 1.f; gs = 0.f;
            }
            scalars[4 + n_groups + t] = w;
            rec += w;
            g_task_losses[t] = lmbda * gL;
            if (g_log_vars != nullptr) g_log_vars[t] = lmbda * gs;
        }
        scalars[0] = lmbda * rec + comp;
        scalars[1] = rec;
        scalars[2] = comp;
        scalars[3] = zbpp;
    }
}

// ---- NonNegativeParametrizer
__global__ void __launch_bounds__(D_THREADS)
nonneg_forward_kernel(const float *__restrict__ p, int64_t n, float bound, float pedestal, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = max_nan(p[i], bound);
        out[i] = v * v - pedestal;
    }
}
__global__ void __launch_bounds__(D_THREADS)
nonneg_backward_kernel(const float *__restrict__ p, const float *__restrict__ g_out, int64_t n, float bound,
                       float *__restrict__ g_p) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = max_nan(p[i], bound);
        const float g = 2.f * v * g_out[i];  // gradient arriving at LowerBound's output
        g_p[i] = lower_bound_grad(p[i], bound, g);
    }
}

static inline unsigned d_blocks(int64_t n, int per_thread) {
    int64_t blocks = (n / per_thread + D_THREADS - 1) / D_THREADS;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace mmnc

using namespace mmnc;

extern "C" int mmnc_distortion_forward(const float *a, const float *b, int64_t n, int kind, float scale, float *out,
                                       void *stream) {
    MMNC_REQUIRE(n >= 0, "distortion_forward: negative size");
    MMNC_REQUIRE(kind == 0 || kind == 1, "distortion_forward: kind must be 0 (mse) or 1 (l1)");
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(a && b && out, "distortion_forward: null pointer");
    if (kind == 0)
        distortion_forward_kernel<0><<<d_blocks(n, 4), D_THREADS, 0, as_stream(stream)>>>(a, b, n, scale, out);
    else
        distortion_forward_kernel<1><<<d_blocks(n, 4), D_THREADS, 0, as_stream(stream)>>>(a, b, n, scale, out);
    return after_launch("distortion_forward_kernel");
}

extern "C" int mmnc_distortion_backward(const float *a, const float *b, int64_t n, int kind, float scale,
                                        const float *g_scalar, float *g_a, void *stream) {
    MMNC_REQUIRE(n >= 0, "distortion_backward: negative size");
    MMNC_REQUIRE(kind == 0 || kind == 1, "distortion_backward: kind must be 0 (mse) or 1 (l1)");
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(a && b && g_scalar && g_a, "distortion_backward: null pointer");
    if (kind == 0)
        distortion_backward_kernel<0><<<d_blocks(n, 4), D_THREADS, 0, as_stream(stream)>>>(a, b, n, scale, g_scalar, g_a);
    else
        distortion_backward_kernel<1><<<d_blocks(n, 4), D_THREADS, 0, as_stream(stream)>>>(a, b, n, scale, g_scalar, g_a);
    return after_launch("distortion_backward_kernel");
}

extern "C" int mmnc_rd_epilogue(const float *lnsum_y, int64_t M, const float *lnsum_z, int64_t N,
                                const int32_t *group_of_channel, int n_groups, const float *group_inv_pixels,
                                const float *group_weight, float z_inv_pixels, float z_weight,
                                const float *task_losses, int T, const float *log_vars, float lmbda, float *scalars,
                                float *g_lnsum_y, float *g_lnsum_z, float *g_task_losses, float *g_log_vars,
                                void *stream) {
    MMNC_REQUIRE(M >= 0 && N >= 0 && T >= 0, "rd_epilogue: negative dimension");
    MMNC_REQUIRE(n_groups >= 0 && n_groups <= RD_MAX_GROUPS, "rd_epilogue: n_groups %d out of range", n_groups);
    MMNC_REQUIRE(scalars && (M == 0 || g_lnsum_y) && (N == 0 || g_lnsum_z) && (T == 0 || g_task_losses),
                 "rd_epilogue: null output");
    MMNC_REQUIRE((M == 0 || (lnsum_y && group_of_channel)) && (N == 0 || lnsum_z) && (T == 0 || task_losses),
                 "rd_epilogue: null input");
    MMNC_REQUIRE(n_groups == 0 || (group_inv_pixels && group_weight), "rd_epilogue: null group tables");
    rd_epilogue_kernel<<<1, RD_THREADS, 0, as_stream(stream)>>>(lnsum_y, M, lnsum_z, N, group_of_channel, n_groups,
                                                                group_inv_pixels, group_weight, z_inv_pixels,
                                                                z_weight, task_losses, T, log_vars, lmbda, scalars,
                                                                g_lnsum_y, g_lnsum_z, g_task_losses, g_log_vars);
    return after_launch("rd_epilogue_kernel");
}

extern "C" int mmnc_nonneg_reparam_forward(const float *p, int64_t n, float bound, float pedestal, float *out,
                                           void *stream) {
    MMNC_REQUIRE(n >= 0, "nonneg_reparam_forward: negative size");
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(p && out, "nonneg_reparam_forward: null pointer");
    nonneg_forward_kernel<<<d_blocks(n, 1), D_THREADS, 0, as_stream(stream)>>>(p, n, bound, pedestal, out);
    return after_launch("nonneg_forward_kernel");
}

extern "C" int mmnc_nonneg_reparam_backward(const float *p, const float *g_out, int64_t n, float bound, float *g_p,
                                            void *stream) {
    MMNC_REQUIRE(n >= 0, "nonneg_reparam_backward: negative size");
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(p && g_out && g_p, "nonneg_reparam_backward: null pointer");
    nonneg_backward_kernel<<<d_blocks(n, 1), D_THREADS, 0, as_stream(stream)>>>(p, g_out, n, bound, g_p);
    return after_launch("nonneg_backward_kernel");
}
