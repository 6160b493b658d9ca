// K1 + K3: GaussianConditional forward / backward and the per-channel ln-likelihood reduction
// (SURVEY.md section 8 rows a2, a5, a6).
//
// Replaces CompressAI's GaussianConditional.forward chain (quantise, LowerBound(scales), abs, 2 x erfc, sub,
// LowerBound(lik): ~16 launches) plus the reference's per-group `log().sum()` (4 launches per group,
// /root/reference/src/models/multi_task_compressor.py:288-291) with one launch per direction.
//
// y is (B, C, Sy), scales / lik are (B, C, Ss) with Sy == Ss or Sy == 1: the reference really produces
// y (B,M,1,1) against scales (B,M,4,4) and relies on torch broadcasting; the
// broadcast is a stride-0 read here and the backward sums g_y over the Ss positions with lane shuffles.
// grid = (C, splits): a block owns one channel so the channel's sum of ln(lik) is a block reduction.
#include "common.cuh"
#include "hd_math.cuh"

namespace mmnc {

constexpr int GC_THREADS = 256;

__global__ void __launch_bounds__(GC_THREADS)
gc_forward_kernel(const float *__restrict__ y, const float *__restrict__ scales, const float *__restrict__ means,
                  int64_t B, int64_t C, int64_t Sy, int64_t Ss, int noise_mode, const float *__restrict__ noise,
                  uint64_t seed, uint64_t offset, float scale_bound, float lik_bound, float *__restrict__ y_hat,
                  float *__restrict__ lik, float *__restrict__ lnsum) {
    __shared__ float red[32];
    const int64_t c = blockIdx.x;
    const int64_t n = B * Ss;
    const bool bcast = (Sy != Ss);
    float acc = 0.f;
    if (noise_mode == MMNC_QUANT_NOISE_PHILOX_DEV) {  // (seed, offset) live in device memory: CUDA-graph friendly
        const uint64_t *st = reinterpret_cast<const uint64_t *>(noise);
        seed = st[0];
        offset += st[1];
        noise_mode = MMNC_QUANT_NOISE_PHILOX;
    }
    auto element = [&](int64_t b, int64_t s) {
        const int64_t ai = (b * C + c) * Ss + s;
        const int64_t yi = bcast ? (b * C + c) : ai;
        const float yv = y[yi];
        const float m = (means != nullptr) ? means[yi] : 0.f;
        float v;
        if (noise_mode == MMNC_QUANT_DEQUANTIZE) v = rintf(yv - m) + m;
        else if (noise_mode == MMNC_QUANT_NOISE_PHILOX) v = yv + philox_uniform_centered(seed, (uint64_t)yi + offset);
        else if (noise_mode == MMNC_QUANT_NOISE_GIVEN) v = yv + noise[yi];
        else v = yv;
        // fp32 without the erfc cancellation (hd_math.cuh): within 1e-6 of the float64 value of the same formula
        float l = gc_likelihood_s(v, m, scales[ai], scale_bound);
        if (lik_bound > 0.f) l = max_nan(l, lik_bound);
        lik[ai] = l;
        if (!bcast || s == 0) y_hat[yi] = v;
        acc += fast_log(l);
    };
    if (n < (1ll << 31)) {  // 32-bit index arithmetic (a 64-bit division per element costs as much as the likelihood)
        const uint32_t n32 = (uint32_t)n, ss32 = (uint32_t)Ss, step = gridDim.y * blockDim.x;
        for (uint32_t e = blockIdx.y * blockDim.x + threadIdx.x; e < n32; e += step) {
            const uint32_t b = e / ss32;
            element((int64_t)b, (int64_t)(e - b * ss32));
            if (e + step < e) break;  // wrap-around guard
        }
    } else {
        for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.y * blockDim.x) {
            const int64_t b = e / Ss;
            element(b, e - b * Ss);
        }
    }
    if (lnsum != nullptr) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) atomicAdd(&lnsum[c], tot);
    }
}

// Vector path (no broadcast, no means, Ss % 4 == 0, 16-byte aligned tensors): a thread owns four consecutive elements
// of one (image, channel) row - 128-bit loads and stores, ONE Philox call for the four noise values, four independent
// likelihood evaluations in flight.  This is the path the declared roofline shape (1024, 192, 16, 16) takes.
__global__ void __launch_bounds__(GC_THREADS)
gc_forward_vec4_kernel(const float *__restrict__ y, const float *__restrict__ scales, uint32_t B, uint32_t C, uint32_t Ss,
                       int noise_mode, const float *__restrict__ noise, uint64_t seed, uint64_t offset,
                       float scale_bound, float lik_bound, float *__restrict__ y_hat, float *__restrict__ lik,
                       float *__restrict__ lnsum) {
    __shared__ float red[32];
    const uint32_t c = blockIdx.x;
    const uint32_t q = Ss >> 2, n4 = B * q;  // groups of four per channel
    float acc = 0.f;
    if (noise_mode == MMNC_QUANT_NOISE_PHILOX_DEV) {
        const uint64_t *st = reinterpret_cast<const uint64_t *>(noise);
        seed = st[0];
        offset += st[1];
        noise_mode = MMNC_QUANT_NOISE_PHILOX;
    }
    const uint32_t step = gridDim.y * blockDim.x;
    for (uint32_t e = blockIdx.y * blockDim.x + threadIdx.x; e < n4; e += step) {
        const uint32_t b = e / q, s4 = e - b * q;
        const int64_t ai = ((int64_t)b * C + c) * Ss + 4 * s4;
        const float4 yv = *reinterpret_cast<const float4 *>(y + ai);
        const float4 sv = *reinterpret_cast<const float4 *>(scales + ai);
        float v[4] = {yv.x, yv.y, yv.z, yv.w};
        const float sc[4] = {sv.x, sv.y, sv.z, sv.w};
        if (noise_mode == MMNC_QUANT_DEQUANTIZE) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = rintf(v[k]);
        } else if (noise_mode == MMNC_QUANT_NOISE_PHILOX) {
            const uint64_t gi = (uint64_t)ai + offset;
            float u[4];
            if ((gi & 3ull) == 0) {
                philox4_centered(seed, gi >> 2, u);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) u[k] = philox_uniform_centered(seed, gi + k);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] += u[k];
        } else if (noise_mode == MMNC_QUANT_NOISE_GIVEN) {
            const float4 nv = *reinterpret_cast<const float4 *>(noise + ai);
            v[0] += nv.x; v[1] += nv.y; v[2] += nv.z; v[3] += nv.w;
        }
        float l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            l[k] = gc_likelihood_s(v[k], 0.f, sc[k], scale_bound);
            if (lik_bound > 0.f) l[k] = max_nan(l[k], lik_bound);
            acc += fast_log(l[k]);
        }
        __stcs(reinterpret_cast<float4 *>(lik + ai), make_float4(l[0], l[1], l[2], l[3]));
        __stcs(reinterpret_cast<float4 *>(y_hat + ai), make_float4(v[0], v[1], v[2], v[3]));
        if (e + step < e) break;
    }
    if (lnsum != nullptr) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) atomicAdd(&lnsum[c], tot);
    }
}

__global__ void __launch_bounds__(GC_THREADS)
gc_backward_vec4_kernel(const float *__restrict__ y_hat, const float *__restrict__ scales, uint32_t B, uint32_t C,
                        uint32_t Ss, const float *__restrict__ g_yhat, const float *__restrict__ g_lik,
                        const float *__restrict__ g_lnsum, float scale_bound, float lik_bound, float *__restrict__ g_y,
                        float *__restrict__ g_scales) {
    const uint32_t c = blockIdx.x;
    const uint32_t q = Ss >> 2, n4 = B * q;
    const float gls = (g_lnsum != nullptr) ? g_lnsum[c] : 0.f;
    const uint32_t step = gridDim.y * blockDim.x;
    for (uint32_t e = blockIdx.y * blockDim.x + threadIdx.x; e < n4; e += step) {
        const uint32_t b = e / q, s4 = e - b * q;
        const int64_t ai = ((int64_t)b * C + c) * Ss + 4 * s4;
        const float4 yv = *reinterpret_cast<const float4 *>(y_hat + ai);
        const float4 sv = *reinterpret_cast<const float4 *>(scales + ai);
        float4 gl4 = make_float4(0.f, 0.f, 0.f, 0.f), gy4 = gl4;
        if (g_lik != nullptr) gl4 = *reinterpret_cast<const float4 *>(g_lik + ai);
        if (g_yhat != nullptr) gy4 = *reinterpret_cast<const float4 *>(g_yhat + ai);
        const float v[4] = {yv.x, yv.y, yv.z, yv.w}, sc[4] = {sv.x, sv.y, sv.z, sv.w};
        const float gl[4] = {gl4.x, gl4.y, gl4.z, gl4.w}, gyh[4] = {gy4.x, gy4.y, gy4.z, gy4.w};
        float oy[4], os[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float raw = gc_likelihood_s(v[k], 0.f, sc[k], scale_bound);
            const float l = (lik_bound > 0.f) ? max_nan(raw, lik_bound) : raw;
            float g = gl[k] + fast_div(gls, l);
            if (lik_bound > 0.f) g = lower_bound_grad(raw, lik_bound, g);
            float dy, dsc;
            gc_likelihood_grad(v[k], 0.f, sc[k], scale_bound, &dy, &dsc);
            oy[k] = g * dy + gyh[k];
            os[k] = lower_bound_grad(sc[k], scale_bound, g * dsc);
        }
        __stcs(reinterpret_cast<float4 *>(g_y + ai), make_float4(oy[0], oy[1], oy[2], oy[3]));
        __stcs(reinterpret_cast<float4 *>(g_scales + ai), make_float4(os[0], os[1], os[2], os[3]));
        if (e + step < e) break;
    }
}

// reduce_mode: 0 = no broadcast, 1 = shuffle over groups of Ss lanes (Ss power of two <= 32), 2 = atomics
__global__ void __launch_bounds__(GC_THREADS)
gc_backward_kernel(const float *__restrict__ y_hat, const float *__restrict__ scales,
                   const float *__restrict__ means, int64_t B, int64_t C, int64_t Sy, int64_t Ss,
                   const float *__restrict__ g_yhat, const float *__restrict__ g_lik,
                   const float *__restrict__ g_lnsum, float scale_bound, float lik_bound, int reduce_mode,
                   float *__restrict__ g_y, float *__restrict__ g_scales) {
    const int64_t c = blockIdx.x;
    const int64_t n = B * Ss;
    const bool bcast = (Sy != Ss);
    const float gls = (g_lnsum != nullptr) ? g_lnsum[c] : 0.f;
    for (int64_t e0 = (int64_t)blockIdx.y * blockDim.x; e0 < n; e0 += (int64_t)gridDim.y * blockDim.x) {
        const int64_t e = e0 + threadIdx.x;
        const bool valid = e < n;
        float gy = 0.f;
        int64_t yi = 0, s = 0;
        if (valid) {
            const int64_t b = e / Ss;
            s = e - b * Ss;
            const int64_t ai = (b * C + c) * Ss + s;
            yi = bcast ? (b * C + c) : ai;
            const float v = y_hat[yi];
            const float m = (means != nullptr) ? means[yi] : 0.f;
            const float sc = scales[ai];
            const float raw = gc_likelihood_s(v, m, sc, scale_bound);
            const float l = (lik_bound > 0.f) ? max_nan(raw, lik_bound) : raw;
            float g = (g_lik != nullptr ? g_lik[ai] : 0.f) + fast_div(gls, l);
            if (lik_bound > 0.f) g = lower_bound_grad(raw, lik_bound, g);
            float dy, dsc;
            gc_likelihood_grad(v, m, sc, scale_bound, &dy, &dsc);
            gy = g * dy;
            g_scales[ai] = lower_bound_grad(sc, scale_bound, g * dsc);
        }
        if (reduce_mode == 0) {
            if (valid) g_y[yi] = gy + (g_yhat != nullptr ? g_yhat[yi] : 0.f);
        } else if (reduce_mode == 1) {
            for (int o = (int)Ss >> 1; o > 0; o >>= 1) gy += __shfl_xor_sync(0xffffffffu, gy, o);
            if (valid && s == 0) g_y[yi] = gy + (g_yhat != nullptr ? g_yhat[yi] : 0.f);
        } else {
            if (valid) {
                if (s == 0 && g_yhat != nullptr) gy += g_yhat[yi];
                atomicAdd(&g_y[yi], gy);
            }
        }
    }
}

// lnsum[c] += sum_{b,s} ln lik[b,c,s]
__global__ void __launch_bounds__(GC_THREADS)
lnsum_forward_kernel(const float *__restrict__ lik, int64_t B, int64_t C, int64_t S, float *__restrict__ lnsum) {
    __shared__ float red[32];
    const int64_t c = blockIdx.x;
    const int64_t n = B * S;
    float acc = 0.f;
    for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.y * blockDim.x) {
        const int64_t b = e / S, s = e - b * S;
        acc += fast_log(lik[(b * C + c) * S + s]);
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(&lnsum[c], tot);
}

__global__ void __launch_bounds__(GC_THREADS)
lnsum_backward_kernel(const float *__restrict__ lik, int64_t n, int64_t C, int64_t S,
                      const float *__restrict__ g_lnsum, float *__restrict__ g_lik) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = (i / S) % C;
        g_lik[i] = g_lnsum[c] / lik[i];
    }
}

// One block per channel whenever the channel is small (every shape of the reference's models: <= 4096 elements per
// channel): its per-channel sum is then a single fixed-order block reduction and the result is bit-reproducible.  Only
// large inputs are split over several blocks, whose partial sums meet in an atomicAdd (order not fixed).
static inline int gc_splits(int64_t C, int64_t n_per_channel, int per_thread = 1) {
    if (n_per_channel <= 8192) return 1;
    const int64_t target_blocks = (int64_t)sm_count() * 8;
    int64_t splits = (target_blocks + C - 1) / C;
    const int64_t max_splits = (n_per_channel / per_thread + GC_THREADS - 1) / GC_THREADS;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    return (int)splits;
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace mmnc

using namespace mmnc;

extern "C" int mmnc_gc_forward(const float *y, const float *scales, const float *means, int64_t B, int64_t C,
                               int64_t Sy, int64_t Ss, int noise_mode, const float *noise, uint64_t seed,
                               uint64_t offset, float scale_bound, float likelihood_bound, float *y_hat,
                               float *lik, float *lnsum, void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && Sy >= 0 && Ss >= 0, "gc_forward: negative dimension");
    MMNC_REQUIRE(Sy == Ss || Sy == 1, "gc_forward: y spatial size must equal the scales' or be 1 (got %lld vs %lld)",
                 (long long)Sy, (long long)Ss);
    MMNC_REQUIRE(noise_mode >= 0 && noise_mode <= 4, "gc_forward: bad noise_mode %d", noise_mode);
    if (B * C * Ss == 0) return MMNC_OK;
    MMNC_REQUIRE(y && scales && y_hat && lik, "gc_forward: null pointer");
    MMNC_REQUIRE((noise_mode != MMNC_QUANT_NOISE_GIVEN && noise_mode != MMNC_QUANT_NOISE_PHILOX_DEV) || noise,
                 "gc_forward: this noise_mode needs the noise pointer");
    if (Sy == Ss && means == nullptr && (Ss & 3) == 0 && B * Ss < (1ll << 31) && C < (1ll << 31) && aligned16(y) &&
        aligned16(scales) && aligned16(y_hat) && aligned16(lik) &&
        (noise_mode != MMNC_QUANT_NOISE_GIVEN || aligned16(noise))) {
        dim3 grid4((unsigned)C, (unsigned)gc_splits(C, B * Ss, 4));
        gc_forward_vec4_kernel<<<grid4, GC_THREADS, 0, as_stream(stream)>>>(y, scales, (uint32_t)B, (uint32_t)C,
                                                                            (uint32_t)Ss, noise_mode, noise, seed, offset,
                                                                            scale_bound, likelihood_bound, y_hat, lik,
                                                                            lnsum);
        return after_launch("gc_forward_vec4_kernel");
    }
    dim3 grid((unsigned)C, (unsigned)gc_splits(C, B * Ss));
    gc_forward_kernel<<<grid, GC_THREADS, 0, as_stream(stream)>>>(y, scales, means, B, C, Sy, Ss, noise_mode,
                                                                   noise, seed, offset, scale_bound,
                                                                   likelihood_bound, y_hat, lik, lnsum);
    return after_launch("gc_forward_kernel");
}

extern "C" int mmnc_gc_backward(const float *y_hat, const float *scales, const float *means, int64_t B, int64_t C,
                                int64_t Sy, int64_t Ss, const float *g_yhat, const float *g_lik,
                                const float *g_lnsum, float scale_bound, float likelihood_bound, float *g_y,
                                float *g_scales, void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && Sy >= 0 && Ss >= 0, "gc_backward: negative dimension");
    MMNC_REQUIRE(Sy == Ss || Sy == 1, "gc_backward: unsupported broadcast");
    if (B * C * Ss == 0) return MMNC_OK;
    MMNC_REQUIRE(y_hat && scales && g_y && g_scales, "gc_backward: null pointer");
    if (Sy == Ss && means == nullptr && (Ss & 3) == 0 && B * Ss < (1ll << 31) && C < (1ll << 31) && aligned16(y_hat) &&
        aligned16(scales) && aligned16(g_y) && aligned16(g_scales) && aligned16(g_yhat) && aligned16(g_lik)) {
        dim3 grid4((unsigned)C, (unsigned)gc_splits(C, B * Ss, 4));
        gc_backward_vec4_kernel<<<grid4, GC_THREADS, 0, as_stream(stream)>>>(y_hat, scales, (uint32_t)B, (uint32_t)C,
                                                                             (uint32_t)Ss, g_yhat, g_lik, g_lnsum,
                                                                             scale_bound, likelihood_bound, g_y,
                                                                             g_scales);
        return after_launch("gc_backward_vec4_kernel");
    }
    int reduce_mode = 0;
    if (Sy != Ss) {
        const bool pow2 = (Ss & (Ss - 1)) == 0;
        reduce_mode = (pow2 && Ss <= 32) ? 1 : 2;
        if (reduce_mode == 2) MMNC_CUDA(cudaMemsetAsync(g_y, 0, sizeof(float) * (size_t)(B * C * Sy), as_stream(stream)));
    }
    dim3 grid((unsigned)C, (unsigned)gc_splits(C, B * Ss));
    gc_backward_kernel<<<grid, GC_THREADS, 0, as_stream(stream)>>>(y_hat, scales, means, B, C, Sy, Ss, g_yhat, g_lik,
                                                                    g_lnsum, scale_bound, likelihood_bound,
                                                                    reduce_mode, g_y, g_scales);
    return after_launch("gc_backward_kernel");
}

extern "C" int mmnc_lnsum_forward(const float *lik, int64_t B, int64_t C, int64_t S, float *lnsum, void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && S >= 0, "lnsum_forward: negative dimension");
    if (B * C * S == 0) return MMNC_OK;
    MMNC_REQUIRE(lik && lnsum, "lnsum_forward: null pointer");
    dim3 grid((unsigned)C, (unsigned)gc_splits(C, B * S));
    lnsum_forward_kernel<<<grid, GC_THREADS, 0, as_stream(stream)>>>(lik, B, C, S, lnsum);
    return after_launch("lnsum_forward_kernel");
}

extern "C" int mmnc_lnsum_backward(const float *lik, int64_t B, int64_t C, int64_t S, const float *g_lnsum,
                                   float *g_lik, void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && S >= 0, "lnsum_backward: negative dimension");
    const int64_t n = B * C * S;
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(lik && g_lnsum && g_lik, "lnsum_backward: null pointer");
    int64_t blocks = (n + GC_THREADS - 1) / GC_THREADS;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    lnsum_backward_kernel<<<(unsigned)blocks, GC_THREADS, 0, as_stream(stream)>>>(lik, n, C, S, g_lnsum, g_lik);
    return after_launch("lnsum_backward_kernel");
}
