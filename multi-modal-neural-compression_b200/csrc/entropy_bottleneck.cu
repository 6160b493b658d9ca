// K1 + K2: EntropyBottleneck forward / backward.
//
// Replaces CompressAI's EntropyBottleneck.forward chain (quantise -> permute -> 2 x 5-layer per-channel
// softplus-matrix MLP -> sign trick -> sigmoid difference -> LowerBound -> permute back; ~78 torch launches,
// reference call site /root/reference/src/models/multi_task_compressor.py:495) with ONE launch per direction.
//
// Layout: x is (B, C, S) NCHW-contiguous; the (C, 1, B*S) permutation CompressAI materialises is only an
// indexing change here.  grid = (C, splits): a block owns one channel, so its 58 transformed parameters sit in
// shared memory and every lane reads them as broadcasts; threads stride over that channel's B*S elements.
// Per-channel reductions (sum of ln lik, 58 parameter gradients) are warp shuffles -> shared -> one atomic per
// block and value.
#include "common.cuh"
#include "hd_math.cuh"

namespace mmnc {

constexpr int EB_THREADS = 128;

__device__ __forceinline__ int64_t eb_addr(int64_t e, int64_t c, int64_t C, int64_t S) {
    const int64_t b = e / S, s = e - b * S;
    return (b * C + c) * S + s;
}

__global__ void __launch_bounds__(EB_THREADS)
eb_forward_kernel(const float *__restrict__ x, int64_t B, int64_t C, int64_t S, const float *__restrict__ params,
                  const float *__restrict__ medians, int noise_mode, const float *__restrict__ noise,
                  uint64_t seed, uint64_t offset, float bound, int form, float *__restrict__ out,
                  float *__restrict__ lik, float *__restrict__ lnsum) {
    __shared__ float P[EB_NP];
    __shared__ float red[32];
    const int64_t c = blockIdx.x;
    if (threadIdx.x < EB_NP) P[threadIdx.x] = eb_transform(threadIdx.x, params[c * EB_NP + threadIdx.x]);
    __syncthreads();
    const float med = medians[c];
    const int64_t n = B * S;
    float acc = 0.f;
    if (noise_mode == MMNC_QUANT_NOISE_PHILOX_DEV) {  // (seed, offset) live in device memory: CUDA-graph friendly
        const uint64_t *st = reinterpret_cast<const uint64_t *>(noise);
        seed = st[0];
        offset += st[1];
        noise_mode = MMNC_QUANT_NOISE_PHILOX;
    }
    (void)form;  // "sign" and "plain" are the same number; the difference-propagating evaluation serves both
    auto element = [&](int64_t a) {
        const float xv = x[a];
        float v;
        if (noise_mode == MMNC_QUANT_DEQUANTIZE) v = rintf(xv - med) + med;
        else if (noise_mode == MMNC_QUANT_NOISE_PHILOX) v = xv + philox_uniform_centered(seed, (uint64_t)a + offset);
        else if (noise_mode == MMNC_QUANT_NOISE_GIVEN) v = xv + noise[a];
        else v = xv;
        // the reference forms v - 0.5 / v + 0.5 in fp32 before the MLP; keep that rounding.  fp32 without the
        // sigmoid cancellation (hd_math.cuh): within ~2e-6 of the float64 value of the formula
        const float tl = v - 0.5f, tu = v + 0.5f;
        float l = eb_likelihood_s(P, tl, tu - tl);
        if (bound > 0.f) l = max_nan(l, bound);
        out[a] = v;
        lik[a] = l;
        acc += fast_log(l);
    };
    if (n < (1ll << 31)) {  // 32-bit index arithmetic
        const uint32_t n32 = (uint32_t)n, s32 = (uint32_t)S, step = gridDim.y * blockDim.x;
        for (uint32_t e = blockIdx.y * blockDim.x + threadIdx.x; e < n32; e += step) {
            const uint32_t b = e / s32;
            element(((int64_t)b * C + c) * S + (int64_t)(e - b * s32));
            if (e + step < e) break;  // wrap-around guard
        }
    } else {
        for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.y * blockDim.x)
            element(eb_addr(e, c, C, S));
    }
    if (lnsum != nullptr) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) atomicAdd(&lnsum[c], tot);
    }
}

// Vector path (S % 4 == 0, 16-byte aligned tensors): four consecutive elements of one (image, channel) row per thread -
// 128-bit loads and stores, one Philox call for the four noise values, four independent MLP evaluations in flight.
__global__ void __launch_bounds__(EB_THREADS)
eb_forward_vec4_kernel(const float *__restrict__ x, uint32_t B, uint32_t C, uint32_t S, const float *__restrict__ params,
                       const float *__restrict__ medians, int noise_mode, const float *__restrict__ noise,
                       uint64_t seed, uint64_t offset, float bound, float *__restrict__ out, float *__restrict__ lik,
                       float *__restrict__ lnsum) {
    __shared__ float P[EB_NP];
    __shared__ float red[32];
    const uint32_t c = blockIdx.x;
    if (threadIdx.x < EB_NP) P[threadIdx.x] = eb_transform(threadIdx.x, params[(int64_t)c * EB_NP + threadIdx.x]);
    __syncthreads();
    const float med = medians[c];
    const uint32_t q = S >> 2, n4 = B * q;
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
        const int64_t a = ((int64_t)b * C + c) * S + 4 * s4;
        const float4 xv = *reinterpret_cast<const float4 *>(x + a);
        float v[4] = {xv.x, xv.y, xv.z, xv.w};
        if (noise_mode == MMNC_QUANT_DEQUANTIZE) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = rintf(v[k] - med) + med;
        } else if (noise_mode == MMNC_QUANT_NOISE_PHILOX) {
            const uint64_t gi = (uint64_t)a + offset;
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
            const float4 nv = *reinterpret_cast<const float4 *>(noise + a);
            v[0] += nv.x; v[1] += nv.y; v[2] += nv.z; v[3] += nv.w;
        }
        float l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float tl = v[k] - 0.5f, tu = v[k] + 0.5f;
            l[k] = eb_likelihood_s(P, tl, tu - tl);
            if (bound > 0.f) l[k] = max_nan(l[k], bound);
            acc += fast_log(l[k]);
        }
        __stcs(reinterpret_cast<float4 *>(out + a), make_float4(v[0], v[1], v[2], v[3]));
        __stcs(reinterpret_cast<float4 *>(lik + a), make_float4(l[0], l[1], l[2], l[3]));
        if (e + step < e) break;
    }
    if (lnsum != nullptr) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) atomicAdd(&lnsum[c], tot);
    }
}

__global__ void __launch_bounds__(EB_THREADS)
eb_backward_kernel(const float *__restrict__ outv, int64_t B, int64_t C, int64_t S,
                   const float *__restrict__ params, const float *__restrict__ g_out,
                   const float *__restrict__ g_lik, const float *__restrict__ g_lnsum, float bound, int form,
                   int train, float *__restrict__ g_x, float *__restrict__ g_params) {
    __shared__ float P[EB_NP];
    __shared__ float Praw[EB_NP];
    __shared__ float red[EB_THREADS / 32][EB_NP];
    const int64_t c = blockIdx.x;
    if (threadIdx.x < EB_NP) {
        const float raw = params[c * EB_NP + threadIdx.x];
        Praw[threadIdx.x] = raw;
        P[threadIdx.x] = eb_transform(threadIdx.x, raw);
    }
    __syncthreads();
    const float gls = (g_lnsum != nullptr) ? g_lnsum[c] : 0.f;
    const int64_t n = B * S;
    float gP[EB_NP];
#pragma unroll
    for (int k = 0; k < EB_NP; ++k) gP[k] = 0.f;
    for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.y * blockDim.x) {
        const int64_t a = eb_addr(e, c, C, S);
        const float v = outv[a];
        EbTrace tl, tu;
        const float lower = eb_logits<true>(P, v - 0.5f, &tl);
        const float upper = eb_logits<true>(P, v + 0.5f, &tu);
        const float raw = eb_likelihood(lower, upper, form);
        const float l = (bound > 0.f) ? max_nan(raw, bound) : raw;
        float g = (g_lik != nullptr ? g_lik[a] : 0.f) + fast_div(gls, l);
        if (bound > 0.f) g = lower_bound_grad(raw, bound, g);
        float dl, du;
        eb_likelihood_grad(lower, upper, form, &dl, &du);
        float gv = eb_logits_backward<true>(P, v - 0.5f, tl, g * dl, gP);
        gv += eb_logits_backward<true>(P, v + 0.5f, tu, g * du, gP);
        // noise mode: out = x + u -> d out / d x = 1.  dequantize mode: round() has zero gradient.
        g_x[a] = train ? (gv + (g_out != nullptr ? g_out[a] : 0.f)) : 0.f;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < EB_NP; ++k) {
        const float s = warp_sum(gP[k]);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < EB_NP) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < EB_THREADS / 32; ++w) s += red[w][threadIdx.x];
        s *= eb_transform_grad(threadIdx.x, Praw[threadIdx.x], P[threadIdx.x]);
        atomicAdd(&g_params[c * EB_NP + threadIdx.x], s);
    }
}

// _logits_cumulative on (C, L) with detached parameters
__global__ void __launch_bounds__(EB_THREADS)
eb_logits_kernel(const float *__restrict__ v, int64_t C, int64_t L, const float *__restrict__ params,
                 float *__restrict__ logits) {
    __shared__ float P[EB_NP];
    const int64_t c = blockIdx.x;
    if (threadIdx.x < EB_NP) P[threadIdx.x] = eb_transform(threadIdx.x, params[c * EB_NP + threadIdx.x]);
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.y * blockDim.x)
        logits[c * L + i] = eb_logits<false>(P, v[c * L + i], nullptr);
}

// aux loss: sum_c sum_{q<3} |F_c(quantiles[c][q]) - target[q]|, gradient w.r.t. quantiles only.
// One warp per channel-triple would waste lanes; instead thread <-> (channel, q) and parameters straight from
// global memory (the whole input is 3*C floats; this kernel exists to collapse ~35 launches into one).
__global__ void __launch_bounds__(EB_THREADS)
eb_aux_loss_kernel(const float *__restrict__ quantiles, int64_t C, const float *__restrict__ params,
                   const float *__restrict__ target3, float *__restrict__ loss, float *__restrict__ g_quantiles) {
    __shared__ float red[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float contrib = 0.f;
    if (i < 3 * C) {
        const int64_t c = i / 3;
        const int q = (int)(i - 3 * c);
        float P[EB_NP];
#pragma unroll
        for (int k = 0; k < EB_NP; ++k) P[k] = eb_transform(k, params[c * EB_NP + k]);
        const float t = quantiles[i];
        EbTrace tr;
        const float f = eb_logits<true>(P, t, &tr);
        const float d = f - target3[q];
        contrib = fabsf(d);
        g_quantiles[i] = eb_logits_backward<false>(P, t, tr, sign_t(d), nullptr);
    }
    const float tot = block_sum(contrib, red);
    if (threadIdx.x == 0) atomicAdd(loss, tot);
}

// One block per channel whenever the channel is small (every shape of the reference's models: <= 4096 elements per
// channel): per-channel sums (ln-likelihood, the 58 parameter gradients) are then single fixed-order block reductions
// and bit-reproducible.  Large inputs are split over several blocks whose partial sums meet in atomicAdds.
static inline int eb_splits(int64_t C, int64_t n_per_channel, int per_thread = 1) {
    if (n_per_channel <= 8192) return 1;
    // enough blocks to cover the machine a few times over, but never more than the work in a channel
    const int64_t target_blocks = (int64_t)sm_count() * 8;
    int64_t splits = (target_blocks + C - 1) / C;
    const int64_t max_splits = (n_per_channel / per_thread + EB_THREADS - 1) / EB_THREADS;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    return (int)splits;
}

}  // namespace mmnc

using namespace mmnc;

extern "C" int mmnc_eb_forward(const float *x, int64_t B, int64_t C, int64_t S, const float *params,
                               const float *medians, int noise_mode, const float *noise, uint64_t seed,
                               uint64_t offset, float likelihood_bound, int likelihood_form, float *out,
                               float *lik, float *lnsum, void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && S >= 0, "eb_forward: negative dimension");
    MMNC_REQUIRE(noise_mode >= 0 && noise_mode <= 4, "eb_forward: bad noise_mode %d", noise_mode);
    MMNC_REQUIRE(likelihood_form == 0 || likelihood_form == 1, "eb_forward: bad likelihood_form");
    if (B * C * S == 0) return MMNC_OK;
    MMNC_REQUIRE(x && params && medians && out && lik, "eb_forward: null pointer");
    MMNC_REQUIRE((noise_mode != MMNC_QUANT_NOISE_GIVEN && noise_mode != MMNC_QUANT_NOISE_PHILOX_DEV) || noise,
                 "eb_forward: this noise_mode needs the noise pointer");
    MMNC_REQUIRE(C <= 2147483647LL, "eb_forward: too many channels");
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(lik) |
                      (noise_mode == MMNC_QUANT_NOISE_GIVEN ? reinterpret_cast<uintptr_t>(noise) : 0)) & 15) == 0;
    if ((S & 3) == 0 && al && B * S < (1ll << 31)) {
        dim3 grid4((unsigned)C, (unsigned)eb_splits(C, B * S, 4));
        eb_forward_vec4_kernel<<<grid4, EB_THREADS, 0, as_stream(stream)>>>(x, (uint32_t)B, (uint32_t)C, (uint32_t)S, params,
                                                                            medians, noise_mode, noise, seed, offset,
                                                                            likelihood_bound, out, lik, lnsum);
        return after_launch("eb_forward_vec4_kernel");
    }
    dim3 grid((unsigned)C, (unsigned)eb_splits(C, B * S));
    eb_forward_kernel<<<grid, EB_THREADS, 0, as_stream(stream)>>>(x, B, C, S, params, medians, noise_mode, noise,
                                                                   seed, offset, likelihood_bound, likelihood_form,
                                                                   out, lik, lnsum);
    return after_launch("eb_forward_kernel");
}

extern "C" int mmnc_eb_backward(const float *out, int64_t B, int64_t C, int64_t S, const float *params,
                                const float *g_out, const float *g_lik, const float *g_lnsum,
                                float likelihood_bound, int likelihood_form, float *g_x, float *g_params,
                                void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && S >= 0, "eb_backward: negative dimension");
    if (B * C * S == 0) return MMNC_OK;
    MMNC_REQUIRE(out && params && g_x && g_params, "eb_backward: null pointer");
    dim3 grid((unsigned)C, (unsigned)eb_splits(C, B * S));
    eb_backward_kernel<<<grid, EB_THREADS, 0, as_stream(stream)>>>(out, B, C, S, params, g_out, g_lik, g_lnsum,
                                                                    likelihood_bound, likelihood_form, 1, g_x,
                                                                    g_params);
    return after_launch("eb_backward_kernel");
}

extern "C" int mmnc_eb_logits(const float *v, int64_t C, int64_t L, const float *params, float *logits,
                              void *stream) {
    MMNC_REQUIRE(C >= 0 && L >= 0, "eb_logits: negative dimension");
    if (C * L == 0) return MMNC_OK;
    MMNC_REQUIRE(v && params && logits, "eb_logits: null pointer");
    dim3 grid((unsigned)C, (unsigned)eb_splits(C, L));
    eb_logits_kernel<<<grid, EB_THREADS, 0, as_stream(stream)>>>(v, C, L, params, logits);
    return after_launch("eb_logits_kernel");
}

extern "C" int mmnc_eb_aux_loss(const float *quantiles, int64_t C, const float *params, const float *target3,
                                float *loss, float *g_quantiles, void *stream) {
    MMNC_REQUIRE(C >= 0, "eb_aux_loss: negative dimension");
    if (C == 0) return MMNC_OK;
    MMNC_REQUIRE(quantiles && params && target3 && loss && g_quantiles, "eb_aux_loss: null pointer");
    const int64_t n = 3 * C;
    eb_aux_loss_kernel<<<(unsigned)((n + EB_THREADS - 1) / EB_THREADS), EB_THREADS, 0, as_stream(stream)>>>(
        quantiles, C, params, target3, loss, g_quantiles);
    return after_launch("eb_aux_loss_kernel");
}
