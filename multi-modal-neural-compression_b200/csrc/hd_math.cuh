// Per-element arithmetic of the rate path, written once as host+device inline functions.
// The CUDA kernels call these; tests/hostcheck compiles the same header with g++ to check the logic against
// the oracle without a GPU (test infrastructure only — the shipped library exports GPU entry points only).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MMNC_HD __host__ __device__ __forceinline__
#else
#define MMNC_HD inline
#endif

namespace mmnc {

// ---------------------------------------------------------------------------------------------- scalars
MMNC_HD float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // F.softplus(beta=1, threshold=20)
MMNC_HD float sigmoid_t(float x) { return 1.f / (1.f + expf(-x)); }
MMNC_HD float sign_t(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
// torch.max(x, bound): unlike fmaxf it propagates a NaN in x (LowerBound must not hide a diverged value)
MMNC_HD float max_nan(float x, float bound) { return (x != x) ? x : fmaxf(x, bound); }

// Device: MUFU-based approximations (ex2.approx / lg2.approx / rcp.approx: 1-2 ulp) where the error budget allows it
// (stated at each use); host (tests/hostcheck): the libm function of the same name.
#if defined(__CUDA_ARCH__)
MMNC_HD float fast_exp(float x) { return __expf(x); }
MMNC_HD float fast_log(float x) { return __logf(x); }
MMNC_HD float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MMNC_HD float fast_div(float a, float b) { return __fdividef(a, b); }
#else
MMNC_HD float fast_exp(float x) { return expf(x); }
MMNC_HD float fast_log(float x) { return logf(x); }
MMNC_HD float fast_rcp(float x) { return 1.f / x; }
MMNC_HD float fast_div(float a, float b) { return a / b; }
#endif

// Philox4x32-10 keyed by a 64-bit seed, counter = 64-bit block index.  One call yields FOUR uniforms: element i of a
// noise stream uses component (i & 3) of block (i >> 2), so a thread that owns four consecutive elements pays for one
// call (the forward kernels' vector path) and a scalar caller gets the same numbers.  Values are U[-0.5, 0.5).
MMNC_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
MMNC_HD void philox4_centered(uint64_t seed, uint64_t block, float out[4]) {
    uint32_t c0 = (uint32_t)block, c1 = (uint32_t)(block >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = (float)(c0 >> 8) * (1.0f / 16777216.0f) - 0.5f;  // 24-bit mantissa, [0,1) - 0.5
    out[1] = (float)(c1 >> 8) * (1.0f / 16777216.0f) - 0.5f;
    out[2] = (float)(c2 >> 8) * (1.0f / 16777216.0f) - 0.5f;
    out[3] = (float)(c3 >> 8) * (1.0f / 16777216.0f) - 0.5f;
}
MMNC_HD float philox_uniform_centered(uint64_t seed, uint64_t index) {
    float u[4];
    philox4_centered(seed, index >> 2, u);
    const uint32_t k = (uint32_t)index & 3u;
    return k == 0 ? u[0] : (k == 1 ? u[1] : (k == 2 ? u[2] : u[3]));
}

// tanh(x) = 1 - 2 / (e^2x + 1) in five instructions (ex2.approx + rcp.approx), ABSOLUTE error ~1.5e-7 everywhere
// (the relative error grows as 1e-7 / |x| towards 0, where tanhf switches to a polynomial at five times the cost).
// The bottleneck network only ever uses tanh additively (h + factor tanh(h), 1 - tanh^2), so the absolute error is
// what reaches the likelihood: measured <= 2e-6 relative on the likelihood (tests/test_kernel_math_hostcheck.py,
// bench.py roofline_likelihood).  The EB kernels evaluate 24 of these per element and are bound by instruction issue
// (profiles/), not by HBM.  Saturates correctly: e = inf -> 1, e = 0 -> -1; NaN propagates.
MMNC_HD float tanh_f(float x) {
    const float e = fast_exp(2.f * x);
    return 1.f - 2.f * fast_rcp(e + 1.f);
}

// ---------------------------------------------------------------------------------------------- EB (A.3)
// Packed per-channel parameter layout (58 floats), see include/mmnc_b200.h.
constexpr int EB_NP = 58;
MMNC_HD bool eb_is_matrix(int k) { return k < 3 || (k >= 9 && k < 18) || (k >= 24 && k < 33) || (k >= 39 && k < 48) || (k >= 54 && k < 57); }
MMNC_HD bool eb_is_factor(int k) { return (k >= 6 && k < 9) || (k >= 21 && k < 24) || (k >= 36 && k < 39) || (k >= 51 && k < 54); }
// raw -> what the MLP consumes: softplus(matrix), bias, tanh(factor)
MMNC_HD float eb_transform(int k, float raw) { return eb_is_matrix(k) ? softplus_t(raw) : (eb_is_factor(k) ? tanhf(raw) : raw); }
// d transformed / d raw, given raw and transformed values
MMNC_HD float eb_transform_grad(int k, float raw, float tr) {
    return eb_is_matrix(k) ? sigmoid_t(raw) : (eb_is_factor(k) ? (1.f - tr * tr) : 1.f);
}

// activations kept for the backward pass of one logits evaluation
struct EbTrace {
    float h[4][3];   // inputs of layers 1..4 (h[0] = output of layer 0's gate, ...)
    float th[4][3];  // tanh(pre-activation) of layers 0..3
};

template <bool kTrace>
MMNC_HD float eb_logits(const float *P, float t, EbTrace *tr) {
    float h[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = P[i] * t + P[3 + i];
        const float ta = tanh_f(a);
        h[i] = a + P[6 + i] * ta;
        if (kTrace) { tr->th[0][i] = ta; tr->h[0][i] = h[i]; }
    }
#pragma unroll
    for (int l = 1; l < 4; ++l) {
        const int base = 9 + (l - 1) * 15;
        float n[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float a = P[base + 3 * i] * h[0];
            a += P[base + 3 * i + 1] * h[1];
            a += P[base + 3 * i + 2] * h[2];
            a += P[base + 9 + i];
            const float ta = tanh_f(a);
            n[i] = a + P[base + 12 + i] * ta;
            if (kTrace) { tr->th[l][i] = ta; }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) { h[i] = n[i]; if (kTrace) tr->h[l][i] = n[i]; }
    }
    float f = P[54] * h[0];
    f += P[55] * h[1];
    f += P[56] * h[2];
    f += P[57];
    return f;
}

// Back-propagates dF through one traced evaluation.  Accumulates gradients w.r.t. the TRANSFORMED parameters
// into gP (58, may be null when parameters are detached) and returns dF/dt * dF.
template <bool kParamGrads>
MMNC_HD float eb_logits_backward(const float *P, float t, const EbTrace &tr, float dF, float *gP) {
    float gh[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        gh[j] = dF * P[54 + j];
        if (kParamGrads) gP[54 + j] += dF * tr.h[3][j];
    }
    if (kParamGrads) gP[57] += dF;
#pragma unroll
    for (int l = 3; l >= 1; --l) {
        const int base = 9 + (l - 1) * 15;
        float ga[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float ta = tr.th[l][i];
            ga[i] = gh[i] * (1.f + P[base + 12 + i] * (1.f - ta * ta));
            if (kParamGrads) {
                gP[base + 12 + i] += gh[i] * ta;
                gP[base + 9 + i] += ga[i];
#pragma unroll
                for (int j = 0; j < 3; ++j) gP[base + 3 * i + j] += ga[i] * tr.h[l - 1][j];
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
            gh[j] = ga[0] * P[base + j] + ga[1] * P[base + 3 + j] + ga[2] * P[base + 6 + j];
    }
    float dt = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float ta = tr.th[0][i];
        const float ga = gh[i] * (1.f + P[6 + i] * (1.f - ta * ta));
        if (kParamGrads) {
            gP[6 + i] += gh[i] * ta;
            gP[3 + i] += ga;
            gP[i] += ga * t;
        }
        dt += ga * P[i];
    }
    return dt;
}

// ---- fp32 evaluation of the forward likelihood WITHOUT the cancellation.
// lik = sigma(F(t + 1/2)) - sigma(F(t - 1/2)) (the "sign" form of CompressAI is the same number, |sigma(s up) - sigma(s lo)|).
// Instead of two independent logits, the lower logit F_l and the DIFFERENCE dF = F_u - F_l are carried through the five
// layers: every layer is monotone (softplus(matrix) >= 0, |tanh(factor)| < 1), so all differences are positive and are
// sums of positive terms - nothing cancels.  tanh(a + da) - tanh(a) = tanh(da) (1 - tanh^2 a) / (1 + tanh a tanh da)
// for small da (direct difference otherwise), and sigma(b) - sigma(a) = e^a expm1(b - a) / ((1 + e^a)(1 + e^b)) on
// the side where a + b <= 0.  Within ~2e-6 of the float64 value of the formula on the bulk (tests/hostcheck), where two
// independent fp32 evaluations differ by up to 1.5e-5; replaces a float64 evaluation that cost 24 double-precision tanh
// per element.
MMNC_HD float eb_tanh_diff(float a, float ta, float da) {
    if (da < 0.25f) {
        // tanh(da) by its Taylor series to da^9 (truncation < 1e-8 below 0.25): five multiply-adds instead of a tanhf
        const float s = da * da;
        float p = 2.186948853615520e-2f;
        p = p * s - 5.396825396825397e-2f;
        p = p * s + 1.333333333333333e-1f;
        p = p * s - 3.333333333333333e-1f;
        const float td = da + da * (p * s);
        return fast_div(td * (1.f - ta * ta), 1.f + ta * td);
    }
    return tanh_f(a + da) - ta;
}
// tl = fp32(v - 1/2) and dt = fp32(v + 1/2) - tl (an exact subtraction; 1 up to the rounding of the two sums, which the
// reference has as well)
MMNC_HD float eb_likelihood_s(const float *P, float tl, float dt) {
    float h[3], dh[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = P[i] * tl + P[3 + i];
        const float da = P[i] * dt;
        const float ta = tanh_f(a);
        h[i] = a + P[6 + i] * ta;
        dh[i] = da + P[6 + i] * eb_tanh_diff(a, ta, da);
    }
#pragma unroll
    for (int l = 1; l < 4; ++l) {
        const int base = 9 + (l - 1) * 15;
        float n[3], dn[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float a = P[base + 3 * i] * h[0];
            a += P[base + 3 * i + 1] * h[1];
            a += P[base + 3 * i + 2] * h[2];
            a += P[base + 9 + i];
            float da = P[base + 3 * i] * dh[0];
            da += P[base + 3 * i + 1] * dh[1];
            da += P[base + 3 * i + 2] * dh[2];
            const float ta = tanh_f(a);
            n[i] = a + P[base + 12 + i] * ta;
            dn[i] = da + P[base + 12 + i] * eb_tanh_diff(a, ta, da);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) { h[i] = n[i]; dh[i] = dn[i]; }
    }
    float fl = P[54] * h[0];
    fl += P[55] * h[1];
    fl += P[56] * h[2];
    fl += P[57];
    float df = P[54] * dh[0];
    df += P[55] * dh[1];
    df += P[56] * dh[2];
    // sigma(fl + df) - sigma(fl) on the side where both arguments lean negative
    float a = fl, b = fl + df;
    if (a + b > 0.f) { const float na = -b; b = -a; a = na; }
    // ea, eb <= 1 (a <= b, a + b <= 0 ... b may be positive only when |a| is larger): ex2.approx is 2 ulp here
    const float ea = fast_exp(a), eb = fast_exp(b);
    if (df > 30.f) return fast_div(eb, 1.f + eb) - fast_div(ea, 1.f + ea);  // far apart: no cancellation (expm1 would overflow)
    return fast_div(ea * expm1f(df), (1.f + ea) * (1.f + eb));
}

// likelihood from the two logits (before the floor); also the partial derivatives w.r.t. (lower, upper)
MMNC_HD float eb_likelihood(float lower, float upper, int form) {
    if (form == 0) {
        const float s = -sign_t(lower + upper);
        return fabsf(sigmoid_t(s * upper) - sigmoid_t(s * lower));
    }
    return sigmoid_t(upper) - sigmoid_t(lower);
}
MMNC_HD void eb_likelihood_grad(float lower, float upper, int form, float *d_lower, float *d_upper) {
    if (form == 0) {
        const float s = -sign_t(lower + upper);
        const float su = sigmoid_t(s * upper), sl = sigmoid_t(s * lower);
        const float sg = sign_t(su - sl);
        *d_upper = sg * s * su * (1.f - su);
        *d_lower = -sg * s * sl * (1.f - sl);
    } else {
        const float su = sigmoid_t(upper), sl = sigmoid_t(lower);
        *d_upper = su * (1.f - su);
        *d_lower = -sl * (1.f - sl);
    }
}

// LowerBound custom gradient (A.2): pass when x >= bound or the gradient is negative
MMNC_HD float lower_bound_grad(float x, float bound, float g) { return (x >= bound || g < 0.f) ? g : 0.f; }

// ---------------------------------------------------------------------------------------------- GC (A.4)
constexpr float GC_CONST = -0.70710678118654752440f;      // -(2 ** -0.5)
constexpr float INV_SQRT_2PI = 0.39894228040143267794f;

MMNC_HD float gc_std_cumulative(float t) { return 0.5f * erfcf(GC_CONST * t); }

MMNC_HD float gc_likelihood(float y_hat, float mean, float scale, float scale_bound) {
    const float sc = max_nan(scale, scale_bound);
    const float v = fabsf(y_hat - mean);
    const float upper = gc_std_cumulative((0.5f - v) / sc);
    const float lower = gc_std_cumulative((-0.5f - v) / sc);
    return upper - lower;
}
// erfc with fractional error ~1.2e-7 (Numerical Recipes' Chebyshev fit `erfcc`: erfc(z) = t exp(-z^2 + P(t)),
// t = 1 / (1 + z / 2)): one reciprocal, nine multiply-adds, one exponential - about a third of erfcf.
MMNC_HD float erfc_f(float x) {
    const float z = fabsf(x);
    const float t = fast_rcp(1.f + 0.5f * z);
    float p = 0.17087277f;
    p = p * t - 0.82215223f;
    p = p * t + 1.48851587f;
    p = p * t - 1.13520398f;
    p = p * t + 0.27886807f;
    p = p * t - 0.18628806f;
    p = p * t + 0.09678418f;
    p = p * t + 0.37409196f;
    p = p * t + 1.00002368f;
    p = p * t - 1.26551223f;
    const float ans = t * fast_exp(p - z * z);
    return x >= 0.f ? ans : 2.f - ans;  // NaN falls through as NaN
}
// fp32 evaluation WITHOUT the cancellation: lik = (1/sqrt(pi)) * integral of exp(-t^2) over [cu, cl], the two erfc
// arguments.  Narrow intervals (cl - cu < 0.5, i.e. scale > 1.41) are integrated with a 5-point Gauss-Legendre rule:
// truncation error < 2e-9 relative where lik > 1e-3 and < 1e-6 down to the 1e-9 floor, and every term is positive, so
// nothing cancels.  Wide intervals use the erfc difference, whose cancellation factor is at most ~4.5 there.  One
// reciprocal of the scale serves both arguments (the reference divides twice; 1 ulp of the argument), exponentials
// are ex2.approx on the device: measured against the float64 value of the formula (tests/test_kernel_math_hostcheck.py
// on the host, tests/test_gpu_parity.py and bench.py's roofline_likelihood on the device) this stays inside 1e-5 on the
// bulk and 5e-5 on the tails.
MMNC_HD float gc_likelihood_s(float y_hat, float mean, float scale, float scale_bound) {
    const float sc = max_nan(scale, scale_bound);
    const float v = fabsf(y_hat - mean);
    const float ninv = GC_CONST * fast_rcp(sc);             // -1 / (sc sqrt 2)
    const float cu = (0.5f - v) * ninv, cl = (-0.5f - v) * ninv;  // cu < cl
    const float d = -ninv;                                  // cl - cu exactly
    if (d < 0.5f) {
        const float h = 0.5f * d, m = cu + h;
        const float s1 = 0.5384693101056831f * h, s2 = 0.9061798459386640f * h;
        float sum = 0.5688888888888889f * fast_exp(-(m * m));
        sum += 0.4786286704993665f * (fast_exp(-((m + s1) * (m + s1))) + fast_exp(-((m - s1) * (m - s1))));
        sum += 0.2369268850561891f * (fast_exp(-((m + s2) * (m + s2))) + fast_exp(-((m - s2) * (m - s2))));
        return 0.56418958354775628695f * h * sum;  // 1 / sqrt(pi)
    }
    return 0.5f * (erfc_f(cu) - erfc_f(cl));  // also the NaN route (d is NaN when the scale is)
}
// d lik / d y_hat and d lik / d (bounded scale)
MMNC_HD void gc_likelihood_grad(float y_hat, float mean, float scale, float scale_bound, float *d_y, float *d_sc) {
    const float sc = max_nan(scale, scale_bound);
    const float d = y_hat - mean;
    const float v = fabsf(d);
    const float inv = fast_rcp(sc);
    const float tu = (0.5f - v) * inv, tl = (-0.5f - v) * inv;
    const float pu = INV_SQRT_2PI * fast_exp(-0.5f * tu * tu), pl = INV_SQRT_2PI * fast_exp(-0.5f * tl * tl);
    *d_y = sign_t(d) * (pl - pu) * inv;
    *d_sc = (tl * pl - tu * pu) * inv;
}

// ---------------------------------------------------------------------------------------------- indexes (A.4)
// idx = (n-1) - #{ t in table[0..n-2] : s <= t } with s = max(scale, bound); NaN -> n-1.
MMNC_HD int gc_scale_index(float scale, float bound, const float *table, int n) {
    const float s = max_nan(scale, bound);
    if (s != s) return n - 1;  // every comparison with NaN is false -> nothing is subtracted
    // first k in [0, n-1) with table[k] >= s  (table ascending)
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (table[mid] >= s) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------- rANS (A.7)
constexpr uint64_t RANS_L = 1ull << 31;
constexpr int RANS_PRECISION = 16;
constexpr int RANS_BYPASS_BITS = 4;
constexpr int RANS_BYPASS_MAX = 15;

// Exact division of the coder state by a symbol frequency without a divide in the sequential loop.
// For 1 < freq < 2^16 let s = ceil(log2 freq) and m = ceil(2^(63+s) / freq) (a 64-bit number, 2^63 <= m < 2^64).
// Then floor(x / freq) == mulhi64(x, m) >> (s - 1) for every x < 2^63: x m / 2^(63+s) = x / freq + x e / (freq 2^(63+s))
// with 0 <= e < freq, and the excess is below 1 / freq because x < 2^63 and freq <= 2^s.  The coder only divides states
// below x_max = 2^47 freq <= 2^63.  m is computed where it is cheap (one thread per SYMBOL, pass 1); the exhaustive
// host check over every freq lives in tests/test_kernel_math_hostcheck.py.
MMNC_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}
MMNC_HD int rans_ceil_log2(uint32_t freq) {  // freq >= 2
#if defined(__CUDA_ARCH__)
    return 32 - __clz((int)(freq - 1u));
#else
    int s = 0;
    while ((1u << s) < freq) ++s;
    return s;
#endif
}
MMNC_HD uint64_t rans_reciprocal(uint32_t freq) {  // 1 < freq < 2^16; two 64-bit divides, no 128-bit arithmetic
    const int s = rans_ceil_log2(freq);
    const uint64_t n_hi = 1ull << (31 + s);          // 2^(63+s) = n_hi * 2^32
    const uint64_t q_hi = n_hi / freq, r1 = n_hi - q_hi * freq;
    const uint64_t n_lo = r1 << 32;                  // r1 < freq < 2^16
    const uint64_t q_lo = n_lo / freq, r2 = n_lo - q_lo * freq;
    return (q_hi << 32) + q_lo + (r2 != 0 ? 1ull : 0ull);
}
MMNC_HD uint64_t rans_div(uint64_t x, uint32_t freq, uint64_t m) {  // floor(x / freq), x < 2^63
    return freq == 1u ? x : (mulhi64(x, m) >> (rans_ceil_log2(freq) - 1));
}

struct RansEnc {
    uint64_t x;
    uint32_t *ptr;  // write pointer, moves backwards
    MMNC_HD void init(uint32_t *end) { x = RANS_L; ptr = end; }
    MMNC_HD void put(uint32_t start, uint32_t freq) {
        const uint64_t x_max = ((RANS_L >> RANS_PRECISION) << 32) * freq;
        if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
        x = ((x / freq) << RANS_PRECISION) + (x % freq) + start;
    }
    // same update with the precomputed reciprocal of freq (rans_reciprocal): the kernels' sequential pass uses this
    MMNC_HD void put_rcp(uint32_t start, uint32_t freq, uint64_t m) {
        const uint64_t x_max = ((RANS_L >> RANS_PRECISION) << 32) * freq;
        if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
        const uint64_t q = rans_div(x, freq, m);
        x = (q << RANS_PRECISION) + (x - q * freq) + start;
    }
    MMNC_HD void put_bits(uint32_t val) {
        const uint64_t x_max = ((RANS_L >> 16) << 32) * (uint64_t)(1u << (16 - RANS_BYPASS_BITS));
        if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
        x = (x << RANS_BYPASS_BITS) | val;
    }
    MMNC_HD void flush() { ptr -= 2; ptr[0] = (uint32_t)x; ptr[1] = (uint32_t)(x >> 32); }
};

// number of 4-bit groups needed for raw (0 for raw == 0)
MMNC_HD int rans_nibbles(uint32_t raw) {
    int n = 0;
    while (n < 8 && (raw >> (n * RANS_BYPASS_BITS)) != 0) ++n;
    return n;
}

// Encodes the escape tail of one symbol (count digits + nibbles) — called BEFORE put() of the symbol itself
// because the coder consumes pushes in reverse order.
MMNC_HD void rans_put_escape_reversed(RansEnc &e, uint32_t raw) {
    const int nb = rans_nibbles(raw);
    for (int j = nb - 1; j >= 0; --j) e.put_bits((raw >> (j * RANS_BYPASS_BITS)) & RANS_BYPASS_MAX);
    // count digits were pushed as: 15, 15, ..., rest ; reversed: rest first
    // (nb <= 8 < 15, so there is exactly one digit; kept general)
    int full = nb / RANS_BYPASS_MAX, rest = nb - full * RANS_BYPASS_MAX;
    e.put_bits((uint32_t)rest);
    for (int k = 0; k < full; ++k) e.put_bits(RANS_BYPASS_MAX);
}

// symbol -> (cdf slot, raw escape payload).  Returns the slot; *raw is meaningful only when slot == max_value.
MMNC_HD int rans_map_symbol(int32_t symbol, int32_t offset, int32_t max_value, uint32_t *raw) {
    int32_t value = symbol - offset;
    *raw = 0;
    if (value < 0) { *raw = (uint32_t)(-2 * value - 1); value = max_value; }
    else if (value >= max_value) { *raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
    return value;
}

struct RansDec {
    uint64_t x;
    const uint8_t *p;    // read cursor (bytes; streams need not be 4-byte aligned)
    const uint8_t *end;
    bool overrun;
    MMNC_HD uint32_t word() {
        if (p + 4 > end) { overrun = true; return 0; }
        const uint32_t w = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        p += 4;
        return w;
    }
    MMNC_HD void init(const uint8_t *begin, const uint8_t *end_) {
        p = begin; end = end_; overrun = false;
        const uint64_t lo = word();
        const uint64_t hi = word();
        x = lo | (hi << 32);
    }
    MMNC_HD uint32_t peek() const { return (uint32_t)(x & 0xFFFFu); }
    MMNC_HD void advance(uint32_t start, uint32_t freq) {
        x = (uint64_t)freq * (x >> RANS_PRECISION) + (x & 0xFFFFu) - start;
        if (x < RANS_L) x = (x << 32) | word();
    }
    MMNC_HD uint32_t get_bits() {
        const uint32_t val = (uint32_t)(x & RANS_BYPASS_MAX);
        x >>= RANS_BYPASS_BITS;
        if (x < RANS_L) x = (x << 32) | word();
        return val;
    }
    MMNC_HD int32_t get_escape(int32_t max_value) {
        int32_t val = (int32_t)get_bits();
        int32_t nb = val;
        while (val == RANS_BYPASS_MAX && !overrun) { val = (int32_t)get_bits(); nb += val; }
        uint32_t raw = 0;
        for (int j = 0; j < nb && j < 8; ++j) raw |= get_bits() << (j * RANS_BYPASS_BITS);
        int32_t value = (int32_t)(raw >> 1);
        return (raw & 1u) ? (-value - 1) : (value + max_value);
    }
};

// Ragged 16-bit tables: a row of `len` CDF entries is stored as uint16 (the final entry 65536 wraps to 0 and is never
// compared).  Slot = last s in [0, len - 2] with row[s] <= cum.
MMNC_HD int rans_find_slot_u16(const uint16_t *row, int len, uint32_t cum) {
    int lo = 0, hi = len - 1;  // first index in [0, len - 1) with row > cum; row[len - 1] stands for 65536 > cum
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((uint32_t)row[mid] > cum) hi = mid; else lo = mid + 1;
    }
    return lo - 1;
}

// s = (first position in cdf[0..len) with value > cum) - 1, by binary search (cdf strictly increasing)
MMNC_HD int rans_find_slot(const int32_t *cdf, int len, uint32_t cum) {
    int lo = 0, hi = len;  // first index with cdf > cum
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((uint32_t)cdf[mid] > cum) hi = mid; else lo = mid + 1;
    }
    return lo - 1;
}

}  // namespace mmnc
