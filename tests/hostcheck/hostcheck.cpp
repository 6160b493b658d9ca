// TEST INFRASTRUCTURE ONLY.  Compiles the host+device inline arithmetic of the CUDA kernels
// (multi-modal-neural-compression_b200/csrc/hd_math.cuh) with g++ so the `-m "not gpu"` suite can check the
// kernels' per-element logic (EB MLP forward/backward, GC likelihood, scale index search, rANS state machine)
// against the oracle on a machine without a GPU.  This object is never linked into libmmnc_b200.so and nothing
// in the product imports it: it is a checker, not a CPU fallback.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../multi-modal-neural-compression_b200/csrc/hd_math.cuh"

using namespace mmnc;

extern "C" {

// x: n values of channel-major layout (C, L); params (C, 58) raw.  Mirrors eb_forward_kernel's element body.
void hc_eb_forward(const float *v, int C, int L, const float *params, float bound, int form, float *lik,
                   float *lower_out, float *upper_out) {
    for (int c = 0; c < C; ++c) {
        float P[EB_NP];
        for (int k = 0; k < EB_NP; ++k) P[k] = eb_transform(k, params[c * EB_NP + k]);
        for (int i = 0; i < L; ++i) {
            const float t = v[c * L + i];
            const float tl = t - 0.5f, tu = t + 0.5f;
            float l = eb_likelihood_s(P, tl, tu - tl);  // what the GPU forward evaluates (both forms are this number)
            (void)form;
            if (bound > 0.f) l = max_nan(l, bound);
            lik[c * L + i] = l;
            if (lower_out) lower_out[c * L + i] = eb_logits<false>(P, tl, nullptr);
            if (upper_out) upper_out[c * L + i] = eb_logits<false>(P, tu, nullptr);
        }
    }
}

// gradient of sum_i g_lik[i] * lik[i] w.r.t. v (C, L) and raw params (C, 58); mirrors eb_backward_kernel.
void hc_eb_backward(const float *v, int C, int L, const float *params, const float *g_lik, float bound, int form,
                    float *g_v, float *g_params) {
    for (int c = 0; c < C; ++c) {
        float P[EB_NP], gP[EB_NP];
        for (int k = 0; k < EB_NP; ++k) { P[k] = eb_transform(k, params[c * EB_NP + k]); gP[k] = 0.f; }
        for (int i = 0; i < L; ++i) {
            const float t = v[c * L + i];
            EbTrace tl, tu;
            const float lo = eb_logits<true>(P, t - 0.5f, &tl), up = eb_logits<true>(P, t + 0.5f, &tu);
            const float raw = eb_likelihood(lo, up, form);
            float g = g_lik[c * L + i];
            if (bound > 0.f) g = lower_bound_grad(raw, bound, g);
            float dl, du;
            eb_likelihood_grad(lo, up, form, &dl, &du);
            float gv = eb_logits_backward<true>(P, t - 0.5f, tl, g * dl, gP);
            gv += eb_logits_backward<true>(P, t + 0.5f, tu, g * du, gP);
            g_v[c * L + i] = gv;
        }
        for (int k = 0; k < EB_NP; ++k)
            g_params[c * EB_NP + k] = gP[k] * eb_transform_grad(k, params[c * EB_NP + k], P[k]);
    }
}

void hc_gc_forward(const float *y_hat, const float *scales, int64_t n, float scale_bound, float lik_bound,
                   float *lik, float *d_y, float *d_sc) {
    for (int64_t i = 0; i < n; ++i) {
        float l = gc_likelihood_s(y_hat[i], 0.f, scales[i], scale_bound);  // what the GPU kernels evaluate
        if (lik_bound > 0.f) l = max_nan(l, lik_bound);
        lik[i] = l;
        gc_likelihood_grad(y_hat[i], 0.f, scales[i], scale_bound, &d_y[i], &d_sc[i]);
    }
}

// the stable fp32 EB likelihood the GPU forward uses (eb_likelihood_s), before the floor; params (C, 58) raw
void hc_eb_likelihood_s(const float *v, int C, int L, const float *params, float *lik) {
    for (int c = 0; c < C; ++c) {
        float P[EB_NP];
        for (int k = 0; k < EB_NP; ++k) P[k] = eb_transform(k, params[c * EB_NP + k]);
        for (int i = 0; i < L; ++i) {
            const float tl = v[c * L + i] - 0.5f, tu = v[c * L + i] + 0.5f;
            lik[c * L + i] = eb_likelihood_s(P, tl, tu - tl);
        }
    }
}

// the stable fp32 likelihood the GPU forward uses (gc_likelihood_s), before the floor
void hc_gc_likelihood_s(const float *y_hat, const float *scales, int64_t n, float scale_bound, float *lik) {
    for (int64_t i = 0; i < n; ++i) lik[i] = gc_likelihood_s(y_hat[i], 0.f, scales[i], scale_bound);
}

void hc_scale_index(const float *scales, int64_t n, const float *table, int len, float bound, int32_t *idx) {
    for (int64_t i = 0; i < n; ++i) idx[i] = gc_scale_index(scales[i], bound, table, len);
}

void hc_philox(uint64_t seed, uint64_t offset, int64_t n, float *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = philox_uniform_centered(seed, (uint64_t)i + offset);
}

// one stream through the same map + reverse-encode logic as rans_map_kernel / rans_encode_kernel
int64_t hc_rans_encode(const int32_t *symbols, const int32_t *indexes, int64_t n, const int32_t *cdf, int n_cdfs,
                       int stride, const int32_t *sizes, const int32_t *offsets, uint8_t *out, int64_t cap_words) {
    std::vector<uint32_t> slab((size_t)cap_words);
    std::vector<uint32_t> st((size_t)n), raw((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const int ci = indexes[i];
        if (ci < 0 || ci >= n_cdfs) return -1;
        const int maxv = sizes[ci] - 2;
        uint32_t r;
        const int slot = rans_map_symbol(symbols[i], offsets[ci], maxv, &r);
        const int32_t *row = cdf + (int64_t)ci * stride;
        st[i] = ((uint32_t)row[slot] & 0xFFFFu) | ((uint32_t)(row[slot + 1] - row[slot]) << 16);
        raw[i] = r;
    }
    RansEnc enc;
    enc.init(slab.data() + cap_words);
    for (int64_t i = n - 1; i >= 0; --i) {
        const uint32_t start = st[i] & 0xFFFFu, range = st[i] >> 16;
        if (range == 0 || enc.ptr - slab.data() < 16) return -2;
        if (start + range == 65536u) rans_put_escape_reversed(enc, raw[i]);
        enc.put(start, range);
    }
    enc.flush();
    const int64_t nbytes = (slab.data() + cap_words - enc.ptr) * 4;
    memcpy(out, enc.ptr, (size_t)nbytes);
    return nbytes;
}

int hc_rans_decode(const uint8_t *bytes, int64_t nbytes, const int32_t *indexes, int64_t n, const int32_t *cdf,
                   int n_cdfs, int stride, const int32_t *sizes, const int32_t *offsets, int32_t *out) {
    if (nbytes < 8) return -1;
    RansDec dec;
    dec.init(bytes, bytes + nbytes);
    for (int64_t i = 0; i < n; ++i) {
        const int ci = indexes[i];
        if (ci < 0 || ci >= n_cdfs) return -3;
        const int32_t *row = cdf + (int64_t)ci * stride;
        const int len = sizes[ci], maxv = len - 2;
        const int slot = rans_find_slot(row, len, dec.peek());
        if (slot < 0 || slot > maxv) return -4;
        dec.advance((uint32_t)row[slot], (uint32_t)(row[slot + 1] - row[slot]));
        int32_t value = slot;
        if (slot == maxv) value = dec.get_escape(maxv);
        out[i] = value + offsets[ci];
        if (dec.overrun) return -2;
    }
    return 0;
}
}
