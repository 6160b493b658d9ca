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

// ragged uint16 view of a (n_cdfs x stride) int32 table, as rans_pack_tables_kernel builds it
static void hc_pack(const int32_t *cdf, int n_cdfs, int stride, const int32_t *sizes, std::vector<int32_t> &row_start,
                    std::vector<uint16_t> &ragged) {
    row_start.assign((size_t)n_cdfs + 1, 0);
    for (int r = 0; r < n_cdfs; ++r) row_start[r + 1] = row_start[r] + (sizes[r] < 0 ? 0 : (sizes[r] > stride ? stride : sizes[r]));
    ragged.assign((size_t)row_start[n_cdfs] + 8, 0);
    for (int r = 0; r < n_cdfs; ++r)
        for (int c = 0; c < sizes[r] && c < stride; ++c)
            ragged[row_start[r] + c] = (uint16_t)((uint32_t)cdf[(int64_t)r * stride + c] & 0xFFFFu);
}

// one stream through the same map + reverse-encode logic as rans_map_kernel / rans_encode_kernel (ragged 16-bit
// table, reciprocal multiply instead of divide, escape payload re-derived from the symbol)
int64_t hc_rans_encode(const int32_t *symbols, const int32_t *indexes, int64_t n, const int32_t *cdf, int n_cdfs,
                       int stride, const int32_t *sizes, const int32_t *offsets, uint8_t *out, int64_t cap_words) {
    std::vector<int32_t> row_start;
    std::vector<uint16_t> ragged;
    hc_pack(cdf, n_cdfs, stride, sizes, row_start, ragged);
    std::vector<uint32_t> slab((size_t)cap_words);
    std::vector<uint32_t> st((size_t)n);
    std::vector<uint64_t> rcp((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const int ci = indexes[i];
        if (ci < 0 || ci >= n_cdfs) return -1;
        const int maxv = sizes[ci] - 2;
        uint32_t r;
        const int slot = rans_map_symbol(symbols[i], offsets[ci], maxv, &r);
        const uint16_t *row = ragged.data() + row_start[ci];
        const uint32_t start = row[slot], range = ((uint32_t)row[slot + 1] - start) & 0xFFFFu;
        st[i] = start | (range << 16);
        rcp[i] = range > 1u ? rans_reciprocal(range) : 0ull;
    }
    RansEnc enc;
    enc.init(slab.data() + cap_words);
    for (int64_t i = n - 1; i >= 0; --i) {
        const uint32_t start = st[i] & 0xFFFFu, range = st[i] >> 16;
        if (range == 0 || enc.ptr - slab.data() < 16) return -2;
        if (start + range == 65536u) {
            uint32_t raw;
            rans_map_symbol(symbols[i], offsets[indexes[i]], sizes[indexes[i]] - 2, &raw);
            rans_put_escape_reversed(enc, raw);
        }
        enc.put_rcp(start, range, rcp[i]);
    }
    enc.flush();
    const int64_t nbytes = (slab.data() + cap_words - enc.ptr) * 4;
    memcpy(out, enc.ptr, (size_t)nbytes);
    return nbytes;
}

int hc_rans_decode(const uint8_t *bytes, int64_t nbytes, const int32_t *indexes, int64_t n, const int32_t *cdf,
                   int n_cdfs, int stride, const int32_t *sizes, const int32_t *offsets, int32_t *out) {
    if (nbytes < 8) return -1;
    std::vector<int32_t> row_start;
    std::vector<uint16_t> ragged;
    hc_pack(cdf, n_cdfs, stride, sizes, row_start, ragged);
    RansDec dec;
    dec.init(bytes, bytes + nbytes);
    for (int64_t i = 0; i < n; ++i) {
        const int ci = indexes[i];
        if (ci < 0 || ci >= n_cdfs) return -3;
        const uint16_t *row = ragged.data() + row_start[ci];
        const int len = sizes[ci], maxv = len - 2;
        const int slot = rans_find_slot_u16(row, len, dec.peek());
        if (slot < 0 || slot > maxv) return -4;
        const uint32_t start = row[slot];
        dec.advance(start, ((uint32_t)row[slot + 1] - start) & 0xFFFFu);
        int32_t value = slot;
        if (slot == maxv) value = dec.get_escape(maxv);
        out[i] = value + offsets[ci];
        if (dec.overrun) return -2;
    }
    return 0;
}

// Exhaustive check of the reciprocal division over EVERY frequency the 16-bit coder can see: for each freq the
// quotient is compared with the hardware divide at the states where an off-by-one would show first (multiples of freq
// and their neighbours, up to the largest state the coder ever divides, 2^47 freq - 1, and up to 2^63 - 1) plus
// `n_random` pseudo-random states.  Returns the number of mismatches (0 = exact).
int64_t hc_rans_reciprocal_check(int n_random) {
    int64_t bad = 0;
    uint64_t lcg = 0x9E3779B97F4A7C15ull;
    for (uint32_t freq = 1; freq < 65536u; ++freq) {
        const uint64_t m = freq > 1u ? rans_reciprocal(freq) : 0ull;
        const uint64_t x_max = ((uint64_t)1 << 47) * freq - 1;
        const uint64_t top = ((uint64_t)1 << 63) - 1;
        const uint64_t probes[] = {0, 1, freq - 1, freq, (uint64_t)freq + 1, x_max, x_max - 1, x_max + 1 - freq, x_max - freq,
                                   top, top - 1, top / freq * freq, top / freq * freq - 1, ((uint64_t)1 << 31),
                                   ((uint64_t)1 << 32) - 1, ((uint64_t)1 << 32), ((uint64_t)1 << 48) - 1};
        for (uint64_t x : probes)
            if (x <= top && rans_div(x, freq, m) != x / freq) ++bad;
        for (int i = 0; i < n_random; ++i) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t k = (lcg >> 1) / freq;  // a quotient below 2^63 / freq
            const uint64_t base = k * freq;        // multiple of freq: the boundary where floor() steps
            for (uint64_t x : {base, base - (base ? 1 : 0), base + freq - 1})
                if (x <= top && rans_div(x, freq, m) != x / freq) ++bad;
        }
    }
    return bad;
}
}
