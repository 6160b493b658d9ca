/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (multi-modal-neural-compression_b200/).  Allowed callers: tests/, __graft_entry__.smoke(),
 * bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the arithmetic restated here lives in the third-party dependency
 * `compressai==1.2.4` (pinned at /root/reference/requirements.txt:10), which is NOT vendored under
 * /root/reference, is not installed in this image and cannot be fetched (no network).  The reference
 * holds no golden vectors or tests for this path.  What follows is a plain-C
 * restatement of the *published* CompressAI 1.2.4 algorithm:
 *
 *   - compressai/cpp_exts/ops/ops.cpp            :: pmf_to_quantized_cdf
 *   - compressai/cpp_exts/rans/rans_interface.cpp :: BufferedRansEncoder::encode_with_indexes / flush,
 *                                                    RansDecoder::decode_with_indexes
 *   - third_party/ryg_rans/rans64.h              :: Rans64EncPut / Rans64EncFlush / Rans64DecInit /
 *                                                    Rans64DecAdvance (64-bit state, 32-bit words)
 *
 * anchored on the reference's own call sites:
 *   /root/reference/src/models/multi_task_compressor.py:509 (ScaleHyperprior.compress -> encode_with_indexes)
 *   /root/reference/src/models/multi_task_compressor.py:543,546 (EB/GC.decompress -> decode_with_indexes)
 *   /root/reference/src/models/multi_task_compressor.py:486-489 (update -> pmf_to_quantized_cdf)
 *
 * Plain C, scalar, single-threaded (exactly like the pybind11 original, which holds the GIL).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PRECISION 16
#define ORC_BYPASS_BITS 4
#define ORC_BYPASS_MAX ((1 << ORC_BYPASS_BITS) - 1)
#define ORC_RANS_L (1ull << 31)

/* ------------------------------------------------------------------------------------------------
 * pmf -> strictly increasing integer CDF with 2^precision total (ops.cpp::pmf_to_quantized_cdf).
 * pmf: n floats.  cdf_out: n + 1 uint32.  Returns 0, or -1 when no frequency can be stolen.
 * ---------------------------------------------------------------------------------------------- */
int orc_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf_out) {
    const int m = n + 1;
    uint32_t *cdf = cdf_out;
    cdf[0] = 0;
    for (int i = 0; i < n; ++i) {
        /* float * int -> float; std::round (half away from zero) on float; then uint32 conversion */
        float scaled = pmf[i] * (float)(1 << precision);
        cdf[i + 1] = (uint32_t)roundf(scaled);
    }
    /* std::accumulate(..., 0): the accumulator is a (32-bit) int */
    int32_t total_i = 0;
    for (int i = 0; i < m; ++i) total_i = (int32_t)((uint32_t)total_i + cdf[i]);
    const uint32_t total = (uint32_t)total_i;
    if (total == 0) return -2;
    for (int i = 0; i < m; ++i) {
        cdf[i] = (uint32_t)((((uint64_t)(1 << precision)) * cdf[i]) / total);
    }
    for (int i = 1; i < m; ++i) cdf[i] += cdf[i - 1]; /* partial_sum */
    cdf[m - 1] = 1u << precision;

    for (int i = 0; i < m - 1; ++i) {
        if (cdf[i] == cdf[i + 1]) {
            /* steal one count from the symbol with the smallest frequency > 1 (first one on ties) */
            uint32_t best_freq = ~0u;
            int best = -1;
            for (int j = 0; j < m - 1; ++j) {
                uint32_t f = cdf[j + 1] - cdf[j];
                if (f > 1 && f < best_freq) {
                    best_freq = f;
                    best = j;
                }
            }
            if (best < 0) return -1;
            if (best < i) {
                for (int j = best + 1; j <= i; ++j) cdf[j]--;
            } else {
                for (int j = i + 1; j <= best; ++j) cdf[j]++;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * rANS encoder.  `cdfs` is the dense (n_cdfs x cdf_stride) int32 table exactly as CompressAI keeps
 * `_quantized_cdf`; `cdfs_sizes` = `_cdf_length`; `offsets` = `_offset`.
 * Output: native-endian bytes, returned through a malloc'ed buffer the caller frees with orc_free.
 * Returns the byte count (>= 8) or a negative error.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint16_t start;
    uint16_t range;
    uint8_t bypass;
} orc_sym_t;

static inline void enc_put(uint64_t *r, uint32_t **pptr, uint32_t start, uint32_t freq, uint32_t scale_bits) {
    uint64_t x = *r;
    uint64_t x_max = ((ORC_RANS_L >> scale_bits) << 32) * freq;
    if (x >= x_max) {
        *pptr -= 1;
        **pptr = (uint32_t)x;
        x >>= 32;
    }
    *r = ((x / freq) << scale_bits) + (x % freq) + start;
}

static inline void enc_put_bits(uint64_t *r, uint32_t **pptr, uint32_t val, uint32_t nbits) {
    uint64_t x = *r;
    uint32_t freq = 1u << (16 - nbits);
    uint64_t x_max = ((ORC_RANS_L >> 16) << 32) * freq;
    if (x >= x_max) {
        *pptr -= 1;
        **pptr = (uint32_t)x;
        x >>= 32;
    }
    *r = (x << nbits) | val;
}

int64_t orc_rans_encode_with_indexes(const int32_t *symbols, const int32_t *indexes, int64_t n,
                                     const int32_t *cdfs, int n_cdfs, int cdf_stride,
                                     const int32_t *cdfs_sizes, const int32_t *offsets,
                                     uint8_t **out_bytes) {
    /* worst case pushes per symbol: 1 + (1 + 8) bypass nibbles for a 32-bit raw value + count digits */
    int64_t cap = n * 2 + 16;
    orc_sym_t *syms = (orc_sym_t *)malloc((size_t)cap * sizeof(orc_sym_t));
    int64_t ns = 0;
    if (!syms) return -1;
#define PUSH(s_, r_, b_)                                                         \
    do {                                                                         \
        if (ns == cap) {                                                         \
            cap *= 2;                                                            \
            syms = (orc_sym_t *)realloc(syms, (size_t)cap * sizeof(orc_sym_t));  \
            if (!syms) return -1;                                                \
        }                                                                        \
        syms[ns].start = (uint16_t)(s_);                                         \
        syms[ns].range = (uint16_t)(r_);                                         \
        syms[ns].bypass = (uint8_t)(b_);                                         \
        ++ns;                                                                    \
    } while (0)

    for (int64_t i = 0; i < n; ++i) {
        const int32_t ci = indexes[i];
        if (ci < 0 || ci >= n_cdfs) {
            free(syms);
            return -3;
        }
        const int32_t *cdf = cdfs + (int64_t)ci * cdf_stride;
        const int32_t max_value = cdfs_sizes[ci] - 2;
        int32_t value = symbols[i] - offsets[ci];
        uint32_t raw = 0;
        if (value < 0) {
            raw = (uint32_t)(-2 * value - 1);
            value = max_value;
        } else if (value >= max_value) {
            raw = (uint32_t)(2 * (value - max_value));
            value = max_value;
        }
        PUSH(cdf[value], cdf[value + 1] - cdf[value], 0);
        if (value == max_value) {
            int32_t n_bypass = 0;
            while (n_bypass * ORC_BYPASS_BITS < 32 && (raw >> (n_bypass * ORC_BYPASS_BITS)) != 0) ++n_bypass;
            int32_t val = n_bypass;
            while (val >= ORC_BYPASS_MAX) {
                PUSH(ORC_BYPASS_MAX, ORC_BYPASS_MAX + 1, 1);
                val -= ORC_BYPASS_MAX;
            }
            PUSH(val, val + 1, 1);
            for (int32_t j = 0; j < n_bypass; ++j) {
                int32_t nib = (int32_t)((raw >> (j * ORC_BYPASS_BITS)) & ORC_BYPASS_MAX);
                PUSH(nib, nib + 1, 1);
            }
        }
    }
#undef PUSH
    /* flush(): original sizes the word buffer at #pushes (UB below 2 pushes); we add the 2 flush words */
    int64_t nwords = ns + 2;
    uint32_t *words = (uint32_t *)malloc((size_t)nwords * sizeof(uint32_t));
    if (!words) {
        free(syms);
        return -1;
    }
    uint32_t *ptr = words + nwords;
    uint64_t x = ORC_RANS_L;
    for (int64_t k = ns - 1; k >= 0; --k) {
        if (!syms[k].bypass)
            enc_put(&x, &ptr, syms[k].start, syms[k].range, ORC_PRECISION);
        else
            enc_put_bits(&x, &ptr, syms[k].start, ORC_BYPASS_BITS);
    }
    ptr -= 2;
    ptr[0] = (uint32_t)(x >> 0);
    ptr[1] = (uint32_t)(x >> 32);
    int64_t nbytes = (int64_t)((words + nwords) - ptr) * (int64_t)sizeof(uint32_t);
    uint8_t *out = (uint8_t *)malloc((size_t)nbytes);
    if (!out) {
        free(syms);
        free(words);
        return -1;
    }
    memcpy(out, ptr, (size_t)nbytes);
    *out_bytes = out;
    free(words);
    free(syms);
    return nbytes;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------------
 * rANS decoder (RansDecoder::decode_with_indexes): linear CDF scan, bypass un-escape.
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t dec_get_bits(uint64_t *r, const uint32_t **pptr, uint32_t nbits) {
    uint64_t x = *r;
    uint32_t val = (uint32_t)(x & ((1u << nbits) - 1));
    x >>= nbits;
    if (x < ORC_RANS_L) {
        x = (x << 32) | **pptr;
        *pptr += 1;
    }
    *r = x;
    return val;
}

int orc_rans_decode_with_indexes(const uint8_t *encoded, int64_t nbytes, const int32_t *indexes, int64_t n,
                                 const int32_t *cdfs, int n_cdfs, int cdf_stride,
                                 const int32_t *cdfs_sizes, const int32_t *offsets, int32_t *out) {
    (void)nbytes;
    const uint32_t *ptr = (const uint32_t *)encoded;
    uint64_t x = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
    ptr += 2;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t ci = indexes[i];
        if (ci < 0 || ci >= n_cdfs) return -3;
        const int32_t *cdf = cdfs + (int64_t)ci * cdf_stride;
        const int32_t max_value = cdfs_sizes[ci] - 2;
        const int32_t offset = offsets[ci];
        const uint32_t cum = (uint32_t)(x & ((1u << ORC_PRECISION) - 1));
        int32_t k = 0;
        const int32_t len = cdfs_sizes[ci];
        while (k < len && !((uint32_t)cdf[k] > cum)) ++k; /* std::find_if(v > cum_freq) */
        const int32_t s = k - 1;
        const uint32_t start = (uint32_t)cdf[s];
        const uint32_t freq = (uint32_t)(cdf[s + 1] - cdf[s]);
        x = (uint64_t)freq * (x >> ORC_PRECISION) + (x & ((1ull << ORC_PRECISION) - 1)) - start;
        if (x < ORC_RANS_L) {
            x = (x << 32) | *ptr;
            ptr += 1;
        }
        int32_t value = s;
        if (value == max_value) {
            int32_t val = (int32_t)dec_get_bits(&x, &ptr, ORC_BYPASS_BITS);
            int32_t n_bypass = val;
            while (val == ORC_BYPASS_MAX) {
                val = (int32_t)dec_get_bits(&x, &ptr, ORC_BYPASS_BITS);
                n_bypass += val;
            }
            uint32_t raw = 0;
            for (int32_t j = 0; j < n_bypass; ++j) {
                val = (int32_t)dec_get_bits(&x, &ptr, ORC_BYPASS_BITS);
                raw |= (uint32_t)val << (j * ORC_BYPASS_BITS);
            }
            value = (int32_t)(raw >> 1);
            if (raw & 1u)
                value = -value - 1;
            else
                value += max_value;
        }
        out[i] = value + offset;
    }
    return 0;
}
