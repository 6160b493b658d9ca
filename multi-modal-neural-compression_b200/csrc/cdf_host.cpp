// (a9) host-side integer CDF construction: the product's own equivalent of CompressAI's
// `pmf_to_quantized_cdf` (C++/pybind11 there; plain C ABI here).  Off the per-step path — it runs once per table
// row when `update()` is called after loading weights (/root/reference/src/models/multi_task_compressor.py:486-489)
// — but its integers decide every bitstream, so the arithmetic follows the published algorithm exactly:
// scale by 2^precision and round half away from zero, renormalise with 64-bit integer division by the 32-bit
// total, prefix-sum, pin the last entry, then repair zero-width bins by stealing from the narrowest bin wider
// than one count (first such bin on ties).
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../../include/mmnc_b200.h"

namespace mmnc { void set_error(const char *fmt, ...); }

extern "C" int mmnc_pmf_to_quantized_cdf_h(const float *pmf_h, int n, int precision, uint32_t *cdf_h) {
    if (pmf_h == nullptr || cdf_h == nullptr || n < 1 || precision < 1 || precision > 16) {
        mmnc::set_error("pmf_to_quantized_cdf: invalid arguments (n=%d precision=%d)", n, precision);
        return MMNC_ERR_INVALID;
    }
    for (int i = 0; i < n; ++i) {
        if (!(pmf_h[i] >= 0.0f) || !std::isfinite(pmf_h[i])) {
            mmnc::set_error("pmf_to_quantized_cdf: pmf[%d] is negative or not finite", i);
            return MMNC_ERR_INVALID;
        }
    }
    const float one = static_cast<float>(1 << precision);
    std::vector<uint32_t> freq(static_cast<size_t>(n) + 1, 0u);
    for (int i = 0; i < n; ++i) freq[i + 1] = static_cast<uint32_t>(std::round(pmf_h[i] * one));
    const uint32_t total = static_cast<uint32_t>(std::accumulate(freq.begin(), freq.end(), 0));
    if (total == 0u) {
        mmnc::set_error("pmf_to_quantized_cdf: pmf sums to zero");
        return MMNC_ERR_INVALID;
    }
    uint32_t running = 0u;
    for (size_t i = 0; i < freq.size(); ++i) {
        running += static_cast<uint32_t>((static_cast<uint64_t>(1u << precision) * freq[i]) / total);
        cdf_h[i] = running;
    }
    cdf_h[n] = 1u << precision;
    const int bins = n;  // cdf has n + 1 entries -> n bins
    for (int i = 0; i < bins; ++i) {
        if (cdf_h[i] != cdf_h[i + 1]) continue;
        int donor = -1;
        uint32_t donor_width = ~0u;
        for (int j = 0; j < bins; ++j) {
            const uint32_t w = cdf_h[j + 1] - cdf_h[j];
            if (w > 1u && w < donor_width) { donor_width = w; donor = j; }
        }
        if (donor < 0) {
            mmnc::set_error("pmf_to_quantized_cdf: no bin to steal from (n=%d too large for precision %d)", n, precision);
            return MMNC_ERR_INVALID;
        }
        if (donor < i) { for (int j = donor + 1; j <= i; ++j) cdf_h[j] -= 1u; }
        else           { for (int j = i + 1; j <= donor; ++j) cdf_h[j] += 1u; }
    }
    return MMNC_OK;
}
