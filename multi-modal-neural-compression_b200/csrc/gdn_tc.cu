// K4 (tensor-core path): GDN / IGDN channel contraction on tcgen05 (placeholder until the kernel lands).
#include "common.cuh"

namespace mmnc {

bool gdn_tc_supported(int64_t, int64_t, int64_t) { return false; }

int gdn_tc_forward(const float *, int64_t, int64_t, int64_t, const float *, const float *, int, int, float *,
                   cudaStream_t) {
    set_error("gdn_tc_forward: not built");
    return MMNC_ERR_UNSUPPORTED;
}

}  // namespace mmnc
