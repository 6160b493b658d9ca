// Channels-last (NHWC) row access shared by the tensor-core GDN kernels that take either memory format.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmnc {

// Channels-last (NHWC) access: a pixel's C channels are one contiguous row, so a thread moves its row with 8- or
// 16-byte vector accesses (`vec` = 4, 2 or 1 floats, chosen on the host from C and the base alignment).  Channels >= C
// read as 0 and are never written.
template <int N>
__device__ __forceinline__ void nhwc_load_row(const float *__restrict__ row, int C, int vec, bool ok, float (&v)[N]) {
#pragma unroll
    for (int c = 0; c < N; ++c) v[c] = 0.f;
    if (!ok) return;
    if (vec == 4) {
#pragma unroll
        for (int c = 0; c < N; c += 4)
            if (c + 4 <= C) {
                const float4 t = __ldcs(reinterpret_cast<const float4 *>(row + c));
                v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
            }
    } else if (vec == 2) {
#pragma unroll
        for (int c = 0; c < N; c += 2)
            if (c + 2 <= C) {
                const float2 t = __ldcs(reinterpret_cast<const float2 *>(row + c));
                v[c] = t.x; v[c + 1] = t.y;
            }
    } else {
#pragma unroll
        for (int c = 0; c < N; ++c)
            if (c < C) v[c] = __ldcs(row + c);
    }
}
template <int N>
__device__ __forceinline__ void nhwc_store_block(float *__restrict__ row, int c0, int C, int vec, const float (&v)[N]) {
    if (vec == 4) {
#pragma unroll
        for (int c = 0; c < N; c += 4)
            if (c0 + c + 4 <= C) __stcs(reinterpret_cast<float4 *>(row + c0 + c), make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
    } else if (vec == 2) {
#pragma unroll
        for (int c = 0; c < N; c += 2)
            if (c0 + c + 2 <= C) __stcs(reinterpret_cast<float2 *>(row + c0 + c), make_float2(v[c], v[c + 1]));
    } else {
#pragma unroll
        for (int c = 0; c < N; ++c)
            if (c0 + c < C) __stcs(row + c0 + c, v[c]);
    }
}


// vector width of a channels-last row access: the row stride is C floats, so C and the bases must allow it
int gdn_nhwc_vec(const void *a, const void *b, const void *c, int64_t C);

}  // namespace mmnc
