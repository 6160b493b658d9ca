// Shared host-side plumbing for the mmnc_b200 C-ABI library: error slot, launch counter, launch macro.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmnc_b200.h"

namespace mmnc {

void set_error(const char *fmt, ...);
void count_launch();
int sm_count();

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Checks the launch that was just issued (cudaGetLastError does not synchronise).
inline int after_launch(const char *what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MMNC_ERR_CUDA;
    }
    return MMNC_OK;
}

#define MMNC_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            mmnc::set_error(__VA_ARGS__); \
            return MMNC_ERR_INVALID;     \
        }                                \
    } while (0)

#define MMNC_CUDA(expr)                                                   \
    do {                                                                  \
        cudaError_t e__ = (expr);                                         \
        if (e__ != cudaSuccess) {                                         \
            mmnc::set_error("%s: %s", #expr, cudaGetErrorString(e__));    \
            return MMNC_ERR_CUDA;                                         \
        }                                                                 \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0.  `scratch` holds >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float *scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
    if (warp == 0) v = warp_sum(v);
    return v;
}

}  // namespace mmnc
