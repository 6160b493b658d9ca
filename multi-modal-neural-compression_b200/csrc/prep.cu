// (f4) Input pipeline on the GPU: raw decoded pixels -> the float NCHW tensors the compressors consume.
//
// The reference converts every sample on the host, one PIL image at a time, inside DataLoader worker processes
// (/root/reference/src/datasets/clevr.py:48-83, /root/reference/src/datasets/transforms.py:39-131) and ships fp32
// tensors to the device: 117 MB per 64-image, 3-task batch.  Here the host only stacks the raw pixels of a batch (uint8
// HWC for rgb / normal / semantic, uint16 HW for depth) into pinned memory - a quarter of the bytes - and three small
// kernels do what `ToTensor`, the 16-bit depth scaling and the semantic class remapping do, bit for bit:
//   rgb, normal : uint8 (B, H, W, Cs) -> float (B, Cd, H, W), value / 255 (true division, like torch's .div(255)),
//                 first Cd channels only (clevr.py:66-67: rgb keeps 3 of 4)
//   depth       : uint16 (B, H, W)    -> float (B, 1, H, W), value / (2^15 - 1)   (transforms.py:121-125)
//   semantic    : uint8 (B, H, W, Cs) -> float (B, 1, H, W), class index of channel 1 through a 256-entry table built
//                 from SEM1_CLASSES (clevr.py:13, 68-79); values outside the table keep their raw value, as the
//                 reference's in-place replacement loop leaves them
// A warp reads 32 consecutive pixels (Cs bytes each) and writes 32 consecutive floats per output plane.
#include "common.cuh"

namespace mmnc {

constexpr int PREP_THREADS = 256;

__global__ void __launch_bounds__(PREP_THREADS)
prep_u8_hwc_kernel(const uint8_t *__restrict__ src, int64_t B, int64_t HW, int Cs, int Cd, float divisor,
                   float *__restrict__ dst) {
    const int64_t n = B * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / HW, p = i - b * HW;
        const uint8_t *px = src + i * Cs;
        float *out = dst + b * Cd * HW + p;
        for (int c = 0; c < Cd; ++c) out[(int64_t)c * HW] = (float)px[c] / divisor;
    }
}

__global__ void __launch_bounds__(PREP_THREADS)
prep_u16_kernel(const uint16_t *__restrict__ src, int64_t n, float divisor, float *__restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i] / divisor;
}

__global__ void __launch_bounds__(PREP_THREADS)
prep_labels_kernel(const uint8_t *__restrict__ src, int64_t n, int Cs, int channel, const float *__restrict__ lut256,
                   float *__restrict__ dst) {
    __shared__ float lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut256[i];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = lut[src[i * Cs + channel]];
}

static inline unsigned prep_blocks(int64_t n) {
    int64_t blocks = (n + PREP_THREADS - 1) / PREP_THREADS;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace mmnc

using namespace mmnc;

extern "C" int mmnc_prep_u8_hwc_to_f32_chw(const uint8_t *src, int64_t B, int64_t HW, int src_channels,
                                           int dst_channels, float divisor, float *dst, void *stream) {
    MMNC_REQUIRE(B >= 0 && HW >= 0, "prep_u8: negative dimension");
    MMNC_REQUIRE(src_channels >= 1 && dst_channels >= 1 && dst_channels <= src_channels, "prep_u8: bad channel counts");
    MMNC_REQUIRE(divisor > 0.f, "prep_u8: divisor must be positive");
    if (B * HW == 0) return MMNC_OK;
    MMNC_REQUIRE(src && dst, "prep_u8: null pointer");
    prep_u8_hwc_kernel<<<prep_blocks(B * HW), PREP_THREADS, 0, as_stream(stream)>>>(src, B, HW, src_channels, dst_channels,
                                                                                   divisor, dst);
    return after_launch("prep_u8_hwc_kernel");
}

extern "C" int mmnc_prep_u16_to_f32(const uint16_t *src, int64_t n, float divisor, float *dst, void *stream) {
    MMNC_REQUIRE(n >= 0 && divisor > 0.f, "prep_u16: bad arguments");
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(src && dst, "prep_u16: null pointer");
    prep_u16_kernel<<<prep_blocks(n), PREP_THREADS, 0, as_stream(stream)>>>(src, n, divisor, dst);
    return after_launch("prep_u16_kernel");
}

extern "C" int mmnc_prep_labels(const uint8_t *src, int64_t n_pixels, int src_channels, int channel,
                                const float *lut256, float *dst, void *stream) {
    MMNC_REQUIRE(n_pixels >= 0 && src_channels >= 1 && channel >= 0 && channel < src_channels, "prep_labels: bad arguments");
    if (n_pixels == 0) return MMNC_OK;
    MMNC_REQUIRE(src && lut256 && dst, "prep_labels: null pointer");
    prep_labels_kernel<<<prep_blocks(n_pixels), PREP_THREADS, 0, as_stream(stream)>>>(src, n_pixels, src_channels, channel,
                                                                                     lut256, dst);
    return after_launch("prep_labels_kernel");
}
