// K1: EntropyModel.quantize / dequantize as stand-alone entry points and
// GaussianConditional.build_indexes (row a10).  Inside the forward kernels quantisation is fused; these exist
// because `quantize`, `dequantize` and `build_indexes` are public methods of the CompressAI modules the
// reference calls (/root/reference/src/models/multi_task_compressor.py:509, 545).
#include "common.cuh"
#include "hd_math.cuh"

namespace mmnc {

constexpr int Q_THREADS = 256;

static inline unsigned q_blocks(int64_t n) {
    int64_t blocks = (n + Q_THREADS - 1) / Q_THREADS;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

__device__ __forceinline__ float mean_at(const float *means, int mode, int64_t i, int64_t C, int64_t S) {
    if (mode == MMNC_MEANS_NONE) return 0.f;
    if (mode == MMNC_MEANS_PER_CHANNEL) return means[(i / S) % C];
    return means[i];
}

__global__ void __launch_bounds__(Q_THREADS)
quantize_noise_kernel(const float *__restrict__ x, int64_t n, int noise_mode, const float *__restrict__ noise,
                      uint64_t seed, uint64_t offset, float *__restrict__ out) {
    if (noise_mode == MMNC_QUANT_NOISE_PHILOX_DEV) {
        const uint64_t *st = reinterpret_cast<const uint64_t *>(noise);
        seed = st[0];
        offset += st[1];
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float u = (noise_mode == MMNC_QUANT_NOISE_GIVEN) ? noise[i]
                                                                : philox_uniform_centered(seed, (uint64_t)i + offset);
        out[i] = x[i] + u;
    }
}

template <typename OutT, bool kAddBack>
__global__ void __launch_bounds__(Q_THREADS)
quantize_round_kernel(const float *__restrict__ x, int64_t n, int64_t C, int64_t S,
                      const float *__restrict__ means, int means_mode, OutT *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float m = mean_at(means, means_mode, i, C, S);
        const float r = rintf(x[i] - m);  // torch.round == round-half-to-even
        out[i] = kAddBack ? (OutT)(r + m) : (OutT)r;
    }
}

__global__ void __launch_bounds__(Q_THREADS)
dequantize_symbols_kernel(const int32_t *__restrict__ sym, int64_t n, int64_t C, int64_t S,
                          const float *__restrict__ means, int means_mode, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)sym[i] + mean_at(means, means_mode, i, C, S);
}

__global__ void __launch_bounds__(Q_THREADS)
build_indexes_kernel(const float *__restrict__ scales, int64_t n, const float *__restrict__ table, int table_len,
                     float bound, int32_t *__restrict__ idx) {
    extern __shared__ float tab[];
    for (int k = threadIdx.x; k < table_len; k += blockDim.x) tab[k] = table[k];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        idx[i] = gc_scale_index(scales[i], bound, tab, table_len);
}

}  // namespace mmnc

using namespace mmnc;

extern "C" int mmnc_quantize_noise(const float *x, int64_t n, int noise_mode, const float *noise, uint64_t seed,
                                   uint64_t offset, float *out, void *stream) {
    MMNC_REQUIRE(n >= 0, "quantize_noise: negative size");
    MMNC_REQUIRE(noise_mode == MMNC_QUANT_NOISE_PHILOX || noise_mode == MMNC_QUANT_NOISE_GIVEN ||
                     noise_mode == MMNC_QUANT_NOISE_PHILOX_DEV, "quantize_noise: bad noise_mode %d", noise_mode);
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(x && out, "quantize_noise: null pointer");
    MMNC_REQUIRE(noise_mode == MMNC_QUANT_NOISE_PHILOX || noise, "quantize_noise: needs the noise pointer");
    quantize_noise_kernel<<<q_blocks(n), Q_THREADS, 0, as_stream(stream)>>>(x, n, noise_mode, noise, seed, offset, out);
    return after_launch("quantize_noise_kernel");
}

static int check_means(const char *who, const float *means, int mode) {
    MMNC_REQUIRE(mode >= 0 && mode <= 2, "%s: bad means_mode %d", who, mode);
    MMNC_REQUIRE(mode == MMNC_MEANS_NONE || means, "%s: means_mode %d needs means", who, mode);
    return MMNC_OK;
}

extern "C" int mmnc_quantize_dequantize(const float *x, int64_t B, int64_t C, int64_t S, const float *means,
                                        int means_mode, float *out, void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && S >= 0, "quantize_dequantize: negative dimension");
    if (int rc = check_means("quantize_dequantize", means, means_mode)) return rc;
    const int64_t n = B * C * S;
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(x && out, "quantize_dequantize: null pointer");
    quantize_round_kernel<float, true><<<q_blocks(n), Q_THREADS, 0, as_stream(stream)>>>(x, n, C, S, means, means_mode, out);
    return after_launch("quantize_round_kernel<float>");
}

extern "C" int mmnc_quantize_symbols(const float *x, int64_t B, int64_t C, int64_t S, const float *means,
                                     int means_mode, int32_t *symbols, void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && S >= 0, "quantize_symbols: negative dimension");
    if (int rc = check_means("quantize_symbols", means, means_mode)) return rc;
    const int64_t n = B * C * S;
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(x && symbols, "quantize_symbols: null pointer");
    quantize_round_kernel<int32_t, false><<<q_blocks(n), Q_THREADS, 0, as_stream(stream)>>>(x, n, C, S, means, means_mode, symbols);
    return after_launch("quantize_round_kernel<int32>");
}

extern "C" int mmnc_dequantize_symbols(const int32_t *symbols, int64_t B, int64_t C, int64_t S, const float *means,
                                       int means_mode, float *out, void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && S >= 0, "dequantize_symbols: negative dimension");
    if (int rc = check_means("dequantize_symbols", means, means_mode)) return rc;
    const int64_t n = B * C * S;
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(symbols && out, "dequantize_symbols: null pointer");
    dequantize_symbols_kernel<<<q_blocks(n), Q_THREADS, 0, as_stream(stream)>>>(symbols, n, C, S, means, means_mode, out);
    return after_launch("dequantize_symbols_kernel");
}

extern "C" int mmnc_build_indexes(const float *scales, int64_t n, const float *scale_table, int table_len,
                                  float scale_bound, int32_t *indexes, void *stream) {
    MMNC_REQUIRE(n >= 0, "build_indexes: negative size");
    MMNC_REQUIRE(table_len >= 1 && table_len <= 8192, "build_indexes: table_len %d out of range", table_len);
    if (n == 0) return MMNC_OK;
    MMNC_REQUIRE(scales && scale_table && indexes, "build_indexes: null pointer");
    build_indexes_kernel<<<q_blocks(n), Q_THREADS, sizeof(float) * table_len, as_stream(stream)>>>(
        scales, n, scale_table, table_len, scale_bound, indexes);
    return after_launch("build_indexes_kernel");
}
