// (f2) Step metrics of the reference's `average_metrics` (/root/reference/src/models/multi_task_compressor.py:359-384).
//
// PSNR of the regression tasks needs nothing new: the distortion kernel already returns sum (x_hat - x)^2 / (B C)
// and PSNR = -10 log10(MSE) once both images are scaled by 255 with data_range 255.  The semantic task is different:
// the reference takes argmax over the 17 class logits, casts to float and computes PSNR / MS-SSIM of the CLASS-ID
// images (data_range 17).  This kernel fuses argmax, the class-id image and its squared error against the target in
// one pass over the logits (18 floats read per pixel instead of materialising argmax, a cast and a subtraction).
#include "common.cuh"

namespace mmnc {

constexpr int AM_THREADS = 256;

// logits (B, K, S), target (B, 1, S) class ids as floats -> labels (B, 1, S) float (optional), sse += sum (argmax - target)^2
__global__ void __launch_bounds__(AM_THREADS)
argmax_sse_kernel(const float *__restrict__ logits, const float *__restrict__ target, int64_t B, int K, int64_t S,
                  float *__restrict__ labels, float *__restrict__ sse) {
    __shared__ float red[32];
    float acc = 0.f;
    const int64_t n = B * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / S, s = i - b * S;
        const float *p = logits + b * K * S + s;  // a warp reads 32 consecutive pixels of one class plane: coalesced
        float best = p[0];
        int arg = 0;
        for (int k = 1; k < K; ++k) {
            const float v = p[(int64_t)k * S];
            if (v > best || (v != v && best == best)) { best = v; arg = k; }  // first maximum wins, NaN counts as maximal (torch.argmax)
        }
        const float lab = (float)arg;
        if (labels != nullptr) labels[i] = lab;
        if (target != nullptr) {
            const float d = lab - target[i];
            acc += d * d;
        }
    }
    if (sse != nullptr) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) atomicAdd(sse, tot);
    }
}

}  // namespace mmnc

using namespace mmnc;

extern "C" int mmnc_argmax_sse(const float *logits, const float *target, int64_t B, int K, int64_t S, float *labels,
                               float *sse, void *stream) {
    MMNC_REQUIRE(B >= 0 && K >= 1 && S >= 0, "argmax_sse: bad dimensions");
    if (B * S == 0) return MMNC_OK;
    MMNC_REQUIRE(logits && (labels || (target && sse)), "argmax_sse: null pointer");
    MMNC_REQUIRE(!sse || target, "argmax_sse: the squared error needs a target");
    int64_t blocks = (B * S + AM_THREADS - 1) / AM_THREADS;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    argmax_sse_kernel<<<(unsigned)blocks, AM_THREADS, 0, as_stream(stream)>>>(logits, target, B, K, S, labels, sse);
    return after_launch("argmax_sse_kernel");
}

// ---------------------------------------------------------------------------------------------------------------
// Per-channel sum over (batch, space) of an NCHW tensor: the bias gradient of every convolution.  torch computes it
// with a generic reduction (4.8 ms per training step of config C2, profiles/); this is a plain streaming reduction:
// grid = (C, splits), float4 loads along the contiguous plane, one partial per block, fixed-order finish.
namespace mmnc {

constexpr int CS_THREADS = 256;

__global__ void __launch_bounds__(CS_THREADS)
channel_sum_kernel(const float *__restrict__ g, int64_t B, int64_t C, int64_t S, float *__restrict__ part) {
    __shared__ float red[32];
    const int64_t c = blockIdx.x;
    const int splits = gridDim.y;
    float acc = 0.f;
    const bool vec = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
    if (vec) {
        const int64_t q = S >> 2, n4 = B * q;
        for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < n4; e += (int64_t)splits * blockDim.x) {
            const int64_t b = e / q, s4 = e - b * q;
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(g + (b * C + c) * S) + s4);
            acc += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        const int64_t n = B * S;
        for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < n; e += (int64_t)splits * blockDim.x) {
            const int64_t b = e / S, s = e - b * S;
            acc += g[(b * C + c) * S + s];
        }
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) part[c * splits + blockIdx.y] = tot;
}

__global__ void __launch_bounds__(CS_THREADS)
channel_sum_finish_kernel(const float *__restrict__ part, int64_t C, int splits, float *__restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[c * splits + k];  // fixed order: bit-reproducible
    out[c] = s;
}

}  // namespace mmnc

extern "C" int64_t mmnc_channel_sum_workspace_floats(int64_t B, int64_t C, int64_t S) {
    if (B <= 0 || C <= 0 || S <= 0) return 1;
    int64_t splits = ((int64_t)sm_count() * 8 + C - 1) / C;
    const int64_t max_splits = (B * S / 4 + CS_THREADS - 1) / CS_THREADS;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 1024) splits = 1024;
    return C * splits;
}

extern "C" int mmnc_channel_sum(const float *g, int64_t B, int64_t C, int64_t S, float *workspace, float *out,
                                void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && S >= 0, "channel_sum: negative dimension");
    if (C == 0) return MMNC_OK;
    MMNC_REQUIRE(out, "channel_sum: null output");
    if (B * S == 0) {
        MMNC_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, as_stream(stream)));
        return MMNC_OK;
    }
    MMNC_REQUIRE(g && workspace, "channel_sum: null pointer");
    const int splits = (int)(mmnc_channel_sum_workspace_floats(B, C, S) / C);
    channel_sum_kernel<<<dim3((unsigned)C, (unsigned)splits), CS_THREADS, 0, as_stream(stream)>>>(g, B, C, S, workspace);
    if (int rc = after_launch("channel_sum_kernel")) return rc;
    channel_sum_finish_kernel<<<(unsigned)((C + CS_THREADS - 1) / CS_THREADS), CS_THREADS, 0, as_stream(stream)>>>(
        workspace, C, splits, out);
    return after_launch("channel_sum_finish_kernel");
}

// ---------------------------------------------------------------------------------------------------------------
// In-place per-channel bias add on an NCHW tensor: x[b, c, s] += bias[c].  torch adds the bias with a generic
// broadcasting element-wise kernel after every cuDNN convolution (3.8 ms per training step of config C2, 2.5 TB/s);
// this one streams float4 along the contiguous planes.
namespace mmnc {

__global__ void __launch_bounds__(256)
bias_add_kernel(float *__restrict__ x, const float *__restrict__ bias, int64_t BC, int64_t C, int64_t S) {
    const bool vec = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    if (vec) {
        const int64_t q = S >> 2, n4 = BC * q;
        float4 *x4 = reinterpret_cast<float4 *>(x);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
            const float b = __ldg(bias + (i / q) % C);
            float4 v = x4[i];
            v.x += b; v.y += b; v.z += b; v.w += b;
            x4[i] = v;
        }
    } else {
        const int64_t n = BC * S;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
            x[i] += __ldg(bias + (i / S) % C);
    }
}

}  // namespace mmnc

extern "C" int mmnc_bias_add(float *x, const float *bias, int64_t B, int64_t C, int64_t S, void *stream) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && S >= 0, "bias_add: negative dimension");
    if (B * C * S == 0) return MMNC_OK;
    MMNC_REQUIRE(x && bias, "bias_add: null pointer");
    const int64_t work = (S % 4 == 0) ? B * C * S / 4 : B * C * S;
    int64_t blocks = (work + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    bias_add_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, bias, B * C, C, S);
    return after_launch("bias_add_kernel");
}
