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
