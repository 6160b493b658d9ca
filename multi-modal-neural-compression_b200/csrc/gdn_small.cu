// K4 for very few channels (C <= 4): the output heads' last IGDNs run on the reconstructed images themselves
// (C = 3 or 1 at 256x256 and 128x128).  The contraction is 9 FMAs per pixel, so these layers are pure streaming:
// thread <-> 4 consecutive pixels, float4 loads / stores per channel, gamma and beta in registers.
// Backward keeps the C (C + 1) partial sums of d gamma / d beta in registers across the grid-stride loop and
// writes one partial per block (summed by gdn_reduce_partials in a fixed order).
#include "common.cuh"
#include "gdn_params.cuh"

namespace mmnc {

constexpr int GS_THREADS = 256;

template <int V> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<1> { using type = float; };

template <int V>
__device__ __forceinline__ void vload(const float *p, float (&v)[V]) {
    if (V == 4) { const float4 t = __ldcs(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1 % V] = t.y; v[2 % V] = t.z; v[3 % V] = t.w; }
    else v[0] = __ldcs(p);
}
template <int V>
__device__ __forceinline__ void vstore(float *p, const float (&v)[V]) {
    if (V == 4) __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1 % V], v[2 % V], v[3 % V]));
    else __stcs(p, v[0]);
}

// kNHWC: channels-last input / output - the V pixels of a thread are C * V consecutive floats (C float4 when V = 4),
// de-interleaved in registers.
template <int C, int V, bool kBackward, bool kNHWC = false>
__global__ void __launch_bounds__(GS_THREADS)
gdn_small_kernel(const float *__restrict__ x, const float *__restrict__ g, int64_t NP, int64_t HW, const GdnParams prm,
                 int inverse, float *__restrict__ out, float *__restrict__ part) {
    float gam[C][C], bet[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
        bet[i] = prm.b(i);
#pragma unroll
        for (int j = 0; j < C; ++j) gam[i][j] = prm.g(i * C + j);
    }
    float acc[C][C + 1];
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j <= C; ++j) acc[i][j] = 0.f;
    const float coef = inverse ? 0.5f : -0.5f;
    const int64_t groups = NP / V;  // HW % V == 0, so a group never straddles two images
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < groups; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t P = q * V;
        const int64_t b = P / HW;
        const int64_t base = b * C * HW + (P - b * HW);
        float xv[C][V], gv[C][V];
        if constexpr (kNHWC) {
            float tx[C][V], tg[C][V];  // flat view: element (pixel v, channel c) sits at v * C + c
#pragma unroll
            for (int k = 0; k < C; ++k) {
                vload<V>(x + P * C + k * V, tx[k]);
                if (kBackward) vload<V>(g + P * C + k * V, tg[k]);
            }
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    xv[c][v] = (&tx[0][0])[v * C + c];
                    if (kBackward) gv[c][v] = (&tg[0][0])[v * C + c];
                }
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                vload<V>(x + base + c * HW, xv[c]);
                if (kBackward) vload<V>(g + base + c * HW, gv[c]);
            }
        }
        float res[C][V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float x2[C], n[C], rs[C];
#pragma unroll
            for (int c = 0; c < C; ++c) x2[c] = xv[c][v] * xv[c][v];
#pragma unroll
            for (int i = 0; i < C; ++i) {
                float s = bet[i];
#pragma unroll
                for (int j = 0; j < C; ++j) s = fmaf(gam[i][j], x2[j], s);
                n[i] = s;
                rs[i] = rsqrtf(s);
            }
            if (!kBackward) {
#pragma unroll
                for (int i = 0; i < C; ++i) res[i][v] = xv[i][v] * (inverse ? n[i] * rs[i] : rs[i]);
            } else {
                float u[C];
#pragma unroll
                for (int i = 0; i < C; ++i) {
                    const float pm1 = inverse ? rs[i] : rs[i] * rs[i] * rs[i];
                    u[i] = coef * gv[i][v] * xv[i][v] * pm1;
#pragma unroll
                    for (int j = 0; j < C; ++j) acc[i][j] = fmaf(u[i], x2[j], acc[i][j]);
                    acc[i][C] += u[i];
                }
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    float t = 0.f;
#pragma unroll
                    for (int i = 0; i < C; ++i) t = fmaf(gam[i][k], u[i], t);
                    res[k][v] = fmaf(2.f * xv[k][v], t, gv[k][v] * (inverse ? n[k] * rs[k] : rs[k]));
                }
            }
        }
        if constexpr (kNHWC) {
            float to[C][V];
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int v = 0; v < V; ++v) (&to[0][0])[v * C + c] = res[c][v];
#pragma unroll
            for (int k = 0; k < C; ++k) vstore<V>(out + P * C + k * V, to[k]);
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) vstore<V>(out + base + c * HW, res[c]);
        }
    }
    if (kBackward) {
        __shared__ float red[GS_THREADS / 32][C * (C + 1)];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < C; ++i)
#pragma unroll
            for (int j = 0; j <= C; ++j) {
                const float s = warp_sum(acc[i][j]);
                if (lane == 0) red[warp][i * (C + 1) + j] = s;
            }
        __syncthreads();
        if (threadIdx.x < C * (C + 1)) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < GS_THREADS / 32; ++w) s += red[w][threadIdx.x];
            part[(int64_t)blockIdx.x * C * (C + 1) + threadIdx.x] = s;
        }
    }
}

int gdn_reduce_partials(const float *part, int ksplit, int C, const GdnParams &prm, float *dgamma, float *dbeta,
                        cudaStream_t s);

bool gdn_small_supported(int64_t C) { return C >= 1 && C <= 4; }

static int small_grid(int64_t NP, int V) {
    int64_t blocks = (NP / V + GS_THREADS - 1) / GS_THREADS;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

size_t gdn_small_backward_workspace(int64_t B, int64_t C, int64_t HW) {
    return sizeof(float) * (size_t)small_grid(B * HW, 1) * C * (C + 1) + 256;
}

template <int C, bool kBackward>
static int launch_small(const float *x, const float *g, int64_t NP, int64_t HW, const GdnParams &prm, int inverse,
                        float *out, float *part, int *grid_out, cudaStream_t s, int nhwc) {
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(g)) % 16 == 0);
    // NCHW: four pixels of one image plane; NHWC: four pixels = 4 C consecutive floats anywhere in the tensor
    const bool vec = al && (nhwc ? (NP % 4 == 0) : (HW % 4 == 0));
    const int grid = small_grid(NP, vec ? 4 : 1);
    *grid_out = grid;
    if (nhwc && C > 1) {  // C = 1: both memory formats are the same bytes
        if (vec) gdn_small_kernel<C, 4, kBackward, true><<<grid, GS_THREADS, 0, s>>>(x, g, NP, HW, prm, inverse, out, part);
        else gdn_small_kernel<C, 1, kBackward, true><<<grid, GS_THREADS, 0, s>>>(x, g, NP, HW, prm, inverse, out, part);
    } else {
        if (vec) gdn_small_kernel<C, 4, kBackward><<<grid, GS_THREADS, 0, s>>>(x, g, NP, HW, prm, inverse, out, part);
        else gdn_small_kernel<C, 1, kBackward><<<grid, GS_THREADS, 0, s>>>(x, g, NP, HW, prm, inverse, out, part);
    }
    return after_launch(kBackward ? "gdn_small_kernel<bwd>" : "gdn_small_kernel<fwd>");
}

int gdn_small_forward(const float *x, int64_t B, int64_t C, int64_t HW, const GdnParams &prm, int inverse, float *y,
                      cudaStream_t s, int nhwc) {
    int grid;
    switch (C) {
        case 1: return launch_small<1, false>(x, nullptr, B * HW, HW, prm, inverse, y, nullptr, &grid, s, nhwc);
        case 2: return launch_small<2, false>(x, nullptr, B * HW, HW, prm, inverse, y, nullptr, &grid, s, nhwc);
        case 3: return launch_small<3, false>(x, nullptr, B * HW, HW, prm, inverse, y, nullptr, &grid, s, nhwc);
        case 4: return launch_small<4, false>(x, nullptr, B * HW, HW, prm, inverse, y, nullptr, &grid, s, nhwc);
        default: set_error("gdn_small_forward: C out of range"); return MMNC_ERR_UNSUPPORTED;
    }
}

int gdn_small_backward(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const GdnParams &prm,
                       int inverse, float *dx, float *dbeta, float *dgamma, void *workspace, size_t workspace_bytes,
                       cudaStream_t s, int nhwc) {
    MMNC_REQUIRE(workspace_bytes >= gdn_small_backward_workspace(B, C, HW), "gdn_backward: workspace too small");
    float *part = static_cast<float *>(workspace);
    int grid = 0, rc;
    switch (C) {
        case 1: rc = launch_small<1, true>(x, g, B * HW, HW, prm, inverse, dx, part, &grid, s, nhwc); break;
        case 2: rc = launch_small<2, true>(x, g, B * HW, HW, prm, inverse, dx, part, &grid, s, nhwc); break;
        case 3: rc = launch_small<3, true>(x, g, B * HW, HW, prm, inverse, dx, part, &grid, s, nhwc); break;
        case 4: rc = launch_small<4, true>(x, g, B * HW, HW, prm, inverse, dx, part, &grid, s, nhwc); break;
        default: set_error("gdn_small_backward: C out of range"); return MMNC_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    return gdn_reduce_partials(part, grid, (int)C, prm, dgamma, dbeta, s);
}

}  // namespace mmnc
