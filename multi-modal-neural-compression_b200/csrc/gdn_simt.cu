// K4 (SIMT path): GDN / IGDN forward and backward with fp32 FMAs.
//
// Replaces compressai.layers.GDN.forward (x**2 -> conv2d 1x1 -> (r)sqrt -> mul: ~10 launches and 9 tensor-sized
// HBM passes, reference sites /root/reference/src/models/multi_task_compressor.py:146-172 and
// disjoint_latent.py:150-154) with one launch forward.  This file is the exact-fp32 path: it serves channel
// counts / spatial sizes the tensor-core kernel (gdn_tc.cu) does not take, and MMNC_GDN_FP32 requests.
//
// Layout: x is (B, C, HW).  Pixels are flattened over (b, hw); thread <-> pixel, so for a fixed channel the 32
// lanes of a warp touch 32 consecutive floats (coalesced whenever HW >= 32).  Output channels are processed in
// chunks of 16 accumulators per thread; the chunk's column block of gamma sits in shared memory and is read as
// broadcast float4s (16 FMAs per 4 LDS.128 + 1 LDG that hits L1 after the first chunk).
#include "common.cuh"
#include "gdn_params.cuh"

namespace mmnc {

constexpr int GDN_TP = 128;  // pixels per block
constexpr int GDN_CH = 16;   // output channels per chunk

// MODE 0: forward            acc_i = sum_j gamma[i][j] x_j^2 ; y_i = x_i * (beta_i + acc_i)^(-+1/2)
// MODE 1: backward, stage 1  same contraction; u_i = p g_i x_i n_i^(p-1) -> U ; dx_i = g_i n_i^p
// MODE 2: backward, stage 2  acc_k = sum_i gamma[i][k] u_i ; dx_k += 2 x_k acc_k
// JS > 1 (small, latency-bound layers): JS threads share a pixel and split the contraction index between them, so the
// serial chain of dependent load rounds is JS times shorter; their partial sums meet in shared memory.
template <int MODE, int JS>
__global__ void __launch_bounds__(GDN_TP *JS)
gdn_simt_kernel(const float *__restrict__ x, const float *__restrict__ g, int64_t NP, int C, int64_t HW,
                const GdnParams prm, int inverse, float *out, float *U) {
    extern __shared__ __align__(16) float W[];  // [C][GDN_CH] (+ [JS - 1][GDN_CH][GDN_TP] partial sums when JS > 1)
    constexpr int THREADS = GDN_TP * JS;
    const int nchunks = (C + GDN_CH - 1) / GDN_CH;
    const float pcoef = inverse ? 0.5f : -0.5f;
    const int px = threadIdx.x % GDN_TP, js = threadIdx.x / GDN_TP;
    const int jn = (C + JS - 1) / JS, j_begin = js * jn, j_end = (j_begin + jn < C) ? j_begin + jn : C;
    float *part = W + (size_t)C * GDN_CH;
    for (int64_t tile = blockIdx.x; tile * GDN_TP < NP; tile += gridDim.x) {
        const int64_t P = tile * GDN_TP + px;
        const bool valid = P < NP;
        const int64_t b = valid ? P / HW : 0;
        const int64_t base = b * C * HW + (valid ? P - b * HW : 0);
        const float *xin = (MODE == 2 ? U : x) + base;
        for (int chunk = blockIdx.y; chunk < nchunks; chunk += gridDim.y) {
            const int i0 = chunk * GDN_CH;
            __syncthreads();
            // loads first (eight per thread in flight), then the re-parametrisation and the shared-memory stores:
            // one dependent L2 round trip per element made this staging loop the longest part of a small layer
            for (int sbase = threadIdx.x; sbase < C * GDN_CH; sbase += THREADS * 8) {
                float raw[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = sbase + u * THREADS;
                    float v = 0.f;
                    if (idx < C * GDN_CH) {
                        if (MODE == 2) {  // W[i][kk] = gamma[i][i0 + kk]
                            const int i = idx / GDN_CH, kk = idx - i * GDN_CH;
                            if (i0 + kk < C) v = prm.gamma[(int64_t)i * C + i0 + kk];
                        } else {          // W[j][ii] = gamma[i0 + ii][j]
                            const int ii = idx / C, j = idx - ii * C;
                            if (i0 + ii < C) v = prm.gamma[(int64_t)(i0 + ii) * C + j];
                        }
                    }
                    raw[u] = v;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int idx = sbase + u * THREADS;
                    if (idx >= C * GDN_CH) break;
                    if (MODE == 2) {
                        const int i = idx / GDN_CH, kk = idx - i * GDN_CH;
                        W[idx] = (i0 + kk < C) ? prm.g_of(raw[u]) : 0.f;
                    } else {
                        const int ii = idx / C, j = idx - ii * C;
                        W[j * GDN_CH + ii] = (i0 + ii < C) ? prm.g_of(raw[u]) : 0.f;
                    }
                }
            }
            __syncthreads();
            float acc[GDN_CH];
#pragma unroll
            for (int k = 0; k < GDN_CH; ++k) acc[k] = 0.f;
            if (valid) {
#pragma unroll 16
                for (int j = j_begin; j < j_end; ++j) {
                    float v = xin[(int64_t)j * HW];
                    if (MODE != 2) v = v * v;
                    const float4 *w4 = reinterpret_cast<const float4 *>(W + j * GDN_CH);
#pragma unroll
                    for (int q = 0; q < GDN_CH / 4; ++q) {
                        const float4 w = w4[q];
                        acc[4 * q + 0] += v * w.x;
                        acc[4 * q + 1] += v * w.y;
                        acc[4 * q + 2] += v * w.z;
                        acc[4 * q + 3] += v * w.w;
                    }
                }
            }
            if (JS > 1) {  // slices 1 .. JS-1 hand their partial sums to slice 0 (fixed order: deterministic)
                if (js > 0) {
#pragma unroll
                    for (int k = 0; k < GDN_CH; ++k) part[((js - 1) * GDN_CH + k) * GDN_TP + px] = acc[k];
                }
                __syncthreads();
                if (js == 0) {
#pragma unroll
                    for (int q = 1; q < JS; ++q)
#pragma unroll
                        for (int k = 0; k < GDN_CH; ++k) acc[k] += part[((q - 1) * GDN_CH + k) * GDN_TP + px];
                }
            }
            if (!valid || js != 0) continue;
            float xi_v[GDN_CH], gi_v[GDN_CH], be_v[GDN_CH];
#pragma unroll
            for (int ii = 0; ii < GDN_CH; ++ii) {  // every load of the epilogue in flight before the first use
                const int i = (i0 + ii < C) ? i0 + ii : C - 1;
                const int64_t a = base + (int64_t)i * HW;
                xi_v[ii] = x[a];
                gi_v[ii] = (MODE == 1) ? g[a] : 0.f;
                be_v[ii] = (MODE != 2) ? prm.beta[i] : 0.f;
            }
#pragma unroll
            for (int ii = 0; ii < GDN_CH; ++ii) {
                const int i = i0 + ii;
                if (i >= C) break;
                const int64_t a = base + (int64_t)i * HW;
                const float xi = xi_v[ii];
                if (MODE == 2) {
                    out[a] += 2.f * xi * acc[ii];
                } else {
                    const float n = prm.b_of(be_v[ii]) + acc[ii];
                    const float rt = sqrtf(n);
                    const float pw = inverse ? rt : 1.f / rt;  // n^p
                    if (MODE == 0) {
                        out[a] = xi * pw;
                    } else {
                        const float gi = gi_v[ii];
                        U[a] = pcoef * gi * xi * (pw / n);
                        out[a] = gi * pw;
                    }
                }
            }
        }
    }
}

// dgamma / dbeta partials: D[i][j] = sum_P U[i][P] * X2[j][P], j in [0, C] with X2[C][P] == 1 (gives dbeta).
// Split-K SGEMM: block = 64 x 64 output tile over a contiguous range of flattened pixels, 256 threads, 4 x 4
// outputs per thread; partial results go to part[ksplit][C][C+1], summed by gdn_reduce_kernel (deterministic).
constexpr int DG_T = 64, DG_K = 16;

__global__ void __launch_bounds__(256)
gdn_dgamma_kernel(const float *__restrict__ U, const float *__restrict__ x, int64_t NP, int C, int64_t HW,
                  int64_t k_per_split, float *__restrict__ part) {
    __shared__ __align__(16) float As[DG_K][DG_T + 4];
    __shared__ __align__(16) float Bs[DG_K][DG_T + 4];
    const int CJ = C + 1;
    const int tiles_j = (CJ + DG_T - 1) / DG_T;
    const int ti = blockIdx.x / tiles_j, tj = blockIdx.x - ti * tiles_j;
    const int i0 = ti * DG_T, j0 = tj * DG_T;
    const int64_t k_begin = (int64_t)blockIdx.y * k_per_split;
    const int64_t k_end = (k_begin + k_per_split < NP) ? k_begin + k_per_split : NP;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lk = threadIdx.x & 15;   // pixel within the K chunk this thread loads
    const int lr = threadIdx.x >> 4;   // first row this thread loads (rows lr, lr+16, lr+32, lr+48)
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    for (int64_t k0 = k_begin; k0 < k_end; k0 += DG_K) {
        const int64_t P = k0 + lk;
        const bool pv = P < k_end;
        const int64_t b = pv ? P / HW : 0;
        const int64_t base = b * C * HW + (pv ? P - b * HW : 0);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = lr + 16 * r;
            const int i = i0 + row, j = j0 + row;
            As[lk][row] = (pv && i < C) ? U[base + (int64_t)i * HW] : 0.f;
            float bv = 0.f;
            if (pv && j < C) { const float t = x[base + (int64_t)j * HW]; bv = t * t; }
            else if (pv && j == C) bv = 1.f;
            Bs[lk][row] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < DG_K; ++kk) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 bq = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] += av[p] * bv[q];
        }
        __syncthreads();
    }
    float *dst = part + (int64_t)blockIdx.y * C * CJ;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = i0 + ty * 4 + p;
        if (i >= C) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + tx * 4 + q;
            if (j < CJ) dst[(int64_t)i * CJ + j] = acc[p][q];
        }
    }
}

// one warp per output element: lanes stride over the partials, then a shuffle tree — a fixed summation order
// Fixed-order sum of the per-CTA partials.  Block = 32 consecutive elements x 8 slices of the partial index: a warp
// reads 128 contiguous bytes of one partial (the old one-warp-per-element mapping fetched 32 sectors per load), four
// independent accumulators per thread keep the loads in flight, and the 8 slices are combined through shared memory.
// (Few elements, e.g. C = 3: EX = 8 elements x 32 slices per block instead, so that the partial index is still spread.)
template <int EX>
__global__ void __launch_bounds__(256)
gdn_reduce_kernel(const float *__restrict__ part, int ksplit, int C, const GdnParams prm,
                  float *__restrict__ dgamma, float *__restrict__ dbeta) {
    constexpr int SL = 256 / EX;  // slices of the partial index
    __shared__ float red[SL][EX + 1];
    const int CJ = C + 1;
    const int64_t n = (int64_t)C * CJ;
    const int ex = threadIdx.x % EX, ky = threadIdx.x / EX;
    const int64_t e = (int64_t)blockIdx.x * EX + ex;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (e < n) {
        const float *p = part + e;
        int k = ky;
        for (; k + 3 * SL < ksplit; k += 4 * SL) {
            a0 += p[(int64_t)k * n];
            a1 += p[(int64_t)(k + SL) * n];
            a2 += p[(int64_t)(k + 2 * SL) * n];
            a3 += p[(int64_t)(k + 3 * SL) * n];
        }
        for (; k < ksplit; k += SL) a0 += p[(int64_t)k * n];
    }
    red[ky][ex] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ky == 0 && e < n) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < SL; ++q) s += red[q][ex];
        const int i = (int)(e / CJ), j = (int)(e - (int64_t)i * CJ);
        if (j < C) dgamma[(int64_t)i * C + j] = prm.dg((int64_t)i * C + j, s); else dbeta[i] = prm.db(i, s);
    }
}

int gdn_reduce_partials(const float *part, int ksplit, int C, const GdnParams &prm, float *dgamma, float *dbeta,
                        cudaStream_t s) {
    const int64_t n = (int64_t)C * (C + 1);
    if (n >= 8192) gdn_reduce_kernel<32><<<(unsigned)((n + 31) / 32), 256, 0, s>>>(part, ksplit, C, prm, dgamma, dbeta);
    else gdn_reduce_kernel<8><<<(unsigned)((n + 7) / 8), 256, 0, s>>>(part, ksplit, C, prm, dgamma, dbeta);
    return after_launch("gdn_reduce_kernel");
}

struct SimtGrid { unsigned tiles, split; };
static inline SimtGrid simt_grid(int64_t NP, int C) {
    const int64_t tiles = (NP + GDN_TP - 1) / GDN_TP;
    const int nchunks = (C + GDN_CH - 1) / GDN_CH;
    const int64_t want = (int64_t)sm_count() * 4;
    int64_t split = 1;
    if (tiles < want) { split = (want + tiles - 1) / tiles; if (split > nchunks) split = nchunks; }
    SimtGrid gr;
    gr.tiles = (unsigned)(tiles < want * 4 ? tiles : want * 4);
    gr.split = (unsigned)split;
    return gr;
}

static inline int dgamma_ksplit(int64_t NP, int C) {
    const int tiles = ((C + DG_T - 1) / DG_T) * ((C + 1 + DG_T - 1) / DG_T);
    int64_t ks = ((int64_t)sm_count() * 2 + tiles - 1) / tiles;
    const int64_t max_ks = (NP + 127) / 128;  // at least 128 pixels per split (small layers are latency-bound: spread them)
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    return (int)ks;
}

// few blocks (small layers): 4 or 2 threads per pixel split the contraction index, see the kernel.  Up to one block
// per SM: 4 (512 threads); up to two per SM: 2 (two 256-thread blocks fit one SM's registers, so still a single wave).
static inline int simt_slices(const SimtGrid &gr) {
    const int64_t blocks = (int64_t)gr.tiles * gr.split;
    return blocks <= sm_count() ? 4 : (blocks <= 2 * (int64_t)sm_count() ? 2 : 1);
}

template <int MODE, int JS>
static int simt_launch_js(const SimtGrid &gr, size_t smem, const float *x, const float *g, int64_t NP, int C, int64_t HW,
                          const GdnParams &prm, int inverse, float *out, float *U, cudaStream_t s, const char *what) {
    smem += sizeof(float) * (size_t)(JS - 1) * GDN_CH * GDN_TP;
    if (smem > 48 * 1024)
        MMNC_CUDA(cudaFuncSetAttribute(gdn_simt_kernel<MODE, JS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gdn_simt_kernel<MODE, JS><<<dim3(gr.tiles, gr.split), GDN_TP * JS, smem, s>>>(x, g, NP, C, HW, prm, inverse, out, U);
    return after_launch(what);
}

template <int MODE>
static int simt_launch(const SimtGrid &gr, size_t smem, const float *x, const float *g, int64_t NP, int C, int64_t HW,
                       const GdnParams &prm, int inverse, float *out, float *U, cudaStream_t s, const char *what) {
    switch (simt_slices(gr)) {
        case 4: return simt_launch_js<MODE, 4>(gr, smem, x, g, NP, C, HW, prm, inverse, out, U, s, what);
        case 2: return simt_launch_js<MODE, 2>(gr, smem, x, g, NP, C, HW, prm, inverse, out, U, s, what);
        default: return simt_launch_js<MODE, 1>(gr, smem, x, g, NP, C, HW, prm, inverse, out, U, s, what);
    }
}

int gdn_simt_forward(const float *x, int64_t B, int64_t C, int64_t HW, const GdnParams &prm, int inverse, float *y,
                     cudaStream_t s) {
    const int64_t NP = B * HW;
    const SimtGrid gr = simt_grid(NP, (int)C);
    const size_t smem = sizeof(float) * (size_t)C * GDN_CH;
    return simt_launch<0>(gr, smem, x, nullptr, NP, (int)C, HW, prm, inverse, y, nullptr, s, "gdn_simt_kernel<fwd>");
}

size_t gdn_simt_backward_workspace(int64_t B, int64_t C, int64_t HW) {
    const int64_t NP = B * HW;
    const int ks = dgamma_ksplit(NP, (int)C);
    return sizeof(float) * ((size_t)(B * C * HW) + (size_t)ks * C * (C + 1)) + 256;
}

int gdn_simt_backward(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const GdnParams &prm,
                      int inverse, float *dx, float *dbeta, float *dgamma, void *workspace, size_t workspace_bytes,
                      cudaStream_t s) {
    const int64_t NP = B * HW;
    MMNC_REQUIRE(workspace_bytes >= gdn_simt_backward_workspace(B, C, HW), "gdn_backward: workspace too small");
    float *U = static_cast<float *>(workspace);
    float *part = U + ((B * C * HW + 63) / 64) * 64;
    const SimtGrid gr = simt_grid(NP, (int)C);
    const size_t smem = sizeof(float) * (size_t)C * GDN_CH;
    if (int rc = simt_launch<1>(gr, smem, x, g, NP, (int)C, HW, prm, inverse, dx, U, s, "gdn_simt_kernel<bwd1>")) return rc;
    if (int rc = simt_launch<2>(gr, smem, x, g, NP, (int)C, HW, prm, inverse, dx, U, s, "gdn_simt_kernel<bwd2>")) return rc;
    const int ks = dgamma_ksplit(NP, (int)C);
    int64_t kper = (NP + ks - 1) / ks;
    kper = (kper + DG_K - 1) / DG_K * DG_K;
    const int tiles = (int)(((C + DG_T - 1) / DG_T) * ((C + 1 + DG_T - 1) / DG_T));
    gdn_dgamma_kernel<<<dim3((unsigned)tiles, (unsigned)ks), 256, 0, s>>>(U, x, NP, (int)C, HW, kper, part);
    if (int rc = after_launch("gdn_dgamma_kernel")) return rc;
    return gdn_reduce_partials(part, ks, (int)C, prm, dgamma, dbeta, s);
}

}  // namespace mmnc
