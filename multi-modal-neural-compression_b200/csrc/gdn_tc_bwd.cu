// K4 backward (tensor-core path): GDN / IGDN gradient as THREE tcgen05 contractions fused in one kernel
// (SURVEY.md section 8 row a8, Appendix A.5).  x and the upstream gradient g are read from HBM once, dx is written
// once — 12 B/element, the algorithmic minimum; the norm is recomputed instead of being saved by the forward.
//
// Per tile of 128 flattened pixels (M = 128 TMEM lanes), channels padded to P = ceil16(C + 1):
//   MMA1  n[p][i]   = sum_j x2[p][j] gamma[i][j]        A = x^2 in TMEM,  B = gamma (smem, K-major)       -> D
//   (epilogue 1)  rs = n^-1/2 ;  u_i = -+1/2 g_i x_i n_i^(p-1) ;  f_i = g_i n_i^p
//   MMA2  t[p][k]   = sum_i u[p][i] gamma[i][k]         A = u in TMEM,    B = gamma^T (smem, K-major)     -> D
//   MMA3  dG[i][j] += sum_p u[p][i] x2[p][j]            A = u^T, B = x2^T : both staged in shared memory with the
//                                                        pixel index as K (K-major, 128-byte swizzle)     -> D3
//   (epilogue 2)  dx_k = f_k + 2 x_k t_k
// The padded channel C carries the constant 1 on the x^2 side, so gamma's padded column holds beta (norm comes back
// as beta + sum) and column C of dG accumulates d beta = sum_p u.  D3 lives in TMEM for the whole kernel; every CTA
// writes its partial (C x (C+1)) once at the end and a small kernel sums the partials in a fixed order.
//
// CTA = 256 threads: thread t and thread t+128 own the same pixel (TMEM lane t % 128) and split the channels
// between them (warp w can touch TMEM lanes 32 (w % 4) .. +31, so two warpgroups can serve one tile) — the
// per-thread register footprint is halved and twice as many loads are in flight.
#include <stdlib.h>

#include "common.cuh"
#include "gdn_params.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"
#include "gdn_nhwc.cuh"

namespace mmnc {

namespace tcb {
constexpr int THREADS = 256;
constexpr int TILE = 128;

// byte offset of (row r, pixel p) in a K-major (K = pixel) 128-byte-swizzled operand whose 8-row groups are 1024 B
// apart and whose 32-pixel K atoms are R8 * 1024 B apart
__device__ __forceinline__ uint32_t sw128_offset(int r, int p, int R8) {
    const int atomk = p >> 5, p32 = p & 31;
    return (uint32_t)(((atomk * R8 + (r >> 3)) << 10) + ((r & 7) << 7) + ((((p32 >> 2) ^ (r & 7)) & 7) << 4) +
                      ((p32 & 3) << 2));
}
}  // namespace tcb

template <int KS, int NK>
__device__ __forceinline__ void v1_ts_chain(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if constexpr (KS < NK) {
        tc::mma_tf32_ts_step<KS * 8, KS * 16>(d, a, b_lo, b_hi, idesc, KS > 0 ? 1u : 0u);
        v1_ts_chain<KS + 1, NK>(d, a, b_lo, b_hi, idesc);
    }
}
template <int KS>
__device__ __forceinline__ void v1_ss_chain(uint32_t d3, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t atom16,
                                            uint32_t idesc, uint32_t first_acc) {
    if constexpr (KS < tcb::TILE / 8) {
        tc::mma_tf32_ss_step<(KS & 3) * 2>(d3, a_lo, b_lo, hi, (uint32_t)(KS >> 2) * atom16, idesc,
                                           KS == 0 ? first_acc : 1u);
        v1_ss_chain<KS + 1>(d3, a_lo, b_lo, hi, atom16, idesc, first_acc);
    }
}

// KH8 = (channels per thread) / 8 = P / 16.  HALF (which half of the channels), kInverse and kFull are template
// parameters so that every channel index below is a compile-time constant: only the last 16 padded channels of
// HALF 1 keep run-time checks against C, everything else is straight-line code with immediate offsets.
struct TileCtx {
    int HW, C, pix, R8, vec;
    uint32_t tmem_base, lane_base, idesc;
    uint8_t *ubuf, *x2buf;
    uint64_t desc_b1, desc_b2, desc_a3, desc_b3;
    uint64_t *mbar;
};

template <int KH8, int HALF, bool kInverse, bool kFull, bool kNHWC = false>
__device__ __forceinline__ void gdn_tc_bwd_tile_body(const float *__restrict__ xb, const float *__restrict__ gb,
                                                float *__restrict__ dxb, int HW, int C, bool valid, int pix,
                                                uint32_t tmem_base, uint32_t lane_base, uint8_t *ubuf,
                                                uint8_t *x2buf, int R8, uint64_t desc_b1, uint64_t desc_b2,
                                                uint64_t desc_a3, uint64_t desc_b3, uint32_t idesc, uint64_t *mbar,
                                                uint32_t &parity, bool first_tile, int vec = 1) {
    using namespace tc;
    using namespace tcb;
    constexpr int KH = KH8 * 8;        // channels handled by this thread
    constexpr int P = KH * 2;          // padded channel count (K and N of the pixel-row MMAs)
    constexpr int c_begin = HALF * KH;
    constexpr int SAFE = P - 16;       // channels below this are always real (C > P - 16)
    constexpr uint32_t a_col = 0, d_col = (uint32_t)P, d3_col = (uint32_t)(2 * P);
    const bool ok = kFull || valid;
    const uint32_t sb = (uint32_t)HW * 4u;  // channel stride in bytes
    // swizzled shared-memory offsets: row c = c_begin + j0 + j with j0, c_begin multiples of 8, so (c & 7) == j and
    // (c >> 3) is a constant: one register per j holds the thread-dependent part, the row group is an immediate
    uint32_t soff[8];
    {
        const int atomk = pix >> 5, p32 = pix & 31;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            soff[j] = (uint32_t)(((atomk * R8 + (c_begin >> 3)) << 10) + (j << 7) + ((((p32 >> 2) ^ j) & 7) << 4) +
                                 ((p32 & 3) << 2));
    }
    const int rows = 8 * R8;  // rows that exist in ubuf / x2buf
    const uint32_t us = smem_u32(ubuf), x2s = smem_u32(x2buf);

    // ---- loads: this thread's channels of x and g, all in flight (32-bit element offsets from a per-pixel base)
    float xv[KH], gv[KH];
    if constexpr (kNHWC) {
        // xb / gb / dxb = this pixel's row; the thread owns channels [c_begin, c_begin + KH) of it (0 beyond C)
        nhwc_load_row<KH>(xb + c_begin, C - c_begin, vec, ok, xv);
        nhwc_load_row<KH>(gb + c_begin, C - c_begin, vec, ok, gv);
    } else {
#pragma unroll
        for (int j = 0; j < KH; ++j) {
            const int c = c_begin + j;
            const int cc = (c < SAFE) ? c : ((c < C) ? c : C - 1);  // padded channels re-read a real one (weights are 0)
            xv[j] = ok ? __ldcs(chan_ptr(xb, sb, cc)) : 0.f;
            gv[j] = ok ? __ldcs(chan_ptr(gb, sb, cc)) : 0.f;
            // compile-time fence every 4 channels: keeps the address arithmetic next to its load instead of letting the
            // scheduler materialise all 2 KH 64-bit addresses first (the loads still issue back to back)
            if ((j & 3) == 3) asm volatile("" ::: "memory");
        }
    }
    // ---- x^2 -> A (TMEM) and -> x2buf (smem, K = pixel); padded channel C is the constant 1
#pragma unroll
    for (int j0 = 0; j0 < KH; j0 += 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c_begin + j0 + j;
            float sq = xv[j0 + j] * xv[j0 + j];
            if (c >= SAFE) sq = (c < C) ? sq : ((c == C) ? 1.f : 0.f);
            v[j] = to_tf32(sq);
            if (c < SAFE || c < rows) st_shared_u32(x2s + soff[j] + (j0 << 7), v[j]);
        }
        tmem_st8(lane_base + a_col + c_begin + j0, v);
    }
    tmem_st_wait();
    fence_async_smem();
    fence_before();
    __syncthreads();
    // ---- MMA1: D = x2 * gamma^T (+ beta through the constant column)
    if (threadIdx.x == 0) {
        fence_after();
        // per-step descriptor offsets are immediates inside the asm (tc_ptx.cuh): nothing 64-bit to hoist and spill
        v1_ts_chain<0, P / 8>(tmem_base + d_col, tmem_base + a_col, (uint32_t)desc_b1, (uint32_t)(desc_b1 >> 32), idesc);
        mma_commit(mbar);
    }
    mbar_wait(mbar, parity);
    parity ^= 1;
    fence_after();
    // ---- epilogue 1: u -> A (TMEM) and ubuf (smem); first term of dx kept in gv
    constexpr float coef = kInverse ? 0.5f : -0.5f;
#pragma unroll
    for (int j0 = 0; j0 < KH; j0 += 8) {
        uint32_t r[8], uu[8];
        tmem_ld8(lane_base + d_col + c_begin + j0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c_begin + j0 + j;
            const float n = __uint_as_float(r[j]);
            const float rs = fast_rsqrt(n);
            const float pw = kInverse ? n * rs : rs;            // n^p
            const float pm1 = kInverse ? rs : rs * rs * rs;     // n^(p-1)
            float u = (coef * gv[j0 + j]) * (xv[j0 + j] * pm1);
            if (c >= SAFE && c >= C) u = 0.f;
            if (!kFull && !valid) u = 0.f;
            gv[j0 + j] = gv[j0 + j] * pw;
            uu[j] = to_tf32(u);
            if (c < SAFE || c < rows) st_shared_u32(us + soff[j] + (j0 << 7), uu[j]);
        }
        tmem_st8(lane_base + a_col + c_begin + j0, uu);
    }
    tmem_st_wait();
    fence_async_smem();
    fence_before();
    __syncthreads();
    // ---- MMA2: D = u * gamma (B = gamma^T tile);  MMA3: D3 += u^T x2 (K = 128 pixels)
    if (threadIdx.x == 0) {
        fence_after();
        v1_ts_chain<0, P / 8>(tmem_base + d_col, tmem_base + a_col, (uint32_t)desc_b2, (uint32_t)(desc_b2 >> 32), idesc);
        v1_ss_chain<0>(tmem_base + d3_col, (uint32_t)desc_a3, (uint32_t)desc_b3, (uint32_t)(desc_a3 >> 32),
                       (uint32_t)R8 * 64u, idesc, first_tile ? 0u : 1u);
        mma_commit(mbar);
    }
    mbar_wait(mbar, parity);
    parity ^= 1;
    fence_after();
    // ---- epilogue 2: dx = g n^p + 2 x t
#pragma unroll
    for (int j0 = 0; j0 < KH; j0 += 8) {
        uint32_t r[8];
        tmem_ld8(lane_base + d_col + c_begin + j0, r);
        tmem_ld_wait();
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c_begin + j0 + j;
            const float out = fmaf(2.f * xv[j0 + j], __uint_as_float(r[j]), gv[j0 + j]);
            o8[j] = out;
            if constexpr (!kNHWC) {
                if (ok && (c < SAFE || c < C)) __stcs(chan_ptr(dxb, sb, c), out);
            }
        }
        if constexpr (kNHWC) {
            if (ok) nhwc_store_block<8>(dxb, c_begin + j0, C, vec, o8);
        }
    }
    fence_before();
    __syncthreads();  // A, D and both smem operands are free again
    fence_after();
}

// the rare partial tile (last tile of the tensor) takes the predicated body out of line so that it does not
// weigh on the register allocation of the steady-state loop
template <int KH8, int HALF, bool kInverse, bool kNHWC>
__device__ __noinline__ void gdn_tc_bwd_tile_partial(const float *xb, const float *gb, float *dxb, bool valid,
                                                     const TileCtx &t, uint32_t &parity, bool first) {
    gdn_tc_bwd_tile_body<KH8, HALF, kInverse, false, kNHWC>(xb, gb, dxb, t.HW, t.C, valid, t.pix, t.tmem_base, t.lane_base,
                                                            t.ubuf, t.x2buf, t.R8, t.desc_b1, t.desc_b2, t.desc_a3,
                                                            t.desc_b3, t.idesc, t.mbar, parity, first, t.vec);
}

template <int KH8, int HALF, bool kInverse, bool kNHWC>
__device__ __forceinline__ bool gdn_tc_bwd_loop(const float *__restrict__ x, const float *__restrict__ g,
                                                float *__restrict__ dx, int64_t NP, int64_t HW, const TileCtx &t) {
    using namespace tcb;
    uint32_t parity = 0;
    bool first = true;
    for (int64_t tile = blockIdx.x; tile * TILE < NP; tile += gridDim.x) {
        const int64_t Pix = tile * TILE + t.pix;
        const bool valid = Pix < NP;
        const int64_t b = valid ? Pix / HW : 0;
        // NCHW: first channel of the pixel inside its image; NHWC: the pixel's row
        const int64_t base = kNHWC ? (valid ? Pix * t.C : 0) : (b * t.C * HW + (valid ? Pix - b * HW : 0));
        if ((tile + 1) * TILE <= NP)
            gdn_tc_bwd_tile_body<KH8, HALF, kInverse, true, kNHWC>(x + base, g + base, dx + base, t.HW, t.C, true, t.pix,
                                                                   t.tmem_base, t.lane_base, t.ubuf, t.x2buf, t.R8,
                                                                   t.desc_b1, t.desc_b2, t.desc_a3, t.desc_b3, t.idesc,
                                                                   t.mbar, parity, first, t.vec);
        else
            gdn_tc_bwd_tile_partial<KH8, HALF, kInverse, kNHWC>(x + base, g + base, dx + base, valid, t, parity, first);
        first = false;
    }
    return first;
}

template <int KH8, bool kInverse, bool kNHWC = false>
__global__ void __launch_bounds__(tcb::THREADS, (KH8 <= 4) ? 2 : 1)
gdn_tc_backward_kernel(const float *__restrict__ x, const float *__restrict__ g, int64_t NP, int64_t HW,
                       const GdnParams prm,
                       float *__restrict__ dx, float *__restrict__ part, int C, uint32_t tmem_cols, int vec) {
    using namespace tc;
    using namespace tcb;
    constexpr int P = KH8 * 16;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    // carve: [ubuf | x2buf] (1024-aligned swizzle atoms) then gamma
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int R8 = (C + 1 + 7) >> 3;                 // 8-row groups actually stored per K atom
    const uint32_t buf_bytes = (uint32_t)R8 * 1024u * 4u;  // 4 K atoms of 32 pixels
    uint8_t *ubuf = smem;
    uint8_t *x2buf = smem + buf_bytes;
    float *Bs = reinterpret_cast<float *>(smem + 2 * buf_bytes);                          // gamma   (N = i, K = j)
    float *Bs2 = reinterpret_cast<float *>(smem + 2 * buf_bytes + (size_t)P * P * 4);    // gamma^T (N = k, K = i)
    const int warp = threadIdx.x >> 5;
    const int half = threadIdx.x >> 7;     // which half of the channels this thread owns
    const int pix = threadIdx.x & 127;     // pixel within the tile = TMEM lane

    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    // gamma tile: Bs[n/8][k/4][n%8][k%4] = tf32(gamma[n][k]); column k = C holds beta; zero elsewhere in the padding
    // Loads first (eight elements = sixteen loads per thread in flight), then the arithmetic: one dependent L2 round
    // trip per element made this loop ~30 us per CTA at C = 100.
    constexpr int kcores = P >> 2;
    constexpr int SU = 8;
    for (int base = threadIdx.x; base < P * P; base += THREADS * SU) {
        float raw[SU], rawt[SU];
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int idx = base + u * THREADS;
            const int n = idx / P, k = idx - n * P;
            float v = 0.f, vt = 0.f;
            if (idx < P * P) {
                if (n < C && k < C) { v = prm.gamma[(int64_t)n * C + k]; vt = prm.gamma[(int64_t)k * C + n]; }
                else if (n < C && k == C) v = prm.beta[n];
            }
            raw[u] = v; rawt[u] = vt;
        }
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int idx = base + u * THREADS;
            if (idx >= P * P) break;
            const int n = idx / P, k = idx - n * P;
            float v = 0.f;
            if (n < C && k < C) v = prm.g_of(raw[u]);
            else if (n < C && k == C) v = prm.b_of(raw[u]);
            else if (n >= C && k == C) v = 1.f;  // padded outputs get norm = 1 (finite)
            const int off = (((n >> 3) * kcores + (k >> 2)) << 5) + ((n & 7) << 2) + (k & 3);
            reinterpret_cast<uint32_t *>(Bs)[off] = to_tf32(v);
            // transposed copy for MMA2 (a K-major operand again; tf32 MN-major reads of the same tile returned zeros
            // on sm_100a, so the transpose is materialised once per CTA instead)
            const float vt = (n < C && k < C) ? prm.g_of(rawt[u]) : 0.f;
            reinterpret_cast<uint32_t *>(Bs2)[off] = to_tf32(vt);
        }
    }
    // rows of the two pixel-major operands that no thread writes (r in [P', 8 R8)) must stay finite: zero them
    for (uint32_t i = threadIdx.x; i < 2 * buf_bytes / 4; i += THREADS) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t bs_addr = smem_u32(Bs);
    // MMA1: B = gamma as (N = out channel, K = in channel), K-major, no swizzle
    const uint64_t desc_b1 = make_desc(bs_addr, 128, (uint32_t)kcores * 128, 0);
    // MMA2: B = gamma^T as (N = in channel k, K = out channel i), same layout
    const uint64_t desc_b2 = make_desc(smem_u32(Bs2), 128, (uint32_t)kcores * 128, 0);
    // MMA3: A = u^T (M = channel, K = pixel), B = x2^T (N = channel, K = pixel): K-major, 128-byte swizzle
    const uint64_t desc_a3 = make_desc(smem_u32(ubuf), 16, 1024, 2);
    const uint64_t desc_b3 = make_desc(smem_u32(x2buf), 16, 1024, 2);
    const uint32_t idesc = make_idesc_ex(P, false, false);
    TileCtx ctx;
    ctx.HW = (int)HW;  // channel stride in elements (< 2^31; per-pixel bases stay 64-bit)
    ctx.C = C; ctx.pix = pix; ctx.R8 = R8; ctx.vec = vec;
    ctx.tmem_base = tmem_base; ctx.lane_base = lane_base; ctx.idesc = idesc;
    ctx.ubuf = ubuf; ctx.x2buf = x2buf;
    ctx.desc_b1 = desc_b1; ctx.desc_b2 = desc_b2; ctx.desc_a3 = desc_a3; ctx.desc_b3 = desc_b3;
    ctx.mbar = &mbar;
    // the two warpgroups run the same tile sequence in lock step, each on its half of the channels
    const bool first = (half == 0) ? gdn_tc_bwd_loop<KH8, 0, kInverse, kNHWC>(x, g, dx, NP, HW, ctx)
                                   : gdn_tc_bwd_loop<KH8, 1, kInverse, kNHWC>(x, g, dx, NP, HW, ctx);
    // ---- this CTA's partial d gamma / d beta: D3 lane i = out channel, column j = in channel (j = C: d beta)
    float *dst = part + (int64_t)blockIdx.x * C * (C + 1);
    if (!first && half == 0) {
#pragma unroll 1
        for (int j0 = 0; j0 < P; j0 += 8) {
            uint32_t r[8];
            tmem_ld8(lane_base + (uint32_t)(2 * P) + j0, r);
            tmem_ld_wait();
            if (pix < C) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j0 + j <= C) dst[(int64_t)pix * (C + 1) + j0 + j] = __uint_as_float(r[j]);
            }
        }
    } else if (first && half == 0 && pix < C) {
        for (int j = 0; j <= C; ++j) dst[(int64_t)pix * (C + 1) + j] = 0.f;  // CTA without tiles
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// fixed-order reduction of per-CTA partials [ksplit][C][C+1] -> d gamma, d beta (gdn_simt.cu)
int gdn_reduce_partials(const float *part, int ksplit, int C, const GdnParams &prm, float *dgamma, float *dbeta,
                        cudaStream_t s);

// gdn_tc_bwd2.cu: the TMA-fed, software-pipelined generation of this kernel (HW % 128 == 0 only)
bool gdn_tc_backward2_supported(const float *x, const float *g, int64_t B, int64_t C, int64_t HW);
size_t gdn_tc_backward2_workspace(int64_t B, int64_t C, int64_t HW);
int gdn_tc_backward2(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const GdnParams &prm,
                     int inverse, float *dx, float *dbeta, float *dgamma, void *workspace, size_t workspace_bytes,
                     cudaStream_t s);

static bool tcb_geometry(int64_t C, int *P, uint32_t *tmem_cols, size_t *smem) {
    if (C < 16 || C > 111) return false;
    *P = (int)((C + 1 + 15) / 16 * 16);
    const int need = 3 * *P;
    if (need > 512) return false;
    uint32_t cols = 32;
    while ((int)cols < need) cols <<= 1;
    *tmem_cols = cols;
    const size_t R8 = (size_t)(C + 1 + 7) / 8;
    // MMA3 reads 128 rows of ubuf / P rows of x2buf: the over-read past the stored rows stays inside this allocation
    // as long as what follows is at least as large; gamma follows and a tail pad covers the rest
    size_t bytes = 2 * R8 * 4096 + 2 * (size_t)*P * *P * 4;
    const size_t overread = (size_t)(3 * R8 + 16) * 1024 + R8 * 4096;  // end of the last K atom of x2buf's M=128 view
    if (bytes < overread) bytes = overread;
    *smem = bytes + 1024 + 256;
    return *smem <= 227 * 1024;
}

bool gdn_tc_backward_supported(int64_t B, int64_t C, int64_t HW) {
    int P;
    uint32_t cols;
    size_t smem;
    return tcb_geometry(C, &P, &cols, &smem) && B * HW >= 128 && HW < (1 << 24);
}

static int tcb_grid(int64_t NP, uint32_t tmem_cols, size_t smem) {
    int per_sm = 512 / (int)tmem_cols;
    const int by_smem = (int)((227 * 1024) / smem);
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    int64_t grid = (int64_t)sm_count() * per_sm;
    const int64_t tiles = (NP + tcb::TILE - 1) / tcb::TILE;
    if (grid > tiles) grid = tiles;
    return (int)grid;
}

size_t gdn_tc_backward_workspace(int64_t B, int64_t C, int64_t HW) {
    int P;
    uint32_t cols;
    size_t smem;
    const size_t v2 = gdn_tc_backward2_workspace(B, C, HW);  // also covers the wide layers only the TMA-fed kernel takes
    if (!tcb_geometry(C, &P, &cols, &smem)) return v2;
    const size_t v1 = sizeof(float) * (size_t)tcb_grid(B * HW, cols, smem) * C * (C + 1) + 256;
    return v1 > v2 ? v1 : v2;
}

int gdn_tc_backward(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const GdnParams &prm,
                    int inverse, float *dx, float *dbeta, float *dgamma, void *workspace, size_t workspace_bytes,
                    cudaStream_t s, int nhwc) {
    if (!nhwc && gdn_tc_backward2_supported(x, g, B, C, HW))
        return gdn_tc_backward2(x, g, B, C, HW, prm, inverse, dx, dbeta, dgamma, workspace, workspace_bytes, s);
    int P;
    uint32_t cols;
    size_t smem;
    if (!tcb_geometry(C, &P, &cols, &smem)) {
        set_error("gdn_tc_backward: C = %lld not supported", (long long)C);
        return MMNC_ERR_UNSUPPORTED;
    }
    const int64_t NP = B * HW;
    const int grid = tcb_grid(NP, cols, smem);
    MMNC_REQUIRE(workspace_bytes >= sizeof(float) * (size_t)grid * C * (C + 1), "gdn_backward: workspace too small");
    // TMEM admits 512 / cols CTAs per SM: make shared memory say the same so no CTA ever spins in tcgen05.alloc
    const int max_ctas = 512 / (int)cols;
    const size_t min_smem = (227 * 1024) / (size_t)(max_ctas + 1) + 1;
    if (smem < min_smem) smem = min_smem;
    using Kernel = void (*)(const float *, const float *, int64_t, int64_t, const GdnParams, float *, float *, int,
                            uint32_t, int);
    Kernel kernel = nullptr;
#define MMNC_CASE(N)                                                                                                          \
    case N:                                                                                                                   \
        if (nhwc) kernel = inverse ? (Kernel)gdn_tc_backward_kernel<N, true, true> : (Kernel)gdn_tc_backward_kernel<N, false, true>; \
        else kernel = inverse ? (Kernel)gdn_tc_backward_kernel<N, true> : (Kernel)gdn_tc_backward_kernel<N, false>;            \
        break;
    switch (P / 16) {
        MMNC_CASE(2) MMNC_CASE(3) MMNC_CASE(4) MMNC_CASE(5) MMNC_CASE(6) MMNC_CASE(7)
        default: break;
    }
#undef MMNC_CASE
    if (kernel == nullptr) {
        set_error("gdn_tc_backward: no kernel instance for C = %lld", (long long)C);
        return MMNC_ERR_UNSUPPORTED;
    }
    // largest shared-memory carve-out: with the default heuristic the SM sometimes keeps a split that fits one CTA
    // only, and the second co-resident CTA (the whole point of the TMEM budget) never lands
    if (int rc = tmah::ensure_dynamic_smem(kernel, smem, true)) return rc;
    float *part = static_cast<float *>(workspace);
    kernel<<<(unsigned)grid, tcb::THREADS, smem, s>>>(x, g, NP, HW, prm, dx, part, (int)C, cols,
                                                      nhwc ? gdn_nhwc_vec(x, g, dx, C) : 1);
    if (int rc = after_launch(nhwc ? "gdn_tc_backward_kernel<nhwc>" : "gdn_tc_backward_kernel")) return rc;
    return gdn_reduce_partials(part, grid, (int)C, prm, dgamma, dbeta, s);
}

}  // namespace mmnc
