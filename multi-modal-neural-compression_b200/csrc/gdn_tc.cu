// K4 (tensor-core path): GDN / IGDN forward with the channel contraction on tcgen05.
//
//   norm[p][i] = beta_i + sum_j x[p][j]^2 * gamma[i][j]        D[128 x N] (+)= A[128 x K] * B[K x N]
//
// One CTA = 128 threads = one tile of 128 flattened pixels at a time (persistent loop over tiles).
//   A = x^2 (tf32)  : written by the threads straight into TENSOR MEMORY with tcgen05.st — thread t owns pixel t,
//                     i.e. TMEM lane t; column k holds channel k.  The global loads behind it are fully coalesced
//                     (a warp reads 32 consecutive pixels of one channel) and no shared-memory operand layout has
//                     to be produced for the NCHW (pixel-contiguous = "MN-major") input.
//   B = gamma (tf32): staged ONCE per CTA in shared memory in the canonical K-major no-swizzle core-matrix layout
//                     (8 rows x 16 bytes per core matrix), consumed through a shared-memory matrix descriptor.
//   D = fp32 accumulator in TMEM, read back with tcgen05.ld for the epilogue y = x * (beta + D)^(-+1/2).
// kind::tf32, cta_group::1, M = 128, N = C rounded up to 16, K = 8 per instruction (C rounded up to 8).
//
// Precision modes: MMNC_GDN_TF32 = one pass (x^2 and gamma rounded to tf32 with round-to-nearest);
// MMNC_GDN_3XTF32 = hi/lo split of both operands, three passes (hi*hi + lo*hi + hi*lo), fp32-class accuracy.
// The tensor pipe has so much headroom over HBM here that even three passes stay memory-bound.
#include "common.cuh"
#include "gdn_params.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"
#include "gdn_nhwc.cuh"

namespace mmnc {

namespace tc {

struct Geometry {
    int C, Kp, Np;        // channels, K padded to 8, N padded to 16
    int a_cols, d_col;    // TMEM columns used by A (hi [+ lo]) and first column of D
    uint32_t tmem_cols;   // power of two >= 32
    size_t b_bytes;       // one B operand (hi); the lo copy follows when 3xTF32
};

}  // namespace tc

// One tile: 128 pixels x all channels.  kFull = the tile has no out-of-range pixel (no per-element predicates).
// KP8 = Kp / 8 is a template parameter so that the pixel's x values stay in registers (fully unrolled): they are
// loaded from HBM once, squared into the A operand, and reused by the epilogue — 8 B/element of traffic, the
// algorithmic minimum.  kBetaInMma: when C is not a multiple of 8 the K padding is free, so one padded A column
// is the constant 1 and the matching B row holds beta: the accumulator comes back as beta + sum gamma x^2.
template <bool k3x, int KP8, bool kBetaInMma, bool kFull, bool kNHWC = false>
__device__ __forceinline__ void gdn_tc_tile(const float *__restrict__ xb, float *__restrict__ yb, int64_t HW, int C,
                                            bool inverse, bool valid, const float *beta_s, uint32_t tmem_base, uint32_t lane_base,
                                            uint32_t d_col, uint64_t desc_hi, uint64_t desc_lo, uint32_t idesc,
                                            uint64_t *mbar, uint32_t parity, int vec = 1) {
    using namespace tc;
    constexpr int Kp = KP8 * 8;
    constexpr int NP16 = (KP8 + 1) / 2;  // Np / 16
    constexpr uint32_t a_hi = 0, a_lo = (uint32_t)Kp;
    // ---- all of the pixel's channels in flight at once (one DRAM round trip per tile), kept in registers.
    // Padded channels (c >= C) re-read the last real channel: their B rows are zero, so any finite value works.
    const uint32_t sb = (uint32_t)HW * 4u;  // channel stride in bytes
    float xv[Kp];
    if constexpr (kNHWC) {
        nhwc_load_row<Kp>(xb, C, vec, kFull || valid, xv);  // xb = this pixel's row
    } else {
#pragma unroll
        for (int c = 0; c < Kp; ++c) {
            const int cc = (c < Kp - 8) ? c : ((c < C) ? c : C - 1);
            xv[c] = (kFull || valid) ? __ldcs(chan_ptr(xb, sb, cc)) : 0.f;
            if ((c & 7) == 7) asm volatile("" ::: "memory");  // keep address arithmetic next to its load (register pressure)
        }
    }
    // ---- A operand: x^2 -> tf32 -> TMEM (lane = pixel, column = channel)
#pragma unroll
    for (int c0 = 0; c0 < Kp; c0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float sq = xv[c0 + j] * xv[c0 + j];
            if (kBetaInMma && c0 + j == Kp - 1) sq = 1.f;  // the constant column that picks up beta
            hi[j] = to_tf32(sq);
            if (k3x) lo[j] = to_tf32(sq - __uint_as_float(hi[j]));
        }
        tmem_st8(lane_base + a_hi + c0, hi);
        if (k3x) tmem_st8(lane_base + a_lo + c0, lo);
    }
    tmem_st_wait();
    fence_before();
    __syncthreads();
    // ---- MMA: one thread issues Kp/8 (x3) instructions, then commits to the mbarrier
    if (threadIdx.x == 0) {
        fence_after();
#pragma unroll
        for (int ks = 0; ks < KP8; ++ks) {
            const uint64_t koff = (uint64_t)((ks * 256) >> 4);  // two core matrices (8 tf32) along K
            mma_tf32_ts(tmem_base + d_col, tmem_base + a_hi + ks * 8, desc_hi + koff, idesc, ks > 0 ? 1u : 0u);
            if (k3x) {
                mma_tf32_ts(tmem_base + d_col, tmem_base + a_lo + ks * 8, desc_hi + koff, idesc, 1);
                mma_tf32_ts(tmem_base + d_col, tmem_base + a_hi + ks * 8, desc_lo + koff, idesc, 1);
            }
        }
        mma_commit(mbar);
    }
    mbar_wait(mbar, parity);
    fence_after();
    // ---- epilogue: y = x * n^(-+1/2) with n = beta + D; thread t writes pixel t of every channel
#pragma unroll
    for (int q = 0; q < NP16; ++q) {
        uint32_t r[16];
        tmem_ld16(lane_base + d_col + q * 16, r);
        tmem_ld_wait();
        float o16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int i = q * 16 + j;
            o16[j] = 0.f;
            if (i < Kp) {
                float n = __uint_as_float(r[j]);
                if (!kBetaInMma) n += beta_s[i];
                const float rs = fast_rsqrt(n);
                const float out = xv[i] * (inverse ? n * rs : rs);
                o16[j] = out;
                if constexpr (!kNHWC) {
                    const bool ok = (kFull || valid) && (i < Kp - 8 || i < C);
                    if (ok) __stcs(chan_ptr(yb, sb, i), out);
                }
            }
        }
        if constexpr (kNHWC) {
            if (kFull || valid) nhwc_store_block<16>(yb, q * 16, C, vec, o16);
        }
    }
    fence_before();
    __syncthreads();  // every lane has drained D before the next tile overwrites A / D
    fence_after();
}

// CTAs per SM that TMEM admits for this instance (columns: A = Kp [x2 when split], D = Np <= Kp + 8; power-of-two alloc)
template <bool k3x, int KP8>
constexpr int tc_fwd_ctas() {
    const int need = KP8 * 8 * (k3x ? 2 : 1) + (KP8 + 1) / 2 * 16;
    return need <= 128 ? 4 : (need <= 256 ? 2 : 1);
}

template <bool k3x, int KP8, bool kBetaInMma, bool kNHWC = false>
__global__ void __launch_bounds__(tc::TILE_M, tc_fwd_ctas<k3x, KP8>())
gdn_tc_forward_kernel(const float *__restrict__ x, int64_t NP, int64_t HW, const GdnParams prm, int inverse,
                      float *__restrict__ y, int C, int Np, int d_col_i,
                      uint32_t tmem_cols, uint32_t b_bytes, int vec) {
    using namespace tc;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    float *Bs_hi = reinterpret_cast<float *>(smem);
    float *Bs_lo = reinterpret_cast<float *>(smem + b_bytes);
    float *beta_s = reinterpret_cast<float *>(smem + b_bytes * (k3x ? 2 : 1));
    constexpr int Kp = KP8 * 8;
    const int warp = threadIdx.x >> 5;

    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    // B operand: Bs[n/8][k/4][n%8][k%4] = tf32(gamma[n][k]) (zero padded; row k = Kp-1 holds beta when kBetaInMma)
    // Loads first, eight per thread in flight, then the arithmetic: the loop used to be one dependent L2 round trip
    // per element (91 of them per thread at C = 100, ~30 us per CTA - most of a small layer's run time).
    constexpr int kcores = Kp >> 2;
    constexpr int SU = 8;
    const int total = Np * Kp;
    for (int base = threadIdx.x; base < total; base += TILE_M * SU) {
        float raw[SU];
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int idx = base + u * TILE_M;
            const int n = idx / Kp, k = idx - n * Kp;
            float v = 0.f;
            if (idx < total) {
                if (kBetaInMma && k == Kp - 1) v = (n < C) ? prm.beta[n] : 0.f;
                else if (n < C && k < C) v = prm.gamma[(int64_t)n * C + k];
            }
            raw[u] = v;
        }
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int idx = base + u * TILE_M;
            if (idx >= total) break;
            const int n = idx / Kp, k = idx - n * Kp;
            float g = (n < C && k < C) ? prm.g_of(raw[u]) : 0.f;
            if (kBetaInMma && k == Kp - 1) g = (n < C) ? prm.b_of(raw[u]) : 1.f;
            const int off = (((n >> 3) * kcores + (k >> 2)) << 5) + ((n & 7) << 2) + (k & 3);
            const uint32_t hi = to_tf32(g);
            reinterpret_cast<uint32_t *>(Bs_hi)[off] = hi;
            if (k3x) reinterpret_cast<uint32_t *>(Bs_lo)[off] = to_tf32(g - __uint_as_float(hi));
        }
    }
    for (int i = threadIdx.x; i < Np; i += TILE_M) beta_s[i] = (i < C) ? prm.b(i) : 1.f;
    fence_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t d_col = (uint32_t)d_col_i;
    const uint32_t idesc = make_idesc(Np);
    const uint32_t lbo = 128, sbo = (uint32_t)kcores * 128;
    const uint64_t desc_hi = make_b_desc(smem_u32(Bs_hi), lbo, sbo);
    const uint64_t desc_lo = make_b_desc(smem_u32(Bs_lo), lbo, sbo);
    uint32_t parity = 0;

    for (int64_t tile = blockIdx.x; tile * TILE_M < NP; tile += gridDim.x) {
        const int64_t P = tile * TILE_M + threadIdx.x;
        const bool full = (tile + 1) * TILE_M <= NP;
        const bool valid = P < NP;
        const int64_t b = valid ? P / HW : 0;
        // NCHW: (image, pixel) -> first channel of that pixel; NHWC: the pixel's row of C channels
        const int64_t base = kNHWC ? (valid ? P * C : 0) : (b * C * HW + (valid ? P - b * HW : 0));
        if (full)
            gdn_tc_tile<k3x, KP8, kBetaInMma, true, kNHWC>(x + base, y + base, HW, C, inverse != 0, true, beta_s, tmem_base,
                                                           lane_base, d_col, desc_hi, desc_lo, idesc, &mbar, parity, vec);
        else
            gdn_tc_tile<k3x, KP8, kBetaInMma, false, kNHWC>(x + base, y + base, HW, C, inverse != 0, valid, beta_s, tmem_base,
                                                            lane_base, d_col, desc_hi, desc_lo, idesc, &mbar, parity, vec);
        parity ^= 1;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// gdn_tc_fwd2.cu: the TMA-in / TMA-out generation of this kernel (single-pass TF32, HW % 128 == 0 only)
bool gdn_tc_forward2_supported(const float *x, const float *y, int64_t B, int64_t C, int64_t HW);
int gdn_tc_forward2(const float *x, int64_t B, int64_t C, int64_t HW, const GdnParams &prm, int inverse, float *y,
                    cudaStream_t s);

static bool tc_geometry(int64_t C, bool k3x, tc::Geometry *g) {
    if (C < 8 || C > 128) return false;  // the pixel's channels live in registers: C <= 128
    g->C = (int)C;
    g->Kp = (int)((C + 7) / 8 * 8);
    g->Np = (int)((C + 15) / 16 * 16);
    g->a_cols = g->Kp * (k3x ? 2 : 1);
    g->d_col = g->a_cols;
    const int need = g->a_cols + g->Np;
    if (need > 512) return false;
    uint32_t cols = 32;
    while ((int)cols < need) cols <<= 1;
    g->tmem_cols = cols;
    g->b_bytes = ((size_t)g->Np * g->Kp * sizeof(float) + 127) / 128 * 128;
    if (g->b_bytes * (k3x ? 2 : 1) + sizeof(float) * g->Np + 1024 > 227 * 1024) return false;  // gamma must fit smem
    return true;
}

bool gdn_tc_supported(int64_t B, int64_t C, int64_t HW, int precision) {
    tc::Geometry g;
    // anything with at least one full 128-pixel tile: staging the B operand costs a few microseconds per CTA
    return tc_geometry(C, precision == MMNC_GDN_3XTF32, &g) && C >= 16 && B * HW >= 128 && HW < (1 << 24);
}

// vector width of a channels-last row access: the row stride is C floats, so C and the base must allow it
int gdn_nhwc_vec(const void *a, const void *b, const void *c, int64_t C) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c);
    if (C % 4 == 0 && (bits & 15) == 0) return 4;
    if (C % 2 == 0 && (bits & 7) == 0) return 2;
    return 1;
}

int gdn_tc_forward(const float *x, int64_t B, int64_t C, int64_t HW, const GdnParams &prm, int inverse, int precision,
                   float *y, cudaStream_t s, int nhwc) {
    tc::Geometry g;
    const bool k3x = (precision == MMNC_GDN_3XTF32);
    if (!nhwc && !k3x && gdn_tc_forward2_supported(x, y, B, C, HW)) return gdn_tc_forward2(x, B, C, HW, prm, inverse, y, s);
    if (!tc_geometry(C, k3x, &g)) {
        set_error("gdn_tc_forward: C = %lld not supported by the tensor-core path", (long long)C);
        return MMNC_ERR_UNSUPPORTED;
    }
    const int64_t NP = B * HW;
    const int64_t tiles = (NP + tc::TILE_M - 1) / tc::TILE_M;
    size_t smem = g.b_bytes * (k3x ? 2 : 1) + sizeof(float) * g.Np;
    // co-resident CTAs per SM are bounded by TMEM (512 columns): pad the shared-memory request so the hardware
    // never schedules a CTA that would then spin inside tcgen05.alloc
    const int max_ctas = 512 / (int)g.tmem_cols;
    const size_t smem_cap = 227 * 1024;
    const size_t min_smem = smem_cap / (size_t)(max_ctas + 1) + 1;
    if (smem < min_smem && max_ctas < 8) smem = min_smem;
    smem = (smem + 127) / 128 * 128;
    if (smem > smem_cap) {
        set_error("gdn_tc_forward: gamma does not fit shared memory for C = %lld", (long long)C);
        return MMNC_ERR_UNSUPPORTED;
    }
    using Kernel = void (*)(const float *, int64_t, int64_t, const GdnParams, int, float *, int, int, int, uint32_t,
                            uint32_t, int);
    Kernel kernel = nullptr;
    const bool beta_in_mma = (C % 8) != 0;
    if (nhwc && k3x) {
        set_error("gdn_tc_forward: the channels-last kernels are single-pass TF32 only");
        return MMNC_ERR_UNSUPPORTED;
    }
#define MMNC_TC_CASE(N)                                                                                        \
    case N:                                                                                                    \
        if (nhwc) kernel = beta_in_mma ? (Kernel)gdn_tc_forward_kernel<false, N, true, true> : (Kernel)gdn_tc_forward_kernel<false, N, false, true>; \
        else if (k3x) kernel = beta_in_mma ? (Kernel)gdn_tc_forward_kernel<true, N, true> : (Kernel)gdn_tc_forward_kernel<true, N, false>;   \
        else     kernel = beta_in_mma ? (Kernel)gdn_tc_forward_kernel<false, N, true> : (Kernel)gdn_tc_forward_kernel<false, N, false>; \
        break;
    switch (g.Kp / 8) {
        MMNC_TC_CASE(2) MMNC_TC_CASE(3) MMNC_TC_CASE(4) MMNC_TC_CASE(5) MMNC_TC_CASE(6) MMNC_TC_CASE(7) MMNC_TC_CASE(8)
        MMNC_TC_CASE(9) MMNC_TC_CASE(10) MMNC_TC_CASE(11) MMNC_TC_CASE(12) MMNC_TC_CASE(13) MMNC_TC_CASE(14)
        MMNC_TC_CASE(15) MMNC_TC_CASE(16)
        default: break;
    }
#undef MMNC_TC_CASE
    if (kernel == nullptr) {
        set_error("gdn_tc_forward: no kernel instance for C = %lld", (long long)C);
        return MMNC_ERR_UNSUPPORTED;
    }
    // largest shared-memory carve-out: with the default heuristic the SM sometimes keeps a split that fits one CTA
    // only, and the second co-resident CTA (the whole point of the TMEM budget) never lands
    if (int rc = tmah::ensure_dynamic_smem(kernel, smem, true)) return rc;
    int64_t grid = (int64_t)sm_count() * (max_ctas < 4 ? max_ctas : 4);
    if (grid > tiles) grid = tiles;
    kernel<<<(unsigned)grid, tc::TILE_M, smem, s>>>(x, NP, HW, prm, inverse, y, g.C, g.Np, g.d_col, g.tmem_cols,
                                                    (uint32_t)g.b_bytes, nhwc ? gdn_nhwc_vec(x, y, nullptr, C) : 1);
    return after_launch(k3x ? "gdn_tc_forward_kernel<3xtf32>" : (nhwc ? "gdn_tc_forward_kernel<tf32, nhwc>" : "gdn_tc_forward_kernel<tf32>"));
}

}  // namespace mmnc
