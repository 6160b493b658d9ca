// K4 forward, second generation: the tcgen05 GDN / IGDN contraction of gdn_tc.cu with TMA on both sides
// (SURVEY.md section 8 row a8, Appendix A.5).  Same math, same 8 B/element.
//
//   * a tile (128 pixels x C channels of the NCHW tensor) arrives with ONE cp.async.bulk.tensor (box 128 x C, no
//     swizzle: row c of the landing buffer is the 128 pixels of channel c, so thread <-> pixel reads are conflict-free)
//     and leaves with one cp.async.bulk.tensor store: y is written IN PLACE over x in the landing buffer.
//   * the compute warps never form a global address and never wait on HBM: x is read from shared memory twice (to
//     build A = x^2 in TMEM, and in the epilogue y = x n^(-+1/2)), so nothing but a 16-column slice of the accumulator
//     lives in registers (~50 registers per thread instead of 168).
//   * persistent CTA per SM with NGROUPS independent compute groups of 128 threads (thread <-> pixel <-> TMEM lane).
//     Every group owns a private ring of NSTAGES landing buffers, its own TMEM columns (A and D), its own mbarriers
//     and its own elected thread that issues loads, MMAs and stores - no cross-group protocol.  With three stages a
//     group has one tile being processed, the next one landing and the previous one draining to HBM.
// Requires HW % 128 == 0 and 16-byte aligned tensors; single-pass TF32 only.  Everything else stays on gdn_tc.cu.
#include <cuda.h>  // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time (no -lcuda)
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gdn_params.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace mmnc {

namespace tcf2 {
constexpr int TILE = 128;  // pixels per tile = TMEM lanes
}  // namespace tcf2

struct Fwd2Ctx {
    int C, n_k;              // channels, tiles of this CTA (all groups together)
    int tiles_per_img;
    uint32_t stage0;         // shared address of group 0 / stage 0; stages are Kp * 512 bytes each (C rows + padding rows)
    uint32_t gamma0;         // shared address of the gamma tile
    uint32_t beta0;          // shared address of beta (Np floats)
    uint32_t full_bar0, mma_bar0;
    uint32_t tmem_base;
    const void *tm_x, *tm_y;
};

__device__ __forceinline__ void fwd2_wait(uint32_t addr, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}

template <int KS, int NK>
__device__ __forceinline__ void fwd2_mma_chain(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if constexpr (KS < NK) {
        tc::mma_tf32_ts_step<KS * 8, KS * 16>(d, a, b_lo, b_hi, idesc, KS > 0 ? 1u : 0u);
        fwd2_mma_chain<KS + 1, NK>(d, a, b_lo, b_hi, idesc);
    }
}

// tile kk (group-local index) of group `group`: global tile number, image and first pixel
__device__ __forceinline__ void fwd2_coords(const volatile Fwd2Ctx &t, int ngroups, int group, int kk, int *b, int *hw0) {
    const uint32_t tile = blockIdx.x + (uint32_t)(kk * ngroups + group) * gridDim.x;
    const uint32_t tpi = (uint32_t)t.tiles_per_img;
    const uint32_t bb = tile / tpi;
    *b = (int)bb;
    *hw0 = (int)((tile - bb * tpi) * tcf2::TILE);
}

// SPLIT threads share a pixel (= a TMEM lane; warp w reaches lanes 32 (w % 4) ..) and take the 8-column blocks of A and
// the 16-column blocks of D round-robin: wide layers fit few groups in shared memory, and one 128-thread group is
// one warp per scheduler - nothing to hide a dependent latency behind.  NGROUPS * SPLIT = 4 keeps 16 warps per SM.
template <int KP8, int NGROUPS, int NSTAGES, int SPLIT, bool kBetaInMma>
__global__ void __launch_bounds__(NGROUPS * SPLIT * 128, 1)
gdn_tc_forward2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y, int ntiles,
                       int tiles_per_img, const GdnParams prm, int inverse, int C, int Np, uint32_t tmem_cols) {
    using namespace tc;
    using namespace tcf2;
    constexpr int Kp = KP8 * 8;
    constexpr int TPG = 128 * SPLIT;
    constexpr int THREADS = NGROUPS * TPG;
    constexpr uint32_t kcores = Kp >> 2;
    constexpr uint32_t GAMMA_HI = desc_hi(kcores * 128u, 0);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full_bar[NGROUPS * NSTAGES];
    __shared__ uint64_t mma_bar[NGROUPS];
    __shared__ uint32_t tmem_base_s;
    __shared__ Fwd2Ctx ctx_s;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // a stage holds Kp rows: the TMA box writes the first C, the padding rows are initialised once (0, and 1 in the row
    // of the constant channel that picks up beta) so that the A fill needs neither predicates nor selects
    const uint32_t stage_bytes = (uint32_t)Kp * 512u;
    float *Bs = reinterpret_cast<float *>(smem + (size_t)NGROUPS * NSTAGES * stage_bytes);
    float *beta_s = Bs + Np * Kp;
    const int warp = threadIdx.x >> 5;
    const int group = threadIdx.x / TPG, tg = threadIdx.x % TPG;
    const int pix = tg & 127, sub = tg >> 7;
    const bool leader = (tg == 0);

    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NGROUPS * NSTAGES; ++s) mbar_init(&full_bar[s], 1);
        for (int q = 0; q < NGROUPS; ++q) mbar_init(&mma_bar[q], 1);
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_y);
        Fwd2Ctx c;
        c.C = C;
        c.n_k = (blockIdx.x < (uint32_t)ntiles) ? (int)(((uint32_t)ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
        c.tiles_per_img = tiles_per_img;
        c.stage0 = smem_u32(smem);
        c.gamma0 = smem_u32(Bs);
        c.beta0 = smem_u32(beta_s);
        c.full_bar0 = smem_u32(&full_bar[0]);
        c.mma_bar0 = smem_u32(&mma_bar[0]);
        c.tmem_base = 0;
        c.tm_x = &tm_x;
        c.tm_y = &tm_y;
        ctx_s = c;
    }
    for (uint32_t i = threadIdx.x; i < (uint32_t)(NGROUPS * NSTAGES) * (uint32_t)(Kp - C) * 128u; i += THREADS) {
        const uint32_t st = i / ((uint32_t)(Kp - C) * 128u), o = i - st * (uint32_t)(Kp - C) * 128u;
        const int r = C + (int)(o >> 7);
        reinterpret_cast<float *>(smem + (size_t)st * stage_bytes)[(size_t)r * 128 + (o & 127u)] =
            (kBetaInMma && r == Kp - 1) ? 1.f : 0.f;
    }
    __syncthreads();
    const volatile Fwd2Ctx &t = ctx_s;
    // ---- this group's tile sequence: group-local tile kk <-> CTA tile kk * NGROUPS + group
    const int n_g = (t.n_k - group + NGROUPS - 1) / NGROUPS;  // tiles of this group (may be <= 0)
    const uint32_t ring0 = t.stage0 + (uint32_t)(group * NSTAGES) * stage_bytes;
    const uint32_t fbar0 = t.full_bar0 + 8u * (uint32_t)(group * NSTAGES);
    const uint32_t bytes = (uint32_t)C * 512u;
    auto issue_load = [&](int kk) {
        int b, hw0;
        fwd2_coords(t, NGROUPS, group, kk, &b, &hw0);
        const uint32_t bar = fbar0 + 8u * (uint32_t)(kk % NSTAGES);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(ring0 + (uint32_t)(kk % NSTAGES) * stage_bytes), "l"(reinterpret_cast<uint64_t>(t.tm_x)), "r"(hw0),
              "r"(0), "r"(b), "r"(bar)
            : "memory");
    };
    // the first tile of every group is on its way while gamma is being staged (small layers are all prologue)
    if (leader && n_g > 0) issue_load(0);
    // B operand: Bs[n/8][k/4][n%8][k%4] = tf32(gamma[n][k]) (zero padded; column k = Kp-1 holds beta when kBetaInMma);
    // loads first, eight per thread in flight, then the arithmetic
    {
        constexpr int SU = 8;
        const int total = Np * Kp;
        for (int base = threadIdx.x; base < total; base += THREADS * SU) {
            float raw[SU];
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const int idx = base + u * THREADS;
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
                const int idx = base + u * THREADS;
                if (idx >= total) break;
                const int n = idx / Kp, k = idx - n * Kp;
                float g = (n < C && k < C) ? prm.g_of(raw[u]) : 0.f;
                if (kBetaInMma && k == Kp - 1) g = (n < C) ? prm.b_of(raw[u]) : 1.f;
                const int off = (((n >> 3) * (int)kcores + (k >> 2)) << 5) + ((n & 7) << 2) + (k & 3);
                reinterpret_cast<uint32_t *>(Bs)[off] = to_tf32(g);
            }
        }
        for (int i = threadIdx.x; i < Np; i += THREADS) beta_s[i] = (i < C) ? prm.b(i) : 1.f;
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Np >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t mbar = t.mma_bar0 + 8u * (uint32_t)group;
    const uint32_t a_base = tmem_base_s + (uint32_t)group * (uint32_t)(Kp + Np);
    const uint32_t d_base = a_base + (uint32_t)Kp;
    const uint32_t lane_sel = ((uint32_t)((warp & 3) * 32)) << 16;
    const uint32_t bar_id = 1u + (uint32_t)group;

    uint32_t parity = 0;
#pragma unroll 1
    for (int kk = 0; kk < n_g; ++kk) {
        const uint32_t xs = ring0 + (uint32_t)(kk % NSTAGES) * stage_bytes + (uint32_t)pix * 4u;
        // the stage tile kk + 1 lands in held tile kk + 1 - NSTAGES, whose store was issued NSTAGES - 1 tiles ago:
        // allow the NSTAGES - 2 younger stores to be still reading, then start the next load a full tile ahead
        if (leader && kk + 1 < n_g) {
            if constexpr (NSTAGES >= 3) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSTAGES - 2) : "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            issue_load(kk + 1);
        }
        fwd2_wait(fbar0 + 8u * (uint32_t)(kk % NSTAGES), (uint32_t)((kk / NSTAGES) & 1));
        // ---- A operand: x^2 -> tf32 -> TMEM (lane = pixel, column = channel); channels >= C are padding.
        //      Thread `sub` of the pixel owns the contiguous blocks [sub NB, sub NB + NB): its first channel is folded
        //      into the bases, every other offset is an immediate, and the loads are unconditional (padding rows of the
        //      stage hold 0 / 1) so that all of them are in flight before the first use.
        {
            constexpr int NB = (KP8 + SPLIT - 1) / SPLIT;
            const int cb = sub * NB * 8;
            const uint32_t xc = xs + (uint32_t)cb * 512u;
            float xv[NB * 8];
#pragma unroll
            for (int i = 0; i < NB * 8; ++i) xv[i] = ld_shared_f32(xc + (uint32_t)i * 512u);  // padding rows: 0 / 1
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                uint32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = to_tf32_fast(xv[i * 8 + j] * xv[i * 8 + j]);
                if (SPLIT == 1 || cb + i * 8 < Kp) tmem_st8(a_base + lane_sel + (uint32_t)(cb + i * 8), v);
            }
        }
        tmem_st_wait();
        fence_before();
        named_bar_sync(bar_id, TPG);
        if (leader) {
            fence_after();
            fwd2_mma_chain<0, KP8>(d_base, a_base, desc_lo(t.gamma0, 128), GAMMA_HI, IDESC);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
        fwd2_wait(mbar, parity);
        parity ^= 1;
        fence_after();
        // ---- epilogue: y = x * n^(-+1/2), n = beta + D, written over x in the landing buffer
        const uint32_t beta_a = t.beta0;
        const int nq = ((Np >> 4) + SPLIT - 1) / SPLIT;  // 16-column blocks of D per thread, contiguous like the A blocks
        const int q_end = min(Np, (sub + 1) * nq * 16);
#pragma unroll 1
        for (int q = sub * nq * 16; q < q_end; q += 16) {
            uint32_t r[16];
            tmem_ld16(d_base + lane_sel + (uint32_t)q, r);
            // everything the block needs from shared memory is loaded before the first y is stored: the compiler
            // cannot prove that a store to the landing buffer and a later load do not alias, and would serialise them
            float xq[16], bv[16];
            const uint32_t xc = xs + (uint32_t)q * 512u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                xq[j] = ld_shared_f32(xc + (uint32_t)j * 512u);  // rows >= C: padding or the next buffer, never stored back
                if (!kBetaInMma) bv[j] = ld_shared_f32(beta_a + (uint32_t)(q + j) * 4u);
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (q + j < C) {
                    float n = __uint_as_float(r[j]);
                    if (!kBetaInMma) n += bv[j];
                    const float rs = fast_rsqrt(n);
                    const float y = xq[j] * (inverse ? n * rs : rs);
                    st_shared_u32(xc + (uint32_t)j * 512u, __float_as_uint(y));
                }
            }
        }
        fence_async_smem();  // y (generic-proxy writes) -> visible to the TMA store
        fence_before();
        named_bar_sync(bar_id, TPG);
        fence_after();
        if (leader) {
            int b, hw0;
            fwd2_coords(t, NGROUPS, group, kk, &b, &hw0);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                             reinterpret_cast<uint64_t>(t.tm_y)),
                         "r"(hw0), "r"(0), "r"(b), "r"(ring0 + (uint32_t)(kk % NSTAGES) * stage_bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // shared memory must outlive the stores
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, tmem_cols);
}

// ------------------------------------------------------------------------------------------------------ host side
struct Fwd2Geometry {
    int Kp, Np, groups, stages, split;
    uint32_t tmem_cols;
    size_t smem;
};

// groups / stages are a function of KP8 = ceil(C / 8) alone (sized for the largest C of the bucket), so that there is
// one kernel instance per bucket: three stages per group whenever they fit, as many groups as fit (at most 4)
constexpr size_t FWD2_BUDGET = 227 * 1024 - 1024 - 512;
constexpr size_t fwd2_gamma_bytes(int kp8) {
    const size_t Kp = (size_t)kp8 * 8, Np = (Kp + 15) / 16 * 16;
    return Np * Kp * 4 + Np * 4;
}
constexpr int fwd2_stages(int kp8) {
    return (fwd2_gamma_bytes(kp8) + 3 * (size_t)kp8 * 8 * 512 <= FWD2_BUDGET) ? 3 : 2;
}
constexpr int fwd2_groups(int kp8) {
    const size_t Kp = (size_t)kp8 * 8, Np = (Kp + 15) / 16 * 16;
    if (fwd2_stages(kp8) == 2) return 1;
    int g = (int)((FWD2_BUDGET - fwd2_gamma_bytes(kp8)) / (3 * Kp * 512));
    if (g > 4) g = 4;
    while (g > 1 && (size_t)g * (Kp + Np) > 512) --g;
    return g;
}

constexpr int fwd2_split(int kp8) {
    const int g = fwd2_groups(kp8);
    return g >= 3 ? 1 : (g == 2 ? 2 : 4);
}

static bool fwd2_geometry(int64_t C, Fwd2Geometry *g) {
    if (C < 16 || C > 128) return false;
    g->Kp = (int)((C + 7) / 8 * 8);
    g->Np = (int)((C + 15) / 16 * 16);
    const int kp8 = g->Kp / 8;
    if (fwd2_gamma_bytes(kp8) + 2 * (size_t)g->Kp * 512 > FWD2_BUDGET) return false;
    g->groups = fwd2_groups(kp8);
    g->stages = fwd2_stages(kp8);
    g->split = fwd2_split(kp8);
    uint32_t cols = 32;
    while ((int)cols < g->groups * (g->Kp + g->Np)) cols <<= 1;
    g->tmem_cols = cols;
    // A sub-thread whose column range runs past Kp, and the epilogue's rows C .. Np-1, read a few rows beyond their
    // stage - never used, but they must stay inside the allocation for the last stage too: gamma follows the stages,
    // and a tail is added where gamma is smaller than the over-read.
    const int split = fwd2_split(kp8);
    const int nb = (kp8 + split - 1) / split;
    const int last_row = (nb * 8 * split > g->Np) ? nb * 8 * split : g->Np;
    const size_t over = (size_t)(last_row - g->Kp) * 512;
    const size_t tail = over > fwd2_gamma_bytes(kp8) ? over - fwd2_gamma_bytes(kp8) : 0;
    g->smem = (size_t)g->groups * g->stages * (size_t)g->Kp * 512 + fwd2_gamma_bytes(kp8) + tail + 1024;
    if (g->smem + 512 > 227 * 1024) return false;
    return true;
}

static int fwd2_mode() {  // MMNC_GDN_FWD = v1 | v2 | auto (default)
    static int mode = []() {
        const char *e = getenv("MMNC_GDN_FWD");
        if (e && !strcmp(e, "v1")) return 1;
        if (e && !strcmp(e, "v2")) return 2;
        return 0;
    }();
    return mode;
}

bool gdn_tc_forward2_supported(const float *x, const float *y, int64_t B, int64_t C, int64_t HW) {
    Fwd2Geometry geo;
    if (fwd2_mode() == 1) return false;
    if (!fwd2_geometry(C, &geo)) return false;
    if (geo.stages < 3) return false;  // two-stage rings (C > 112) measured slower than gdn_tc.cu (0.53 vs 0.61 of the roof)
    if (HW % tcf2::TILE != 0 || HW >= (1 << 24) || B >= (1 << 24) || B * HW / tcf2::TILE >= (1ll << 31)) return false;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return false;
    return tmah::encode_tiled() != nullptr;
}

static int fwd2_make_map(CUtensorMap *m, const float *p, int64_t B, int64_t C, int64_t HW) {
    return tmah::tensor_map_3d(m, p, (uint64_t)HW, (uint64_t)C, (uint64_t)B, (uint64_t)HW * 4, (uint64_t)C * HW * 4,
                               (uint32_t)tcf2::TILE, (uint32_t)C, 1, CU_TENSOR_MAP_SWIZZLE_NONE, "gdn_tc_forward2");
}

int gdn_tc_forward2(const float *x, int64_t B, int64_t C, int64_t HW, const GdnParams &prm, int inverse, float *y,
                    cudaStream_t s) {
    Fwd2Geometry geo;
    if (!fwd2_geometry(C, &geo)) {
        set_error("gdn_tc_forward2: C = %lld not supported", (long long)C);
        return MMNC_ERR_UNSUPPORTED;
    }
    const int64_t ntiles = B * HW / tcf2::TILE;
    int64_t grid = sm_count();
    if (grid > ntiles) grid = ntiles;
    CUtensorMap tm_x, tm_y;
    if (int rc = fwd2_make_map(&tm_x, x, B, C, HW)) return rc;
    if (int rc = fwd2_make_map(&tm_y, y, B, C, HW)) return rc;
    using Kernel = void (*)(const CUtensorMap, const CUtensorMap, int, int, const GdnParams, int, int, int, uint32_t);
    Kernel kernel = nullptr;
    const bool bim = (C % 8) != 0;
    // threads per pixel: measured on B200 (GDN(50/64)@256^2, GDN(100/128)@128^2): 2 for two groups, 4 for one group
    const int split = geo.split;
#define MMNC_F2_CASE(N)                                                                                                 \
    case N:                                                                                                             \
        kernel = bim ? (Kernel)gdn_tc_forward2_kernel<N, fwd2_groups(N), fwd2_stages(N), fwd2_split(N), true>           \
                     : (Kernel)gdn_tc_forward2_kernel<N, fwd2_groups(N), fwd2_stages(N), fwd2_split(N), false>;         \
        break;
    switch (geo.Kp / 8) {
        MMNC_F2_CASE(2) MMNC_F2_CASE(3) MMNC_F2_CASE(4) MMNC_F2_CASE(5) MMNC_F2_CASE(6) MMNC_F2_CASE(7) MMNC_F2_CASE(8)
        MMNC_F2_CASE(9) MMNC_F2_CASE(10) MMNC_F2_CASE(11) MMNC_F2_CASE(12) MMNC_F2_CASE(13) MMNC_F2_CASE(14)
        default: break;  // C > 112 leaves room for two stages only: measured slower than gdn_tc.cu, not instantiated
    }
#undef MMNC_F2_CASE
    if (kernel == nullptr) {
        set_error("gdn_tc_forward2: no kernel instance for C = %lld", (long long)C);
        return MMNC_ERR_UNSUPPORTED;
    }
    if (int rc = tmah::ensure_dynamic_smem(kernel, geo.smem)) return rc;
    kernel<<<(unsigned)grid, geo.groups * split * 128, geo.smem, s>>>(tm_x, tm_y, (int)ntiles, (int)(HW / tcf2::TILE), prm,
                                                                   inverse, (int)C, geo.Np, geo.tmem_cols);
    return after_launch("gdn_tc_forward2_kernel");
}

}  // namespace mmnc
