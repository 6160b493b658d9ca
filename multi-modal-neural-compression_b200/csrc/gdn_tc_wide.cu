// K4 (wide layers): GDN / IGDN with 129 .. 256 channels on tcgen05 (the IGDN(256) sites of
// BASELINE config C3).  Single-pass TF32, NCHW.
//
// With 256 channels nothing of the <= 128 channel kernels' residency survives: gamma alone is 256 KB (more than shared
// memory), a pixel's channels are 256 registers, and A (256 columns) + D (256 columns) fill tensor memory, so the
// d gamma accumulator of the fused backward has no room.  The layer is therefore split differently:
//
//   forward   y = x n^p,  n = beta + x^2 gamma^T                        gdn_wide_forward_kernel   (8 B / element)
//   backward  u = p g x n^(p-1),  dx = g n^p + 2 x (u gamma)            gdn_wide_dx_kernel        (16 B / element: x, g, dx
//                                                                                                  and u, which goes to HBM)
//             d gamma = u^T x^2,  d beta = sum u                        gdn_wide_dgamma_kernel    (8 B / element: u, x)
//
// Pixel-tile kernels (forward, dx): one CTA = 128 pixels = the 128 TMEM lanes, FOUR threads per pixel, each owning 64
// channels (64 TMEM columns of A and of D).  x (and g) are read straight from global memory - a warp reads 32 consecutive
// pixels of one channel, fully coalesced - x^2 / u go to TMEM with tcgen05.st as the A operand.  The B operand (gamma, then
// gamma^T) is STREAMED: a pack kernel writes both once per call in the K-major core-matrix layout, cut into K chunks of 32
// contraction channels (32 KB each); the MMA-issuing thread moves the chunks with cp.async.bulk through a ring of six
// shared-memory slots (full / empty mbarriers; tcgen05.commit releases a slot when the MMAs that read it retire).  The
// packed images are 2 x 256 KB and stay in L2.
//
// d gamma kernel: a split-K GEMM over pixels.  u and x are K-major for it as they lie in NCHW (K = pixel, contiguous), so
// TMA boxes of [32 pixels x C channels] with the 128-byte swizzle are the operands as they land; the CTA squares the x box
// in place, sums the rows of the u box (d beta), and issues 2 (halves of the 256 output channels) x 4 (K = 8 pixels)
// MMAs with N = 256 into a 2 x 256 column accumulator that lives in TMEM for the whole kernel.  Every CTA writes one
// [C][C + 1] partial; gdn_reduce_partials (fixed order) adds them and applies the re-parametrisation's chain rule.
#include <string.h>

#include "common.cuh"
#include "gdn_params.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace mmnc {

namespace tcw {

constexpr int TILE = 128;                      // pixels per tile = TMEM lanes
constexpr int P = 256;                         // padded channel count: N of both contractions, TMEM columns of A and of D
constexpr int KC = 32;                         // contraction channels per streamed chunk
constexpr int SLOTS = 6;                       // shared-memory ring of chunks
constexpr uint32_t CHUNK_BYTES = P * KC * 4;   // 32 KB
constexpr int QC = 64;                         // channels per thread (four threads per pixel)
constexpr int COMPUTE = 4 * TILE;              // compute threads
constexpr int THREADS = COMPUTE;
constexpr uint32_t B_HI = tc::desc_hi((KC / 4) * 128u, 0);  // K-major, no swizzle: SBO = the KC / 4 cores of an 8-row group
constexpr uint32_t PIX_HI = tc::desc_hi(1024u, 2);          // K-major (K = pixel), 128-byte swizzle
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

constexpr int PX = 32;                         // d gamma kernel: pixels per chunk = one 128-byte swizzle atom along K
constexpr int STAGES = 3;
constexpr uint32_t BOX_BYTES = P * PX * 4;     // one landed operand: 256 rows x 128 B
constexpr int DG_THREADS = 256;

__device__ __forceinline__ void wait_bar(uint32_t addr, uint32_t parity) {
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
__device__ __forceinline__ void commit_bar(uint32_t addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void expect_bytes(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// The ring is driven by ONE thread (the MMA-issuing leader), whose state lives in shared memory so that it costs the
// other 511 threads no registers.  Chunk q of the CTA's sequence goes to slot q % SLOTS; the sequence repeats the call's
// `per_tile` chunks (gamma, then gamma^T for the backward) once per tile.
struct Ring {
    uint32_t base, full0, empty0;    // shared addresses: slots, "chunk landed" barriers, "slot free" barriers
    const uint32_t *packed;
    int per_tile, total;             // chunks per tile, chunks of this CTA's whole run
    int q_prod, q_cons, c;           // produced / consumed so far; position inside the per-tile sequence
    uint32_t pslot, pphase, cslot, cphase;
};

__device__ __forceinline__ void produce_one(Ring *r) {
    const uint32_t slot = r->pslot;
    // the MMAs that read the slot's previous occupant have retired (tcgen05.commit arrives on the barrier)
    if (r->q_prod >= SLOTS) wait_bar(r->empty0 + 8u * slot, r->pphase ^ 1u);
    const uint32_t bar = r->full0 + 8u * slot, dst = r->base + slot * CHUNK_BYTES;
    const uint64_t src = reinterpret_cast<uint64_t>(r->packed) + (uint64_t)r->c * CHUNK_BYTES;
    expect_bytes(bar, CHUNK_BYTES);
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + i * (CHUNK_BYTES / 4)), "l"(src + i * (CHUNK_BYTES / 4)), "r"(CHUNK_BYTES / 4), "r"(bar)
                     : "memory");
    if (++r->c == r->per_tile) r->c = 0;
    if (++r->pslot == SLOTS) { r->pslot = 0; r->pphase ^= 1u; }
    ++r->q_prod;
}

// Fill the ring as far as it goes.  Called when every issued MMA has retired (all slots free): never blocks.
__device__ __forceinline__ void top_up(Ring *r) {
    while (r->q_prod < r->total && r->q_prod < r->q_cons + SLOTS) produce_one(r);
}

template <int KS>
__device__ __forceinline__ void chunk_mma(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t first_acc) {
    if constexpr (KS < KC / 8) {
        tc::mma_tf32_ts_step<KS * 8, KS * 16>(d, a, b_lo, B_HI, IDESC, KS == 0 ? first_acc : 1u);
        chunk_mma<KS + 1>(d, a, b_lo, first_acc);
    }
}

// One contraction (leader only): D[128 x 256] = A[128 x 32 n_kc] * B, chunk by chunk.  Everything already in the ring is
// issued back to back; when the ring runs dry (n_kc > SLOTS) the remaining chunks are requested in order, each as soon as
// the slot it reuses is released - by then the tensor pipe still has several chunks queued, which covers the L2 latency.
__device__ __forceinline__ void contract(uint32_t tmem_base, Ring *r, uint32_t mma_bar, int n_kc) {
    int remaining = n_kc;
#pragma unroll 1
    for (int kc = 0; kc < n_kc; ++kc, --remaining) {
        if (r->q_cons == r->q_prod) {
            const int m = remaining < SLOTS ? remaining : SLOTS;
            for (int i = 0; i < m; ++i) produce_one(r);
        }
        const uint32_t slot = r->cslot;
        wait_bar(r->full0 + 8u * slot, r->cphase);
        chunk_mma<0>(tmem_base + P, tmem_base + (uint32_t)(kc * KC), tc::desc_lo(r->base + slot * CHUNK_BYTES, 128), kc > 0 ? 1u : 0u);
        commit_bar(r->empty0 + 8u * slot);
        if (++r->cslot == SLOTS) { r->cslot = 0; r->cphase ^= 1u; }
        ++r->q_cons;
    }
    commit_bar(mma_bar);
}

struct Setup {
    uint32_t mma_bar, tmem_base;
    int n_tiles;  // of this CTA
};

}  // namespace tcw

// Both images of the streamed operand, chunked: out[chunk][n / 8][kl / 4][n % 8][kl % 4] with kl = k - 32 chunk.
//   image 0 (chunks 0 .. n_kc)        gamma    N = out channel i, K = in channel j
//   image 1 (chunks n_kc .. 2 n_kc)   gamma^T  N = in channel j,  K = out channel i
__global__ void __launch_bounds__(256)
gdn_wide_pack_kernel(const GdnParams prm, int C, int n_kc, int images, uint32_t *__restrict__ out) {
    using namespace tcw;
    const int per_image = n_kc * KC * P;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < per_image * images; idx += gridDim.x * blockDim.x) {
        const int image = idx / per_image, e = idx - image * per_image;
        const int n = e / (n_kc * KC), k = e - n * (n_kc * KC);  // consecutive threads walk K: coalesced reads of gamma rows
        float v = 0.f;
        if (n < C && k < C) v = image ? prm.g((int64_t)k * C + n) : prm.g((int64_t)n * C + k);
        const int chunk = k / KC, kl = k - chunk * KC;
        const int off = (((n >> 3) * (KC / 4) + (kl >> 2)) << 5) + ((n & 7) << 2) + (kl & 3);
        out[(size_t)(image * n_kc + chunk) * (CHUNK_BYTES / 4) + off] = tc::to_tf32(v);
    }
}

// Shared start-up of the two pixel-tile kernels.
__device__ __forceinline__ void wide_setup(tcw::Setup &st, tcw::Ring *ring, uint8_t *smem_raw, uint64_t *bars, uint32_t *tmem_slot,
                                           float *beta_s, const GdnParams &prm, int C, int64_t NP, const uint32_t *packed,
                                           int per_tile) {
    using namespace tc;
    using namespace tcw;
    const int64_t tiles = (NP + TILE - 1) / TILE;
    st.n_tiles = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    if ((threadIdx.x >> 5) == 0) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * SLOTS + 1; ++i) mbar_init(&bars[i], 1);
        ring->base = (smem_u32(smem_raw) + 127u) & ~127u;
        ring->full0 = smem_u32(&bars[0]);
        ring->empty0 = smem_u32(&bars[SLOTS]);
        ring->packed = packed;
        ring->per_tile = per_tile;
        ring->total = per_tile * st.n_tiles;
        ring->q_prod = ring->q_cons = ring->c = 0;
        ring->pslot = ring->pphase = ring->cslot = ring->cphase = 0u;
        top_up(ring);  // the first chunks are on their way while the tile's x is loaded
    }
    for (int i = threadIdx.x; i < P; i += THREADS) beta_s[i] = (i < C) ? prm.b(i) : 1.f;
    fence_before();
    __syncthreads();
    fence_after();
    st.mma_bar = smem_u32(&bars[2 * SLOTS]);
    st.tmem_base = *tmem_slot;
}

template <bool kInverse>
__global__ void __launch_bounds__(tcw::THREADS, 1)
gdn_wide_forward_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t NP, int64_t HW, int C, const GdnParams prm,
                        const uint32_t *__restrict__ packed, int n_kc) {
    using namespace tc;
    using namespace tcw;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * SLOTS + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ float beta_s[P];
    __shared__ Ring ring;
    Setup st;
    wide_setup(st, &ring, smem_raw, bars, &tmem_slot, beta_s, prm, C, NP, packed, n_kc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const int q4 = warp >> 2;                          // which 64 channels
        const int pl = ((warp & 3) << 5) | lane;           // pixel of the tile = TMEM lane
        const int creal = min(QC, max(0, C - q4 * QC));    // real channels among this thread's 64
        const uint32_t lane_a = st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(q4 * QC);
        const uint32_t lane_d = lane_a + (uint32_t)P;
        const uint32_t sb = (uint32_t)HW * 4u;
        uint32_t parity = 0;
#pragma unroll 1
        for (int t = 0; t < st.n_tiles; ++t) {
            const int64_t pix = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE + pl;
            const bool valid = pix < NP;
            const int64_t b = valid ? pix / HW : 0;
            const int64_t e0 = (b * C + q4 * QC) * HW + (valid ? pix - b * HW : 0);
            const float *xb = x + e0;
            float *yb = y + e0;
            float xv[QC];
#pragma unroll
            for (int c = 0; c < QC; ++c) xv[c] = (valid && c < creal) ? __ldcs(chan_ptr(xb, sb, c)) : 0.f;
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += 16) {
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = to_tf32_fast(xv[c0 + j] * xv[c0 + j]);
                tmem_st16(lane_a + c0, v);
            }
            tmem_st_wait();
            fence_before();
            named_bar_sync(1, COMPUTE);
            if (threadIdx.x == 0) {
                fence_after();
                contract(st.tmem_base, &ring, st.mma_bar, n_kc);
            }
            wait_bar(st.mma_bar, parity);
            parity ^= 1u;
            fence_after();
            if (threadIdx.x == 0) top_up(&ring);  // every slot is free: the next contraction's chunks land during the epilogue
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(lane_d + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float n = __uint_as_float(r[j]) + beta_s[q4 * QC + c0 + j];
                    const float rs = fast_rsqrt(n);
                    const float out = xv[c0 + j] * (kInverse ? n * rs : rs);
                    if (valid && c0 + j < creal) __stcs(chan_ptr(yb, sb, c0 + j), out);
                }
            }
            // every lane has drained D before the next tile's MMA overwrites it
            fence_before();
            named_bar_sync(1, COMPUTE);
            fence_after();
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(st.tmem_base, 512);
}

// dx and u.  Registers: only ONE 64-channel vector stays live per thread - g, turned in place into f = g n^p by epilogue 1
// and consumed by epilogue 2; x is read three times (A fill: HBM; the two epilogues: L2), eight channels at a time and one
// block ahead.  g is requested while MMA1 runs.
template <bool kInverse>
__global__ void __launch_bounds__(tcw::THREADS, 1)
gdn_wide_dx_kernel(const float *__restrict__ x, const float *__restrict__ g, float *__restrict__ dx, float *__restrict__ U,
                   int64_t NP, int64_t HW, int C, const GdnParams prm, const uint32_t *__restrict__ packed, int n_kc) {
    using namespace tc;
    using namespace tcw;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * SLOTS + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ float beta_s[P];
    __shared__ Ring ring;
    Setup st;
    wide_setup(st, &ring, smem_raw, bars, &tmem_slot, beta_s, prm, C, NP, packed, 2 * n_kc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr float coef = kInverse ? 0.5f : -0.5f;
    {
        const int q4 = warp >> 2;
        const int pl = ((warp & 3) << 5) | lane;
        const int creal = min(QC, max(0, C - q4 * QC));
        const uint32_t lane_a = st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(q4 * QC);
        const uint32_t lane_d = lane_a + (uint32_t)P;
        const uint32_t sb = (uint32_t)HW * 4u;
        const float *beta_q = beta_s + q4 * QC;
        uint32_t parity = 0;
#pragma unroll 1
        for (int t = 0; t < st.n_tiles; ++t) {
            const int64_t pix = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE + pl;
            const bool valid = pix < NP;
            const int64_t b = valid ? pix / HW : 0;
            const int64_t e0 = (b * C + q4 * QC) * HW + (valid ? pix - b * HW : 0);
            const float *xb = x + e0, *gb = g + e0;
            float *dxb = dx + e0, *ub = U + e0;
            int cr = creal;
            asm volatile("" : "+r"(cr));  // keeps the 64 channel predicates from being hoisted out of the tile loop
            // ---- A = x^2
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += 16) {
                float xv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) xv[j] = (valid && c0 + j < cr) ? __ldg(chan_ptr(xb, sb, c0 + j)) : 0.f;
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = to_tf32_fast(xv[j] * xv[j]);
                tmem_st16(lane_a + c0, v);
            }
            tmem_st_wait();
            fence_before();
            named_bar_sync(1, COMPUTE);
            if (threadIdx.x == 0) {
                fence_after();
                contract(st.tmem_base, &ring, st.mma_bar, n_kc);  // n - beta = x^2 gamma^T
            }
            // ---- g on its way while MMA1 runs
            float gf[QC];
            asm volatile("" : "+r"(cr));
#pragma unroll
            for (int c = 0; c < QC; ++c) gf[c] = (valid && c < cr) ? __ldcs(chan_ptr(gb, sb, c)) : 0.f;
            float xn[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) xn[j] = (valid && j < cr) ? __ldg(chan_ptr(xb, sb, j)) : 0.f;
            wait_bar(st.mma_bar, parity);
            parity ^= 1u;
            fence_after();
            if (threadIdx.x == 0) top_up(&ring);  // every slot is free: the next contraction's chunks land during the epilogue
            // ---- epilogue 1: u -> A and -> HBM (the d gamma kernel's operand); f = g n^p replaces g
            asm volatile("" : "+r"(cr));
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += 8) {
                float xc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) xc[j] = xn[j];
                if (c0 + 8 < QC) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) xn[j] = (valid && c0 + 8 + j < cr) ? __ldg(chan_ptr(xb, sb, c0 + 8 + j)) : 0.f;
                }
                uint32_t r[8], uu[8];
                tmem_ld8(lane_d + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float n = __uint_as_float(r[j]) + beta_q[c0 + j];
                    const float rs = fast_rsqrt(n);
                    const float pw = kInverse ? n * rs : rs;         // n^p
                    const float pm1 = kInverse ? rs : rs * rs * rs;  // n^(p-1)
                    const float gv = gf[c0 + j];
                    uu[j] = to_tf32_fast((coef * gv) * (xc[j] * pm1));  // padding channels: g = 0, so u = 0
                    gf[c0 + j] = gv * pw;
                    if (valid && c0 + j < cr) chan_ptr(ub, sb, c0 + j)[0] = __uint_as_float(uu[j]);
                }
                tmem_st8(lane_a + c0, uu);
            }
            tmem_st_wait();
            fence_before();
            named_bar_sync(1, COMPUTE);
            if (threadIdx.x == 0) {
                fence_after();
                contract(st.tmem_base, &ring, st.mma_bar, n_kc);  // t = u gamma
            }
            asm volatile("" : "+r"(cr));
#pragma unroll
            for (int j = 0; j < 8; ++j) xn[j] = (valid && j < cr) ? __ldg(chan_ptr(xb, sb, j)) : 0.f;
            wait_bar(st.mma_bar, parity);
            parity ^= 1u;
            fence_after();
            if (threadIdx.x == 0) top_up(&ring);  // every slot is free: the next contraction's chunks land during the epilogue
            // ---- epilogue 2: dx = f + 2 x t
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += 8) {
                float xc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) xc[j] = xn[j];
                if (c0 + 8 < QC) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) xn[j] = (valid && c0 + 8 + j < cr) ? __ldg(chan_ptr(xb, sb, c0 + 8 + j)) : 0.f;
                }
                uint32_t r[8];
                tmem_ld8(lane_d + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float out = fmaf(2.f * xc[j], __uint_as_float(r[j]), gf[c0 + j]);
                    if (valid && c0 + j < cr) __stcs(chan_ptr(dxb, sb, c0 + j), out);
                }
            }
            // No barrier here: the next tile's barrier (after its A fill) orders these TMEM reads of D before the next MMA1
            // overwrites it, and A was last read by MMA2, which has retired.
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(st.tmem_base, 512);
}

// d gamma / d beta partials of one CTA over its share of the 32-pixel chunks.
__global__ void __launch_bounds__(tcw::DG_THREADS, 1)
gdn_wide_dgamma_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_x, int n_chunks,
                       int chunks_per_img, int C, float *__restrict__ part) {
    using namespace tc;
    using namespace tcw;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], free_bar[STAGES], done_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t stage0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *stage0_p = smem_raw + (stage0 - smem_u32(smem_raw));
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&free_bar[s], 1); }
        mbar_init(&done_bar, 1);
        tma_prefetch_desc(&tm_u);
        tma_prefetch_desc(&tm_x);
    }
    // rows >= C of every landing buffer are never written by the TMA boxes (C rows): zero them once
    {
        float4 *z = reinterpret_cast<float4 *>(stage0_p);
        for (int i = threadIdx.x; i < (int)(STAGES * 2 * BOX_BYTES / 16); i += DG_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int n_my = (n_chunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t full0 = smem_u32(&full_bar[0]), free0 = smem_u32(&free_bar[0]);
    const uint32_t box_bytes = (uint32_t)C * (PX * 4u);
    auto issue = [&](int k) {
        const int s = k % STAGES;
        const int chunk = (int)blockIdx.x + k * (int)gridDim.x;
        const int b = chunk / chunks_per_img, hw0 = (chunk - b * chunks_per_img) * PX;
        const uint32_t bar = full0 + 8u * (uint32_t)s, ub = stage0 + (uint32_t)s * 2u * BOX_BYTES;
        expect_bytes(bar, 2u * box_bytes);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(ub), "l"(reinterpret_cast<uint64_t>(&tm_u)), "r"(hw0), "r"(0), "r"(b), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(ub + BOX_BYTES), "l"(reinterpret_cast<uint64_t>(&tm_x)), "r"(hw0), "r"(0), "r"(b), "r"(bar) : "memory");
    };
    if (threadIdx.x == 0)
        for (int k = 0; k < STAGES && k < n_my; ++k) issue(k);
    float dbeta = 0.f;
    const int row = threadIdx.x;  // this thread's channel for d beta
    const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
#pragma unroll 1
    for (int k = 0; k < n_my; ++k) {
        const int s = k % STAGES;
        wait_bar(full0 + 8u * (uint32_t)s, (uint32_t)((k / STAGES) & 1));
        uint8_t *ubuf = stage0_p + (size_t)s * 2 * BOX_BYTES;
        // x -> x^2 (tf32) in place: an element-wise pass does not care about the swizzle
        float4 *xq = reinterpret_cast<float4 *>(ubuf + BOX_BYTES);
#pragma unroll
        for (int i = 0; i < (int)(BOX_BYTES / 16 / DG_THREADS); ++i) {
            float4 v = xq[threadIdx.x + i * DG_THREADS];
            v.x = __uint_as_float(to_tf32_fast(v.x * v.x));
            v.y = __uint_as_float(to_tf32_fast(v.y * v.y));
            v.z = __uint_as_float(to_tf32_fast(v.z * v.z));
            v.w = __uint_as_float(to_tf32_fast(v.w * v.w));
            xq[threadIdx.x + i * DG_THREADS] = v;
        }
        // d beta: the 32 pixels of this thread's row of u (the swizzle permutes 16-byte pieces inside the row only; the
        // rotation keeps the eight threads of a quarter warp on different banks)
        {
            const uint8_t *urow = ubuf + row_off;
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 v = *reinterpret_cast<const float4 *>(urow + (((j + row) & 7) << 4));
                a += (v.x + v.y) + (v.z + v.w);
            }
            dbeta += a;
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after();
            const uint32_t ub = stage0 + (uint32_t)s * 2u * BOX_BYTES, xb = ub + BOX_BYTES;
            const uint32_t b_lo = desc_lo(xb, 16);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t a_lo = desc_lo(ub + (uint32_t)h * (BOX_BYTES / 2), 16);
                const uint32_t d = tmem_base + (uint32_t)(h * P);
                mma_tf32_ss_step<0>(d, a_lo, b_lo, PIX_HI, 0u, IDESC, k > 0 ? 1u : 0u);
                mma_tf32_ss_step<2>(d, a_lo, b_lo, PIX_HI, 0u, IDESC, 1u);
                mma_tf32_ss_step<4>(d, a_lo, b_lo, PIX_HI, 0u, IDESC, 1u);
                mma_tf32_ss_step<6>(d, a_lo, b_lo, PIX_HI, 0u, IDESC, 1u);
            }
            commit_bar(free0 + 8u * (uint32_t)s);
            // the stage of the previous chunk is free once its MMAs retired: it takes the chunk two ahead
            if (k >= 1 && k + 2 < n_my) {
                wait_bar(free0 + 8u * (uint32_t)((k - 1) % STAGES), (uint32_t)(((k - 1) / STAGES) & 1));
                issue(k + 2);
            }
        }
    }
    if (threadIdx.x == 0) commit_bar(smem_u32(&done_bar));
    wait_bar(smem_u32(&done_bar), 0u);
    fence_after();
    // ---- the CTA's partial: [C][C + 1], column C = d beta
    float *mine = part + (size_t)blockIdx.x * C * (C + 1);
    {
        const int h = warp >> 2, lane_row = ((warp & 3) << 5) | lane, i = h * 128 + lane_row;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(h * P);
        float *dst = mine + (size_t)i * (C + 1);
        for (int j0 = 0; j0 < C; j0 += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + j0, r);
            tmem_ld_wait();
            if (i < C) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j0 + j < C) dst[j0 + j] = __uint_as_float(r[j]);
            }
        }
    }
    if (row < C) mine[(size_t)row * (C + 1) + C] = dbeta;
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// gdn_simt.cu
int gdn_reduce_partials(const float *part, int ksplit, int C, const GdnParams &prm, float *dgamma, float *dbeta,
                        cudaStream_t s);

// ------------------------------------------------------------------------------------------------------ host side
static bool wide_enabled() {
    static const bool on = []() { const char *e = getenv("MMNC_GDN_WIDE"); return !(e && !strcmp(e, "0")); }();
    return on;
}
static int wide_kchunks(int64_t C) { return (int)((C + tcw::KC - 1) / tcw::KC); }
static size_t align128(size_t v) { return (v + 127) / 128 * 128; }

static bool wide_shape_ok(int64_t B, int64_t C, int64_t HW) {
    return wide_enabled() && C > 128 && C <= tcw::P && B * HW >= 4096 && HW < (1 << 24) && B < (1 << 24) &&
           B * HW / tcw::PX < (1ll << 31) && B * C < (1ll << 31);
}

size_t gdn_tc_wide_forward_workspace(int64_t B, int64_t C, int64_t HW) {
    if (!wide_shape_ok(B, C, HW)) return 0;
    return (size_t)wide_kchunks(C) * tcw::CHUNK_BYTES + 256;
}

bool gdn_tc_wide_forward_supported(int64_t B, int64_t C, int64_t HW, const void *workspace, size_t workspace_bytes) {
    return wide_shape_ok(B, C, HW) && workspace != nullptr && workspace_bytes >= gdn_tc_wide_forward_workspace(B, C, HW);
}

static constexpr size_t WIDE_RING_SMEM = (size_t)tcw::SLOTS * tcw::CHUNK_BYTES + 128;

int gdn_tc_wide_forward(const float *x, int64_t B, int64_t C, int64_t HW, const GdnParams &prm, int inverse, float *y,
                        void *workspace, size_t workspace_bytes, cudaStream_t s) {
    MMNC_REQUIRE(gdn_tc_wide_forward_supported(B, C, HW, workspace, workspace_bytes), "gdn_tc_wide_forward: unsupported call");
    const int n_kc = wide_kchunks(C);
    uint32_t *packed = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(workspace) + 127) / 128 * 128);
    const int elems = n_kc * tcw::KC * tcw::P;
    gdn_wide_pack_kernel<<<(elems + 255) / 256, 256, 0, s>>>(prm, (int)C, n_kc, 1, packed);
    if (int rc = after_launch("gdn_wide_pack_kernel")) return rc;
    auto kernel = inverse ? gdn_wide_forward_kernel<true> : gdn_wide_forward_kernel<false>;
    if (int rc = tmah::ensure_dynamic_smem(kernel, WIDE_RING_SMEM)) return rc;
    const int64_t NP = B * HW, tiles = (NP + tcw::TILE - 1) / tcw::TILE;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    kernel<<<(unsigned)grid, tcw::THREADS, WIDE_RING_SMEM, s>>>(x, y, NP, HW, (int)C, prm, packed, n_kc);
    return after_launch("gdn_wide_forward_kernel");
}

bool gdn_tc_wide_backward_supported(const float *x, const float *g, int64_t B, int64_t C, int64_t HW) {
    (void)g;
    if (!wide_shape_ok(B, C, HW) || HW % tcw::PX != 0) return false;
    if (reinterpret_cast<uintptr_t>(x) & 15) return false;
    return tmah::encode_tiled() != nullptr;
}

// [partials | packed gamma, gamma^T | u]
size_t gdn_tc_wide_backward_workspace(int64_t B, int64_t C, int64_t HW) {
    if (!wide_shape_ok(B, C, HW)) return 0;
    return align128(sizeof(float) * (size_t)sm_count() * C * (C + 1)) + align128(2 * (size_t)wide_kchunks(C) * tcw::CHUNK_BYTES) +
           align128(sizeof(float) * (size_t)(B * C * HW)) + 256;
}

int gdn_tc_wide_backward(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const GdnParams &prm, int inverse,
                         float *dx, float *dbeta, float *dgamma, void *workspace, size_t workspace_bytes, cudaStream_t s) {
    MMNC_REQUIRE(gdn_tc_wide_backward_supported(x, g, B, C, HW), "gdn_tc_wide_backward: unsupported call");
    MMNC_REQUIRE(workspace_bytes >= gdn_tc_wide_backward_workspace(B, C, HW), "gdn_backward: workspace too small");
    const int n_kc = wide_kchunks(C);
    char *base = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(workspace) + 127) / 128 * 128);
    float *part = reinterpret_cast<float *>(base);
    uint32_t *packed = reinterpret_cast<uint32_t *>(base + align128(sizeof(float) * (size_t)sm_count() * C * (C + 1)));
    float *U = reinterpret_cast<float *>(reinterpret_cast<char *>(packed) + align128(2 * (size_t)n_kc * tcw::CHUNK_BYTES));
    const int elems = 2 * n_kc * tcw::KC * tcw::P;
    gdn_wide_pack_kernel<<<(elems + 255) / 256, 256, 0, s>>>(prm, (int)C, n_kc, 2, packed);
    if (int rc = after_launch("gdn_wide_pack_kernel")) return rc;
    {
        auto kernel = inverse ? gdn_wide_dx_kernel<true> : gdn_wide_dx_kernel<false>;
        if (int rc = tmah::ensure_dynamic_smem(kernel, WIDE_RING_SMEM)) return rc;
        const int64_t NP = B * HW, tiles = (NP + tcw::TILE - 1) / tcw::TILE;
        int64_t grid = sm_count();
        if (grid > tiles) grid = tiles;
        kernel<<<(unsigned)grid, tcw::THREADS, WIDE_RING_SMEM, s>>>(x, g, dx, U, NP, HW, (int)C, prm, packed, n_kc);
        if (int rc = after_launch("gdn_wide_dx_kernel")) return rc;
    }
    CUtensorMap tm_u, tm_x;
    if (int rc = tmah::tensor_map_3d(&tm_u, U, (uint64_t)HW, (uint64_t)C, (uint64_t)B, (uint64_t)HW * 4, (uint64_t)C * HW * 4,
                                     tcw::PX, (uint32_t)C, 1, CU_TENSOR_MAP_SWIZZLE_128B, "gdn_tc_wide_backward"))
        return rc;
    if (int rc = tmah::tensor_map_3d(&tm_x, x, (uint64_t)HW, (uint64_t)C, (uint64_t)B, (uint64_t)HW * 4, (uint64_t)C * HW * 4,
                                     tcw::PX, (uint32_t)C, 1, CU_TENSOR_MAP_SWIZZLE_128B, "gdn_tc_wide_backward"))
        return rc;
    const size_t dg_smem = (size_t)tcw::STAGES * 2 * tcw::BOX_BYTES + 1024;
    if (int rc = tmah::ensure_dynamic_smem(gdn_wide_dgamma_kernel, dg_smem)) return rc;
    const int64_t n_chunks = B * HW / tcw::PX;
    int64_t grid = sm_count();
    if (grid > n_chunks) grid = n_chunks;
    gdn_wide_dgamma_kernel<<<(unsigned)grid, tcw::DG_THREADS, dg_smem, s>>>(tm_u, tm_x, (int)n_chunks, (int)(HW / tcw::PX), (int)C, part);
    if (int rc = after_launch("gdn_wide_dgamma_kernel")) return rc;
    return gdn_reduce_partials(part, (int)grid, (int)C, prm, dgamma, dbeta, s);
}

}  // namespace mmnc
