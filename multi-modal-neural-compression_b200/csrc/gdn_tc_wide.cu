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
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

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
// the same arrive delivered to the barrier at this offset in every CTA of `mask` (thread-block cluster)
__device__ __forceinline__ void commit_bar_multicast(uint32_t addr, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(addr), "h"(mask) : "memory");
}
__device__ __forceinline__ void expect_bytes(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

#ifdef MMNC_WIDE_PROFILE
__device__ unsigned long long mmnc_wide_prof[32];
#define MMNC_TICK(i)                                                                  \
    do {                                                                              \
        if (threadIdx.x == 0 && blockIdx.x == 0) {                                    \
            const long long now__ = clock64();                                        \
            mmnc_wide_prof[i] += (unsigned long long)(now__ - tick__);                \
            tick__ = now__;                                                           \
        }                                                                             \
    } while (0)
#define MMNC_TICK_INIT() long long tick__ = clock64()
#else
#define MMNC_TICK(i) do { } while (0)
#define MMNC_TICK_INIT() do { } while (0)
#endif

// The ring has NO shared state besides its mbarriers.  Chunk q of the CTA's sequence (the call's `per_tile` chunks - gamma,
// then gamma^T for the backward - repeated once per tile) lives in slot q % SLOTS, and the sequence is known in advance, so
//   * the CONSUMER (the MMA-issuing leader, thread 0) only counts: wait "landed", issue four MMAs, tcgen05.commit to "free";
//   * every slot has its own PRODUCER: lane 0 of warp 1 + slot.  It requests the slot's next chunk whenever the slot is free
//     and chunks remain - at start-up, and from inside every wait for a contraction (where it polls its "free" barrier next
//     to the contraction's barrier instead of only spinning), so a slot is refilled the moment the MMAs that read it retire.
// The first versions kept a ring descriptor in shared memory and had the leader request chunks: ~660 cycles per request on
// the very thread that feeds the tensor pipe (128 cycles per MMA when nothing starves it, tools/probes/mma_rate_probe.cu),
// and six requests in a row at the top of its epilogue with 511 threads waiting for it at the next barrier.
//
// Thread-block clusters (kCS = 2, optional): every CTA of a cluster walks the SAME chunk sequence (same number of tiles),
// each producer loads 1 / kCS of its chunk and MULTICASTS it into the same slot of every CTA of the cluster, so the L2
// reads of gamma drop by kCS.  A slot may be overwritten when all kCS consumers have released it: the "free" barriers
// count kCS arrivals, and each CTA's tcgen05.commit is multicast to all of them.  (Not the default: wide_cluster_size().)
struct RingAddr { uint32_t base, full0, empty0; };  // shared addresses: slots, "chunk landed" barriers, "slot free" barriers

__device__ __forceinline__ bool test_bar(uint32_t addr, uint32_t parity) {  // non-blocking
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done != 0;
}

// Producer of slot `slot`: request its chunk number `round` (CTA chunk q = round * SLOTS + slot) if the slot is free.
template <int kCS>
__device__ __forceinline__ bool try_produce(const RingAddr &ra, int slot, int &round, int total, int per_tile,
                                            const uint32_t *packed, uint32_t rank) {
    const int q = round * SLOTS + slot;
    if (q >= total) return false;
    // the MMAs (of every CTA of the cluster) that read the slot's previous occupant have retired
    if (round > 0 && !test_bar(ra.empty0 + 8u * (uint32_t)slot, (uint32_t)((round - 1) & 1))) return false;
    const int c = q % per_tile;
    const uint32_t bar = ra.full0 + 8u * (uint32_t)slot;
    expect_bytes(bar, CHUNK_BYTES);  // the whole chunk lands here, whoever sends the pieces
    constexpr uint32_t SLICE = CHUNK_BYTES / kCS, PIECES = (kCS == 1) ? 4 : 2, PIECE = SLICE / PIECES;
    const uint32_t dst = ra.base + (uint32_t)slot * CHUNK_BYTES + rank * SLICE;
    const uint64_t src = reinterpret_cast<uint64_t>(packed) + (uint64_t)c * CHUNK_BYTES + rank * SLICE;
#pragma unroll
    for (uint32_t j = 0; j < PIECES; ++j) {
        if constexpr (kCS == 1)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + j * PIECE), "l"(src + j * PIECE), "r"(PIECE), "r"(bar)
                         : "memory");
        else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                         ::"r"(dst + j * PIECE), "l"(src + j * PIECE), "r"(PIECE), "r"(bar), "h"((uint16_t)((1u << kCS) - 1u))
                         : "memory");
    }
    ++round;
    return true;
}

// Every thread's wait for a contraction; the slot producers keep their slots filled while they wait.  When the
// contraction's barrier completes, every chunk it read has been released, so one more attempt refills the slot.
template <int kCS>
__device__ __forceinline__ void wait_contraction(uint32_t mma_bar, uint32_t parity, const RingAddr &ra, int my_slot, int &round,
                                                 int total, int per_tile, const uint32_t *packed, uint32_t rank) {
    if (my_slot < 0) {
        wait_bar(mma_bar, parity);
        return;
    }
    while (!test_bar(mma_bar, parity)) try_produce<kCS>(ra, my_slot, round, total, per_tile, packed, rank);
    try_produce<kCS>(ra, my_slot, round, total, per_tile, packed, rank);
}

template <int KS>
__device__ __forceinline__ void chunk_mma(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t first_acc) {
    if constexpr (KS < KC / 8) {
        tc::mma_tf32_ts_step<KS * 8, KS * 16>(d, a, b_lo, B_HI, IDESC, KS == 0 ? first_acc : 1u);
        chunk_mma<KS + 1>(d, a, b_lo, first_acc);
    }
}

// One contraction (leader only): D[128 x 256] = A[128 x 32 n_kc] * B, chunk by chunk.  q_cons = chunks consumed so far.
template <int kCS>
__device__ __forceinline__ void contract(uint32_t tmem_base, const RingAddr &ra, uint32_t mma_bar, int n_kc, int &q_cons) {
    uint32_t slot = (uint32_t)(q_cons % SLOTS), phase = (uint32_t)((q_cons / SLOTS) & 1);
#ifdef MMNC_WIDE_PROFILE
    long long waited__ = 0;
#endif
#pragma unroll 1
    for (int kc = 0; kc < n_kc; ++kc) {
#ifdef MMNC_WIDE_PROFILE
        const long long w0__ = clock64();
#endif
        wait_bar(ra.full0 + 8u * slot, phase);
#ifdef MMNC_WIDE_PROFILE
        waited__ += clock64() - w0__;  // (a register: a global update per chunk would sit in the issue loop itself)
#endif
        chunk_mma<0>(tmem_base + P, tmem_base + (uint32_t)(kc * KC), tc::desc_lo(ra.base + slot * CHUNK_BYTES, 128), kc > 0 ? 1u : 0u);
        if constexpr (kCS == 1) commit_bar(ra.empty0 + 8u * slot);
        else commit_bar_multicast(ra.empty0 + 8u * slot, (uint16_t)((1u << kCS) - 1u));
        if (++slot == SLOTS) { slot = 0; phase ^= 1u; }
    }
    commit_bar(mma_bar);
    q_cons += n_kc;
#ifdef MMNC_WIDE_PROFILE
    if (blockIdx.x == 0) mmnc_wide_prof[20] += (unsigned long long)waited__;
#endif
}

struct Setup {
    uint32_t mma_bar, tmem_base;
    int n_tiles;       // of this CTA (the same for every CTA of a cluster)
    int64_t tiles;     // of the whole call
    RingAddr ra;
    int my_slot;       // the ring slot this thread produces for, or -1
    int round;         // chunks this producer has requested
    int total;         // chunks of this CTA's whole run
    uint32_t rank;     // of this CTA in its cluster
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

// Shared start-up of the two pixel-tile kernels.  Every CTA walks ceil(tiles / grid) tiles; the few indices past the end
// are clamped to the last tile (it is computed twice, with identical results), so that the CTAs of a cluster consume
// identical chunk sequences and the kFull instances need no per-tile predicate.
template <int kCS>
__device__ __forceinline__ void wide_setup(tcw::Setup &st, uint8_t *smem_raw, uint64_t *bars, uint32_t *tmem_slot, float *beta_s,
                                           const GdnParams &prm, int C, int64_t NP, const uint32_t *packed, int per_tile) {
    using namespace tc;
    using namespace tcw;
    st.tiles = (NP + TILE - 1) / TILE;
    st.n_tiles = (int)((st.tiles + gridDim.x - 1) / gridDim.x);
    st.total = per_tile * st.n_tiles;
    st.ra.base = (smem_u32(smem_raw) + 127u) & ~127u;
    st.ra.full0 = smem_u32(&bars[0]);
    st.ra.empty0 = smem_u32(&bars[SLOTS]);
    st.rank = (kCS > 1) ? cluster_rank() : 0u;
    const int warp = threadIdx.x >> 5;
    st.my_slot = ((threadIdx.x & 31) == 0 && warp >= 1 && warp <= SLOTS) ? warp - 1 : -1;
    st.round = 0;
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 0) {
        for (int i = 0; i < SLOTS; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[SLOTS + i], kCS); }
        mbar_init(&bars[2 * SLOTS], 1);
    }
    for (int i = threadIdx.x; i < P; i += THREADS) beta_s[i] = (i < C) ? prm.b(i) : 1.f;
    fence_before();
    __syncthreads();
    if constexpr (kCS > 1) cluster_sync_all();  // every CTA's barriers exist before anyone multicasts into them
    fence_after();
    // the first chunks are on their way while the tile's x is loaded
    if (st.my_slot >= 0) try_produce<kCS>(st.ra, st.my_slot, st.round, st.total, per_tile, packed, st.rank);
    st.mma_bar = smem_u32(&bars[2 * SLOTS]);
    st.tmem_base = *tmem_slot;
}

// kFull: C = 192 or 256 and no ragged tile - every channel / pixel predicate folds away.  kHW != 0: H*W known at compile time, so
// the channel stride turns every per-channel address into an immediate offset (otherwise one 64-bit multiply-add per
// access: a third of the backward's instructions, measured with ncu).
template <bool kInverse, bool kFull, int kHW, int kCS>
__global__ void __launch_bounds__(tcw::THREADS, 1)
gdn_wide_forward_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t NP, int64_t HW_rt, int C, const GdnParams prm,
                        const uint32_t *__restrict__ packed, int n_kc) {
    using namespace tc;
    using namespace tcw;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * SLOTS + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ float beta_s[P];
    Setup st;
    wide_setup<kCS>(st, smem_raw, bars, &tmem_slot, beta_s, prm, C, NP, packed, n_kc);
    int q_cons = 0;  // (leader) chunks consumed
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const int q4 = warp >> 2;                          // which 64 channels
        const int pl = ((warp & 3) << 5) | lane;           // pixel of the tile = TMEM lane
        const int creal = kFull ? QC : min(QC, max(0, C - q4 * QC));    // real channels among this thread's 64
        const uint32_t lane_a = st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(q4 * QC);
        const uint32_t lane_d = lane_a + (uint32_t)P;
        const int64_t HW = kHW ? (int64_t)kHW : HW_rt;
        const uint32_t sb = (uint32_t)HW * 4u;
        uint32_t parity = 0;
        // element offset of this thread's first channel in tile number t of the CTA (indices past the end: the last tile)
        auto tile_offset = [&](int t, bool *valid) -> int64_t {
            int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            if (tile >= st.tiles) tile = st.tiles - 1;
            const int64_t pix = tile * TILE + pl;
            *valid = kFull || pix < NP;
            const int64_t b = *valid ? pix / HW : 0;
            return (b * C + q4 * QC) * HW + (*valid ? pix - b * HW : 0);
        };
        // x lives in registers from the A fill to the epilogue.  The epilogue hands every 16-channel block, the moment it
        // has used it, to the NEXT tile's loads: x of tile t + 1 crosses HBM underneath the epilogue's stores, for no
        // extra register and no extra traffic (phase probe: the A fill waited 7.7 k of a tile's 21.5 k cycles for x).
        // kFull covers C = 192 as well as 256: a warp's 64 channels are then all real or all absent, and a warp without
        // channels only keeps the barriers company (the MMAs never read A columns past 32 n_kc = C).
        const bool active = !kFull || q4 * QC < C;
        float xv[QC];
        bool valid;
        int64_t e0 = tile_offset(0, &valid);
#pragma unroll
        for (int c = 0; c < QC; ++c) xv[c] = (active && (kFull || (valid && c < creal))) ? __ldcs(chan_ptr(x + e0, sb, c)) : 0.f;
        MMNC_TICK_INIT();
#pragma unroll 1
        for (int t = 0; t < st.n_tiles; ++t) {
            float *yb = y + e0;
            const bool valid_t = valid;
            if (active) {
#pragma unroll
                for (int c0 = 0; c0 < QC; c0 += 16) {
                    uint32_t v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = to_tf32_fast(xv[c0 + j] * xv[c0 + j]);
                    tmem_st16(lane_a + c0, v);
                }
                tmem_st_wait();
            }
            MMNC_TICK(0);
            fence_before();
            named_bar_sync(1, COMPUTE);
            MMNC_TICK(1);
            if (threadIdx.x == 0) {
                fence_after();
                contract<kCS>(st.tmem_base, st.ra, st.mma_bar, n_kc, q_cons);
            }
            MMNC_TICK(2);
            const bool more = t + 1 < st.n_tiles;
            if (more) e0 = tile_offset(t + 1, &valid);
            const float *xnext = x + e0;
            wait_contraction<kCS>(st.mma_bar, parity, st.ra, st.my_slot, st.round, st.total, n_kc, packed, st.rank);
            MMNC_TICK(3);
            parity ^= 1u;
            fence_after();
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += 16) {
                if (!active) break;
                uint32_t r[16];
                tmem_ld16(lane_d + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float n = __uint_as_float(r[j]) + beta_s[q4 * QC + c0 + j];
                    const float rs = fast_rsqrt(n);
                    const float out = xv[c0 + j] * (kInverse ? n * rs : rs);
                    if (kFull || (valid_t && c0 + j < creal)) __stcs(chan_ptr(yb, sb, c0 + j), out);
                }
                if (more) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        xv[c0 + j] = (kFull || (valid && c0 + j < creal)) ? __ldcs(chan_ptr(xnext, sb, c0 + j)) : 0.f;
                }
            }
            MMNC_TICK(4);
            // every lane has drained D before the next tile's MMA overwrites it
            fence_before();
            named_bar_sync(1, COMPUTE);
            fence_after();
            MMNC_TICK(5);
        }
    }
    fence_before();
    __syncthreads();
    if constexpr (kCS > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
    if (warp == 0) tmem_dealloc(st.tmem_base, 512);
}

// dx and u.  Registers: only ONE 64-channel vector stays live per thread - g, turned in place into f = g n^p by epilogue 1
// and consumed by epilogue 2; x is read three times (A fill: HBM; the two epilogues: L2), eight channels at a time and one
// block ahead.  g is requested while MMA1 runs.
template <bool kInverse, bool kFull, int kHW, int kCS>
__global__ void __launch_bounds__(tcw::THREADS, 1)
gdn_wide_dx_kernel(const float *__restrict__ x, const float *__restrict__ g, float *__restrict__ dx, float *__restrict__ U,
                   int64_t NP, int64_t HW_rt, int C, const GdnParams prm, const uint32_t *__restrict__ packed, int n_kc) {
    using namespace tc;
    using namespace tcw;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * SLOTS + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ float beta_s[P];
    Setup st;
#ifdef MMNC_WIDE_PROFILE
    const long long cta_t0__ = clock64();
#endif
    wide_setup<kCS>(st, smem_raw, bars, &tmem_slot, beta_s, prm, C, NP, packed, 2 * n_kc);
    int q_cons = 0;  // (leader) chunks consumed
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr float coef = kInverse ? 0.5f : -0.5f;
    {
        const int q4 = warp >> 2;
        const int pl = ((warp & 3) << 5) | lane;
        const int creal = kFull ? QC : min(QC, max(0, C - q4 * QC));
        const uint32_t lane_a = st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(q4 * QC);
        const uint32_t lane_d = lane_a + (uint32_t)P;
        const int64_t HW = kHW ? (int64_t)kHW : HW_rt;
        const uint32_t sb = (uint32_t)HW * 4u;
        const float *beta_q = beta_s + q4 * QC;
        uint32_t parity = 0;
#define MMNC_CH(c) (kFull || (valid && (c) < cr))
        const bool active = !kFull || q4 * QC < C;  // see the forward kernel
        MMNC_TICK_INIT();
#pragma unroll 1
        for (int t = 0; t < st.n_tiles; ++t) {
            int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            if (tile >= st.tiles) tile = st.tiles - 1;
            const int64_t pix = tile * TILE + pl;
            const bool valid = kFull || pix < NP;
            const int64_t b = valid ? pix / HW : 0;
            const int64_t e0 = (b * C + q4 * QC) * HW + (valid ? pix - b * HW : 0);
            const float *xb = x + e0, *gb = g + e0;
            float *dxb = dx + e0, *ub = U + e0;
            int cr = creal;
            asm volatile("" : "+r"(cr));  // keeps the 64 channel predicates from being hoisted out of the tile loop
            // ---- A = x^2.  (Handing x of the next tile from epilogue 2 to this fill through registers, as the forward kernel
            //      does, was measured: the fill drops from 7.2 k to 2.1 k cycles, but the 64 extra live values spill and
            //      epilogue 2 goes from 8.9 k to 20.9 k - 0.55 -> 0.61 ms.)
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += 16) {
                if (!active) break;
                float xv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) xv[j] = MMNC_CH(c0 + j) ? __ldg(chan_ptr(xb, sb, c0 + j)) : 0.f;
                uint32_t v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = to_tf32_fast(xv[j] * xv[j]);
                tmem_st16(lane_a + c0, v);
            }
            tmem_st_wait();
            MMNC_TICK(8);
            fence_before();
            named_bar_sync(1, COMPUTE);
            MMNC_TICK(9);
            if (threadIdx.x == 0) {
                fence_after();
                contract<kCS>(st.tmem_base, st.ra, st.mma_bar, n_kc, q_cons);  // n - beta = x^2 gamma^T
            }
            MMNC_TICK(10);
            // ---- g on its way while MMA1 runs
            float gf[QC];
            asm volatile("" : "+r"(cr));
#pragma unroll
            for (int c = 0; c < QC; ++c) gf[c] = (active && MMNC_CH(c)) ? __ldcs(chan_ptr(gb, sb, c)) : 0.f;
            float xn[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) xn[j] = (active && MMNC_CH(j)) ? __ldg(chan_ptr(xb, sb, j)) : 0.f;
            wait_contraction<kCS>(st.mma_bar, parity, st.ra, st.my_slot, st.round, st.total, 2 * n_kc, packed, st.rank);
            MMNC_TICK(11);
            parity ^= 1u;
            fence_after();
            // ---- epilogue 1: u -> A and -> HBM (the d gamma kernel's operand); f = g n^p replaces g
            asm volatile("" : "+r"(cr));
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += 8) {
                if (!active) break;
                float xc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) xc[j] = xn[j];
                if (c0 + 8 < QC) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) xn[j] = MMNC_CH(c0 + 8 + j) ? __ldg(chan_ptr(xb, sb, c0 + 8 + j)) : 0.f;
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
                    if (MMNC_CH(c0 + j)) chan_ptr(ub, sb, c0 + j)[0] = __uint_as_float(uu[j]);
                }
                tmem_st8(lane_a + c0, uu);
            }
            MMNC_TICK(12);
            tmem_st_wait();
            fence_before();
            named_bar_sync(1, COMPUTE);
            MMNC_TICK(13);
            if (threadIdx.x == 0) {
                fence_after();
                contract<kCS>(st.tmem_base, st.ra, st.mma_bar, n_kc, q_cons);  // t = u gamma
            }
            MMNC_TICK(14);
            asm volatile("" : "+r"(cr));
#pragma unroll
            for (int j = 0; j < 8; ++j) xn[j] = (active && MMNC_CH(j)) ? __ldg(chan_ptr(xb, sb, j)) : 0.f;
            wait_contraction<kCS>(st.mma_bar, parity, st.ra, st.my_slot, st.round, st.total, 2 * n_kc, packed, st.rank);
            MMNC_TICK(15);
            parity ^= 1u;
            fence_after();
            // ---- epilogue 2: dx = f + 2 x t.  Each block also asks L2 for the same channels of the NEXT tile's x, so that the
            //      next A fill finds them there (no register, no extra HBM traffic).
            const float *xpf = nullptr;
            if (kFull && active && t + 1 < st.n_tiles) {  // (the generic instances pay more for the address arithmetic than they gain)
                int64_t tile_n = (int64_t)blockIdx.x + (int64_t)(t + 1) * gridDim.x;
                if (tile_n >= st.tiles) tile_n = st.tiles - 1;
                const int64_t pix_n = tile_n * TILE + pl;
                if (kFull || pix_n < NP) {
                    const int64_t bn = pix_n / HW;
                    xpf = x + (bn * C + q4 * QC) * HW + (pix_n - bn * HW);
                }
            }
#pragma unroll
            for (int c0 = 0; c0 < QC; c0 += 8) {
                if (!active) break;
                float xc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) xc[j] = xn[j];
                if (c0 + 8 < QC) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) xn[j] = MMNC_CH(c0 + 8 + j) ? __ldg(chan_ptr(xb, sb, c0 + 8 + j)) : 0.f;
                }
                uint32_t r[8];
                tmem_ld8(lane_d + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float out = fmaf(2.f * xc[j], __uint_as_float(r[j]), gf[c0 + j]);
                    if (MMNC_CH(c0 + j)) __stcs(chan_ptr(dxb, sb, c0 + j), out);
                }
                if (xpf != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (kFull || c0 + j < cr) asm volatile("prefetch.global.L2 [%0];" ::"l"(chan_ptr(xpf, sb, c0 + j)));
                }
            }
            MMNC_TICK(16);
            // No barrier here: the next tile's barrier (after its A fill) orders these TMEM reads of D before the next MMA1
            // overwrites it, and A was last read by MMA2, which has retired.
        }
#undef MMNC_CH
    }
#ifdef MMNC_WIDE_PROFILE
    if (threadIdx.x == 0) {  // slowest CTA and the sum over CTAs of the dx kernel's wall cycles
        const unsigned long long el = (unsigned long long)(clock64() - cta_t0__);
        atomicMax(&mmnc_wide_prof[30], el);
        atomicAdd(&mmnc_wide_prof[31], el);
    }
#endif
    fence_before();
    __syncthreads();
    if constexpr (kCS > 1) cluster_sync_all();
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

// Cluster size of the pixel-tile kernels: MMNC_GDN_WIDE_CLUSTER = 1 (default) | 2.  Measured on B200 at 64 x 256 x 64 x 64:
// forward 0.184 / 0.183 ms, backward 0.548 / 0.536 ms for 1 / 2 (and 4 was slower: 0.223 / 0.631 ms) - multicast halves the
// L2 reads of gamma, but the contraction is bound by the shared-memory array itself (every chunk byte is written once by
// the bulk copy and read once by the MMAs: ~130 B / clk at the full TF32 rate against 128 B / clk), not by L2.
static int wide_cluster_size() {
    static const int cs = []() {
        const char *e = getenv("MMNC_GDN_WIDE_CLUSTER");
        return (e && atoi(e) == 2) ? 2 : 1;
    }();
    return cs;
}

// Launch `kernel` as clusters of `cs` CTAs, as many clusters as can be resident at once (one CTA per SM).
template <typename... Args>
static int launch_clustered(void (*kernel)(Args...), int cs, int64_t tiles, cudaStream_t s, const char *what, Args... args) {
    if (int rc = tmah::ensure_dynamic_smem(kernel, WIDE_RING_SMEM)) return rc;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(tcw::THREADS, 1, 1);
    cfg.dynamicSmemBytes = WIDE_RING_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (cs > 1) ? 1 : 0;
    int clusters = sm_count() / cs;
    if (cs > 1) {
        // (kernel, cluster size, device) -> clusters that fit at once; GPCs whose SM count is not a multiple of the
        // cluster size leave a few SMs idle
        static std::mutex mu;
        static std::unordered_map<uint64_t, int> cache;
        int dev = 0;
        cudaGetDevice(&dev);
        const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(kernel) * 64u + (uint64_t)(dev & 15) * 4u + (uint64_t)(cs >> 1);
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it == cache.end()) {
            cfg.gridDim = dim3((unsigned)(sm_count() / cs * cs), 1, 1);
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
                cudaGetLastError();
                n = sm_count() / cs / 2;  // conservative: half the machine
            }
            it = cache.emplace(key, n).first;
        }
        if (clusters > it->second) clusters = it->second;
    }
    const int64_t need = (tiles + cs - 1) / cs;
    if (clusters > need) clusters = (int)need;
    cfg.gridDim = dim3((unsigned)(clusters * cs), 1, 1);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    count_launch();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MMNC_ERR_CUDA;
    }
    return MMNC_OK;
}

// instance tables: [inverse][shape class: 0 generic, 1 full tiles + C a multiple of 64, 2 = 1 with H*W = 1024, 3 = 1 with H*W = 4096]
#define MMNC_WIDE_ROW(K, I, CS) { K<I, false, 0, CS>, K<I, true, 0, CS>, K<I, true, 1024, CS>, K<I, true, 4096, CS> }
#define MMNC_WIDE_TABLE(K, CS) { MMNC_WIDE_ROW(K, false, CS), MMNC_WIDE_ROW(K, true, CS) }
static int wide_shape_class(int64_t B, int64_t C, int64_t HW) {
    if (C % tcw::QC != 0 || (B * HW) % tcw::TILE != 0) return 0;  // 192 or 256 channels, no ragged tile
    return HW == 4096 ? 3 : (HW == 1024 ? 2 : 1);
}

int gdn_tc_wide_forward(const float *x, int64_t B, int64_t C, int64_t HW, const GdnParams &prm, int inverse, float *y,
                        void *workspace, size_t workspace_bytes, cudaStream_t s) {
    MMNC_REQUIRE(gdn_tc_wide_forward_supported(B, C, HW, workspace, workspace_bytes), "gdn_tc_wide_forward: unsupported call");
    const int n_kc = wide_kchunks(C);
    uint32_t *packed = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(workspace) + 127) / 128 * 128);
    const int elems = n_kc * tcw::KC * tcw::P;
    gdn_wide_pack_kernel<<<(elems + 255) / 256, 256, 0, s>>>(prm, (int)C, n_kc, 1, packed);
    if (int rc = after_launch("gdn_wide_pack_kernel")) return rc;
    using Kernel = void (*)(const float *, float *, int64_t, int64_t, int, const GdnParams, const uint32_t *, int);
    static const Kernel k1[2][4] = MMNC_WIDE_TABLE(gdn_wide_forward_kernel, 1);
    static const Kernel k2[2][4] = MMNC_WIDE_TABLE(gdn_wide_forward_kernel, 2);
    const int cs = wide_cluster_size(), cls = wide_shape_class(B, C, HW);
    const Kernel kernel = (cs == 1 ? k1 : k2)[inverse ? 1 : 0][cls];
    const int64_t NP = B * HW, tiles = (NP + tcw::TILE - 1) / tcw::TILE;
    return launch_clustered(kernel, cs, tiles, s, "gdn_wide_forward_kernel", x, y, NP, HW, (int)C, prm,
                            (const uint32_t *)packed, n_kc);
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
        using Kernel = void (*)(const float *, const float *, float *, float *, int64_t, int64_t, int, const GdnParams,
                                const uint32_t *, int);
        static const Kernel k1[2][4] = MMNC_WIDE_TABLE(gdn_wide_dx_kernel, 1);
        static const Kernel k2[2][4] = MMNC_WIDE_TABLE(gdn_wide_dx_kernel, 2);
        const int cs = wide_cluster_size(), cls = wide_shape_class(B, C, HW);
        const Kernel kernel = (cs == 1 ? k1 : k2)[inverse ? 1 : 0][cls];
        const int64_t NP = B * HW, tiles = (NP + tcw::TILE - 1) / tcw::TILE;
        if (int rc = launch_clustered(kernel, cs, tiles, s, "gdn_wide_dx_kernel", x, g, dx, U, NP, HW, (int)C, prm,
                                      (const uint32_t *)packed, n_kc))
            return rc;
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

#ifdef MMNC_WIDE_PROFILE
// tools/probes/wide_phase_probe.py: per-phase clock totals of CTA 0's leader thread (read and reset)
extern "C" int mmnc_wide_profile_read(unsigned long long *out32) {
    unsigned long long zero[32] = {0};
    if (cudaMemcpyFromSymbol(out32, mmnc::tcw::mmnc_wide_prof, sizeof(zero)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(mmnc::tcw::mmnc_wide_prof, zero, sizeof(zero)) != cudaSuccess) return -1;
    return 0;
}
#endif
