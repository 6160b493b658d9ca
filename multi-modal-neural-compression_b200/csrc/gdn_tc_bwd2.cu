// K4 backward, second generation: the three-contraction GDN / IGDN gradient of gdn_tc_bwd.cu, fed by TMA and
// software-pipelined inside one persistent CTA per SM.
//
// What changed against gdn_tc_bwd.cu (same math, same 12 B/element):
//   * x and g tiles arrive by cp.async.bulk.tensor (rank-3 map over the NCHW tensor, box = 32 pixels x C channels,
//     128-byte swizzle).  A box lands as [channel][32 pixels] rows of 128 B in 1024-byte swizzle atoms, which IS the
//     K-major (K = pixel) operand layout of the d gamma contraction: the threads square x / turn g into u IN PLACE, so
//     the landing buffers double as the MMA3 operands and no second shared-memory copy exists.
//   * NSTAGES landing stages are shared by NGROUPS compute groups of 256 threads (two threads per pixel = TMEM lane,
//     each owning half of the channels); group q owns tiles q, q + NGROUPS, ... of the CTA's sequence and has its own
//     TMEM columns (A, D, its own d gamma partial and - with two groups - a parking area for f = g n^p), its own
//     MMA-issuing thread and its own mbarriers, so one group's epilogues overlap the other group's MMAs while the
//     loads of the tiles after next are in flight.  C <= 63: 2 groups x 3 stages; C = 64 .. 79: 1 x 2; wider: 1 x 1
//     (gamma and gamma^T alone take 100 KB at C = 100).
//   * barriers per stage: x and g complete separately (the A fill and MMA1 start when x has landed, g is first needed
//     in epilogue 1); per group: MMA1 / MMA2 done (alternating phases of one mbarrier) and MMA3 done (the stage may
//     be refilled, d gamma is current).  Epilogue 2 runs while MMA3 is still executing.  The refill of a stage is
//     issued by the group that retired it: at the end of the tile, during the next tile's MMA1 wait (1 group x 2
//     stages), or - single stage - as soon as polling between the blocks of epilogue 2 sees MMA3 retired.
//   * no global-load address arithmetic or load latency in the compute warps; every loop invariant lives in shared
//     memory, descriptors are assembled inside the MMA asm blocks: no spills (see Bwd2Ctx and tc_ptx.cuh for why
//     that matters more than usual here).
//   * wide layers (112 <= C <= 128: gamma and gamma^T would take 166 KB next to a 139 KB landing stage) STREAM the
//     contraction operand: one shared-memory buffer holds gamma for MMA1 and is refilled with gamma^T for MMA2 (and
//     back) by cp.async.bulk copies of a pre-packed image of both tiles (a small pack kernel writes them, already in
//     the core-matrix layout, into the workspace: 2 x 81 KB that stay in L2).  The refill is issued by the leader the
//     moment the MMA that read the buffer has retired and lands behind the epilogue that follows.
//   * layers with 80 <= C <= 111 used to run with a single landing stage (two do not fit next to gamma and gamma^T), so
//     every tile waited for its own x to arrive.  Streaming the gamma operand frees enough shared memory for a SECOND
//     x buffer: x of tile k + 1 lands while tile k computes (its A fill and MMA1 need nothing else), g follows as soon
//     as MMA3 of tile k has released the single u buffer (kXPF).  Used when a CTA has at least three tiles.
// Requires HW % 128 == 0 (a 128-pixel tile never straddles two images) and 16-byte aligned tensors; everything else
// stays on gdn_tc_bwd.cu.
#include <cuda.h>  // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time (no -lcuda)
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gdn_params.cuh"
#include "tc_ptx.cuh"
#include "tma_host.cuh"

namespace mmnc {

namespace tcb2 {
constexpr int TILE = 128;  // pixels per tile = TMEM lanes
}  // namespace tcb2

// Everything a compute group needs.  The kernel runs 512 threads per SM (128 registers each) and x / g of a tile
// take 64 of them, so loop invariants must not sit in registers: the struct lives in SHARED memory and is read
// through a volatile reference at the point of use (an LDS costs ~30 cycles; a spilled register comes back from L2
// in ~340, because the L1 is a few KB once the shared-memory carve-out is at its maximum).  Descriptors, TMEM
// columns and buffer addresses are derived from these few values where they are needed.
struct Bwd2Ctx {
    int C, R8, n_k;          // channels, 8-row groups per K atom, tiles of this CTA
    int tiles_per_img;
    uint32_t sb;             // channel stride in bytes
    uint32_t stage0;         // shared address of stage 0 ([u | x2] per stage, R8 * 4096 bytes each)
    uint32_t gamma0;         // shared address of the gamma tile; gamma^T follows P * P * 4 bytes later
    uint32_t full_bar0, mma_bar0, free_bar0;  // shared addresses of the mbarrier arrays
    uint32_t bfull;          // (streamed operand) shared address of the "gamma buffer filled" mbarrier
    uint32_t tmem_base;
    const void *tm_x, *tm_g;
    const float *bsrc;       // (streamed operand) packed gamma tile in global memory; gamma^T follows P * P floats later
    float *dx;
};

// Streamed operand: refill the single gamma buffer with tile `which` (0 = gamma, 1 = gamma^T) of the packed image.
template <int P>
__device__ __forceinline__ void bwd2_load_b(const volatile Bwd2Ctx &t, int which) {
    constexpr uint32_t BYTES = (uint32_t)(P * P * 4), PIECE = BYTES / 4;  // P is a multiple of 16: PIECE % 256 == 0
    static_assert(P % 16 == 0, "four 16-byte aligned pieces");
    const uint32_t bar = t.bfull, dst = t.gamma0;
    const uint64_t src = reinterpret_cast<uint64_t>(t.bsrc) + (uint64_t)which * BYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BYTES) : "memory");
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + i * PIECE), "l"(src + i * PIECE), "r"(PIECE), "r"(bar)
                     : "memory");
}

__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
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

__device__ __forceinline__ bool mbar_test_addr(uint32_t addr, uint32_t parity) {  // non-blocking
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

template <int NSTAGES>
__device__ __forceinline__ void bwd2_issue_tile(const volatile Bwd2Ctx &t, int k) {
    const int s = k % NSTAGES;
    const uint32_t tile = blockIdx.x + (uint32_t)k * gridDim.x;  // < 2^31 tiles (checked on the host)
    const int b = (int)(tile / (uint32_t)t.tiles_per_img);
    const int hw0 = (int)(tile - (uint32_t)b * (uint32_t)t.tiles_per_img) * tcb2::TILE;
    // two barriers per stage: the x boxes are issued first and complete on their own barrier, so that the A fill and
    // MMA1 start while g (first needed in epilogue 1) is still landing
    const uint32_t bar_x = t.full_bar0 + 16u * (uint32_t)s, bar_g = bar_x + 8u;
    const uint32_t half_bytes = (uint32_t)t.C * 512u;  // 4 boxes x C rows x 128 B per tensor
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_x), "r"(half_bytes) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_g), "r"(half_bytes) : "memory");
    const uint32_t R8 = (uint32_t)t.R8, buf_bytes = R8 * 4096u;
    const uint32_t ub = t.stage0 + (uint32_t)s * 2u * buf_bytes, xb = ub + buf_bytes;
    const uint64_t tmx = reinterpret_cast<uint64_t>(t.tm_x), tmg = reinterpret_cast<uint64_t>(t.tm_g);
#pragma unroll
    for (int a = 0; a < 4; ++a)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(xb + (uint32_t)a * R8 * 1024u), "l"(tmx), "r"(hw0 + 32 * a), "r"(0), "r"(b), "r"(bar_x)
            : "memory");
#pragma unroll
    for (int a = 0; a < 4; ++a)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(ub + (uint32_t)a * R8 * 1024u), "l"(tmg), "r"(hw0 + 32 * a), "r"(0), "r"(b), "r"(bar_g)
            : "memory");
}

// x-prefetch mode (kXPF): shared memory is [u | x2 (even tiles) | x2 (odd tiles)]; x of tile k lands in buffer k & 1 on
// its own barrier (full_bar[0] / full_bar[2]), g of every tile in the single u buffer on full_bar[1].
__device__ __forceinline__ void bwd2_tile_coords(const volatile Bwd2Ctx &t, int k, int *b, int *hw0) {
    const uint32_t tile = blockIdx.x + (uint32_t)k * gridDim.x;
    const uint32_t bb = tile / (uint32_t)t.tiles_per_img;
    *b = (int)bb;
    *hw0 = (int)(tile - bb * (uint32_t)t.tiles_per_img) * tcb2::TILE;
}
__device__ __forceinline__ void bwd2_issue_x(const volatile Bwd2Ctx &t, int k) {
    int b, hw0;
    bwd2_tile_coords(t, k, &b, &hw0);
    const uint32_t R8 = (uint32_t)t.R8, buf_bytes = R8 * 4096u;
    const uint32_t bar = t.full_bar0 + 16u * (uint32_t)(k & 1), dst = t.stage0 + buf_bytes * (1u + (uint32_t)(k & 1));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)t.C * 512u) : "memory");
    const uint64_t tm = reinterpret_cast<uint64_t>(t.tm_x);
#pragma unroll
    for (int a = 0; a < 4; ++a)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(dst + (uint32_t)a * R8 * 1024u), "l"(tm), "r"(hw0 + 32 * a), "r"(0), "r"(b), "r"(bar)
            : "memory");
}
__device__ __forceinline__ void bwd2_issue_g(const volatile Bwd2Ctx &t, int k) {
    int b, hw0;
    bwd2_tile_coords(t, k, &b, &hw0);
    const uint32_t R8 = (uint32_t)t.R8;
    const uint32_t bar = t.full_bar0 + 8u, dst = t.stage0;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)t.C * 512u) : "memory");
    const uint64_t tm = reinterpret_cast<uint64_t>(t.tm_g);
#pragma unroll
    for (int a = 0; a < 4; ++a)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(dst + (uint32_t)a * R8 * 1024u), "l"(tm), "r"(hw0 + 32 * a), "r"(0), "r"(b), "r"(bar)
            : "memory");
}
// the single u buffer and the x buffer of tile k are free (MMA3 of tile k has retired): g of the next tile, x of the one after
__device__ __forceinline__ void bwd2_refill_xpf(const volatile Bwd2Ctx &t, int k) {
    if (k + 1 < t.n_k) bwd2_issue_g(t, k + 1);
    if (k + 2 < t.n_k) bwd2_issue_x(t, k + 2);
}

// compile-time unrolled MMA chains (every per-step offset is an immediate inside the asm block, see tc_ptx.cuh)
template <int KS, int NK>
__device__ __forceinline__ void mma_ts_chain(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if constexpr (KS < NK) {
        tc::mma_tf32_ts_step<KS * 8, KS * 16>(d, a, b_lo, b_hi, idesc, KS > 0 ? 1u : 0u);
        mma_ts_chain<KS + 1, NK>(d, a, b_lo, b_hi, idesc);
    }
}
template <int KS>
__device__ __forceinline__ void mma_ss_chain(uint32_t d3, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t atom16,
                                             uint32_t idesc, uint32_t first_acc) {
    if constexpr (KS < tcb2::TILE / 8) {
        tc::mma_tf32_ss_step<(KS & 3) * 2>(d3, a_lo, b_lo, hi, (uint32_t)(KS >> 2) * atom16, idesc,
                                           KS == 0 ? first_acc : 1u);
        mma_ss_chain<KS + 1>(d3, a_lo, b_lo, hi, atom16, idesc, first_acc);
    }
}

// tcgen05.ld / st of W = 8 or 16 consecutive columns
template <int W>
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, uint32_t (&r)[W]) {
    if constexpr (W == 16) tc::tmem_ld16(taddr, r); else tc::tmem_ld8(taddr, r);
}
template <int W>
__device__ __forceinline__ void tmem_stw(uint32_t taddr, const uint32_t (&r)[W]) {
    if constexpr (W == 16) tc::tmem_st16(taddr, r); else tc::tmem_st8(taddr, r);
}

// One compute group's share of the CTA's tile sequence.  Thread (pix, half) owns channels [half KH, half KH + KH) of
// pixel pix.  `half` is a run-time value on purpose: one copy of the (fully unrolled) tile body instead of two keeps
// the instruction footprint inside the instruction cache (ncu: 21 % of the stalls were instruction fetches with two
// copies).  Channel offsets are folded into the per-thread TMEM / shared / global bases; only the last 16 channels of
// a thread can be padding, and a 16-bit mask says which.
template <int KH8, int NGROUPS, int NSTAGES, int TPP, bool kInverse, bool kStream, bool kXPF>
__device__ __forceinline__ bool bwd2_group_loop(const volatile Bwd2Ctx &t, int group, int tg, uint32_t tmem_base_in) {
    static_assert(!kXPF || (kStream && NGROUPS == 1 && NSTAGES == 1), "x prefetch: one group, one u buffer, streamed gamma");
    using namespace tc;
    using namespace tcb2;
    constexpr int P = KH8 * 16;   // padded channel count
    // TPP threads share a pixel (a TMEM lane) and own contiguous channel ranges of KH channels: P / 2 each for two
    // threads; for four, P / 4 rounded up to 8 (the last thread's range then runs past P: those blocks are skipped).
    // Measured on B200: four threads per pixel is SLOWER - on the wide, single-group layers (with f parked in TMEM to
    // fit 128 registers; GDN(100)@128^2: 0.444 vs 0.390 ms) and on the two-group ones (1024 threads x 64 registers;
    // GDN(50)@256^2: 0.621 vs 0.539 ms) - so every instance uses two.
    constexpr int KH = (TPP == 2) ? P / 2 : (P / 4 + 7) / 8 * 8;
    constexpr int TPG = 128 * TPP;
    // first local channel that may be padding: with two threads only the last 16 channels of a thread (C > P - 16)
    constexpr int TAIL0 = (TPP == 2 && KH > 16) ? KH - 16 : 0;
    constexpr float coef = kInverse ? 0.5f : -0.5f;
    // two groups = 128 registers per thread: g is then re-read from its landing buffer in epilogue 1 and f = g n^p
    // is parked in spare TMEM columns until epilogue 2, so that only x stays in registers across the MMA waits
    constexpr bool PARK = (NGROUPS == 2);
    constexpr int KG = PARK ? 1 : KH;
    constexpr int W = (KH % 16 == 0 && !PARK) ? 16 : 8;  // columns per tcgen05.ld / st
    constexpr int W1 = PARK ? 8 : W;            // epilogue 1 is where the 128-register budget is tightest
    // refill a stage during the next tile's MMA1 wait instead of at the end of its own tile.  Measured on B200: helps the
    // single-group double-stage layers (GDN(64)@256^2 0.817 -> 0.783 ms), hurts two groups on three stages (the
    // refilled tile is the OTHER group's next one and loses its head start: GDN(50)@256^2 0.527 -> 0.669 ms)
    constexpr bool DEFER = (NGROUPS == 1 && NSTAGES == 2);
    constexpr uint32_t kcores = P >> 2;
    constexpr uint32_t GAMMA_HI = desc_hi(kcores * 128u, 0);  // K-major, no swizzle: SBO = one 8-row group of cores
    constexpr uint32_t PIX_HI = desc_hi(1024u, 2);            // K-major (K = pixel), 128-byte swizzle
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int pix = tg & 127;
    const int c_begin = (tg >> 7) * KH;
    const bool leader = (tg == 0);
    // The asynchronous refills are issued by OTHER threads than the one that issues the MMAs: each costs its thread a few
    // hundred cycles (context reads from shared memory, coordinates, four to eight bulk / TMA instructions), the whole
    // group waits at the next barrier for its slowest thread, and the MMA issuer is on the critical path anyway.
    const bool loader_b = kStream ? (tg == 32) : false;              // gamma / gamma^T buffer
    const bool loader_t = (tg == 64);                                // landing stages
    const uint32_t bar_id = 1u + (uint32_t)group;
    const uint32_t tmem_base = tmem_base_in;
    const uint32_t a_base = tmem_base + (uint32_t)(group * 2 * P);
    const uint32_t lane_a = a_base + (((uint32_t)(((threadIdx.x >> 5) & 3) * 32)) << 16) + (uint32_t)c_begin;
    const uint32_t lane_d = lane_a + (uint32_t)P;
    const uint32_t lane_f = lane_a + (uint32_t)((NGROUPS * 3 - group) * P);  // columns after every A / D / D3 region
    const uint32_t mbar = t.mma_bar0 + 8u * (uint32_t)group;    // MMA1 done / MMA2 done (alternating phases)
    const uint32_t fbar = t.free_bar0 + 8u * (uint32_t)group;  // MMA3 done: the stage may be refilled, D3 is current
    // local channels TAIL0 + i with i < one_i are real, i == one_i is the constant-1 channel C, the rest is padding
    const int one_i = t.C - c_begin - TAIL0;
    // swizzled offset of (row c, pixel pix): rows of one 8-row group differ only in the XOR of address bits 4..6 with
    // (c & 7), and every base below has zeros there, so   addr(c) = ((base | q << 4) ^ ((c & 7) << 4)) + row-group
    // immediate : one register per buffer instead of eight precomputed offsets
    uint32_t bq;
    {
        const int atomk = pix >> 5, p32 = pix & 31, R8 = t.R8;
        bq = (uint32_t)(((atomk * R8 + (c_begin >> 3)) << 10) + ((p32 >> 2) << 4) + ((p32 & 3) << 2));
    }
    // `oi` is one_i laundered through an empty asm before every phase: otherwise ptxas hoists the sixteen padding
    // selects out of the tile loop, keeps them in registers for the whole kernel and spills them
#define MMNC_REAL(j) ((j) < TAIL0 || (j) - TAIL0 < oi)
#define MMNC_FRESH_OI() asm volatile("" : "+r"(oi))
    // The landing buffers store 8 R8 >= C + 1 rows per K atom; rows >= C are never written by the TMA box and hold 0
    // (row C of the x buffer: 1, the constant channel), so x and g are loaded and x^2 / u written back WITHOUT per-channel
    // predicates or selects.  Only a thread's last 8-channel block can lie beyond the stored rows (P - 8 R8 is 0 or 8).
    const bool last_ok = (c_begin + KH <= 8 * (int)t.R8);
#define MMNC_ROW(j) ((j) < KH - 8 || last_ok)
#define MMNC_SOFF(base, j) (((base) ^ (uint32_t)(((j) & 7) << 4)) + (uint32_t)((((j) >> 3) << 10) + (((j) & 7) << 7)))
    uint32_t parity = 0;
    bool first = true;
    int oi = one_i, pending = -1;
#pragma unroll 1
    for (int k = group; k < t.n_k; k += NGROUPS) {
        const int s = k % NSTAGES;
        const uint32_t buf_bytes = (uint32_t)t.R8 * 4096u;
        const uint32_t us = t.stage0 + (kXPF ? 0u : (uint32_t)s * 2u * buf_bytes);
        const uint32_t xs = kXPF ? t.stage0 + buf_bytes * (1u + (uint32_t)(k & 1)) : us + buf_bytes;
        const uint32_t uq = us + bq, xq = xs + bq;
        const uint32_t bar_x = t.full_bar0 + 16u * (uint32_t)(kXPF ? (k & 1) : s);
        const uint32_t bar_g = kXPF ? t.full_bar0 + 8u : bar_x + 8u;
        const uint32_t fpar = (uint32_t)(((kXPF ? (k >> 1) : (k / NSTAGES))) & 1), gpar = kXPF ? (uint32_t)(k & 1) : fpar;
        mbar_wait_addr(bar_x, fpar);
        // ---- this thread's channels of x out of the landing buffer (conflict-free: a warp reads one 128 B row)
        float xv[KH], gv[KG];
        MMNC_FRESH_OI();
#pragma unroll
        for (int j = 0; j < KH; ++j) xv[j] = MMNC_ROW(j) ? ld_shared_f32(MMNC_SOFF(xq, j)) : 0.f;
        // ---- x^2 -> A (TMEM) and back into the landing buffer (MMA3's B operand); padded channel C is the constant 1,
        //      whose shared-memory row was written once at start-up and is never touched by the TMA box (C rows)
#pragma unroll
        for (int j0 = 0; j0 < KH; j0 += W) {
            if (TPP > 2 && c_begin + j0 >= P) break;  // (warp-uniform) columns past the A region belong to D
            uint32_t v[W];
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const uint32_t sq = to_tf32_fast(xv[j0 + j] * xv[j0 + j]);
                if (MMNC_ROW(j0 + j)) st_shared_u32(MMNC_SOFF(xq, j0 + j), sq);  // padding rows get their 0 / 1 back
                v[j] = sq;
            }
            tmem_stw<W>(lane_a + j0, v);
        }
        tmem_st_wait();
        fence_async_smem();
        fence_before();
        named_bar_sync(bar_id, TPG);
        // ---- MMA1: D = x2 * gamma^T (+ beta through the constant column)
        if (leader) {
            fence_after();
            if constexpr (kStream) mbar_wait_addr(t.bfull, 0u);  // fills alternate gamma (phase 0) / gamma^T (phase 1)
            mma_ts_chain<0, P / 8>(a_base + P, a_base, desc_lo(t.gamma0, 128), GAMMA_HI, IDESC);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
        // A refill that could not be issued at the end of the previous tile (its MMA3 had not retired yet) is issued
        // here, while the whole group waits for MMA1 anyway.
        if (DEFER && loader_t && pending >= 0) {
            mbar_wait_addr(fbar, (uint32_t)(((pending - group) / NGROUPS) & 1));
            if (pending + NSTAGES < t.n_k) bwd2_issue_tile<NSTAGES>(t, pending + NSTAGES);
            pending = -1;
        }
        // g has had the A fill and the MMA1 issue to land behind x
        mbar_wait_addr(bar_g, gpar);
        if constexpr (!PARK) {
            MMNC_FRESH_OI();
#pragma unroll
            for (int j = 0; j < KH; ++j) gv[j] = MMNC_ROW(j) ? ld_shared_f32(MMNC_SOFF(uq, j)) : 0.f;
        }
        mbar_wait_addr(mbar, parity);
        parity ^= 1;
        fence_after();
        // MMA1 has retired: the buffer it read is refilled with gamma^T while epilogue 1 runs
        if (loader_b) bwd2_load_b<P>(t, 1);
        // ---- epilogue 1: u -> A (TMEM) and over g in the landing buffer (MMA3's A operand); f = g n^p kept for later
        MMNC_FRESH_OI();
#pragma unroll
        for (int j0 = 0; j0 < KH; j0 += W1) {
            if (TPP > 2 && c_begin + j0 >= P) break;
            uint32_t r[W1], uu[W1], ff[W1];
            float g8[W1];
            tmem_ldw<W1>(lane_d + j0, r);
#pragma unroll
            for (int j = 0; j < W1; ++j) {
                if constexpr (PARK)
                    g8[j] = MMNC_ROW(j0 + j) ? ld_shared_f32(MMNC_SOFF(uq, j0 + j)) : 0.f;
                else
                    g8[j] = gv[(j0 + j) % KG];
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < W1; ++j) {
                const float n = __uint_as_float(r[j]);
                const float rs = fast_rsqrt(n);
                const float pw = kInverse ? n * rs : rs;         // n^p
                const float pm1 = kInverse ? rs : rs * rs * rs;  // n^(p-1)
                const float u = (coef * g8[j]) * (xv[j0 + j] * pm1);
                const float f = g8[j] * pw;
                if constexpr (PARK) ff[j] = __float_as_uint(f); else gv[(j0 + j) % KG] = f;
                uu[j] = to_tf32_fast(u);  // padding channels: g = 0 there, so u = 0
                if (MMNC_ROW(j0 + j)) st_shared_u32(MMNC_SOFF(uq, j0 + j), uu[j]);
            }
            tmem_stw<W1>(lane_a + j0, uu);
            if constexpr (PARK) tmem_stw<W1>(lane_f + j0, ff);
        }
        tmem_st_wait();
        fence_async_smem();
        fence_before();
        named_bar_sync(bar_id, TPG);
        // ---- MMA2: D = u * gamma (B = gamma^T tile), committed on its own so that epilogue 2 overlaps MMA3;
        //      MMA3: D3 += u^T x2 (K = 128 pixels of this stage), committed to the "stage free" barrier
        if (leader) {
            fence_after();
            if constexpr (kStream) mbar_wait_addr(t.bfull, 1u);
            mma_ts_chain<0, P / 8>(a_base + P, a_base, desc_lo(t.gamma0 + (kStream ? 0u : (uint32_t)(P * P * 4)), 128), GAMMA_HI, IDESC);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
            mma_ss_chain<0>(tmem_base + (uint32_t)(NGROUPS * 2 * P + group * P), desc_lo(us, 16), desc_lo(xs, 16), PIX_HI,
                            (uint32_t)t.R8 * 64u, IDESC, first ? 0u : 1u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(fbar) : "memory");
        }
        mbar_wait_addr(mbar, parity);
        parity ^= 1;
        fence_after();
        // MMA2 has retired: gamma comes back for the next tile's MMA1 while epilogue 2 runs (nothing is left in flight
        // when the CTA has no further tile)
        if (loader_b && k + NGROUPS < t.n_k) bwd2_load_b<P>(t, 0);
        // ---- epilogue 2: dx = g n^p + 2 x t
        bool refilled = false;
        {
            const uint32_t tile = blockIdx.x + (uint32_t)k * gridDim.x;
            const uint32_t tpi = (uint32_t)t.tiles_per_img, sb = t.sb;
            const uint32_t b = tile / tpi;
            const uint32_t hw0 = (tile - b * tpi) * TILE;
            MMNC_FRESH_OI();
            float *dxb = t.dx + ((int64_t)(b * (uint32_t)t.C + (uint32_t)c_begin) * (int64_t)(sb >> 2) + hw0 + pix);
#pragma unroll
            for (int j0 = 0; j0 < KH; j0 += W) {
                if (TPP > 2 && c_begin + j0 >= P) break;
                uint32_t r[W], ff[W];
                tmem_ldw<W>(lane_d + j0, r);
                if constexpr (PARK) tmem_ldw<W>(lane_f + j0, ff);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const float f = PARK ? __uint_as_float(ff[j]) : gv[(j0 + j) % KG];
                    const float out = fmaf(2.f * xv[j0 + j], __uint_as_float(r[j]), f);
                    if (MMNC_REAL(j0 + j)) __stcs(chan_ptr(dxb, sb, j0 + j), out);
                }
                // single stage: the next tile's load can only start when MMA3 has retired, and every cycle until then
                // is exposed - poll between the blocks of this epilogue instead of finishing it first
                if (NSTAGES == 1 && loader_t && !refilled && mbar_test_addr(fbar, (uint32_t)(((k - group) / NGROUPS) & 1))) {
                    if constexpr (kXPF) bwd2_refill_xpf(t, k);
                    else if (k + NSTAGES < t.n_k) bwd2_issue_tile<NSTAGES>(t, k + NSTAGES);
                    refilled = true;
                }
            }
        }
        // No barrier here: the next tile's first barrier (after its A fill) already orders every thread's TMEM reads
        // of this tile before the next MMA1 overwrites D, and A was last read by MMA2, which has completed.
        // The stage is free once MMA3 (its last reader) has retired: refill it with the tile NSTAGES ahead, now or
        // (DEFER) while the group waits for the next tile's MMA1.
        if (loader_t) {
            if (DEFER) {
                pending = k;
            } else if (!refilled) {
                mbar_wait_addr(fbar, (uint32_t)(((k - group) / NGROUPS) & 1));
                if constexpr (kXPF) bwd2_refill_xpf(t, k);
                else if (k + NSTAGES < t.n_k) bwd2_issue_tile<NSTAGES>(t, k + NSTAGES);
            }
        }
        first = false;
    }
    // the last tile of the group: wait for its MMA3 (D3 is final after that) and hand its stage on if anyone needs it
    if (DEFER && loader_t && pending >= 0) {
        mbar_wait_addr(fbar, (uint32_t)(((pending - group) / NGROUPS) & 1));
        if (pending + NSTAGES < t.n_k) bwd2_issue_tile<NSTAGES>(t, pending + NSTAGES);
    }
    // the refilling thread has seen the last MMA3 retire; after this barrier D3 is final for the whole group
    fence_before();
    named_bar_sync(bar_id, TPG);
    fence_after();
#undef MMNC_REAL
#undef MMNC_ROW
#undef MMNC_FRESH_OI
#undef MMNC_SOFF
    return first;
}

template <int KH8, int NGROUPS, int NSTAGES, int TPP, bool kInverse, bool kStream, bool kXPF = false>
__global__ void __launch_bounds__(NGROUPS * TPP * 128, 1)
gdn_tc_backward2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                        int ntiles, int tiles_per_img, int HW, const GdnParams prm, float *__restrict__ dx,
                        float *__restrict__ part, int C, uint32_t tmem_cols, const float *__restrict__ packed_b) {
    using namespace tc;
    using namespace tcb2;
    constexpr int P = KH8 * 16;
    constexpr int TPG = 128 * TPP;
    constexpr int THREADS = NGROUPS * TPG;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int NBUFS = kXPF ? 3 : 2 * NSTAGES;  // landing buffers: [u | x2] per stage, or [u | x2 | x2] with x prefetch
    __shared__ uint64_t full_bar[kXPF ? 4 : 2 * NSTAGES];  // [stage][x, g]; x prefetch: x even, g, x odd, unused
    __shared__ uint64_t mma_bar[NGROUPS];
    __shared__ uint64_t free_bar[NGROUPS];
    __shared__ uint64_t bfull_bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int R8 = (C + 1 + 7) >> 3;
    const uint32_t buf_bytes = (uint32_t)R8 * 4096u;  // 4 K atoms of 32 pixels, R8 swizzle atoms of 8 rows each
    const uint32_t stage_bytes = 2u * buf_bytes;      // [u | x2]
    float *Bs = reinterpret_cast<float *>(smem + (size_t)NBUFS * buf_bytes);  // gamma   (N = i, K = j)
    float *Bs2 = Bs + P * P;                                                        // gamma^T (N = k, K = i)
    const int warp = threadIdx.x >> 5;
    const int group = threadIdx.x / TPG, tg = threadIdx.x % TPG;

    __shared__ Bwd2Ctx ctx_s;
    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (threadIdx.x == 0) {
        for (int s = 0; s < (kXPF ? 4 : 2 * NSTAGES); ++s) mbar_init(&full_bar[s], 1);
        for (int q = 0; q < NGROUPS; ++q) mbar_init(&mma_bar[q], 1);
        for (int q = 0; q < NGROUPS; ++q) mbar_init(&free_bar[q], 1);
        mbar_init(&bfull_bar, 1);
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_g);
        Bwd2Ctx c;
        c.C = C; c.R8 = R8;
        c.tiles_per_img = tiles_per_img;
        c.n_k = (blockIdx.x < (uint32_t)ntiles) ? (int)(((uint32_t)ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
        c.sb = (uint32_t)HW * 4u;
        c.stage0 = smem_u32(smem);
        c.gamma0 = smem_u32(Bs);
        c.full_bar0 = smem_u32(&full_bar[0]);
        c.mma_bar0 = smem_u32(&mma_bar[0]);
        c.free_bar0 = smem_u32(&free_bar[0]);
        c.bfull = smem_u32(&bfull_bar);
        c.tmem_base = 0;  // read from tmem_base_s once the allocation is visible
        c.tm_x = &tm_x; c.tm_g = &tm_g;
        c.bsrc = packed_b;
        c.dx = dx;
        ctx_s = c;
    }
    __syncthreads();
    const volatile Bwd2Ctx &ctx = ctx_s;
    // the first NSTAGES tiles are on their way while gamma is being staged (small layers are all prologue); the TMA
    // box writes rows < C only, the loop below initialises rows >= C only
    if (threadIdx.x == 0) {
        const int n_k = ctx.n_k;
        if constexpr (kXPF) {
            if (n_k > 0) { bwd2_issue_x(ctx, 0); bwd2_issue_g(ctx, 0); }
            if (n_k > 1) bwd2_issue_x(ctx, 1);
        } else {
            const int pre = n_k < NSTAGES ? n_k : NSTAGES;
            for (int k = 0; k < pre; ++k) bwd2_issue_tile<NSTAGES>(ctx, k);
        }
        if (kStream && n_k > 0) bwd2_load_b<P>(ctx, 0);  // gamma for the first MMA1
    }
    // gamma tiles, as in gdn_tc_bwd.cu: Bs[n/8][k/4][n%8][k%4] = tf32(gamma[n][k]); column k = C holds beta
    // Loads first (eight elements = sixteen loads per thread in flight), then the arithmetic: one dependent L2 round
    // trip per element made this loop ~30 us per CTA at C = 100.
    constexpr int kcores = P >> 2;
    constexpr int SU = 8;
    for (int base = threadIdx.x; base < (kStream ? 0 : P * P); base += THREADS * SU) {  // streamed: packed by a kernel
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
    // landing stages: rows the TMA box never writes (C <= r < 8 R8) are 0, except row C of every x2 buffer = the constant 1
    {
        const int pad_rows = 8 * R8 - C;                                   // 1 .. 8
        const uint32_t per_buf = 4u * (uint32_t)pad_rows * 32u;            // 4 K atoms x pad rows x 32 words
        const uint32_t words = (uint32_t)NBUFS * per_buf;
        for (uint32_t i = threadIdx.x; i < words; i += THREADS) {
            const uint32_t buf = i / per_buf, o = i - buf * per_buf;
            const uint32_t atom = o / ((uint32_t)pad_rows * 32u), o2 = o - atom * (uint32_t)pad_rows * 32u;
            const int r = C + (int)(o2 >> 5);
            const uint32_t w = o2 & 31u;
            const uint32_t byte = buf * buf_bytes + ((atom * (uint32_t)R8 + (uint32_t)(r >> 3)) << 10) + ((uint32_t)(r & 7) << 7) + (w << 2);
            const bool is_x2 = kXPF ? (buf >= 1u) : ((buf & 1u) != 0u);
            *reinterpret_cast<uint32_t *>(smem + byte) = (is_x2 && r == C) ? 0x3f800000u : 0u;
        }
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const bool first = bwd2_group_loop<KH8, NGROUPS, NSTAGES, TPP, kInverse, kStream, kXPF>(ctx, group, tg, tmem_base_s);
    // ---- this group's partial d gamma / d beta: D3 lane i = out channel, column j = in channel (j = C: d beta)
    float *dst = part + ((int64_t)blockIdx.x * NGROUPS + group) * C * (C + 1);
    const int pix = tg & 127;
    if ((tg >> 7) == 0) {
        if (!first) {
            const uint32_t lane_d3 = tmem_base_s + (((uint32_t)((warp & 3) * 32)) << 16) +
                                     (uint32_t)(NGROUPS * 2 * P + group * P);
#pragma unroll 1
            for (int j0 = 0; j0 < P; j0 += 8) {
                uint32_t r[8];
                tmem_ld8(lane_d3 + j0, r);
                tmem_ld_wait();
                if (pix < C) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j0 + j <= C) dst[(int64_t)pix * (C + 1) + j0 + j] = __uint_as_float(r[j]);
                }
            }
        } else if (pix < C) {
            for (int j = 0; j <= C; ++j) dst[(int64_t)pix * (C + 1) + j] = 0.f;  // group without tiles
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, tmem_cols);
}

// Streamed operand: both contraction tiles in the K-major core-matrix layout the MMAs read, written once per call.
//   out[0 .. P*P)      gamma   (N = out channel i, K = in channel j); column K = C holds beta, rows >= C hold 1 there
//   out[P*P .. 2 P*P)  gamma^T (N = in channel k, K = out channel i)
__global__ void __launch_bounds__(256)
gdn_pack_gamma_kernel(const GdnParams prm, int C, int P, uint32_t *__restrict__ out) {
    const int kcores = P >> 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P * P; idx += gridDim.x * blockDim.x) {
        const int n = idx / P, k = idx - n * P;
        float v = 0.f, vt = 0.f;
        if (n < C && k < C) { v = prm.g(n * C + k); vt = prm.g(k * C + n); }
        else if (n < C && k == C) v = prm.b(n);
        else if (n >= C && k == C) v = 1.f;  // padded outputs get norm = 1 (finite)
        const int off = (((n >> 3) * kcores + (k >> 2)) << 5) + ((n & 7) << 2) + (k & 3);
        out[off] = tc::to_tf32(v);
        out[P * P + off] = tc::to_tf32(vt);
    }
}

// fixed-order reduction of per-group partials [ksplit][C][C+1] -> d gamma, d beta (gdn_simt.cu)
int gdn_reduce_partials(const float *part, int ksplit, int C, const GdnParams &prm, float *dgamma, float *dbeta,
                        cudaStream_t s);

// ------------------------------------------------------------------------------------------------------ host side
struct Bwd2Geometry {
    int P, groups, stages;
    bool stream;          // one gamma buffer refilled from a packed global image instead of two resident tiles
    uint32_t tmem_cols;
    size_t smem;
};

static bool bwd2_geometry(int64_t C, Bwd2Geometry *g) {
    if (C < 16 || C > 128) return false;
    const int P = (int)((C + 1 + 15) / 16 * 16);
    const size_t R8 = (size_t)(C + 1 + 7) / 8;
    g->stream = C > 111;
    const size_t stage = 2 * R8 * 4096, gam = (g->stream ? 1 : 2) * (size_t)P * P * 4;
    const size_t budget = 227 * 1024 - 1024 - 256;  // alignment slack + static shared memory
    // two groups need x and g of a tile in 128 registers per thread: P <= 64 (32 channels per thread)
    int groups = (P <= 64) ? 2 : 1;
    if (g->stream && stage + gam > budget) return false;
    int stages = 0;
    for (;;) {
        const int want = (groups == 2) ? 3 : 2;
        for (int s = want; s >= groups; --s)
            if ((size_t)s * stage + gam <= budget) { stages = s; break; }
        if (stages || groups == 1) break;
        groups = 1;
    }
    if (!stages) {
        if (stage + gam > budget) return false;
        stages = 1;
    }
    g->P = P; g->groups = groups; g->stages = stages;
    const int need = groups * (groups == 2 ? 4 : 3) * P;  // A, D, D3 per group (+ the parked f when two groups share the SM)
    uint32_t cols = 32;
    while ((int)cols < need) cols <<= 1;
    if (cols > 512) return false;
    g->tmem_cols = cols;
    // MMA3 views 128 rows (A) / P rows (B) of buffers that store 8 R8 rows per K atom: the over-read of the last K
    // atom of the last stage must stay inside the allocation (gamma follows the stages; pad if that is not enough)
    size_t bytes = (size_t)stages * stage + gam;
    const size_t last_u = (size_t)(stages - 1) * stage, last_x = last_u + R8 * 4096;
    const size_t over_a = last_u + (3 * R8 + 16) * 1024, over_b = last_x + (3 * R8 + (size_t)P / 8) * 1024;
    if (bytes < over_a) bytes = over_a;
    if (bytes < over_b) bytes = over_b;
    g->smem = bytes + 1024;
    return g->smem + 256 <= 227 * 1024;
}

static int bwd2_mode() {  // MMNC_GDN_BWD = v1 | v2 | auto (default)
    static int mode = []() {
        const char *e = getenv("MMNC_GDN_BWD");
        if (e && !strcmp(e, "v1")) return 1;
        if (e && !strcmp(e, "v2")) return 2;
        return 0;
    }();
    return mode;
}

bool gdn_tc_backward2_supported(const float *x, const float *g, int64_t B, int64_t C, int64_t HW) {
    Bwd2Geometry geo;
    if (bwd2_mode() == 1) return false;
    if (!bwd2_geometry(C, &geo)) return false;
    if (HW % tcb2::TILE != 0 || HW >= (1 << 24) || B >= (1 << 24) || B * HW / tcb2::TILE >= (1ll << 31) ||
        B * C >= (1ll << 31))
        return false;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(g) & 15)) return false;
    return tmah::encode_tiled() != nullptr;
}

// x-prefetch variant of the single-stage layers (80 <= C <= 111): [u | x2 | x2] + ONE streamed gamma buffer.  Worth it
// only when a CTA walks through several tiles (the pack kernel and the per-tile refills of gamma cost a little).
static bool bwd2_xpf_geometry(int64_t C, int64_t ntiles, const Bwd2Geometry &geo, size_t *smem) {
    static const bool enabled = []() { const char *e = getenv("MMNC_GDN_XPF"); return !(e && !strcmp(e, "0")); }();
    if (!enabled || geo.stream || geo.groups != 1 || geo.stages != 1) return false;
    if (ntiles < 3 * (int64_t)sm_count()) return false;
    const size_t R8 = (size_t)(C + 1 + 7) / 8, buf = R8 * 4096, gam = (size_t)geo.P * geo.P * 4;
    size_t bytes = 3 * buf + gam;
    const size_t over_b = 2 * buf + (3 * R8 + (size_t)geo.P / 8) * 1024, over_a = (3 * R8 + 16) * 1024;
    if (bytes < over_b) bytes = over_b;
    if (bytes < over_a) bytes = over_a;
    *smem = bytes + 1024;
    return *smem + 256 <= 227 * 1024;
}

bool gdn_tc_backward2_prefetches(int64_t B, int64_t C, int64_t HW) {
    Bwd2Geometry geo;
    size_t smem;
    return bwd2_geometry(C, &geo) && bwd2_xpf_geometry(C, B * HW / tcb2::TILE, geo, &smem);
}

bool gdn_tc_backward2_streams(int64_t C) {
    Bwd2Geometry geo;
    return bwd2_geometry(C, &geo) && geo.stream;
}

size_t gdn_tc_backward2_workspace(int64_t B, int64_t C, int64_t HW) {
    Bwd2Geometry geo;
    if (!bwd2_geometry(C, &geo)) return 0;
    (void)B; (void)HW;
    const bool may_stream = geo.stream || (geo.groups == 1 && geo.stages == 1);  // streamed gamma or the x-prefetch variant
    const size_t packed = may_stream ? 2 * sizeof(float) * (size_t)geo.P * geo.P + 256 : 0;
    return sizeof(float) * (size_t)sm_count() * geo.groups * C * (C + 1) + 256 + packed;
}

static int make_map(CUtensorMap *m, const float *p, int64_t B, int64_t C, int64_t HW) {
    return tmah::tensor_map_3d(m, p, (uint64_t)HW, (uint64_t)C, (uint64_t)B, (uint64_t)HW * 4, (uint64_t)C * HW * 4, 32,
                               (uint32_t)C, 1, CU_TENSOR_MAP_SWIZZLE_128B, "gdn_tc_backward2");
}

int gdn_tc_backward2(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const GdnParams &prm,
                     int inverse, float *dx, float *dbeta, float *dgamma, void *workspace, size_t workspace_bytes,
                     cudaStream_t s) {
    Bwd2Geometry geo;
    if (!bwd2_geometry(C, &geo)) {
        set_error("gdn_tc_backward2: C = %lld not supported", (long long)C);
        return MMNC_ERR_UNSUPPORTED;
    }
    const int64_t ntiles = B * HW / tcb2::TILE;
    int64_t grid = sm_count();
    if (grid > ntiles) grid = ntiles;
    const int ksplit = (int)grid * geo.groups;
    MMNC_REQUIRE(workspace_bytes >= sizeof(float) * (size_t)ksplit * C * (C + 1), "gdn_backward: workspace too small");
    CUtensorMap tm_x, tm_g;
    if (int rc = make_map(&tm_x, x, B, C, HW)) return rc;
    if (int rc = make_map(&tm_g, g, B, C, HW)) return rc;
    using Kernel = void (*)(const CUtensorMap, const CUtensorMap, int, int, int, const GdnParams, float *, float *,
                            int, uint32_t, const float *);
    Kernel kernel = nullptr;
    // two threads per pixel everywhere: four (1024 threads x 64 registers on the two-group instances, 512 x 128 with f
    // parked in TMEM on the single-group ones) was measured slower on every layer shape (see the kernel's comment)
    const int tpp = 2;
#define MMNC_PICK(N, G, S) (inverse ? (Kernel)gdn_tc_backward2_kernel<N, G, S, 2, true, false> : (Kernel)gdn_tc_backward2_kernel<N, G, S, 2, false, false>)
#define MMNC_PICK_STREAM(N) (inverse ? (Kernel)gdn_tc_backward2_kernel<N, 1, 1, 2, true, true> : (Kernel)gdn_tc_backward2_kernel<N, 1, 1, 2, false, true>)
#define MMNC_PICK_XPF(N) (inverse ? (Kernel)gdn_tc_backward2_kernel<N, 1, 1, 2, true, true, true> : (Kernel)gdn_tc_backward2_kernel<N, 1, 1, 2, false, true, true>)
    size_t xpf_smem = 0;
    const bool xpf = bwd2_xpf_geometry(C, ntiles, geo, &xpf_smem);
    if (xpf) { geo.smem = xpf_smem; geo.stream = true; }
    const int key = (xpf ? 2000 : (geo.stream ? 1000 : 0)) + (geo.P / 16) * 100 + geo.groups * 10 + geo.stages;
    switch (key) {
        case 223: kernel = MMNC_PICK(2, 2, 3); break;
        case 323: kernel = MMNC_PICK(3, 2, 3); break;
        case 423: kernel = MMNC_PICK(4, 2, 3); break;
        case 512: kernel = MMNC_PICK(5, 1, 2); break;
        case 611: kernel = MMNC_PICK(6, 1, 1); break;
        case 711: kernel = MMNC_PICK(7, 1, 1); break;
        case 1811: kernel = MMNC_PICK_STREAM(8); break;   // C = 112 .. 127
        case 1911: kernel = MMNC_PICK_STREAM(9); break;   // C = 128
        case 2611: kernel = MMNC_PICK_XPF(6); break;      // C = 80 .. 95, long tile sequences
        case 2711: kernel = MMNC_PICK_XPF(7); break;      // C = 96 .. 111, long tile sequences
        default: break;
    }
#undef MMNC_PICK
#undef MMNC_PICK_STREAM
#undef MMNC_PICK_XPF
    if (kernel == nullptr) {
        set_error("gdn_tc_backward2: no kernel instance for C = %lld (P %d, groups %d, stages %d)", (long long)C, geo.P,
                  geo.groups, geo.stages);
        return MMNC_ERR_UNSUPPORTED;
    }
    if (int rc = tmah::ensure_dynamic_smem(kernel, geo.smem)) return rc;
    float *part = static_cast<float *>(workspace);
    const float *packed_b = nullptr;
    if (geo.stream) {
        // the packed tiles live behind the partials (128-byte aligned: the bulk copies need 16)
        const size_t part_bytes = (sizeof(float) * (size_t)ksplit * C * (C + 1) + 127) / 128 * 128;
        const size_t need = part_bytes + 2 * sizeof(float) * (size_t)geo.P * geo.P;
        MMNC_REQUIRE(workspace_bytes >= need, "gdn_backward: workspace too small for the packed gamma tiles");
        MMNC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "gdn_backward: workspace must be 16-byte aligned");
        uint32_t *dst = reinterpret_cast<uint32_t *>(static_cast<char *>(workspace) + part_bytes);
        gdn_pack_gamma_kernel<<<(geo.P * geo.P + 255) / 256, 256, 0, s>>>(prm, (int)C, geo.P, dst);
        if (int rc = after_launch("gdn_pack_gamma_kernel")) return rc;
        packed_b = reinterpret_cast<const float *>(dst);
    }
    kernel<<<(unsigned)grid, geo.groups * tpp * 128, geo.smem, s>>>(tm_x, tm_g, (int)ntiles, (int)(HW / tcb2::TILE), (int)HW,
                                                                   prm, dx, part, (int)C, geo.tmem_cols, packed_b);
    if (int rc = after_launch("gdn_tc_backward2_kernel")) return rc;
    return gdn_reduce_partials(part, ksplit, (int)C, prm, dgamma, dbeta, s);
}

}  // namespace mmnc
