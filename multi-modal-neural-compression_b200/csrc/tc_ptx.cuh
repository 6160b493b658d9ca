// Inline-PTX wrappers for the Blackwell tensor-core path: tcgen05 (alloc / mma / commit / ld / st / fences),
// mbarrier, and the shared-memory matrix / instruction descriptors.  sm_100a only.
#pragma once
#include <stdint.h>

namespace mmnc {
namespace tc {

constexpr int TILE_M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
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
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void st_shared_u32(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one MUFU.RSQ (<= 2 ulp); n >= beta_min > 0 so flushing denormals is harmless.  torch's CUDA rsqrt is the same op.
__device__ __forceinline__ float fast_rsqrt(float v) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// base + c * stride_bytes with a 32-bit stride: one IMAD.WIDE.U32 per address instead of a sign-extended 64-bit add chain
__device__ __forceinline__ const float *chan_ptr(const float *base, uint32_t stride_bytes, int c) {
    return reinterpret_cast<const float *>(reinterpret_cast<const char *>(base) + (uint64_t)stride_bytes * (uint32_t)c);
}
__device__ __forceinline__ float *chan_ptr(float *base, uint32_t stride_bytes, int c) {
    return reinterpret_cast<float *>(reinterpret_cast<char *>(base) + (uint64_t)stride_bytes * (uint32_t)c);
}

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// Same rounding (nearest, ties away from zero, on the sign-magnitude bit pattern) in two integer instructions.
// cvt.rna.tf32.f32 has no SASS instruction on sm_100a: ptxas expands it into ~6 instructions with Inf / NaN guards.
// Here Inf stays Inf (mantissa 0), a NaN stays a NaN or becomes Inf only if its payload is all ones - irrelevant for
// the finite activations this is used on.
__device__ __forceinline__ uint32_t to_tf32_fast(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }

// shared-memory matrix descriptor: K-major, no swizzle.  Core matrix = 8 rows x 16 B stored as 128 contiguous
// bytes; LBO = byte distance between core matrices adjacent in K, SBO = between 8-row groups (both >> 4).
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version: Blackwell
    return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// instruction descriptor for kind::tf32: D = F32, A = B = TF32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}


// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

// general shared-memory matrix descriptor (see make_b_desc): layout_type 0 = no swizzle, 2 = 128-byte swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout_type & 7u) << 61;
    return d;
}

// instruction descriptor, kind::tf32, D = F32, M = 128; a_mn / b_mn select MN-major ("transposed") operands
__device__ __forceinline__ uint32_t make_idesc_ex(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}


// ---- TMA (cp.async.bulk.tensor) and friends -------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one box of a rank-3 tensor map -> shared memory; completion is counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const void *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// named barrier over a subset of the CTA (id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}

// ---- MMA issue with the shared-memory descriptors assembled INSIDE the asm block ---------------------------------
// A descriptor differs from step to step only in its 14-bit address field.  Building {lo + step, hi} inside a volatile
// asm keeps ptxas from hoisting a dozen loop-invariant 64-bit descriptors out of the tile loop and spilling them:
// with the shared-memory carve-out at its maximum the L1 is a few KB, a spilled descriptor comes back from L2
// (~340 cycles per LDL, measured with ncu), and the single MMA-issuing thread serialises those.
template <int A_COL_STEP, int B_STEP16>
__device__ __forceinline__ void mma_tf32_ts_step(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 ta, lo;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "add.u32 ta, %1, %6;\n\t"
        "add.u32 lo, %2, %7;\n\t"
        "mov.b64 db, {lo, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ta], db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "n"(A_COL_STEP), "n"(B_STEP16)
        : "memory");
}
// both operands in shared memory; `off16` is a run-time offset (in 16-byte units) added to both address fields
template <int STEP16>
__device__ __forceinline__ void mma_tf32_ss_step(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                                 uint32_t off16, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 la, lb;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "add.u32 la, %1, %4;\n\t"
        "add.u32 lb, %2, %4;\n\t"
        "add.u32 la, la, %7;\n\t"
        "add.u32 lb, lb, %7;\n\t"
        "mov.b64 da, {la, %3};\n\t"
        "mov.b64 db, {lb, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(off16), "r"(idesc), "r"(accumulate), "n"(STEP16)
        : "memory");
}
// the two 32-bit halves of make_desc()
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((layout_type & 7u) << 29);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}

}  // namespace tc
}  // namespace mmnc
