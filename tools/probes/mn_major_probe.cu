// Probe: does tcgen05.mma kind::tf32 accept MN-major ("transposed") shared-memory operands with the 128-byte swizzle on
// sm_100a?  This is the layout a TMA box of [pixel rows x 32 channels] of a channels-last (NHWC) tensor lands in, and
// what a d-gamma contraction over the pixel index would need for both operands.  Build: see tools/probes/Makefile.
//   D[M=128 x N=32] = sum_k A[m][k] * B[k][n],  K = 32 "pixels";
//   memory: A = 4 chunks (32 channels each) of [32 pixel rows][128 B], B = 1 chunk, rows XOR-swizzled in 16-byte units.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../multi-modal-neural-compression_b200/csrc/tc_ptx.cuh"

using namespace mmnc::tc;

constexpr int KPIX = 32, MCH = 128, NCH = 32;

__global__ void __launch_bounds__(128) probe_kernel(const float *a_g, const float *b_g, float *d_g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *As = reinterpret_cast<float *>(smem);                       // 4 chunks x 32 rows x 128 B = 16 KB
    float *Bs = reinterpret_cast<float *>(smem + 4 * KPIX * 128);      // 1 chunk  x 32 rows x 128 B =  4 KB
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 32);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    // a_g[k][m] (pixel-major, channel contiguous = NHWC), b_g[k][n]
    for (int i = threadIdx.x; i < KPIX * MCH; i += 128) {
        const int k = i / MCH, m = i % MCH, chunk = m >> 5, c = m & 31;
        const int unit = (c >> 2) ^ (k & 7);
        As[chunk * (KPIX * 32) + k * 32 + unit * 4 + (c & 3)] = __uint_as_float(to_tf32(a_g[i]));
    }
    for (int i = threadIdx.x; i < KPIX * NCH; i += 128) {
        const int k = i / NCH, c = i % NCH;
        const int unit = (c >> 2) ^ (k & 7);
        Bs[k * 32 + unit * 4 + (c & 3)] = __uint_as_float(to_tf32(b_g[i]));
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_ex(NCH, true, true);
        for (int kg = 0; kg < KPIX / 8; ++kg) {
            // MN-major, SWIZZLE_128B: LBO = stride between 32-element MN chunks, SBO = stride between 8-row K groups
            const uint64_t da = make_desc(smem_u32(As) + kg * 1024, KPIX * 128, 1024, 2);
            const uint64_t db = make_desc(smem_u32(Bs) + kg * 1024, KPIX * 128, 1024, 2);
            mma_tf32_ss(tmem, da, db, idesc, kg > 0 ? 1u : 0u);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after();
    uint32_t r[16];
    for (int q = 0; q < 2; ++q) {
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + q * 16, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) d_g[threadIdx.x * NCH + q * 16 + j] = __uint_as_float(r[j]);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
    std::vector<float> a(KPIX * MCH), b(KPIX * NCH), d(MCH * NCH), want(MCH * NCH, 0.f);
    srand(1);
    for (auto &v : a) v = (float)(rand() % 17 - 8) / 8.f;   // exactly representable in tf32
    for (auto &v : b) v = (float)(rand() % 13 - 6) / 4.f;
    for (int m = 0; m < MCH; ++m)
        for (int n = 0; n < NCH; ++n)
            for (int k = 0; k < KPIX; ++k) want[m * NCH + n] += a[k * MCH + m] * b[k * NCH + n];
    float *da, *db, *dd;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, d.size() * 4);
    const int smem = 4 * KPIX * 128 + KPIX * 128 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(da, db, dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxabs = 0;
    int zeros = 0;
    for (size_t i = 0; i < d.size(); ++i) {
        maxerr = fmax(maxerr, fabs(d[i] - want[i]));
        maxabs = fmax(maxabs, fabs(want[i]));
        zeros += d[i] == 0.f;
    }
    printf("MN-major SW128 tf32 SS MMA: max |err| %.3g (max |want| %.3g), %d of %zu outputs are exactly 0 -> %s\n", maxerr,
           maxabs, zeros, d.size(), maxerr < 1e-3 ? "WORKS" : "DOES NOT MATCH");
    printf("d[0..3] = %g %g %g %g ; want %g %g %g %g\n", d[0], d[1], d[2], d[3], want[0], want[1], want[2], want[3]);
    return maxerr < 1e-3 ? 0 : 1;
}
