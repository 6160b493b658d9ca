// How many SM cycles does one tcgen05.mma kind::tf32 (M = 128, N = 256, K = 8) take when nothing else runs?
// One CTA per SM issues REPS x 32 MMAs back to back (A in TMEM or in shared memory, B in shared memory, K-major no-swizzle
// core-matrix layout, garbage data) and one commit; clock64 around issue + completion.  The wide GDN kernels see ~210
// cycles per MMA inside their contractions (tools/probes/wide_phase_probe.py); the nominal TF32 rate would be 131.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multi-modal-neural-compression_b200/csrc \
//        -o tools/probes/mma_rate_probe tools/probes/mma_rate_probe.cu && gpurun -- tools/probes/mma_rate_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tc_ptx.cuh"

using namespace mmnc::tc;

template <int N, bool kSS, int kCommitEvery = 0, int kThreads = 128, bool kSpin = false, bool kTryWait = false>
__global__ void __launch_bounds__(kThreads, 1) probe(long long *out, int reps, const float *src, int stream_chunks, float *sink = nullptr, int ldst = 0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, done_bar;
    __shared__ uint64_t side_bar[8];  // targets of the intermediate commits (never waited on)
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x == 0) { mbar_init(&done_bar, 1); asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&done_bar)) : "memory"); mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&side_bar[i], 1); }
    for (int i = threadIdx.x; i < 48 * 1024; i += kThreads) reinterpret_cast<float *>(smem)[i] = 1.0f;  // 192 KB of finite data
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint32_t b0 = smem_u32(smem);
        const uint32_t hi = desc_hi(1024u, 0);  // 8 cores of 128 B per 8-row group: a 32-wide K chunk
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int ks = 0; ks < 32; ++ks) {
                if (kTryWait && (ks & 3) == 0) mbar_wait(&done_bar, 0);  // a barrier that completed long ago: what does the poll cost?
                // walk through 6 x 32 KB of B like the ring does (a different chunk every four K steps)
                const uint32_t b = b0 + (uint32_t)(((r * 8 + ks / 4) % 6) * 32768) + (uint32_t)((ks & 3) * 256);
                if (kSS)
                    mma_tf32_ss(tmem + 256, make_desc(b0 + 6 * 32768 - 32768, 128, 1024, 0), make_desc(b, 128, 1024, 0), idesc,
                                ks > 0 ? 1u : 0u);
                else
                    mma_tf32_ts(tmem + 256, tmem + (uint32_t)(ks * 8), make_desc(b, 128, 1024, 0), idesc, ks > 0 ? 1u : 0u);
                if (kCommitEvery && (ks % kCommitEvery) == kCommitEvery - 1) mma_commit(&side_bar[(ks / kCommitEvery) & 7]);
            }
        }
        const long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        (void)hi;
    } else if (threadIdx.x == 32 && stream_chunks > 0) {
        // a second thread keeps `stream_chunks` 32 KB bulk copies per 32 MMAs landing in the SAME shared-memory region the
        // MMAs read (contents are all ones either way): the traffic of the ring refills
        __shared__ uint64_t cbar;
        mbar_init(&cbar, 1);
        uint32_t par = 0;
        for (int r = 0; r < reps; ++r)
            for (int c = 0; c < stream_chunks; ++c) {
                mbar_arrive_expect_tx(&cbar, 32768);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem) + (uint32_t)(((r * stream_chunks + c) % 6) * 32768)),
                               "l"(reinterpret_cast<uint64_t>(src) + (uint64_t)(((r * stream_chunks + c) % 16) * 32768)),
                               "r"(32768), "r"(smem_u32(&cbar)) : "memory");
                mbar_wait(&cbar, par);
                par ^= 1;
            }
    } else if (ldst != 0 && threadIdx.x >= 32) {
        // the other warps stream 128 KB of global stores (ldst = 1) or loads (ldst = 2) per 32 MMAs, like the epilogues
        float acc = 0.f;
        const int per = 32768 / (kThreads - 32);
        float *mine = sink + (size_t)blockIdx.x * 32768;
        for (int r = 0; r < reps; ++r)
            for (int i = 0; i < per; ++i) {
                const int idx = i * (kThreads - 32) + (threadIdx.x - 32);
                if (ldst == 1) __stcs(mine + idx, (float)r);
                else acc += __ldcs(mine + idx);
            }
        if (acc == 123.f) out[1] = 0;
    } else if (kSpin) {
        mbar_wait(&bar, 0);  // what the compute threads of the GDN kernels do while the leader issues: poll the barrier
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, bool kSS, int kCommitEvery = 0, int kThreads = 128, bool kSpin = false, bool kTryWait = false>
static void run(const char *what, int grid, int stream_chunks = 0, int ldst = 0) {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    float *src, *sink;
    cudaMalloc(&src, 16 * 32768);
    cudaMalloc(&sink, (size_t)256 * 32768 * 4);
    cudaMemset(sink, 0, (size_t)256 * 32768 * 4);
    {
        float *ones = new float[16 * 8192];
        for (int i = 0; i < 16 * 8192; ++i) ones[i] = 1.0f;
        cudaMemcpy(src, ones, 16 * 32768, cudaMemcpyHostToDevice);
        delete[] ones;
    }
    const int reps = 64;
    const size_t smem = 193 * 1024 + 1024;
    cudaFuncSetAttribute(probe<N, kSS, kCommitEvery, kThreads, kSpin, kTryWait>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) {
        probe<N, kSS, kCommitEvery, kThreads, kSpin, kTryWait><<<grid, kThreads, smem>>>(d, reps, src, stream_chunks, sink, ldst);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-44s grid %3d: issue %7.1f cycles / MMA, issue + drain %7.1f cycles / MMA\n", what, grid,
           (double)h[0] / (reps * 32), (double)h[1] / (reps * 32));
    cudaFree(d);
    cudaFree(src);
    cudaFree(sink);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<256, false>("tf32 M128 N256 K8, A in TMEM, B in smem", 1);
    run<256, false>("tf32 M128 N256 K8, A in TMEM, B in smem", sms);
    run<128, false>("tf32 M128 N128 K8, A in TMEM, B in smem", sms);
    run<256, true>("tf32 M128 N256 K8, A and B in smem", sms);
    run<128, true>("tf32 M128 N128 K8, A and B in smem", sms);
    run<256, false, 4>("... A in TMEM, a tcgen05.commit every 4 MMAs", sms);
    run<256, false, 1>("... A in TMEM, a tcgen05.commit every MMA", sms);
    run<256, false, 4, 512, false>("... 512 threads, 511 at a barrier", sms);
    run<256, false, 4, 512, true>("... 512 threads, 511 polling the mbarrier", sms);
    run<256, false, 4, 256, true>("... 256 threads, 255 polling the mbarrier", sms);
    run<256, false, 4, 128, false>("... + 2 bulk copies of 32 KB per 32 MMAs", sms, 2);
    run<256, false, 4, 128, false>("... + 8 bulk copies of 32 KB per 32 MMAs", sms, 8);
    run<256, false, 4, 512, false>("... 512 threads storing 128 KB per 32 MMAs", sms, 0, 1);
    run<256, false, 4, 512, false>("... 512 threads loading 128 KB per 32 MMAs", sms, 0, 2);
    run<256, false, 4, 512, true, true>("... a try_wait on a completed barrier per 4 MMAs, 511 polling", sms);
    return 0;
}
