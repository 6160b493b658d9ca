// K5: batched rANS encode / decode, bit-exact with CompressAI 1.2.4's RansEncoder / RansDecoder
// (format in Appendix A.7).
//
// Replaces the per-image Python loop + five .tolist() marshals + pybind11 call of EntropyModel.compress /
// decompress (reference call sites /root/reference/src/models/multi_task_compressor.py:509, 543, 546).
//
// The format is a single 64-bit rANS state per image stream, so the only parallelism is across streams:
//   pass 1 (fully parallel, thread <-> symbol): symbol -> (cdf start, range, escape payload), written
//           TRANSPOSED as staging[position][stream] so that pass 2 reads are coalesced across streams;
//   pass 2 (thread <-> stream): walks positions last -> first, pushes words backwards into the stream's slab;
//   compact: exclusive scan of byte counts + gather into one buffer, so the host does a single D2H.
// Decoding is thread <-> stream with a binary search of the CDF row instead of CompressAI's linear scan.
#include "common.cuh"
#include "hd_math.cuh"

namespace mmnc {

constexpr int RANS_MAP_THREADS = 256;
constexpr int RANS_STREAM_THREADS = 32;

__device__ __forceinline__ int32_t stream_index(const int32_t *indexes, int64_t channel_period, int n_cdfs,
                                                int64_t stream, int64_t n_sym, int64_t pos) {
    if (indexes != nullptr) return indexes[stream * n_sym + pos];
    return (int32_t)((pos / channel_period) % n_cdfs);
}

__global__ void __launch_bounds__(RANS_MAP_THREADS)
rans_map_kernel(const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, int64_t channel_period,
                int64_t n_streams, int64_t n_sym, const int32_t *__restrict__ cdf, int n_cdfs, int cdf_stride,
                const int32_t *__restrict__ cdf_sizes, const int32_t *__restrict__ offsets,
                uint2 *__restrict__ staging, int32_t *__restrict__ nbytes) {
    const int64_t total = n_streams * n_sym;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        // i enumerates (position, stream) with stream fastest so that the staging write is coalesced
        const int64_t pos = i / n_streams, stream = i - pos * n_streams;
        const int32_t ci = stream_index(indexes, channel_period, n_cdfs, stream, n_sym, pos);
        uint2 e = make_uint2(0u, 0u);
        if (ci < 0 || ci >= n_cdfs) {
            nbytes[stream] = -1;  // malformed index: flag the stream (pass 2 keeps the flag)
        } else {
            const int32_t max_value = cdf_sizes[ci] - 2;
            uint32_t raw;
            const int slot = rans_map_symbol(symbols[stream * n_sym + pos], offsets[ci], max_value, &raw);
            const int32_t *row = cdf + (int64_t)ci * cdf_stride;
            const uint32_t start = (uint32_t)row[slot];
            const uint32_t range = (uint32_t)(row[slot + 1] - row[slot]);
            e = make_uint2((start & 0xFFFFu) | (range << 16), raw);
        }
        staging[i] = e;
    }
}

__global__ void __launch_bounds__(RANS_STREAM_THREADS)
rans_encode_kernel(const uint2 *__restrict__ staging, int64_t n_streams, int64_t n_sym, uint32_t *__restrict__ slabs,
                   int64_t slab_words, int32_t *__restrict__ nbytes) {
    const int64_t stream = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (stream >= n_streams) return;
    if (nbytes[stream] < 0) return;
    uint32_t *slab = slabs + stream * slab_words;
    RansEnc enc;
    enc.init(slab + slab_words);
    bool ok = true;
    for (int64_t pos = n_sym - 1; pos >= 0; --pos) {
        const uint2 e = staging[pos * n_streams + stream];
        const uint32_t start = e.x & 0xFFFFu, range = e.x >> 16;
        if (range == 0u || enc.ptr - slab < 16) { ok = false; break; }
        if (start + range == 65536u) rans_put_escape_reversed(enc, e.y);  // the escape slot is the last one
        enc.put(start, range);
    }
    if (!ok) { nbytes[stream] = -2; return; }
    enc.flush();
    nbytes[stream] = (int32_t)((slab + slab_words - enc.ptr) * (int64_t)sizeof(uint32_t));
}

// exclusive scan of max(nbytes, 0) into offsets[0..n]; single block of 1024 threads, chunked
__global__ void __launch_bounds__(1024)
rans_scan_kernel(const int32_t *__restrict__ nbytes, int64_t n, int64_t *__restrict__ offsets) {
    __shared__ int64_t warp_excl[32];
    __shared__ int64_t chunk_total;
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = (i < n && nbytes[i] > 0) ? (int64_t)nbytes[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_excl[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int64_t w = warp_excl[lane];
            int64_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_excl[lane] = winc - w;
            if (lane == 31) chunk_total = winc;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        if (i < n) offsets[i] = carry + warp_excl[warp] + (incl - v);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = carry_s;
}

__global__ void __launch_bounds__(128)
rans_gather_kernel(const uint32_t *__restrict__ slabs, int64_t slab_words, const int32_t *__restrict__ nbytes,
                   const int64_t *__restrict__ offsets, uint8_t *__restrict__ packed, int64_t capacity) {
    const int64_t stream = blockIdx.x;
    const int32_t nb = nbytes[stream];
    if (nb <= 0) return;
    const int64_t off = offsets[stream];
    if (off + nb > capacity) return;
    const int64_t words = nb >> 2;
    const uint32_t *src = slabs + stream * slab_words + (slab_words - words);
    uint32_t *dst = reinterpret_cast<uint32_t *>(packed + off);  // offsets are multiples of 4, packed is aligned
    for (int64_t w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
}

__global__ void __launch_bounds__(RANS_STREAM_THREADS)
rans_decode_kernel(const uint8_t *__restrict__ packed, const int64_t *__restrict__ offsets,
                   const int32_t *__restrict__ indexes, int64_t channel_period, int64_t n_streams, int64_t n_sym,
                   const int32_t *__restrict__ cdf, int n_cdfs, int cdf_stride, const int32_t *__restrict__ cdf_sizes,
                   const int32_t *__restrict__ cdf_offsets, int32_t *__restrict__ symbols,
                   int32_t *__restrict__ status) {
    const int64_t stream = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (stream >= n_streams) return;
    const int64_t b0 = offsets[stream], b1 = offsets[stream + 1];
    int32_t st = 0;
    if (b1 - b0 < 8) { status[stream] = -1; return; }
    RansDec dec;
    dec.init(packed + b0, packed + b1);
    for (int64_t pos = 0; pos < n_sym; ++pos) {
        const int32_t ci = stream_index(indexes, channel_period, n_cdfs, stream, n_sym, pos);
        if (ci < 0 || ci >= n_cdfs) { st = -3; break; }
        const int32_t *row = cdf + (int64_t)ci * cdf_stride;
        const int32_t len = cdf_sizes[ci];
        const int32_t max_value = len - 2;
        const int slot = rans_find_slot(row, len, dec.peek());
        if (slot < 0 || slot > max_value) { st = -4; break; }
        const uint32_t start = (uint32_t)row[slot];
        dec.advance(start, (uint32_t)(row[slot + 1] - row[slot]));
        int32_t value = slot;
        if (slot == max_value) value = dec.get_escape(max_value);
        symbols[stream * n_sym + pos] = value + cdf_offsets[ci];
        if (dec.overrun) { st = -2; break; }
    }
    status[stream] = st;
}

}  // namespace mmnc

using namespace mmnc;

extern "C" int64_t mmnc_rans_slab_words(int64_t n_sym) {
    // every symbol adds at most 16 bits (probability >= 2^-16) plus, when escaped, at most 9 nibbles;
    // the state starts with 31 bits and the flush writes 2 words.  +16 words of headroom (checked in-kernel).
    return (n_sym * 52 + 31) / 32 + 4 + 16;
}

extern "C" int mmnc_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int64_t channel_period,
                                      int64_t n_streams, int64_t n_sym, const int32_t *cdf, int n_cdfs,
                                      int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                                      void *staging, uint32_t *slabs, int64_t slab_words, int32_t *nbytes,
                                      void *stream) {
    MMNC_REQUIRE(n_streams >= 0 && n_sym >= 0, "rans_encode_batch: negative size");
    MMNC_REQUIRE(n_cdfs > 0 && cdf_stride > 1, "rans_encode_batch: empty CDF table (run update() first)");
    MMNC_REQUIRE(indexes != nullptr || channel_period > 0, "rans_encode_batch: need indexes or channel_period");
    MMNC_REQUIRE(slab_words >= mmnc_rans_slab_words(n_sym), "rans_encode_batch: slab_words too small");
    if (n_streams == 0) return MMNC_OK;
    MMNC_REQUIRE(cdf && cdf_sizes && offsets && slabs && nbytes && (n_sym == 0 || (symbols && staging)),
                 "rans_encode_batch: null pointer");
    cudaStream_t s = as_stream(stream);
    MMNC_CUDA(cudaMemsetAsync(nbytes, 0, sizeof(int32_t) * (size_t)n_streams, s));
    const int64_t total = n_streams * n_sym;
    if (total > 0) {
        int64_t blocks = (total + RANS_MAP_THREADS - 1) / RANS_MAP_THREADS;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        rans_map_kernel<<<(unsigned)blocks, RANS_MAP_THREADS, 0, s>>>(symbols, indexes, channel_period, n_streams,
                                                                      n_sym, cdf, n_cdfs, cdf_stride, cdf_sizes,
                                                                      offsets, static_cast<uint2 *>(staging), nbytes);
        if (int rc = after_launch("rans_map_kernel")) return rc;
    }
    rans_encode_kernel<<<(unsigned)((n_streams + RANS_STREAM_THREADS - 1) / RANS_STREAM_THREADS),
                         RANS_STREAM_THREADS, 0, s>>>(static_cast<const uint2 *>(staging), n_streams, n_sym, slabs,
                                                      slab_words, nbytes);
    return after_launch("rans_encode_kernel");
}

extern "C" int mmnc_rans_compact(const uint32_t *slabs, int64_t slab_words, const int32_t *nbytes,
                                 int64_t n_streams, int64_t *offsets, uint8_t *packed, int64_t packed_capacity,
                                 void *stream) {
    MMNC_REQUIRE(n_streams >= 0, "rans_compact: negative size");
    MMNC_REQUIRE(offsets, "rans_compact: null offsets");
    cudaStream_t s = as_stream(stream);
    if (n_streams == 0) {
        MMNC_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int64_t), s));
        return MMNC_OK;
    }
    MMNC_REQUIRE(slabs && nbytes && packed, "rans_compact: null pointer");
    rans_scan_kernel<<<1, 1024, 0, s>>>(nbytes, n_streams, offsets);
    if (int rc = after_launch("rans_scan_kernel")) return rc;
    rans_gather_kernel<<<(unsigned)n_streams, 128, 0, s>>>(slabs, slab_words, nbytes, offsets, packed,
                                                          packed_capacity);
    return after_launch("rans_gather_kernel");
}

extern "C" int mmnc_rans_decode_batch(const uint8_t *packed, const int64_t *offsets, const int32_t *indexes,
                                      int64_t channel_period, int64_t n_streams, int64_t n_sym, const int32_t *cdf,
                                      int n_cdfs, int cdf_stride, const int32_t *cdf_sizes,
                                      const int32_t *cdf_offsets, int32_t *symbols, int32_t *status, void *stream) {
    MMNC_REQUIRE(n_streams >= 0 && n_sym >= 0, "rans_decode_batch: negative size");
    MMNC_REQUIRE(n_cdfs > 0 && cdf_stride > 1, "rans_decode_batch: empty CDF table (run update() first)");
    MMNC_REQUIRE(indexes != nullptr || channel_period > 0, "rans_decode_batch: need indexes or channel_period");
    if (n_streams == 0) return MMNC_OK;
    MMNC_REQUIRE(packed && offsets && cdf && cdf_sizes && cdf_offsets && status && (n_sym == 0 || symbols),
                 "rans_decode_batch: null pointer");
    rans_decode_kernel<<<(unsigned)((n_streams + RANS_STREAM_THREADS - 1) / RANS_STREAM_THREADS),
                         RANS_STREAM_THREADS, 0, as_stream(stream)>>>(packed, offsets, indexes, channel_period,
                                                                      n_streams, n_sym, cdf, n_cdfs, cdf_stride,
                                                                      cdf_sizes, cdf_offsets, symbols, status);
    return after_launch("rans_decode_kernel");
}
