// K5: batched rANS encode / decode, bit-exact with CompressAI 1.2.4's RansEncoder / RansDecoder
// (format in Appendix A.7).
//
// Replaces the per-image Python loop + five .tolist() marshals + pybind11 call of EntropyModel.compress /
// decompress (reference call sites /root/reference/src/models/multi_task_compressor.py:509, 543, 546).
//
// The format is a single 64-bit rANS state per image stream, so the only parallelism is across streams:
//   tables : the (n_cdfs x stride) int32 CDF table becomes a RAGGED uint16 table once per update() (53 KB for the
//            64 x 3133 Gaussian table instead of 802 KB): small enough to live in shared memory;
//   pass 1 (fully parallel, thread <-> symbol): symbol -> (cdf start : 16 | range : 16) and the exact 64-bit reciprocal
//           of the range.  The ragged table is staged in shared memory when the batch re-uses it often enough;
//   pass 2 (WARP <-> stream): walks positions last -> first.  The coder state is a chain of dependent integer
//           operations, so a stream cannot use more than one thread's worth of arithmetic - but the other 31 lanes
//           take the memory latency off the chain: the warp fetches 32 staged entries with one coalesced load and
//           hands them to the (lane-uniform) state update by shuffle.  The update itself is a multiply-high + shift
//           instead of a 64-bit divide and modulo (hd_math.cuh: rans_div); escapes (rare) re-derive their payload from
//           the symbol itself; lane 0 pushes the words backwards into the stream's slab;
//   compact: exclusive scan of byte counts + gather into one buffer, so the host does one D2H.
// Decoding is WARP <-> stream as well, with the ragged table in shared memory: the 32 lanes probe 32 pivots of the CDF
// row at once, so a 3133-entry row is resolved in three rounds of (shared load, ballot) instead of twelve dependent
// binary-search steps (CompressAI scans linearly); CDF indexes are fetched 32 at a time, stream words one ahead,
// decoded symbols leave in coalesced groups of 32.
#include "common.cuh"
#include "hd_math.cuh"
#include "tma_host.cuh"

namespace mmnc {

constexpr int RANS_MAP_THREADS = 256;
constexpr int RANS_DECODE_THREADS = 128;
constexpr int RANS_WARP_BLOCK = 128;  // sequential passes: one warp per stream, four streams per block
constexpr size_t RANS_SMEM_TABLE_MAX = 200 * 1024;

struct RansTables {
    const uint16_t *ragged;     // concatenated rows, cdf values modulo 2^16
    const int32_t *row_start;   // n_cdfs + 1 entries
    const int32_t *sizes;       // entries per row (CompressAI's _cdf_length)
    const int32_t *offsets;     // CompressAI's _offset
    int n_cdfs;
    int ragged_len;
};

__device__ __forceinline__ int32_t stream_index(const int32_t *indexes, int64_t channel_period, int n_cdfs,
                                                int64_t stream, int64_t n_sym, int64_t pos) {
    if (indexes != nullptr) return indexes[stream * n_sym + pos];
    return (int32_t)((pos / channel_period) % n_cdfs);
}

// RansEnc::put_bits for the warp-per-stream pass: lane-uniform state, lane 0 stores
__device__ __forceinline__ void warp_put_bits(uint64_t &x, uint32_t *&ptr, int lane, uint32_t val) {
    const uint64_t x_max = ((RANS_L >> 16) << 32) * (uint64_t)(1u << (16 - RANS_BYPASS_BITS));
    if (x >= x_max) { --ptr; if (lane == 0) *ptr = (uint32_t)x; x >>= 32; }
    x = (x << RANS_BYPASS_BITS) | val;
}

// ---------------------------------------------------------------------------------------------- ragged tables
__global__ void __launch_bounds__(256)
rans_pack_tables_kernel(const int32_t *__restrict__ cdf, const int32_t *__restrict__ sizes, int n_cdfs, int stride,
                        int32_t *__restrict__ row_start, uint16_t *__restrict__ ragged, int ragged_capacity) {
    __shared__ int total_s;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int r = 0; r < n_cdfs; ++r) {
            row_start[r] = acc;
            int len = sizes[r];
            len = len < 0 ? 0 : (len > stride ? stride : len);
            acc += len;
        }
        row_start[n_cdfs] = acc;
        total_s = acc;
    }
    __syncthreads();
    if (total_s > ragged_capacity) return;  // the host sized the buffer from the same lengths; defensive
    for (int64_t i = threadIdx.x; i < (int64_t)n_cdfs * stride; i += blockDim.x) {
        const int r = (int)(i / stride), c = (int)(i - (int64_t)r * stride);
        if (c < sizes[r]) ragged[row_start[r] + c] = (uint16_t)((uint32_t)cdf[i] & 0xFFFFu);
    }
}

// ---------------------------------------------------------------------------------------------- encode, pass 1
template <bool kSmemTable>
__global__ void __launch_bounds__(RANS_MAP_THREADS)
rans_map_kernel(const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, int64_t channel_period,
                int64_t n_streams, int64_t n_sym, const RansTables tb, uint32_t *__restrict__ stage_sr,
                uint64_t *__restrict__ stage_rcp) {
    extern __shared__ __align__(16) uint8_t rans_smem[];
    const uint16_t *table = tb.ragged;
    if (kSmemTable) {
        // cooperative 16-byte copies (the ragged buffer is padded to a multiple of 8 entries by the host)
        const uint4 *src = reinterpret_cast<const uint4 *>(tb.ragged);
        uint4 *dst = reinterpret_cast<uint4 *>(rans_smem);
        const int nvec = (tb.ragged_len + 7) >> 3;
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        table = reinterpret_cast<const uint16_t *>(rans_smem);
    }
    const int64_t total = n_streams * n_sym;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t stream = i / n_sym, pos = i - stream * n_sym;  // staging is [stream][position]
        const int32_t ci = stream_index(indexes, channel_period, tb.n_cdfs, stream, n_sym, pos);
        uint32_t sr = 0u;
        uint64_t rcp = 0ull;
        if (ci >= 0 && ci < tb.n_cdfs) {  // a malformed index leaves range = 0: pass 2 flags the stream
            const int32_t max_value = tb.sizes[ci] - 2;
            uint32_t raw;
            const int slot = rans_map_symbol(symbols[stream * n_sym + pos], tb.offsets[ci], max_value, &raw);
            if (max_value >= 0) {
                const uint16_t *row = table + tb.row_start[ci];
                const uint32_t start = row[slot];
                const uint32_t range = ((uint32_t)row[slot + 1] - start) & 0xFFFFu;  // the final 65536 is stored as 0
                sr = start | (range << 16);
                if (range > 1u) rcp = rans_reciprocal(range);
            }
        }
        stage_sr[i] = sr;
        stage_rcp[i] = rcp;
    }
}

// ---------------------------------------------------------------------------------------------- encode, pass 2
__global__ void __launch_bounds__(RANS_WARP_BLOCK)
rans_encode_kernel(const uint32_t *__restrict__ stage_sr, const uint64_t *__restrict__ stage_rcp,
                   const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, int64_t channel_period,
                   int64_t n_streams, int64_t n_sym, const RansTables tb, uint32_t *__restrict__ slabs,
                   int64_t slab_words, int32_t *__restrict__ nbytes) {
    const int lane = threadIdx.x & 31;
    const int64_t stream = (int64_t)blockIdx.x * (RANS_WARP_BLOCK / 32) + (threadIdx.x >> 5);
    if (stream >= n_streams) return;  // whole warps leave together
    uint32_t *slab = slabs + stream * slab_words;
    const uint32_t *sr_row = stage_sr + stream * n_sym;
    const uint64_t *rcp_row = stage_rcp + stream * n_sym;
    RansEnc enc;  // every lane carries the same state; only lane 0 stores
    uint32_t *ptr = slab + slab_words;
    enc.x = RANS_L;
    bool ok = true;
    for (int64_t hi = n_sym; hi > 0 && ok; hi -= 32) {
        // lane l holds the entry of position hi - 1 - l: one coalesced load per 32 symbols
        const int64_t mine = hi - 1 - lane;
        const uint32_t sr_l = mine >= 0 ? sr_row[mine] : 0u;
        const uint64_t rcp_l = mine >= 0 ? rcp_row[mine] : 0ull;
        const int count = hi < 32 ? (int)hi : 32;
#pragma unroll 4
        for (int k = 0; k < count; ++k) {
            const uint32_t sr = __shfl_sync(0xffffffffu, sr_l, k);
            const uint64_t rcp = __shfl_sync(0xffffffffu, rcp_l, k);
            const int64_t pos = hi - 1 - k;
            const uint32_t start = sr & 0xFFFFu, range = sr >> 16;
            if (range == 0u || ptr - slab < 16) { ok = false; break; }
            if (start + range == 65536u) {
                // the escape slot is the last one of its row; its payload is re-derived from the symbol (rare path)
                const int32_t ci = stream_index(indexes, channel_period, tb.n_cdfs, stream, n_sym, pos);
                uint32_t raw;
                rans_map_symbol(symbols[stream * n_sym + pos], tb.offsets[ci], tb.sizes[ci] - 2, &raw);
                const int nb = rans_nibbles(raw);
                // same order as rans_put_escape_reversed: payload nibbles high -> low, then the count digit(s)
                for (int j = nb - 1; j >= 0; --j) warp_put_bits(enc.x, ptr, lane, (raw >> (j * RANS_BYPASS_BITS)) & RANS_BYPASS_MAX);
                const int full = nb / RANS_BYPASS_MAX, rest = nb - full * RANS_BYPASS_MAX;
                warp_put_bits(enc.x, ptr, lane, (uint32_t)rest);
                for (int q = 0; q < full; ++q) warp_put_bits(enc.x, ptr, lane, RANS_BYPASS_MAX);
            }
            // RansEnc::put_rcp with the store predicated on lane 0
            const uint64_t x_max = ((RANS_L >> RANS_PRECISION) << 32) * range;
            if (enc.x >= x_max) { --ptr; if (lane == 0) *ptr = (uint32_t)enc.x; enc.x >>= 32; }
            const uint64_t q = rans_div(enc.x, range, rcp);
            enc.x = (q << RANS_PRECISION) + (enc.x - q * range) + start;
        }
    }
    if (lane != 0) return;
    if (!ok) { nbytes[stream] = -2; return; }
    ptr -= 2;
    ptr[0] = (uint32_t)enc.x;
    ptr[1] = (uint32_t)(enc.x >> 32);
    nbytes[stream] = (int32_t)((slab + slab_words - ptr) * (int64_t)sizeof(uint32_t));
}

// exclusive scan of max(nbytes, 0) into offsets[0..n]; single block of 1024 threads, chunked.  The per-stream byte
// counts / status codes are copied behind the offsets (meta = int64 offsets[n + 1] | int32 nbytes[n]) so that the host
// fetches everything it needs to size the second copy with ONE device-to-host transfer.
__global__ void __launch_bounds__(1024)
rans_scan_kernel(const int32_t *__restrict__ nbytes, int64_t n, int64_t *__restrict__ offsets) {
    __shared__ int64_t warp_excl[32];
    __shared__ int64_t chunk_total;
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int32_t *status_out = reinterpret_cast<int32_t *>(offsets + n + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const int32_t nb = (i < n) ? nbytes[i] : 0;
        if (i < n) status_out[i] = nb;
        const int64_t v = nb > 0 ? (int64_t)nb : 0;
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

// one warp per stream, four streams per block
__global__ void __launch_bounds__(128)
rans_gather_kernel(const uint32_t *__restrict__ slabs, int64_t slab_words, const int32_t *__restrict__ nbytes,
                   const int64_t *__restrict__ offsets, int64_t n_streams, uint8_t *__restrict__ packed,
                   int64_t capacity) {
    const int64_t stream = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (stream >= n_streams) return;
    const int32_t nb = nbytes[stream];
    if (nb <= 0) return;
    const int64_t off = offsets[stream];
    if (off + nb > capacity) return;
    const int64_t words = nb >> 2;
    const uint32_t *src = slabs + stream * slab_words + (slab_words - words);
    uint32_t *dst = reinterpret_cast<uint32_t *>(packed + off);  // offsets are multiples of 4, packed is aligned
    for (int64_t w = threadIdx.x & 31; w < words; w += 32) dst[w] = src[w];
}

// ---------------------------------------------------------------------------------------------- decode
// The stream bytes are whole 32-bit words at 4-byte aligned offsets (the host pads); one word is always prefetched.
struct RansDecW {
    uint64_t x;
    const uint32_t *p, *end;
    uint32_t next;
    bool overrun;
    __device__ __forceinline__ uint32_t word() {
        const uint32_t w = next;
        if (p >= end) { overrun = true; return 0u; }
        ++p;
        next = (p < end) ? *p : 0u;
        return w;
    }
    __device__ __forceinline__ void init(const uint32_t *begin, const uint32_t *end_) {
        p = begin; end = end_; overrun = false;
        next = (p < end) ? *p : 0u;
        const uint64_t lo = word();
        const uint64_t hi = word();
        x = lo | (hi << 32);
    }
    __device__ __forceinline__ uint32_t peek() const { return (uint32_t)(x & 0xFFFFu); }
    __device__ __forceinline__ void advance(uint32_t start, uint32_t freq) {
        x = (uint64_t)freq * (x >> RANS_PRECISION) + (x & 0xFFFFu) - start;
        if (x < RANS_L) x = (x << 32) | word();
    }
    __device__ __forceinline__ uint32_t get_bits() {
        const uint32_t val = (uint32_t)(x & RANS_BYPASS_MAX);
        x >>= RANS_BYPASS_BITS;
        if (x < RANS_L) x = (x << 32) | word();
        return val;
    }
    __device__ __forceinline__ int32_t get_escape(int32_t max_value) {
        int32_t val = (int32_t)get_bits();
        int32_t nb = val;
        while (val == RANS_BYPASS_MAX && !overrun) { val = (int32_t)get_bits(); nb += val; }
        uint32_t raw = 0;
        for (int j = 0; j < nb && j < 8; ++j) raw |= get_bits() << (j * RANS_BYPASS_BITS);
        const int32_t value = (int32_t)(raw >> 1);
        return (raw & 1u) ? (-value - 1) : (value + max_value);
    }
};

// Slot of `cum` in a ragged row (last s in [0, len - 2] with row[s] <= cum), found by the whole warp: every round the
// lanes probe 32 evenly spaced pivots of the remaining range and a ballot counts how many lie at or below cum.
__device__ __forceinline__ int warp_find_slot(const uint16_t *row, int len, uint32_t cum, int lane) {
    int lo = 0, hi = len - 1;  // invariant: row[lo] <= cum < row[hi], row[len - 1] standing for 65536
    while (hi - lo > 1) {
        const int step = (hi - lo + 31) >> 5;
        const int idx = lo + (lane + 1) * step;
        const bool le = idx < hi && (uint32_t)row[idx] <= cum;
        const int k = __popc(__ballot_sync(0xffffffffu, le));  // pivots are increasing: the first k lanes say yes
        const int nlo = lo + k * step;
        hi = min(hi, nlo + step);
        lo = nlo;
    }
    return lo;
}

template <bool kSmemTable>
__global__ void __launch_bounds__(RANS_DECODE_THREADS)
rans_decode_kernel(const uint8_t *__restrict__ packed, const int64_t *__restrict__ offsets,
                   const int32_t *__restrict__ lengths, const int32_t *__restrict__ indexes, int64_t channel_period,
                   int64_t n_streams, int64_t n_sym, const RansTables tb, int32_t *__restrict__ symbols,
                   int32_t *__restrict__ status) {
    extern __shared__ __align__(16) uint8_t rans_smem[];
    const uint16_t *table = tb.ragged;
    const int32_t *row_start = tb.row_start, *sizes = tb.sizes, *offs = tb.offsets;
    if (kSmemTable) {
        // [ragged table | row_start | sizes | offsets]: every lookup of the symbol loop is a shared-memory read
        const uint4 *src = reinterpret_cast<const uint4 *>(tb.ragged);
        uint4 *dst = reinterpret_cast<uint4 *>(rans_smem);
        const int nvec = (tb.ragged_len + 7) >> 3;
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = src[i];
        int32_t *small = reinterpret_cast<int32_t *>(rans_smem + (size_t)nvec * 16);
        for (int i = threadIdx.x; i < tb.n_cdfs; i += blockDim.x) {
            small[i] = tb.row_start[i];
            small[tb.n_cdfs + i] = tb.sizes[i];
            small[2 * tb.n_cdfs + i] = tb.offsets[i];
        }
        __syncthreads();
        table = reinterpret_cast<const uint16_t *>(rans_smem);
        row_start = small; sizes = small + tb.n_cdfs; offs = small + 2 * tb.n_cdfs;
    }
    const int lane = threadIdx.x & 31;
    const int64_t stream = (int64_t)blockIdx.x * (RANS_DECODE_THREADS / 32) + (threadIdx.x >> 5);
    if (stream >= n_streams) return;  // whole warps leave together (after the block-wide staging barrier)
    const int64_t b0 = offsets[stream];
    const int32_t len_bytes = lengths[stream];
    int32_t st = 0;
    if (len_bytes < 8 || (len_bytes & 3)) { if (lane == 0) status[stream] = -1; return; }
    RansDecW dec;  // lane-uniform state
    dec.init(reinterpret_cast<const uint32_t *>(packed + b0), reinterpret_cast<const uint32_t *>(packed + b0 + len_bytes));
    int32_t *out_row = symbols + stream * n_sym;
    for (int64_t p0 = 0; p0 < n_sym && st == 0; p0 += 32) {
        // the CDF indexes do not depend on the decoder state: one coalesced load per 32 symbols
        const int32_t ci_l = (p0 + lane < n_sym) ? stream_index(indexes, channel_period, tb.n_cdfs, stream, n_sym, p0 + lane) : 0;
        const int count = (n_sym - p0 < 32) ? (int)(n_sym - p0) : 32;
        int32_t mine = 0;
        for (int k = 0; k < count; ++k) {
            const int32_t ci = __shfl_sync(0xffffffffu, ci_l, k);
            if (ci < 0 || ci >= tb.n_cdfs) { st = -3; break; }
            const int32_t len = sizes[ci];
            const int32_t max_value = len - 2;
            if (max_value < 0) { st = -4; break; }
            const uint16_t *row = table + row_start[ci];
            const int slot = warp_find_slot(row, len, dec.peek(), lane);
            if (slot < 0 || slot > max_value) { st = -4; break; }
            const uint32_t start = row[slot];
            dec.advance(start, ((uint32_t)row[slot + 1] - start) & 0xFFFFu);
            int32_t value = slot;
            if (slot == max_value) value = dec.get_escape(max_value);
            if (lane == k) mine = value + offs[ci];
            if (dec.overrun) { st = -2; break; }
        }
        if (st == 0 && lane < count) out_row[p0 + lane] = mine;  // 32 decoded symbols leave together
    }
    if (lane == 0) status[stream] = st;
}

static RansTables make_tables(const uint16_t *ragged, int64_t ragged_len, const int32_t *row_start,
                              const int32_t *sizes, const int32_t *offsets, int n_cdfs) {
    RansTables t;
    t.ragged = ragged; t.row_start = row_start; t.sizes = sizes; t.offsets = offsets;
    t.n_cdfs = n_cdfs; t.ragged_len = (int)ragged_len;
    return t;
}

}  // namespace mmnc

using namespace mmnc;

extern "C" int64_t mmnc_rans_slab_words(int64_t n_sym) {
    // every symbol adds at most 16 bits (probability >= 2^-16) plus, when escaped, at most 9 nibbles;
    // the state starts with 31 bits and the flush writes 2 words.  +16 words of headroom (checked in-kernel).
    return (n_sym * 52 + 31) / 32 + 4 + 16;
}

extern "C" int mmnc_rans_pack_tables(const int32_t *cdf, const int32_t *cdf_sizes, int n_cdfs, int cdf_stride,
                                     int32_t *row_start, uint16_t *ragged, int64_t ragged_capacity, void *stream) {
    MMNC_REQUIRE(n_cdfs > 0 && cdf_stride > 1, "rans_pack_tables: empty CDF table (run update() first)");
    MMNC_REQUIRE(cdf && cdf_sizes && row_start && ragged, "rans_pack_tables: null pointer");
    MMNC_REQUIRE(ragged_capacity > 0 && ragged_capacity < (1ll << 30), "rans_pack_tables: bad capacity");
    rans_pack_tables_kernel<<<1, 256, 0, as_stream(stream)>>>(cdf, cdf_sizes, n_cdfs, cdf_stride, row_start, ragged,
                                                             (int)ragged_capacity);
    return after_launch("rans_pack_tables_kernel");
}

extern "C" int mmnc_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int64_t channel_period,
                                      int64_t n_streams, int64_t n_sym, const uint16_t *ragged_cdf,
                                      int64_t ragged_len, const int32_t *row_start, const int32_t *cdf_sizes,
                                      const int32_t *offsets, int n_cdfs, void *staging, uint32_t *slabs,
                                      int64_t slab_words, int32_t *nbytes, void *stream) {
    MMNC_REQUIRE(n_streams >= 0 && n_sym >= 0, "rans_encode_batch: negative size");
    MMNC_REQUIRE(n_cdfs > 0 && ragged_len > 0, "rans_encode_batch: empty CDF table (run update() first)");
    MMNC_REQUIRE(indexes != nullptr || channel_period > 0, "rans_encode_batch: need indexes or channel_period");
    MMNC_REQUIRE(slab_words >= mmnc_rans_slab_words(n_sym), "rans_encode_batch: slab_words too small");
    if (n_streams == 0) return MMNC_OK;
    MMNC_REQUIRE(ragged_cdf && row_start && cdf_sizes && offsets && slabs && nbytes && (n_sym == 0 || (symbols && staging)),
                 "rans_encode_batch: null pointer");
    MMNC_REQUIRE((reinterpret_cast<uintptr_t>(ragged_cdf) & 15) == 0 && (reinterpret_cast<uintptr_t>(staging) & 15) == 0,
                 "rans_encode_batch: ragged table and staging must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    const RansTables tb = make_tables(ragged_cdf, ragged_len, row_start, cdf_sizes, offsets, n_cdfs);
    const int64_t total = n_streams * n_sym;
    // staging = [uint64 reciprocal x total | uint32 start:range x total]
    uint64_t *stage_rcp = static_cast<uint64_t *>(staging);
    uint32_t *stage_sr = reinterpret_cast<uint32_t *>(stage_rcp + total);
    if (total > 0) {
        // enough symbols per CTA that copying the table into shared memory pays for itself: each CTA maps at least
        // as many symbols as the table has entries
        const size_t table_bytes = ((size_t)ragged_len + 7) / 8 * 16;
        int64_t blocks = (total + RANS_MAP_THREADS * 8 - 1) / (RANS_MAP_THREADS * 8);
        const int64_t cap = (int64_t)sm_count() * 4;
        if (blocks > cap) blocks = cap;
        const bool smem_table = table_bytes <= RANS_SMEM_TABLE_MAX && total >= blocks * ragged_len;
        if (smem_table) {
            if (int rc = tmah::ensure_dynamic_smem(rans_map_kernel<true>, table_bytes)) return rc;
            rans_map_kernel<true><<<(unsigned)blocks, RANS_MAP_THREADS, table_bytes, s>>>(
                symbols, indexes, channel_period, n_streams, n_sym, tb, stage_sr, stage_rcp);
        } else {
            blocks = (total + RANS_MAP_THREADS - 1) / RANS_MAP_THREADS;
            if (blocks > cap * 4) blocks = cap * 4;
            rans_map_kernel<false><<<(unsigned)blocks, RANS_MAP_THREADS, 0, s>>>(
                symbols, indexes, channel_period, n_streams, n_sym, tb, stage_sr, stage_rcp);
        }
        if (int rc = after_launch("rans_map_kernel")) return rc;
    }
    constexpr int per_block = RANS_WARP_BLOCK / 32;
    rans_encode_kernel<<<(unsigned)((n_streams + per_block - 1) / per_block), RANS_WARP_BLOCK, 0, s>>>(
        stage_sr, stage_rcp, symbols, indexes, channel_period, n_streams, n_sym, tb, slabs, slab_words, nbytes);
    return after_launch("rans_encode_kernel");
}

extern "C" int mmnc_rans_compact(const uint32_t *slabs, int64_t slab_words, const int32_t *nbytes,
                                 int64_t n_streams, int64_t *meta, uint8_t *packed, int64_t packed_capacity,
                                 void *stream) {
    MMNC_REQUIRE(n_streams >= 0, "rans_compact: negative size");
    MMNC_REQUIRE(meta, "rans_compact: null meta");
    cudaStream_t s = as_stream(stream);
    if (n_streams == 0) {
        MMNC_CUDA(cudaMemsetAsync(meta, 0, sizeof(int64_t), s));
        return MMNC_OK;
    }
    MMNC_REQUIRE(slabs && nbytes && packed, "rans_compact: null pointer");
    rans_scan_kernel<<<1, 1024, 0, s>>>(nbytes, n_streams, meta);
    if (int rc = after_launch("rans_scan_kernel")) return rc;
    rans_gather_kernel<<<(unsigned)((n_streams + 3) / 4), 128, 0, s>>>(slabs, slab_words, nbytes, meta, n_streams, packed,
                                                                      packed_capacity);
    return after_launch("rans_gather_kernel");
}

extern "C" int mmnc_rans_decode_batch(const uint8_t *packed, const int64_t *offsets, const int32_t *lengths,
                                      const int32_t *indexes, int64_t channel_period, int64_t n_streams,
                                      int64_t n_sym, const uint16_t *ragged_cdf, int64_t ragged_len,
                                      const int32_t *row_start, const int32_t *cdf_sizes, const int32_t *cdf_offsets,
                                      int n_cdfs, int32_t *symbols, int32_t *status, void *stream) {
    MMNC_REQUIRE(n_streams >= 0 && n_sym >= 0, "rans_decode_batch: negative size");
    MMNC_REQUIRE(n_cdfs > 0 && ragged_len > 0, "rans_decode_batch: empty CDF table (run update() first)");
    MMNC_REQUIRE(indexes != nullptr || channel_period > 0, "rans_decode_batch: need indexes or channel_period");
    if (n_streams == 0) return MMNC_OK;
    MMNC_REQUIRE(packed && offsets && lengths && ragged_cdf && row_start && cdf_sizes && cdf_offsets && status &&
                     (n_sym == 0 || symbols),
                 "rans_decode_batch: null pointer");
    MMNC_REQUIRE((reinterpret_cast<uintptr_t>(ragged_cdf) & 15) == 0 && (reinterpret_cast<uintptr_t>(packed) & 3) == 0,
                 "rans_decode_batch: ragged table must be 16-byte aligned, packed 4-byte aligned");
    const RansTables tb = make_tables(ragged_cdf, ragged_len, row_start, cdf_sizes, cdf_offsets, n_cdfs);
    const size_t table_bytes = ((size_t)ragged_len + 7) / 8 * 16 + (size_t)n_cdfs * 12;
    constexpr int per_block = RANS_DECODE_THREADS / 32;  // one warp per stream
    const unsigned blocks = (unsigned)((n_streams + per_block - 1) / per_block);
    // the row search is a chain of dependent loads: from shared memory unless the table is too large for it or the
    // batch is so small that copying the table costs more than it saves
    const bool smem_table = table_bytes <= RANS_SMEM_TABLE_MAX && n_sym * 12 >= (int64_t)(table_bytes / 2048);
    if (smem_table) {
        if (int rc = tmah::ensure_dynamic_smem(rans_decode_kernel<true>, table_bytes)) return rc;
        rans_decode_kernel<true><<<blocks, RANS_DECODE_THREADS, table_bytes, as_stream(stream)>>>(
            packed, offsets, lengths, indexes, channel_period, n_streams, n_sym, tb, symbols, status);
    } else {
        rans_decode_kernel<false><<<blocks, RANS_DECODE_THREADS, 0, as_stream(stream)>>>(
            packed, offsets, lengths, indexes, channel_period, n_streams, n_sym, tb, symbols, status);
    }
    return after_launch("rans_decode_kernel");
}
