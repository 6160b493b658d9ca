// Host-side helpers shared by the TMA-fed kernels: the run-time resolved cuTensorMapEncodeTiled entry point, a
// per-thread cache of encoded tensor maps, and a once-per-(kernel, device) setter for the dynamic shared-memory
// attribute.  Encoding a rank-3 map costs a driver call (~1 us) and every GDN site needs two (forward) or two
// (backward) per launch; the caching allocator hands the same addresses back step after step, so in steady state a
// training step encodes nothing ("per-launch cuTensorMapEncodeTiled x2 + cudaFuncSetAttribute").
#pragma once
#include <cuda.h>  // CUtensorMap + enums only; the driver entry points are resolved at run time (no -lcuda)
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace mmnc {
namespace tmah {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline void *driver_entry(const char *name) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

inline EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_entry("cuTensorMapEncodeTiled"));
    return fn;
}

// cuTensorMapEncodeTiled is a driver-API call and wants a current context on the calling thread.  Runtime-API-only
// threads (torch's autograd worker on device 0 never calls cudaSetDevice) may not have one bound yet.
inline void bind_primary_context() {
    typedef CUresult (*GetCurrentFn)(CUcontext *);
    static GetCurrentFn get_current = reinterpret_cast<GetCurrentFn>(driver_entry("cuCtxGetCurrent"));
    static thread_local bool bound = false;
    if (bound) return;
    CUcontext cur = nullptr;
    if (get_current && get_current(&cur) == CUDA_SUCCESS && cur != nullptr) { bound = true; return; }
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);  // CUDA 12: initialises and binds the primary context
    cudaGetLastError();
    bound = true;
}

struct MapKey {
    uint64_t ptr, d0, d1, d2, s0, s1;
    uint32_t b0, b1, b2, swizzle;
    bool operator==(const MapKey &o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const {
        uint64_t h = 0x9e3779b97f4a7c15ull;
        const uint64_t w[8] = {k.ptr, k.d0, k.d1, k.d2, k.s0, k.s1, ((uint64_t)k.b0 << 32) | k.b1,
                               ((uint64_t)k.b2 << 32) | k.swizzle};
        for (uint64_t v : w) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
    }
};

// fp32 rank-3 map [d0 (contiguous), d1, d2] with byte strides s0 (of d1) and s1 (of d2) and box b0 x b1 x b2.
// A tensor map is a pure function of these arguments, so a cached copy is always valid for the same key.
inline int tensor_map_3d(CUtensorMap *out, const float *p, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s0,
                         uint64_t s1, uint32_t b0, uint32_t b1, uint32_t b2, CUtensorMapSwizzle swizzle,
                         const char *who) {
    static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = reinterpret_cast<uint64_t>(p);
    key.d0 = d0; key.d1 = d1; key.d2 = d2; key.s0 = s0; key.s1 = s1;
    key.b0 = b0; key.b1 = b1; key.b2 = b2; key.swizzle = (uint32_t)swizzle;
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return MMNC_OK; }
    EncodeTiledFn fn = encode_tiled();
    if (!fn) { set_error("%s: cuTensorMapEncodeTiled is not available", who); return MMNC_ERR_UNSUPPORTED; }
    bind_primary_context();
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {s0, s1};
    const cuuint32_t box[3] = {b0, b1, b2};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(p), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
        return MMNC_ERR_CUDA;
    }
    if (cache.size() >= 4096) cache.clear();  // bounded: a long-running process with drifting addresses starts over
    cache.emplace(key, *out);
    return MMNC_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize sticks to the function (per device, process-wide) and is an upper
// bound: raise it when a launch needs more than any earlier one did, never per launch.
inline int ensure_dynamic_smem_impl(const void *kernel, size_t bytes, bool max_carveout) {
    static std::mutex mu;
    static std::unordered_map<uint64_t, size_t> granted;  // (function, device) -> bytes already allowed
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(kernel) * 64u + (uint64_t)(dev & 63);
    std::lock_guard<std::mutex> lock(mu);
    auto it = granted.find(key);
    if (it != granted.end() && it->second >= bytes) return MMNC_OK;
    MMNC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (max_carveout)
        MMNC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    granted[key] = bytes;
    return MMNC_OK;
}
template <typename K>
inline int ensure_dynamic_smem(K kernel, size_t bytes, bool max_carveout = false) {
    return ensure_dynamic_smem_impl(reinterpret_cast<const void *>(kernel), bytes, max_carveout);
}

}  // namespace tmah
}  // namespace mmnc
