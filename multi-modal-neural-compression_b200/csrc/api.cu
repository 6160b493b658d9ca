// Library state: version, error slot, launch counter, device query; GDN dispatch between the SIMT and the
// tensor-core kernels.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "gdn_params.cuh"

namespace mmnc {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        else cudaGetLastError();
        cached_dev = dev;
    }
    return cached;
}

// gdn_simt.cu
int gdn_simt_forward(const float *, int64_t, int64_t, int64_t, const GdnParams &, int, float *, cudaStream_t);
size_t gdn_simt_backward_workspace(int64_t, int64_t, int64_t);
int gdn_simt_backward(const float *, const float *, int64_t, int64_t, int64_t, const GdnParams &, int, float *, float *,
                      float *, void *, size_t, cudaStream_t);
// gdn_small.cu
bool gdn_small_supported(int64_t C);
size_t gdn_small_backward_workspace(int64_t B, int64_t C, int64_t HW);
int gdn_small_forward(const float *, int64_t, int64_t, int64_t, const GdnParams &, int, float *, cudaStream_t, int nhwc = 0);
int gdn_small_backward(const float *, const float *, int64_t, int64_t, int64_t, const GdnParams &, int, float *, float *,
                       float *, void *, size_t, cudaStream_t, int nhwc = 0);
// gdn_tc.cu
bool gdn_tc_supported(int64_t B, int64_t C, int64_t HW, int precision);
int gdn_tc_forward(const float *, int64_t, int64_t, int64_t, const GdnParams &, int, int, float *, cudaStream_t, int nhwc = 0);
// gdn_tc_fwd2.cu
bool gdn_tc_forward2_supported(const float *x, const float *y, int64_t B, int64_t C, int64_t HW);
// gdn_tc_bwd2.cu
bool gdn_tc_backward2_supported(const float *x, const float *g, int64_t B, int64_t C, int64_t HW);
bool gdn_tc_backward2_streams(int64_t C);
bool gdn_tc_backward2_prefetches(int64_t B, int64_t C, int64_t HW);
int gdn_tc_backward2(const float *, const float *, int64_t, int64_t, int64_t, const GdnParams &, int, float *, float *,
                     float *, void *, size_t, cudaStream_t);
// gdn_tc_wide.cu
size_t gdn_tc_wide_forward_workspace(int64_t B, int64_t C, int64_t HW);
bool gdn_tc_wide_forward_supported(int64_t B, int64_t C, int64_t HW, const void *workspace, size_t workspace_bytes);
int gdn_tc_wide_forward(const float *, int64_t, int64_t, int64_t, const GdnParams &, int, float *, void *, size_t, cudaStream_t);
bool gdn_tc_wide_backward_supported(const float *x, const float *g, int64_t B, int64_t C, int64_t HW);
size_t gdn_tc_wide_backward_workspace(int64_t B, int64_t C, int64_t HW);
int gdn_tc_wide_backward(const float *, const float *, int64_t, int64_t, int64_t, const GdnParams &, int, float *, float *,
                         float *, void *, size_t, cudaStream_t);
// gdn_tc_bwd.cu
bool gdn_tc_backward_supported(int64_t B, int64_t C, int64_t HW);
size_t gdn_tc_backward_workspace(int64_t B, int64_t C, int64_t HW);
int gdn_tc_backward(const float *, const float *, int64_t, int64_t, int64_t, const GdnParams &, int, float *, float *,
                    float *, void *, size_t, cudaStream_t, int nhwc = 0);

}  // namespace mmnc

using namespace mmnc;

extern "C" int mmnc_version(void) { return 100; }
extern "C" const char *mmnc_last_error(void) { return g_error; }
extern "C" uint64_t mmnc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int mmnc_device_sm_count(void) { return sm_count(); }

// Channels-last (NHWC) tensors are taken natively by the streaming kernels (C <= 4) and by the tensor-core kernels with
// per-thread row access (forward C <= 128, backward C <= 111, single-pass TF32); everything else is NCHW only and the
// caller converts (mmnc_gdn_nhwc_supported tells which).
static bool gdn_nhwc_ok(int64_t B, int64_t C, int64_t HW, int precision, int backward) {
    if (gdn_small_supported(C)) return true;
    if (!(precision == MMNC_GDN_AUTO || precision == MMNC_GDN_TF32)) return false;
    return backward ? gdn_tc_backward_supported(B, C, HW) : gdn_tc_supported(B, C, HW, MMNC_GDN_TF32);
}

extern "C" int mmnc_gdn_nhwc_supported(int64_t B, int64_t C, int64_t HW, int precision, int backward) {
    return gdn_nhwc_ok(B, C, HW, precision, backward) ? 1 : 0;
}

static int gdn_forward_impl(const float *x, int64_t B, int64_t C, int64_t HW, const GdnParams &prm, int inverse,
                            int precision, float *y, void *stream, int nhwc = 0, void *workspace = nullptr,
                            size_t workspace_bytes = 0) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && HW >= 0, "gdn_forward: negative dimension");
    MMNC_REQUIRE(precision >= 0 && precision <= 3, "gdn_forward: bad precision %d", precision);
    if (B * C * HW == 0) return MMNC_OK;
    MMNC_REQUIRE(x && prm.beta && prm.gamma && y, "gdn_forward: null pointer");
    MMNC_REQUIRE(C <= 8192, "gdn_forward: C = %lld too large", (long long)C);
    // `precision` names the arithmetic the caller accepts.  Tensor cores are used when the shape suits the tcgen05
    // kernel in that arithmetic; everything else runs on the fp32 SIMT kernel, which is at least as accurate.
    if (nhwc) {
        MMNC_REQUIRE(gdn_nhwc_ok(B, C, HW, precision, 0), "gdn_forward: no channels-last kernel for C = %lld at this precision",
                     (long long)C);
        if (gdn_small_supported(C)) return gdn_small_forward(x, B, C, HW, prm, inverse, y, as_stream(stream), 1);
        return gdn_tc_forward(x, B, C, HW, prm, inverse, MMNC_GDN_TF32, y, as_stream(stream), 1);
    }
    if (gdn_small_supported(C)) return gdn_small_forward(x, B, C, HW, prm, inverse, y, as_stream(stream));  // fp32
    const int want = (precision == MMNC_GDN_AUTO) ? MMNC_GDN_TF32 : precision;
    if (want != MMNC_GDN_FP32 && gdn_tc_supported(B, C, HW, want))
        return gdn_tc_forward(x, B, C, HW, prm, inverse, want, y, as_stream(stream));
    if (want == MMNC_GDN_TF32 && gdn_tc_wide_forward_supported(B, C, HW, workspace, workspace_bytes))  // 129 <= C <= 256
        return gdn_tc_wide_forward(x, B, C, HW, prm, inverse, y, workspace, workspace_bytes, as_stream(stream));
    return gdn_simt_forward(x, B, C, HW, prm, inverse, y, as_stream(stream));
}

static GdnParams gdn_raw(const float *beta, const float *gamma, float bb, float gb, float ped) {
    GdnParams p;
    p.beta = beta; p.gamma = gamma; p.beta_bound = bb; p.gamma_bound = gb; p.pedestal = ped; p.raw = 1;
    return p;
}

extern "C" int mmnc_gdn_forward(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta,
                                const float *gamma, int inverse, int precision, float *y, void *stream) {
    return gdn_forward_impl(x, B, C, HW, gdn_effective(beta, gamma), inverse, precision, y, stream);
}

extern "C" int mmnc_gdn_forward_raw(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta_raw,
                                    const float *gamma_raw, float beta_bound, float gamma_bound, float pedestal,
                                    int inverse, int precision, float *y, void *stream) {
    return gdn_forward_impl(x, B, C, HW, gdn_raw(beta_raw, gamma_raw, beta_bound, gamma_bound, pedestal), inverse,
                            precision, y, stream);
}

extern "C" size_t mmnc_gdn_forward_workspace_bytes(int64_t B, int64_t C, int64_t HW, int precision) {
    if (B <= 0 || C <= 0 || HW <= 0) return 0;
    if (!(precision == MMNC_GDN_AUTO || precision == MMNC_GDN_TF32)) return 0;
    return gdn_tc_wide_forward_workspace(B, C, HW);
}

extern "C" int mmnc_gdn_forward_ws(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta,
                                   const float *gamma, int inverse, int precision, float *y, void *workspace,
                                   size_t workspace_bytes, void *stream) {
    return gdn_forward_impl(x, B, C, HW, gdn_effective(beta, gamma), inverse, precision, y, stream, 0, workspace,
                            workspace_bytes);
}

extern "C" int mmnc_gdn_forward_raw_ws(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta_raw,
                                       const float *gamma_raw, float beta_bound, float gamma_bound, float pedestal,
                                       int inverse, int precision, float *y, void *workspace, size_t workspace_bytes,
                                       void *stream) {
    return gdn_forward_impl(x, B, C, HW, gdn_raw(beta_raw, gamma_raw, beta_bound, gamma_bound, pedestal), inverse,
                            precision, y, stream, 0, workspace, workspace_bytes);
}

extern "C" int mmnc_gdn_forward_raw_nhwc(const float *x, int64_t B, int64_t C, int64_t HW, const float *beta_raw,
                                         const float *gamma_raw, float beta_bound, float gamma_bound, float pedestal,
                                         int inverse, int precision, float *y, void *stream) {
    return gdn_forward_impl(x, B, C, HW, gdn_raw(beta_raw, gamma_raw, beta_bound, gamma_bound, pedestal), inverse,
                            precision, y, stream, 1);
}

extern "C" size_t mmnc_gdn_backward_workspace_bytes(int64_t B, int64_t C, int64_t HW, int precision) {
    (void)precision;
    if (B <= 0 || C <= 0 || HW <= 0) return 256;
    if (gdn_small_supported(C)) return gdn_small_backward_workspace(B, C, HW);
    // the caller may not know yet whether its tensors will be aligned for the TMA-fed kernel: cover every candidate
    const size_t a = gdn_simt_backward_workspace(B, C, HW), b = gdn_tc_backward_workspace(B, C, HW);
    const bool tf32 = (precision == MMNC_GDN_AUTO || precision == MMNC_GDN_TF32);
    const bool tc = tf32 && gdn_tc_backward_supported(B, C, HW);
    const size_t w = tf32 ? gdn_tc_wide_backward_workspace(B, C, HW) : 0;  // 0 unless 129 <= C <= 256
    const size_t m = tc ? b : (a > b ? a : b);
    return m > w ? m : w;
}

extern "C" int mmnc_gdn_backward_variant(const float *x, const float *g, int64_t B, int64_t C, int64_t HW,
                                         int precision) {
    if (gdn_small_supported(C)) return 0;
    if (precision == MMNC_GDN_AUTO || precision == MMNC_GDN_TF32) {
        if (gdn_tc_backward2_supported(x, g, B, C, HW))
            return gdn_tc_backward2_streams(C) ? 4 : (gdn_tc_backward2_prefetches(B, C, HW) ? 5 : 3);
        if (gdn_tc_backward_supported(B, C, HW)) return 2;
        if (gdn_tc_wide_backward_supported(x, g, B, C, HW)) return 6;
    }
    return 1;
}

extern "C" int mmnc_gdn_forward_variant(const float *x, const float *y, int64_t B, int64_t C, int64_t HW,
                                        int precision) {
    if (gdn_small_supported(C)) return 0;
    const int want = (precision == MMNC_GDN_AUTO) ? MMNC_GDN_TF32 : precision;
    if (want == MMNC_GDN_TF32 && !gdn_tc_supported(B, C, HW, want) && gdn_tc_wide_forward_workspace(B, C, HW) > 0) return 6;
    if (want == MMNC_GDN_FP32 || !gdn_tc_supported(B, C, HW, want)) return 1;
    return (want == MMNC_GDN_TF32 && gdn_tc_forward2_supported(x, y, B, C, HW)) ? 3 : 2;
}

static int gdn_backward_impl(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const GdnParams &prm,
                             int inverse, int precision, float *dx, float *dbeta, float *dgamma, void *workspace,
                             size_t workspace_bytes, void *stream, int nhwc = 0) {
    MMNC_REQUIRE(B >= 0 && C >= 0 && HW >= 0, "gdn_backward: negative dimension");
    MMNC_REQUIRE(precision >= 0 && precision <= 3, "gdn_backward: bad precision %d", precision);
    MMNC_REQUIRE(dbeta && dgamma, "gdn_backward: null pointer");
    if (B * C * HW == 0) {
        if (C > 0) {
            MMNC_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * (size_t)C, as_stream(stream)));
            MMNC_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * (size_t)(C * C), as_stream(stream)));
        }
        return MMNC_OK;
    }
    MMNC_REQUIRE(x && g && prm.beta && prm.gamma && dx && workspace, "gdn_backward: null pointer");
    MMNC_REQUIRE(C <= 8192, "gdn_backward: C = %lld too large", (long long)C);
    // single-pass TF32 on the tensor cores when the caller accepts it (auto / tf32) and the shape suits the kernel;
    // fp32 and 3xtf32 requests, small problems and unusual channel counts run the exact fp32 SIMT kernels
    if (nhwc) {
        MMNC_REQUIRE(gdn_nhwc_ok(B, C, HW, precision, 1), "gdn_backward: no channels-last kernel for C = %lld at this precision",
                     (long long)C);
        if (gdn_small_supported(C))
            return gdn_small_backward(x, g, B, C, HW, prm, inverse, dx, dbeta, dgamma, workspace, workspace_bytes,
                                      as_stream(stream), 1);
        return gdn_tc_backward(x, g, B, C, HW, prm, inverse, dx, dbeta, dgamma, workspace, workspace_bytes,
                               as_stream(stream), 1);
    }
    if (gdn_small_supported(C))
        return gdn_small_backward(x, g, B, C, HW, prm, inverse, dx, dbeta, dgamma, workspace, workspace_bytes,
                                  as_stream(stream));
    if (precision == MMNC_GDN_AUTO || precision == MMNC_GDN_TF32) {
        if (gdn_tc_backward2_supported(x, g, B, C, HW))  // TMA-fed: C <= 128, H*W a multiple of 128, aligned tensors
            return gdn_tc_backward2(x, g, B, C, HW, prm, inverse, dx, dbeta, dgamma, workspace, workspace_bytes,
                                    as_stream(stream));
        if (gdn_tc_backward_supported(B, C, HW))
            return gdn_tc_backward(x, g, B, C, HW, prm, inverse, dx, dbeta, dgamma, workspace, workspace_bytes,
                                   as_stream(stream));
        if (gdn_tc_wide_backward_supported(x, g, B, C, HW) && workspace_bytes >= gdn_tc_wide_backward_workspace(B, C, HW))
            return gdn_tc_wide_backward(x, g, B, C, HW, prm, inverse, dx, dbeta, dgamma, workspace, workspace_bytes,
                                        as_stream(stream));
    }
    return gdn_simt_backward(x, g, B, C, HW, prm, inverse, dx, dbeta, dgamma, workspace, workspace_bytes,
                             as_stream(stream));
}

extern "C" int mmnc_gdn_backward(const float *x, const float *g, int64_t B, int64_t C, int64_t HW, const float *beta,
                                 const float *gamma, int inverse, int precision, float *dx, float *dbeta,
                                 float *dgamma, void *workspace, size_t workspace_bytes, void *stream) {
    return gdn_backward_impl(x, g, B, C, HW, gdn_effective(beta, gamma), inverse, precision, dx, dbeta, dgamma,
                             workspace, workspace_bytes, stream);
}

extern "C" int mmnc_gdn_backward_raw(const float *x, const float *g, int64_t B, int64_t C, int64_t HW,
                                     const float *beta_raw, const float *gamma_raw, float beta_bound,
                                     float gamma_bound, float pedestal, int inverse, int precision, float *dx,
                                     float *dbeta_raw, float *dgamma_raw, void *workspace, size_t workspace_bytes,
                                     void *stream) {
    return gdn_backward_impl(x, g, B, C, HW, gdn_raw(beta_raw, gamma_raw, beta_bound, gamma_bound, pedestal), inverse,
                             precision, dx, dbeta_raw, dgamma_raw, workspace, workspace_bytes, stream);
}

extern "C" int mmnc_gdn_backward_raw_nhwc(const float *x, const float *g, int64_t B, int64_t C, int64_t HW,
                                          const float *beta_raw, const float *gamma_raw, float beta_bound,
                                          float gamma_bound, float pedestal, int inverse, int precision, float *dx,
                                          float *dbeta_raw, float *dgamma_raw, void *workspace, size_t workspace_bytes,
                                          void *stream) {
    return gdn_backward_impl(x, g, B, C, HW, gdn_raw(beta_raw, gamma_raw, beta_bound, gamma_bound, pedestal), inverse,
                             precision, dx, dbeta_raw, dgamma_raw, workspace, workspace_bytes, stream, 1);
}
