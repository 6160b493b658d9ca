// (f2) MS-SSIM of the reference's `average_metrics` (/root/reference/src/models/multi_task_compressor.py:359-384 calls
// pytorch_msssim.ms_ssim on every task, every step): ONE scale of the SSIM computation per launch.
//
// pytorch_msssim's _ssim filters five maps (x, y, x^2, y^2, x y) with a separable 11-tap Gaussian ('valid'), forms the
// contrast-structure and SSIM maps and averages them; as stock torch ops that is ten depth-wise convolutions and ~20
// element-wise passes per scale (21 ms for a (256, 3, 256, 256) pair on B200).  Here a block owns a 32 x 32 tile of the
// OUTPUT: it stages the 42 x 42 input tiles of x and y in shared memory (scaled on load: the reference multiplies both
// images by 255 first), runs the horizontal pass of all five maps into shared memory, the vertical pass in registers,
// and reduces the two maps to one partial sum pair per tile; a second tiny kernel adds a plane's tiles in a fixed order.
// The same block also writes the 2 x 2 average-pooled planes that are the next scale's input, from the tile it already
// holds.  HBM traffic: 8 B / element read (+ 2 B / element for the pooled planes) per scale.
#include "common.cuh"

namespace mmnc {

namespace ssim {
constexpr int T = 32;           // output tile edge
constexpr int K = 11;           // window
constexpr int IN = T + K - 1;   // 42: input tile edge
constexpr int THREADS = 256;
struct Window { float g[K]; };
}  // namespace ssim

__global__ void __launch_bounds__(ssim::THREADS)
ssim_tile_kernel(const float *__restrict__ x, const float *__restrict__ y, int H, int W, float scale, float c1, float c2,
                 const ssim::Window win, float2 *__restrict__ partial, float *__restrict__ px, float *__restrict__ py) {
    using namespace ssim;
    __shared__ float sx[IN][IN + 1], sy[IN][IN + 1];
    __shared__ float hz[5][IN][T + 1];
    __shared__ float red[2][THREADS / 32];
    const int tx = blockIdx.x, ty = blockIdx.y;
    const int64_t plane = blockIdx.z;
    const float *xp = x + plane * (int64_t)H * W, *yp = y + plane * (int64_t)H * W;
    const int x0 = tx * T, y0 = ty * T;
    for (int idx = threadIdx.x; idx < IN * IN; idx += THREADS) {
        const int r = idx / IN, c = idx - r * IN;
        const int gy = y0 + r, gx = x0 + c;
        const bool in = gy < H && gx < W;
        sx[r][c] = in ? xp[(int64_t)gy * W + gx] * scale : 0.f;
        sy[r][c] = in ? yp[(int64_t)gy * W + gx] * scale : 0.f;
    }
    __syncthreads();
    // ---- next scale's input: 2 x 2 averages of the tile's own 32 x 32 input pixels; the last tile of a row / column also
    //      covers what is left of the image (at most 10 more pixels: its staged tile is 42 wide)
    if (px != nullptr) {
        const int rows = (ty == (int)gridDim.y - 1) ? H - y0 : T, cols = (tx == (int)gridDim.x - 1) ? W - x0 : T;
        const int pr = rows >> 1, pc = cols >> 1, W2 = W >> 1, H2 = H >> 1;
        float *pxp = px + plane * (int64_t)H2 * W2, *pyp = py + plane * (int64_t)H2 * W2;
        for (int idx = threadIdx.x; idx < pr * pc; idx += THREADS) {
            const int r = idx / pc, c = idx - r * pc;
            const int64_t o = (int64_t)((y0 >> 1) + r) * W2 + (x0 >> 1) + c;
            pxp[o] = 0.25f * ((sx[2 * r][2 * c] + sx[2 * r][2 * c + 1]) + (sx[2 * r + 1][2 * c] + sx[2 * r + 1][2 * c + 1]));
            pyp[o] = 0.25f * ((sy[2 * r][2 * c] + sy[2 * r][2 * c + 1]) + (sy[2 * r + 1][2 * c] + sy[2 * r + 1][2 * c + 1]));
        }
    }
    // ---- horizontal pass of the five maps.  Register tiling: a work item is FOUR adjacent outputs of one row - 14 loads of x
    //      and of y (and their three products, formed once) feed 4 x 11 taps; the first version loaded 2 values per tap
    //      and was bound by shared-memory instructions (90 per output pixel; now ~33).
    for (int item = threadIdx.x; item < IN * (T / 4); item += THREADS) {
        const int r = item / (T / 4), c0 = (item - r * (T / 4)) * 4;
        float xv[K + 3], yv[K + 3];
#pragma unroll
        for (int j = 0; j < K + 3; ++j) { xv[j] = sx[r][c0 + j]; yv[j] = sy[r][c0 + j]; }
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, aa[4] = {0.f, 0.f, 0.f, 0.f},
              bb[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < K + 3; ++j) {
            const float xx = xv[j] * xv[j], yy = yv[j] * yv[j], xy = xv[j] * yv[j];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int k = j - o;  // tap index of value j for output o
                if (k >= 0 && k < K) {
                    const float w = win.g[k];
                    a[o] = fmaf(w, xv[j], a[o]);
                    b[o] = fmaf(w, yv[j], b[o]);
                    aa[o] = fmaf(w, xx, aa[o]);
                    bb[o] = fmaf(w, yy, bb[o]);
                    ab[o] = fmaf(w, xy, ab[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            hz[0][r][c0 + o] = a[o]; hz[1][r][c0 + o] = b[o]; hz[2][r][c0 + o] = aa[o]; hz[3][r][c0 + o] = bb[o];
            hz[4][r][c0 + o] = ab[o];
        }
    }
    __syncthreads();
    // ---- vertical pass + the two maps: one thread = one column x four adjacent rows (14 loads per map for 4 x 11 taps);
    //      outputs beyond the 'valid' domain (H - 10) x (W - 10) do not count
    float s_ssim = 0.f, s_cs = 0.f;
    const int OH = H - (K - 1), OW = W - (K - 1);
    static_assert(THREADS == T * (T / 4), "one vertical work item per thread");
    {
        const int c = threadIdx.x % T, r0 = (threadIdx.x / T) * 4;
        float acc[5][4];
#pragma unroll
        for (int m = 0; m < 5; ++m)
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[m][o] = 0.f;
#pragma unroll
        for (int j = 0; j < K + 3; ++j) {
            float v[5];
#pragma unroll
            for (int m = 0; m < 5; ++m) v[m] = hz[m][r0 + j][c];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int k = j - o;
                if (k >= 0 && k < K) {
                    const float w = win.g[k];
#pragma unroll
                    for (int m = 0; m < 5; ++m) acc[m][o] = fmaf(w, v[m], acc[m][o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            if (y0 + r0 + o >= OH || x0 + c >= OW) continue;
            const float mu1 = acc[0][o], mu2 = acc[1][o];
            const float m11 = mu1 * mu1, m22 = mu2 * mu2, m12 = mu1 * mu2;
            const float s1 = acc[2][o] - m11, s2 = acc[3][o] - m22, s12 = acc[4][o] - m12;
            const float cs = (2.f * s12 + c2) / (s1 + s2 + c2);
            s_cs += cs;
            s_ssim += ((2.f * m12 + c1) / (m11 + m22 + c1)) * cs;
        }
    }
    s_ssim = warp_sum(s_ssim);
    s_cs = warp_sum(s_cs);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = s_ssim; red[1][warp] = s_cs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < THREADS / 32; ++w) { a += red[0][w]; b += red[1][w]; }
        partial[(plane * gridDim.y + ty) * gridDim.x + tx] = make_float2(a, b);
    }
}

// one warp per plane: the tiles' partial sums in a fixed order -> the two spatial means
__global__ void __launch_bounds__(32)
ssim_finish_kernel(const float2 *__restrict__ partial, int tiles, float inv_count, float *__restrict__ ssim_mean,
                   float *__restrict__ cs_mean) {
    const int64_t plane = blockIdx.x;
    float a = 0.f, b = 0.f;
    for (int t = threadIdx.x; t < tiles; t += 32) {
        const float2 v = partial[plane * tiles + t];
        a += v.x; b += v.y;
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) { ssim_mean[plane] = a * inv_count; cs_mean[plane] = b * inv_count; }
}

}  // namespace mmnc

using namespace mmnc;

extern "C" size_t mmnc_ssim_workspace_floats(int64_t planes, int H, int W) {
    if (planes <= 0 || H < ssim::K || W < ssim::K) return 2;
    const int64_t tiles = (int64_t)((H - ssim::K + 1 + ssim::T - 1) / ssim::T) * ((W - ssim::K + 1 + ssim::T - 1) / ssim::T);
    return (size_t)(2 * planes * tiles);
}

extern "C" int mmnc_ssim_scale(const float *x, const float *y, int64_t planes, int H, int W, float scale, float c1, float c2,
                               float sigma, float *workspace, float *ssim_mean, float *cs_mean, float *px, float *py,
                               void *stream) {
    MMNC_REQUIRE(planes >= 0 && H >= ssim::K && W >= ssim::K, "ssim_scale: the images must be at least 11 x 11");
    MMNC_REQUIRE(planes < 65536, "ssim_scale: at most 65535 planes per call");
    if (planes == 0) return MMNC_OK;
    MMNC_REQUIRE(x && y && workspace && ssim_mean && cs_mean, "ssim_scale: null pointer");
    MMNC_REQUIRE((px == nullptr) == (py == nullptr), "ssim_scale: both pooled outputs or neither");
    MMNC_REQUIRE(px == nullptr || (H % 2 == 0 && W % 2 == 0), "ssim_scale: fused pooling needs even H and W");
    MMNC_REQUIRE(sigma > 0.f, "ssim_scale: sigma must be positive");
    ssim::Window win;
    double sum = 0.0, g[ssim::K];
    for (int i = 0; i < ssim::K; ++i) {
        const double d = i - ssim::K / 2;
        // the window in fp32 like pytorch_msssim's _fspecial_gauss_1d: exp in float, normalised in float
        g[i] = (double)expf((float)(-(d * d) / (2.0 * (double)sigma * (double)sigma)));
        sum += g[i];
    }
    for (int i = 0; i < ssim::K; ++i) win.g[i] = (float)(g[i] / (double)(float)sum);
    const int OH = H - ssim::K + 1, OW = W - ssim::K + 1;
    const dim3 grid((unsigned)((OW + ssim::T - 1) / ssim::T), (unsigned)((OH + ssim::T - 1) / ssim::T), (unsigned)planes);
    float2 *partial = reinterpret_cast<float2 *>(workspace);
    MMNC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "ssim_scale: workspace must be 8-byte aligned");
    ssim_tile_kernel<<<grid, ssim::THREADS, 0, as_stream(stream)>>>(x, y, H, W, scale, c1, c2, win, partial, px, py);
    if (int rc = after_launch("ssim_tile_kernel")) return rc;
    ssim_finish_kernel<<<(unsigned)planes, 32, 0, as_stream(stream)>>>(partial, (int)(grid.x * grid.y),
                                                                       1.f / ((float)OH * (float)OW), ssim_mean, cs_mean);
    return after_launch("ssim_finish_kernel");
}

// All five scales in one call (H and W multiples of 16, so that every pooling is exact 2 x 2): ten launches, no host work
// in between.  means: [5][2][planes] = per scale the SSIM means, then the contrast means.
// workspace: mmnc_ms_ssim_workspace_floats floats = tile partials of scale 0 + the pooled planes of scales 1 .. 4.
extern "C" size_t mmnc_ms_ssim_workspace_floats(int64_t planes, int H, int W) {
    if (planes <= 0 || H <= 0 || W <= 0) return 2;
    size_t pooled = 0;
    for (int i = 1; i < 5; ++i) pooled += 2 * (size_t)planes * (size_t)(H >> i) * (size_t)(W >> i);
    return mmnc_ssim_workspace_floats(planes, H, W) + pooled + 8;
}

extern "C" int mmnc_ms_ssim(const float *x, const float *y, int64_t planes, int H, int W, float scale, float c1, float c2,
                            float sigma, float *workspace, float *means, void *stream) {
    MMNC_REQUIRE(planes >= 0 && H > 0 && W > 0, "ms_ssim: bad dimensions");
    MMNC_REQUIRE(H % 16 == 0 && W % 16 == 0, "ms_ssim: H and W must be multiples of 16 (use mmnc_ssim_scale otherwise)");
    MMNC_REQUIRE((H >> 4) >= ssim::K && (W >> 4) >= ssim::K, "ms_ssim: the image is too small for five scales");
    if (planes == 0) return MMNC_OK;
    MMNC_REQUIRE(x && y && workspace && means, "ms_ssim: null pointer");
    size_t part = mmnc_ssim_workspace_floats(planes, H, W);
    part = (part + 1) / 2 * 2;  // keeps the pooled planes 8-byte aligned like the partials
    float *pool = workspace + part;
    const float *cx = x, *cy = y;
    for (int i = 0; i < 5; ++i) {
        const int h = H >> i, w = W >> i;
        float *nx = nullptr, *ny = nullptr;
        if (i < 4) {
            nx = pool;
            ny = pool + (size_t)planes * (size_t)(h >> 1) * (size_t)(w >> 1);
            pool = ny + (size_t)planes * (size_t)(h >> 1) * (size_t)(w >> 1);
        }
        if (int rc = mmnc_ssim_scale(cx, cy, planes, h, w, i == 0 ? scale : 1.f, c1, c2, sigma, workspace,
                                     means + (size_t)(2 * i) * planes, means + (size_t)(2 * i + 1) * planes, nx, ny, stream))
            return rc;
        cx = nx;
        cy = ny;
    }
    return MMNC_OK;
}

