// SSIM / MS-SSIM statistics and their gradient on the device.
//
// The reference's loss (hnerv_utils.py:338-395, 'Fusion*' types) calls pytorch_msssim==0.2.1 (requirements.txt), which
// is not vendored in the reference tree; its published algorithm is restated in oracle/msssim_oracle.py and these
// kernels are checked against that restatement (parity unpinned: no golden vectors exist for it).
//
// One pyramid level, per (batch, channel) plane, 11-tap separable Gaussian window (sigma 1.5), 'valid' extent:
//     mu_x = G*x, mu_y = G*y, s_xx = G*x^2 - mu_x^2, s_yy = G*y^2 - mu_y^2, s_xy = G*xy - mu_x mu_y
//     cs   = (2 s_xy + C2) / (s_xx + s_yy + C2)        ssim = (2 mu_x mu_y + C1) / (mu_x^2 + mu_y^2 + C1) * cs
// forward : sums of ssim and cs over the valid region (f64 atomics)            -> ssim_stats_kernel
// backward: F = g_s * ssim + g_c * cs per pixel (g_* = upstream scalar / N_valid per plane);
//     P_xx = dF/d(G*x^2), P_xy = dF/d(G*xy), P_mu = dF/d(mu_x) incl. the mu_x inside s_xx, s_xy  -> ssim_partials_kernel
//     dF/dx = G^T*P_mu + 2x G^T*P_xx + y G^T*P_xy   (G^T* = 'full' correlation with the same window) -> ssim_grad_kernel
// Every kernel is a shared-memory tiled separable filter: HBM traffic is the maps themselves.
#include "common.cuh"

namespace bnerv {

constexpr int SS_WIN = 11;
constexpr int SS_R = 5;
constexpr int SS_TW = 32;            // output tile width
constexpr int SS_TH = 16;            // output tile height
constexpr int SS_IW = SS_TW + 2 * SS_R;
constexpr int SS_IH = SS_TH + 2 * SS_R;
constexpr int SS_THREADS = SS_TW * SS_TH / 2;     // 256 threads, two output rows each

__constant__ float c_gauss[SS_WIN];

struct SsimPix { float mu_x, mu_y, sxx, syy, sxy; };

// Loads the (TH+10) x (TW+10) input window of x and y whose top-left INPUT pixel is (h0, w0) (zero outside the
// plane), filters the five moment maps horizontally into shared memory, and leaves the vertical pass to the caller.
__device__ __forceinline__ void ssim_tile_moments(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                  int h0, int w0, float (*sx)[SS_IW], float (*sy)[SS_IW],
                                                  float (*hm)[SS_IH][SS_TW]) {
    for (int i = threadIdx.x; i < SS_IH * SS_IW; i += blockDim.x) {
        const int r = i / SS_IW, c = i - r * SS_IW;
        const int h = h0 + r, w = w0 + c;
        const bool in = (h >= 0 && h < H && w >= 0 && w < W);
        sx[r][c] = in ? x[static_cast<size_t>(h) * W + w] : 0.0f;
        sy[r][c] = in ? y[static_cast<size_t>(h) * W + w] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SS_IH * SS_TW; i += blockDim.x) {
        const int r = i / SS_TW, c = i - r * SS_TW;
        float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
#pragma unroll
        for (int k = 0; k < SS_WIN; ++k) {
            const float g = c_gauss[k], xv = sx[r][c + k], yv = sy[r][c + k];
            a0 = fmaf(g, xv, a0); a1 = fmaf(g, yv, a1);
            a2 = fmaf(g, xv * xv, a2); a3 = fmaf(g, yv * yv, a3); a4 = fmaf(g, xv * yv, a4);
        }
        hm[0][r][c] = a0; hm[1][r][c] = a1; hm[2][r][c] = a2; hm[3][r][c] = a3; hm[4][r][c] = a4;
    }
    __syncthreads();
}

__device__ __forceinline__ SsimPix ssim_vertical(float (*hm)[SS_IH][SS_TW], int r, int c) {
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
#pragma unroll
    for (int k = 0; k < SS_WIN; ++k) {
        const float g = c_gauss[k];
        a0 = fmaf(g, hm[0][r + k][c], a0); a1 = fmaf(g, hm[1][r + k][c], a1);
        a2 = fmaf(g, hm[2][r + k][c], a2); a3 = fmaf(g, hm[3][r + k][c], a3); a4 = fmaf(g, hm[4][r + k][c], a4);
    }
    SsimPix p;
    p.mu_x = a0; p.mu_y = a1;
    p.sxx = a2 - a0 * a0; p.syy = a3 - a1 * a1; p.sxy = a4 - a0 * a1;
    return p;
}

// grid (tiles_x, tiles_y, planes); valid output extent (H-10) x (W-10); stats[plane] = {sum ssim, sum cs} (f64)
__global__ void __launch_bounds__(SS_THREADS) ssim_stats_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                                float C1, float C2, double* __restrict__ stats) {
    __shared__ float sx[SS_IH][SS_IW], sy[SS_IH][SS_IW];
    __shared__ float hm[5][SS_IH][SS_TW];
    const int plane = blockIdx.z;
    const int oh0 = blockIdx.y * SS_TH, ow0 = blockIdx.x * SS_TW;      // output coords == input coords of the window's top-left
    const size_t off = static_cast<size_t>(plane) * H * W;
    ssim_tile_moments(x + off, y + off, H, W, oh0, ow0, sx, sy, hm);
    const int Ho = H - 2 * SS_R, Wo = W - 2 * SS_R;
    float s_ssim = 0.0f, s_cs = 0.0f;
    for (int i = threadIdx.x; i < SS_TH * SS_TW; i += blockDim.x) {
        const int r = i / SS_TW, c = i - r * SS_TW;
        if (oh0 + r < Ho && ow0 + c < Wo) {
            const SsimPix p = ssim_vertical(hm, r, c);
            const float cs = (2.0f * p.sxy + C2) / (p.sxx + p.syy + C2);
            const float l = (2.0f * p.mu_x * p.mu_y + C1) / (p.mu_x * p.mu_x + p.mu_y * p.mu_y + C1);
            s_ssim += l * cs;
            s_cs += cs;
        }
    }
    __shared__ float red[2][SS_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_ssim += __shfl_xor_sync(0xffffffffu, s_ssim, o);
        s_cs += __shfl_xor_sync(0xffffffffu, s_cs, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_ssim; red[1][threadIdx.x >> 5] = s_cs; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0.0;
        for (int w = 0; w < SS_THREADS / 32; ++w) v += red[threadIdx.x][w];
        atomicAdd(stats + 2 * plane + threadIdx.x, v);
    }
}

// P maps at the valid extent: part[0] = P_mu, part[1] = P_xx, part[2] = P_xy, each [planes][Ho][Wo]
// gw[plane] = {g_s, g_c}: upstream gradients of the plane's ssim / cs MEANS already divided by the pixel count
__global__ void __launch_bounds__(SS_THREADS) ssim_partials_kernel(const float* __restrict__ x, const float* __restrict__ y, int H,
                                                                   int W, float C1, float C2, const float* __restrict__ gw,
                                                                   float* __restrict__ part, size_t part_stride) {
    __shared__ float sx[SS_IH][SS_IW], sy[SS_IH][SS_IW];
    __shared__ float hm[5][SS_IH][SS_TW];
    const int plane = blockIdx.z;
    const int oh0 = blockIdx.y * SS_TH, ow0 = blockIdx.x * SS_TW;
    const size_t off = static_cast<size_t>(plane) * H * W;
    ssim_tile_moments(x + off, y + off, H, W, oh0, ow0, sx, sy, hm);
    const int Ho = H - 2 * SS_R, Wo = W - 2 * SS_R;
    const float gs = gw[2 * plane], gc = gw[2 * plane + 1];
    float* pp = part + static_cast<size_t>(plane) * Ho * Wo;
    for (int i = threadIdx.x; i < SS_TH * SS_TW; i += blockDim.x) {
        const int r = i / SS_TW, c = i - r * SS_TW;
        if (oh0 + r < Ho && ow0 + c < Wo) {
            const SsimPix p = ssim_vertical(hm, r, c);
            const float A1 = 2.0f * p.mu_x * p.mu_y + C1, B1 = p.mu_x * p.mu_x + p.mu_y * p.mu_y + C1;
            const float A2 = 2.0f * p.sxy + C2, B2 = p.sxx + p.syy + C2;
            const float l = A1 / B1, cs = A2 / B2;
            const float k = gs * l + gc;                         // dF/d cs
            const float pxx = -k * A2 / (B2 * B2);
            const float pxy = 2.0f * k / B2;
            const float pmu = gs * cs * (2.0f * p.mu_y / B1 - 2.0f * p.mu_x * A1 / (B1 * B1)) - 2.0f * p.mu_x * pxx - p.mu_y * pxy;
            const size_t o = static_cast<size_t>(oh0 + r) * Wo + (ow0 + c);
            pp[o] = pmu; pp[part_stride + o] = pxx; pp[2 * part_stride + o] = pxy;
        }
    }
}

// dx[h,w] (+)= sum_{a,b} g[a] g[b] (P_mu + 2 x[h,w] P_xx + y[h,w] P_xy)[h-a, w-b]  over the valid P extent
// grid (tiles_x, tiles_y, planes) over the INPUT extent
__global__ void __launch_bounds__(SS_THREADS) ssim_grad_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                               const float* __restrict__ part, size_t part_stride, int accumulate,
                                                               float* __restrict__ dx) {
    __shared__ float sp[3][SS_IH][SS_IW];
    __shared__ float hm[3][SS_IH][SS_TW];
    const int plane = blockIdx.z;
    const int h0 = blockIdx.y * SS_TH, w0 = blockIdx.x * SS_TW;
    const int Ho = H - 2 * SS_R, Wo = W - 2 * SS_R;
    const float* pp = part + static_cast<size_t>(plane) * Ho * Wo;
    // P window: output pixel (h, w) gathers P[h - a][w - b], a, b in 0..10  ->  P rows h0-10 .. h0+TH-1
    for (int i = threadIdx.x; i < SS_IH * SS_IW; i += blockDim.x) {
        const int r = i / SS_IW, c = i - r * SS_IW;
        const int ph = h0 - 2 * SS_R + r, pw = w0 - 2 * SS_R + c;
        const bool in = (ph >= 0 && ph < Ho && pw >= 0 && pw < Wo);
        const size_t o = static_cast<size_t>(ph) * Wo + pw;
#pragma unroll
        for (int q = 0; q < 3; ++q) sp[q][r][c] = in ? pp[q * part_stride + o] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SS_IH * SS_TW; i += blockDim.x) {
        const int r = i / SS_TW, c = i - r * SS_TW;
        float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int k = 0; k < SS_WIN; ++k) {                       // P[.., w - b] with b = 10 - k  <->  window column c + k
            const float g = c_gauss[SS_WIN - 1 - k];
            a0 = fmaf(g, sp[0][r][c + k], a0); a1 = fmaf(g, sp[1][r][c + k], a1); a2 = fmaf(g, sp[2][r][c + k], a2);
        }
        hm[0][r][c] = a0; hm[1][r][c] = a1; hm[2][r][c] = a2;
    }
    __syncthreads();
    const size_t off = static_cast<size_t>(plane) * H * W;
    for (int i = threadIdx.x; i < SS_TH * SS_TW; i += blockDim.x) {
        const int r = i / SS_TW, c = i - r * SS_TW;
        const int h = h0 + r, w = w0 + c;
        if (h < H && w < W) {
            float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
            for (int k = 0; k < SS_WIN; ++k) {
                const float g = c_gauss[SS_WIN - 1 - k];
                a0 = fmaf(g, hm[0][r + k][c], a0); a1 = fmaf(g, hm[1][r + k][c], a1); a2 = fmaf(g, hm[2][r + k][c], a2);
            }
            const size_t o = off + static_cast<size_t>(h) * W + w;
            const float v = a0 + 2.0f * x[o] * a1 + y[o] * a2;
            dx[o] = accumulate ? dx[o] + v : v;
        }
    }
}

static int upload_window(float sigma) {
    static float cur_dev[32];                       // __constant__ memory is per device
    static bool init = false;
    if (!init) { for (float& v : cur_dev) v = -1.0f; init = true; }
    int dev = 0;
    cudaGetDevice(&dev);
    float& cur = cur_dev[dev & 31];
    if (cur == sigma) return 0;
    float g[SS_WIN];
    double sum = 0.0;
    for (int i = 0; i < SS_WIN; ++i) {
        const double d = i - SS_WIN / 2;
        g[i] = static_cast<float>(exp(-(d * d) / (2.0 * sigma * sigma)));
        sum += g[i];
    }
    for (int i = 0; i < SS_WIN; ++i) g[i] = static_cast<float>(g[i] / sum);
    cudaError_t e = cudaMemcpyToSymbol(c_gauss, g, sizeof(g));
    if (e != cudaSuccess) return set_error(static_cast<int>(e), "ssim window upload: %s", cudaGetErrorString(e));
    cur = sigma;
    return 0;
}

}  // namespace bnerv

using namespace bnerv;

static int ssim_check(const void* x, const void* y, int planes, int H, int W, const char* who) {
    if (!x || !y) return set_error(BNERV_E_BADARG, "%s: null pointer", who);
    if (planes <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "%s: non-positive size", who);
    if (H <= 2 * SS_R || W <= 2 * SS_R) return set_error(BNERV_E_UNSUPPORTED, "%s: plane %dx%d smaller than the 11x11 window", who, H, W);
    if (planes > 65535) return set_error(BNERV_E_UNSUPPORTED, "%s: more than 65535 planes", who);
    return 0;
}

extern "C" int bnerv_ssim_stats(const float* x, const float* y, int planes, int H, int W, float C1, float C2, double* stats,
                                void* stream) {
    int rc = ssim_check(x, y, planes, H, W, "ssim_stats");
    if (rc) return rc;
    if (!stats) return set_error(BNERV_E_BADARG, "ssim_stats: null pointer");
    if ((rc = upload_window(1.5f))) return rc;
    const int Ho = H - 2 * SS_R, Wo = W - 2 * SS_R;
    dim3 grid((Wo + SS_TW - 1) / SS_TW, (Ho + SS_TH - 1) / SS_TH, planes);
    ssim_stats_kernel<<<grid, SS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, y, H, W, C1, C2, stats);
    return check_launch("ssim_stats_kernel");
}

extern "C" int bnerv_ssim_grad(const float* x, const float* y, int planes, int H, int W, float C1, float C2, const float* gw,
                               float* scratch, int accumulate, float* dx, void* stream) {
    int rc = ssim_check(x, y, planes, H, W, "ssim_grad");
    if (rc) return rc;
    if (!gw || !scratch || !dx) return set_error(BNERV_E_BADARG, "ssim_grad: null pointer");
    if ((rc = upload_window(1.5f))) return rc;
    const int Ho = H - 2 * SS_R, Wo = W - 2 * SS_R;
    const size_t stride = static_cast<size_t>(planes) * Ho * Wo;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 g1((Wo + SS_TW - 1) / SS_TW, (Ho + SS_TH - 1) / SS_TH, planes);
    ssim_partials_kernel<<<g1, SS_THREADS, 0, st>>>(x, y, H, W, C1, C2, gw, scratch, stride);
    if ((rc = check_launch("ssim_partials_kernel"))) return rc;
    dim3 g2((W + SS_TW - 1) / SS_TW, (H + SS_TH - 1) / SS_TH, planes);
    ssim_grad_kernel<<<g2, SS_THREADS, 0, st>>>(x, y, H, W, scratch, stride, accumulate, dx);
    return check_launch("ssim_grad_kernel");
}

extern "C" size_t bnerv_ssim_scratch_floats(int planes, int H, int W) {
    if (planes <= 0 || H <= 2 * SS_R || W <= 2 * SS_R) return 0;
    return 3 * static_cast<size_t>(planes) * (H - 2 * SS_R) * (W - 2 * SS_R);
}
