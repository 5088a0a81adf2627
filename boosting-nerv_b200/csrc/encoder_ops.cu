// Forward of the ConvNeXt encoder of HNeRV_Boost (encoder half; model_blocks.py:223-320,
// model_hnerv.py:189,230-234): what forward_encoder() runs per frame in evaluate() and when the embeddings of a sequence
// are extracted for the compression path.  Inference only - training keeps the torch module (autograd).
//
//   stage i:  [LayerNorm channels_first]  ->  conv k = s, stride s (non-overlapping patches)  ->  [LayerNorm, stage 0]
//             -> blocks:  x + gamma * pwconv2(gelu_erf(pwconv1(LayerNorm(dwconv7x7(x)))))
//
// Arithmetic: f32 on the CUDA cores with exact erff - the embedding feeds the whole decoder, so it keeps the reference's
// own precision (parity gate 1e-5) instead of f16 tensor-core operands; its 8.4 GFLOP per 1080p frame are 0.2 % of the
// decoder's work.  Layout between kernels: channels-last f32 [B][H][W][C] (LayerNorm and the point-wise MLP are per-pixel
// reductions over C); the patch conv gathers its A operand straight from the NCHW frame (stage 0) or the previous NHWC
// map (with the preceding LayerNorm applied on the fly from per-pixel statistics), so no im2col buffer exists.
// One templated register-tiled SGEMM (128x64x16 or 64x64x16 block tiles) serves the patch conv and both point-wise layers.
#include <algorithm>
#include "common.cuh"

namespace bnerv {

// ---- per-pixel LayerNorm statistics over C (biased variance, two passes like F.layer_norm) -----------------------
__global__ void __launch_bounds__(256) enc_ln_stats_kernel(const float* __restrict__ x, size_t pixels, int C, float eps,
                                                           float2* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const size_t warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    for (size_t p = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; p < pixels; p += warps) {   // warp-uniform
        const float* row = x + p * C;
        float s = 0.0f;
        for (int c = lane; c < C; c += 32) s += row[c];
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        const float mu = s / static_cast<float>(C);
        float v = 0.0f;
        for (int c = lane; c < C; c += 32) { const float dlt = row[c] - mu; v += dlt * dlt; }
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) stats[p] = make_float2(mu, rsqrtf(v / static_cast<float>(C) + eps));
    }
}

// LayerNorm in place on a channels-last map (the one after the stage-0 patch conv, model_blocks.py:288-291)
__global__ void __launch_bounds__(256) enc_ln_inplace_kernel(float* __restrict__ x, size_t pixels, int C, float eps,
                                                             const float* __restrict__ w, const float* __restrict__ b) {
    const int lane = threadIdx.x & 31;
    const size_t warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    for (size_t p = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; p < pixels; p += warps) {   // warp-uniform
        float* row = x + p * C;
        float s = 0.0f;
        for (int c = lane; c < C; c += 32) s += row[c];
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        const float mu = s / static_cast<float>(C);
        float v = 0.0f;
        for (int c = lane; c < C; c += 32) { const float dlt = row[c] - mu; v += dlt * dlt; }
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        const float rstd = rsqrtf(v / static_cast<float>(C) + eps);
        for (int c = lane; c < C; c += 32) row[c] = (row[c] - mu) * rstd * __ldg(w + c) + __ldg(b + c);
    }
}

// ---- depthwise 7x7, padding 3, channels-last (Block.dwconv, model_blocks.py:236) ---------------------------------
// A thread owns a strip of DW_TW output pixels of one row and one channel: each of the 7 input rows is loaded once
// (DW_TW + 6 values) and reused by the 7 column taps from registers - 12 loads per output instead of 49.  Neighbouring
// threads are neighbouring channels, so every load / store is a contiguous run of C floats.
constexpr int DW_TW = 8;

__global__ void __launch_bounds__(256) enc_dwconv7_kernel(const float* __restrict__ x, int B, int H, int W, int C,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* __restrict__ y) {
    extern __shared__ float wt[];                       // [49][C]: tap-major so that neighbouring channels read neighbouring words
    for (int k = threadIdx.x; k < 49 * C; k += blockDim.x) {
        const int c = k / 49, t = k - c * 49;
        wt[t * C + c] = __ldg(w + k);
    }
    __syncthreads();
    const int strips = (W + DW_TW - 1) / DW_TW;
    const size_t total = static_cast<size_t>(B) * H * strips * C;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        size_t q = idx / C;
        const int w0 = static_cast<int>(q % strips) * DW_TW;
        q /= strips;
        const int hq = static_cast<int>(q % H);
        const size_t b = q / H;
        float acc[DW_TW];
        const float bv = __ldg(bias + c);
#pragma unroll
        for (int o = 0; o < DW_TW; ++o) acc[o] = bv;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const int hh = hq + r - 3;
            if (hh < 0 || hh >= H) continue;
            const float* row = x + ((b * H + hh) * W) * C + c;
            float v[DW_TW + 6];
#pragma unroll
            for (int t = 0; t < DW_TW + 6; ++t) {
                const int ww = w0 - 3 + t;
                v[t] = (ww >= 0 && ww < W) ? __ldg(row + static_cast<size_t>(ww) * C) : 0.0f;
            }
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                const float wgt = wt[(r * 7 + s) * C + c];
#pragma unroll
                for (int o = 0; o < DW_TW; ++o) acc[o] = fmaf(v[o + s], wgt, acc[o]);
            }
        }
        float* out = y + ((b * H + hq) * W) * C + c;
#pragma unroll
        for (int o = 0; o < DW_TW; ++o)
            if (w0 + o < W) out[static_cast<size_t>(w0 + o) * C] = acc[o];
    }
}

// ---- SGEMM  out[p][n] = epi(sum_k A[p][k] * Wt[n][k] + bias[n]) ---------------------------------------------------
enum { ENC_A_ROWS = 0, ENC_A_PATCH = 1 };
enum { ENC_EPI_BIAS = 0, ENC_EPI_GELU = 1, ENC_EPI_SCALE_RESID = 2 };

struct EncGemm {
    const float* a;            // ROWS: [M][K] row-major.  PATCH: the input map (NCHW or NHWC)
    const float2* stats;       // per-row (ROWS) / per-input-pixel (PATCH) LayerNorm (mean, rstd), or null: no LayerNorm on A
    const float* ln_w;         // [K] (ROWS) / [Cin] (PATCH)
    const float* ln_b;
    const float* w;            // [N][K] row-major (nn.Linear weight / OIHW conv weight flattened)
    const float* bias;         // [N]
    const float* gamma;        // [N], EPI_SCALE_RESID
    float* out;                // [M][N]; EPI_SCALE_RESID: out += gamma * (acc + bias), i.e. the residual is read from `out`
    int M, N, K;
    int Cin, H, W, s, Ho, Wo, nchw;    // PATCH geometry
};

constexpr int EG_BN = 64, EG_BK = 16;       // block tile: (16 * TM) rows x 64 columns x 16 k; 16 x 16 threads, TM x 4 outputs each

template <int AMODE>
__device__ __forceinline__ float enc_load_a(const EncGemm& g, int p, int k) {
    if (p >= g.M || k >= g.K) return 0.0f;
    if (AMODE == ENC_A_ROWS) {
        float v = __ldg(g.a + static_cast<size_t>(p) * g.K + k);
        if (g.stats) { const float2 st = __ldg(g.stats + p); v = (v - st.x) * st.y * __ldg(g.ln_w + k) + __ldg(g.ln_b + k); }
        return v;
    } else {
        const int ss = g.s * g.s;
        const int c = k / ss, ij = k - c * ss, i = ij / g.s, j = ij - i * g.s;
        const int wo = p % g.Wo, ho = (p / g.Wo) % g.Ho, b = p / (g.Wo * g.Ho);
        const int hi = ho * g.s + i, wi = wo * g.s + j;
        const size_t pin = (static_cast<size_t>(b) * g.H + hi) * g.W + wi;
        float v = g.nchw ? __ldg(g.a + ((static_cast<size_t>(b) * g.Cin + c) * g.H + hi) * g.W + wi) : __ldg(g.a + pin * g.Cin + c);
        if (g.stats) { const float2 st = __ldg(g.stats + pin); v = (v - st.x) * st.y * __ldg(g.ln_w + c) + __ldg(g.ln_b + c); }
        return v;
    }
}

template <int AMODE, int EPI, int TM>
__global__ void __launch_bounds__(256) enc_gemm_kernel(EncGemm g) {
    constexpr int BM = 16 * TM;
    __shared__ __align__(16) float As[EG_BK][BM + 4];
    __shared__ __align__(16) float Bs[EG_BK][EG_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * EG_BN;
    const int tm = (tid >> 4) * TM, tn = (tid & 15) * 4;
    float acc[TM][4] = {};
    for (int k0 = 0; k0 < g.K; k0 += EG_BK) {
#pragma unroll
        for (int r = 0; r < BM / 16; ++r) {                     // BM x 16 elements of A: consecutive threads walk k
            const int e = tid + r * 256;
            const int kk = e & 15, mm = e >> 4;
            As[kk][mm] = enc_load_a<AMODE>(g, m0 + mm, k0 + kk);
        }
#pragma unroll
        for (int r = 0; r < EG_BN / 16; ++r) {                  // 64 x 16 elements of W
            const int e = tid + r * 256;
            const int kk = e & 15, nn = e >> 4;
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < g.N && k < g.K) ? __ldg(g.w + static_cast<size_t>(n) * g.K + k) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < EG_BK; ++kk) {
            float a4[TM];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 av = *reinterpret_cast<const float4*>(&As[kk][tm + i]);
                a4[i] = av.x; a4[i + 1] = av.y; a4[i + 2] = av.z; a4[i + 3] = av.w;
            }
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tn]);
            const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int p = m0 + tm + i;
        if (p >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn + j;
            if (n >= g.N) continue;
            float v = acc[i][j] + __ldg(g.bias + n);
            float* o = g.out + static_cast<size_t>(p) * g.N + n;
            if (EPI == ENC_EPI_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));      // nn.GELU() exact form
            if (EPI == ENC_EPI_SCALE_RESID) v = *o + __ldg(g.gamma + n) * v;
            *o = v;
        }
    }
}

__global__ void __launch_bounds__(256) enc_nhwc_to_nchw_kernel(const float* __restrict__ x, int B, int H, int W, int C,
                                                               float* __restrict__ y) {
    const size_t total = static_cast<size_t>(B) * H * W * C;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t hw = static_cast<size_t>(H) * W;
        const size_t q = idx % hw, c = (idx / hw) % C, b = idx / (hw * C);      // idx enumerates the NCHW output
        y[idx] = __ldg(x + (b * hw + q) * C + c);
    }
}

template <int AMODE, int EPI>
static int enc_launch_gemm(const EncGemm& g, cudaStream_t st, const char* what) {
    const unsigned nt = (g.N + EG_BN - 1) / EG_BN;
    if (static_cast<size_t>((g.M + 127) / 128) * nt >= 2 * 148) {       // enough 128-row tiles for two waves: the wider micro-tile
        dim3 grid((g.M + 127) / 128, nt);
        enc_gemm_kernel<AMODE, EPI, 8><<<grid, 256, 0, st>>>(g);
    } else {
        dim3 grid((g.M + 63) / 64, nt);
        enc_gemm_kernel<AMODE, EPI, 4><<<grid, 256, 0, st>>>(g);
    }
    return check_launch(what);
}

static unsigned enc_blocks(size_t work_items, int per_block) {
    return static_cast<unsigned>(std::min<size_t>((work_items + per_block - 1) / per_block, 148u * 16u));
}

}  // namespace bnerv

using namespace bnerv;

extern "C" size_t bnerv_convnext_stage_work_floats(int B, int Hin, int Win, int s, int Cout) {
    if (B <= 0 || Hin <= 0 || Win <= 0 || s <= 0 || Cout <= 0) return 0;
    const size_t pin = static_cast<size_t>(B) * Hin * Win, pout = static_cast<size_t>(B) * (Hin / s) * (Win / s);
    return 2 * pin + 2 * pout + pout * Cout + pout * 4 * Cout;      // input stats | block stats | dwconv map | hidden map
}

extern "C" int bnerv_convnext_stage_fwd(const bnerv_convnext_stage* sg, const float* x, int x_is_nchw, int B, int Hin, int Win,
                                        float* y_nhwc, float* work, void* stream) {
    if (!sg || !x || !y_nhwc || !work) return set_error(BNERV_E_BADARG, "convnext_stage_fwd: null pointer");
    if (!sg->down_w || !sg->down_b) return set_error(BNERV_E_BADARG, "convnext_stage_fwd: null down-conv weights");
    if (B <= 0 || Hin <= 0 || Win <= 0 || sg->Cin <= 0 || sg->Cout <= 0 || sg->s <= 0 || sg->n_blocks < 0)
        return set_error(BNERV_E_BADARG, "convnext_stage_fwd: non-positive size");
    if ((sg->ln_in_w == nullptr) != (sg->ln_in_b == nullptr) || (sg->ln_out_w == nullptr) != (sg->ln_out_b == nullptr))
        return set_error(BNERV_E_BADARG, "convnext_stage_fwd: LayerNorm weight without bias (or the reverse)");
    if (sg->ln_in_w && x_is_nchw) return set_error(BNERV_E_UNSUPPORTED, "convnext_stage_fwd: LayerNorm before the conv needs a channels-last input");
    if (sg->n_blocks > 0 && !sg->blocks) return set_error(BNERV_E_BADARG, "convnext_stage_fwd: null block array");
    const int s = sg->s, Ho = Hin / s, Wo = Win / s, C = sg->Cout;
    if (Ho <= 0 || Wo <= 0) return set_error(BNERV_E_BADARG, "convnext_stage_fwd: input smaller than one %dx%d patch", s, s);
    if (49 * C * sizeof(float) > 48 * 1024) return set_error(BNERV_E_UNSUPPORTED, "convnext_stage_fwd: %d channels (depthwise weights exceed 48 KB)", C);
    const size_t pin = static_cast<size_t>(B) * Hin * Win, pout = static_cast<size_t>(B) * Ho * Wo;
    if (pout >= (1u << 31) / 4u / static_cast<size_t>(C)) return set_error(BNERV_E_UNSUPPORTED, "convnext_stage_fwd: map too large for 32-bit row indices");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float2* stats_in = reinterpret_cast<float2*>(work);
    float2* stats_blk = reinterpret_cast<float2*>(work + 2 * pin);
    float* dmap = work + 2 * pin + 2 * pout;
    float* hidden = dmap + pout * C;
    const float eps = 1e-6f;                                 // every LayerNorm of the encoder (model_blocks.py:237,283,290)

    if (sg->ln_in_w) {
        enc_ln_stats_kernel<<<enc_blocks(pin * 32, 256), 256, 0, st>>>(x, pin, sg->Cin, eps, stats_in);
        if (int rc = check_launch("enc_ln_stats_kernel")) return rc;
    }
    EncGemm g{};
    g.a = x; g.stats = sg->ln_in_w ? stats_in : nullptr; g.ln_w = sg->ln_in_w; g.ln_b = sg->ln_in_b;
    g.w = sg->down_w; g.bias = sg->down_b; g.out = y_nhwc;
    g.M = static_cast<int>(pout); g.N = C; g.K = sg->Cin * s * s;
    g.Cin = sg->Cin; g.H = Hin; g.W = Win; g.s = s; g.Ho = Ho; g.Wo = Wo; g.nchw = x_is_nchw;
    if (int rc = enc_launch_gemm<ENC_A_PATCH, ENC_EPI_BIAS>(g, st, "enc_gemm_kernel(patch conv)")) return rc;
    if (sg->ln_out_w) {
        enc_ln_inplace_kernel<<<enc_blocks(pout * 32, 256), 256, 0, st>>>(y_nhwc, pout, C, eps, sg->ln_out_w, sg->ln_out_b);
        if (int rc = check_launch("enc_ln_inplace_kernel")) return rc;
    }
    for (int bi = 0; bi < sg->n_blocks; ++bi) {
        const bnerv_convnext_block& blk = sg->blocks[bi];
        if (!blk.dw_w || !blk.dw_b || !blk.ln_w || !blk.ln_b || !blk.pw1_w || !blk.pw1_b || !blk.pw2_w || !blk.pw2_b || !blk.gamma)
            return set_error(BNERV_E_BADARG, "convnext_stage_fwd: null weight in block %d", bi);
        enc_dwconv7_kernel<<<enc_blocks(static_cast<size_t>(B) * Ho * ((Wo + DW_TW - 1) / DW_TW) * C, 256), 256, 49 * C * sizeof(float), st>>>(y_nhwc, B, Ho, Wo, C, blk.dw_w, blk.dw_b, dmap);
        if (int rc = check_launch("enc_dwconv7_kernel")) return rc;
        enc_ln_stats_kernel<<<enc_blocks(pout * 32, 256), 256, 0, st>>>(dmap, pout, C, eps, stats_blk);
        if (int rc = check_launch("enc_ln_stats_kernel")) return rc;
        EncGemm g1{};
        g1.a = dmap; g1.stats = stats_blk; g1.ln_w = blk.ln_w; g1.ln_b = blk.ln_b; g1.w = blk.pw1_w; g1.bias = blk.pw1_b; g1.out = hidden;
        g1.M = static_cast<int>(pout); g1.N = 4 * C; g1.K = C;
        if (int rc = enc_launch_gemm<ENC_A_ROWS, ENC_EPI_GELU>(g1, st, "enc_gemm_kernel(pwconv1)")) return rc;
        EncGemm g2{};
        g2.a = hidden; g2.w = blk.pw2_w; g2.bias = blk.pw2_b; g2.gamma = blk.gamma; g2.out = y_nhwc;
        g2.M = static_cast<int>(pout); g2.N = C; g2.K = 4 * C;
        if (int rc = enc_launch_gemm<ENC_A_ROWS, ENC_EPI_SCALE_RESID>(g2, st, "enc_gemm_kernel(pwconv2)")) return rc;
    }
    return 0;
}

extern "C" int bnerv_nhwc_to_nchw(const float* x_nhwc, int B, int H, int W, int C, float* y_nchw, void* stream) {
    if (!x_nhwc || !y_nchw) return set_error(BNERV_E_BADARG, "nhwc_to_nchw: null pointer");
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0) return set_error(BNERV_E_BADARG, "nhwc_to_nchw: non-positive size");
    const size_t total = static_cast<size_t>(B) * H * W * C;
    enc_nhwc_to_nchw_kernel<<<enc_blocks(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x_nhwc, B, H, W, C, y_nchw);
    return check_launch("enc_nhwc_to_nchw_kernel");
}
