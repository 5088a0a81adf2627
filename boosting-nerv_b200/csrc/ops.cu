// Small kernels around the tensor-core conv: weight packing, layout conversion at the model boundary,
// the TAT (SFT) affine-parameter MLP, the 1x1 stem MLP layer, standalone PixelShuffle, and an f32
// CUDA-core fused conv on the reference's own layouts (exact-arithmetic cross-check path).
#include <algorithm>
#include "common.cuh"

namespace bnerv {

// ---------------------------------------------------------------------------------------------
// weight packing: OIHW f32 -> [tap][Kp/8][Np][8] f16, PixelShuffle folded into the row order
// ---------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int s, int cout_p,
                                   int cin_p, __half* __restrict__ wp) {
    const int np = s * s * cout_p;
    const size_t total = static_cast<size_t>(k) * k * cin_p * np;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int kk = idx & 7;
        size_t r = idx >> 3;
        const int n = r % np;
        r /= np;
        const int kg  = r % (cin_p >> 3);
        const int tap = r / (cin_p >> 3);
        int c, i, j;
        packed_row_to_cij(n, s, cout_p, c, i, j);
        const int ci = kg * 8 + kk;
        float v = 0.0f;
        if (c < Cout && ci < Cin) {
            const int o = c * s * s + i * s + j;  // reference output channel feeding (c, i, j)
            v = w[(static_cast<size_t>(o) * Cin + ci) * k * k + tap];
        }
        wp[idx] = __float2half_rn(v);
    }
}

// Quantised ingest (lib/transform_ops.py:239-251 Scale_T.decode, lib/quant_ops.py:40): the effective weight is
// dequant = round(w / scale) * scale; given the integer codes and the scale(s) the f32 product is formed here exactly
// as torch forms it, so the packed f16 weights are bit-identical to packing the materialised dequant_w.
template <typename CodeT>
__global__ void pack_weight_q_kernel(const CodeT* __restrict__ q, const float* __restrict__ scale, int scale_per_channel,
                                     int Cout, int Cin, int k, int s, int cout_p, int cin_p, __half* __restrict__ wp) {
    const int np = s * s * cout_p;
    const size_t total = static_cast<size_t>(k) * k * cin_p * np;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int kk = idx & 7;
        size_t r = idx >> 3;
        const int n = r % np;
        r /= np;
        const int kg  = r % (cin_p >> 3);
        const int tap = r / (cin_p >> 3);
        int c, i, j;
        packed_row_to_cij(n, s, cout_p, c, i, j);
        const int ci = kg * 8 + kk;
        float v = 0.0f;
        if (c < Cout && ci < Cin) {
            const int o = c * s * s + i * s + j;
            v = __fmul_rn(static_cast<float>(q[(static_cast<size_t>(o) * Cin + ci) * k * k + tap]), scale[scale_per_channel ? o : 0]);
        }
        wp[idx] = __float2half_rn(v);
    }
}

template <typename CodeT>
__global__ void pack_bias_q_kernel(const CodeT* __restrict__ q, const float* __restrict__ scale, int scale_per_channel,
                                   int Cout, int s, int cout_p, float* __restrict__ bp) {
    const int np = s * s * cout_p;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < np; n += gridDim.x * blockDim.x) {
        int c, i, j;
        packed_row_to_cij(n, s, cout_p, c, i, j);
        const int o = c * s * s + i * s + j;
        bp[n] = (q != nullptr && c < Cout) ? __fmul_rn(static_cast<float>(q[o]), scale[scale_per_channel ? o : 0]) : 0.0f;
    }
}

// head weights for bnerv_head_conv3: [Kp/8][32][8] f16, column n = tap*Cout + c holds W[c][k][tap]
__global__ void pack_head_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int cin_p, __half* __restrict__ wp) {
    const int total = cin_p * 32;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int kk = idx & 7;
        const int n = (idx >> 3) & 31;
        const int kg = idx >> 8;
        const int ci = kg * 8 + kk;
        float v = 0.0f;
        if (n < 9 * Cout && ci < Cin) {
            const int tap = n / Cout, c = n - tap * Cout;
            v = w[(static_cast<size_t>(c) * Cin + ci) * 9 + tap];
        }
        wp[idx] = __float2half_rn(v);
    }
}

__global__ void pack_bias_kernel(const float* __restrict__ bias, int Cout, int s, int cout_p, float* __restrict__ bp) {
    const int np = s * s * cout_p;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < np; n += gridDim.x * blockDim.x) {
        int c, i, j;
        packed_row_to_cij(n, s, cout_p, c, i, j);
        bp[n] = (bias != nullptr && c < Cout) ? bias[c * s * s + i * s + j] : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// layout conversion
// ---------------------------------------------------------------------------------------------
__global__ void nchw_to_c8_kernel(const float* __restrict__ x, int B, int C, int H, int W, int cp, __half* __restrict__ y) {
    const size_t hw = static_cast<size_t>(H) * W;
    const size_t total = static_cast<size_t>(B) * (cp >> 3) * hw;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t p = idx % hw;
        const size_t r = idx / hw;
        const int g = r % (cp >> 3);
        const int b = r / (cp >> 3);
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            v[k] = (c < C) ? x[(static_cast<size_t>(b) * C + c) * hw + p] : 0.0f;
        }
        uint4 o;
        o.x = pack_h2_sat(v[0], v[1]); o.y = pack_h2_sat(v[2], v[3]);
        o.z = pack_h2_sat(v[4], v[5]); o.w = pack_h2_sat(v[6], v[7]);
        reinterpret_cast<uint4*>(y)[idx] = o;
    }
}

__global__ void c8_to_nchw_kernel(const __half* __restrict__ x, int B, int C, int H, int W, int cp, float* __restrict__ y) {
    const size_t hw = static_cast<size_t>(H) * W;
    const size_t total = static_cast<size_t>(B) * (cp >> 3) * hw;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t p = idx % hw;
        const size_t r = idx / hw;
        const int g = r % (cp >> 3);
        const int b = r / (cp >> 3);
        const uint4 u = reinterpret_cast<const uint4*>(x)[idx];
        const float2 a0 = unpack_h2(u.x), a1 = unpack_h2(u.y), a2 = unpack_h2(u.z), a3 = unpack_h2(u.w);
        const float v[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            if (c < C) y[(static_cast<size_t>(b) * C + c) * hw + p] = v[k];
        }
    }
}

__global__ void pixel_shuffle_kernel(const float* __restrict__ x, int B, int C, int H, int W, int s, float* __restrict__ y) {
    const int Ho = H * s, Wo = W * s;
    const size_t total = static_cast<size_t>(B) * C * Ho * Wo;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int wo = idx % Wo;
        size_t r = idx / Wo;
        const int ho = r % Ho;
        r /= Ho;
        const int c = r % C;
        const int b = r / C;
        const int h = ho / s, i = ho - h * s, w = wo / s, j = wo - w * s;
        y[idx] = x[((static_cast<size_t>(b) * C * s * s + (c * s * s + i * s + j)) * H + h) * W + w];
    }
}

// ---------------------------------------------------------------------------------------------
// TAT / SFT affine parameters: one block per (layer, batch element)
// ---------------------------------------------------------------------------------------------
__global__ void sft_affine_kernel(const bnerv_sft_layer* __restrict__ layers, const float* __restrict__ e, int ch_t) {
    extern __shared__ float sm[];           // e[ch_t] | hs[ch_t] | hh[ch_t]
    float* se = sm;
    float* hs = sm + ch_t;
    float* hh = sm + 2 * ch_t;
    const bnerv_sft_layer L = layers[blockIdx.x];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < ch_t; i += blockDim.x) se[i] = e[static_cast<size_t>(b) * ch_t + i];
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * ch_t; i += blockDim.x) {
        const bool shift = i >= ch_t;
        const int r = shift ? i - ch_t : i;
        const float* w = (shift ? L.wh0 : L.ws0) + static_cast<size_t>(r) * ch_t;
        float acc = (shift ? L.bh0 : L.bs0)[r];
        for (int j = 0; j < ch_t; ++j) acc = fmaf(w[j], se[j], acc);
        (shift ? hh : hs)[r] = fmaxf(acc, 0.0f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < L.Cp; c += blockDim.x) {
        float g = 0.0f, be = 0.0f;
        if (c < L.C) {
            const float* w1 = L.ws1 + static_cast<size_t>(c) * ch_t;
            const float* w2 = L.wh1 + static_cast<size_t>(c) * ch_t;
            float a1 = L.bs1[c], a2 = L.bh1[c];
            for (int j = 0; j < ch_t; ++j) {
                a1 = fmaf(w1[j], hs[j], a1);
                a2 = fmaf(w2[j], hh[j], a2);
            }
            g  = a1 + 1.0f;                 // model_blocks.py:105  x * (scale + 1) + shift
            be = a2;
        }
        L.g1p[static_cast<size_t>(b) * L.Cp + c]  = g;
        L.beta[static_cast<size_t>(b) * L.Cp + c] = be;
    }
}

// ---------------------------------------------------------------------------------------------
// y = act(W x + b): one warp per output row, x staged in shared memory
// ---------------------------------------------------------------------------------------------
__global__ void linear_act_kernel(const float* __restrict__ x, int B, int Cin, const float* __restrict__ w,
                                  const float* __restrict__ bias, int Cout, int act, float* __restrict__ y) {
    extern __shared__ float sx[];           // [B][Cin]
    for (int i = threadIdx.x; i < B * Cin; i += blockDim.x) sx[i] = x[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int o = blockIdx.x * warps_per_block + (threadIdx.x >> 5); o < Cout; o += gridDim.x * warps_per_block) {
        const float* wr = w + static_cast<size_t>(o) * Cin;
        for (int b = 0; b < B; ++b) {
            float acc = 0.0f;
            for (int i = lane; i < Cin; i += 32) acc = fmaf(wr[i], sx[b * Cin + i], acc);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lane == 0) y[static_cast<size_t>(b) * Cout + o] = apply_act_precise(acc + (bias ? bias[o] : 0.0f), act);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The stem of a frame in two launches (NeRV_Boost: model_nerv.py:47-52; the time MLP of HNeRV_Boost / ENeRV_Boost):
//   linear_pair_kernel<true> : x = PE(t) = cat(sin(t*bases), cos(t*bases)) built in shared memory (model_blocks.py:120-126,
//                              the f32 product and libdevice sinf/cosf torch's own kernels evaluate), then the FIRST layer
//                              of two MLPs that share it;
//   linear_pair_kernel<false>: the next layer of both MLPs; problem 0 may also write its output as the C8 f16 map
//                              [B][Cp/8][hw][8] the conv cascade reads (channel = o / hw, pixel = o % hw: the .view(B, C, h, w)
//                              of model_nerv.py:50) - the layout pass of the cascade input fused into the producer.
// Same per-output arithmetic as linear_act_kernel (one warp per output row, lane-strided FMAs, xor-shuffle tree).
// ---------------------------------------------------------------------------------------------
struct LinearProblem {
    const float* x;       // [B][Cin] (ignored when the input is the position encoding)
    const float* w;       // [Cout][Cin]
    const float* bias;    // [Cout] or null
    float* y;             // [B][Cout] or null
    __half* y_c8;         // C8 map or null
    int Cin, Cout, act, hw, cp;
};

template <bool PE>
__global__ void __launch_bounds__(256) linear_pair_kernel(LinearProblem p0, LinearProblem p1, int B, const float* __restrict__ t,
                                                          const float* __restrict__ bases, int levels, int blocks0) {
    extern __shared__ float sx[];           // [B][Cin]
    const bool second = static_cast<int>(blockIdx.x) >= blocks0;
    const LinearProblem& p = second ? p1 : p0;
    const int Cin = p.Cin;
    if (PE) {
        for (int i = threadIdx.x; i < B * levels; i += blockDim.x) {
            const int b = i / levels, l = i - b * levels;
            const float ang = __fmul_rn(t[b], bases[l]);
            sx[b * Cin + l] = sinf(ang);
            sx[b * Cin + levels + l] = cosf(ang);
        }
    } else {
        for (int i = threadIdx.x; i < B * Cin; i += blockDim.x) sx[i] = p.x[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int blk = second ? blockIdx.x - blocks0 : blockIdx.x;
    const int nblk = second ? gridDim.x - blocks0 : blocks0;
    for (int o = blk * warps_per_block + (threadIdx.x >> 5); o < p.Cout; o += nblk * warps_per_block) {
        const float* wr = p.w + static_cast<size_t>(o) * Cin;
        for (int b = 0; b < B; ++b) {
            float acc = 0.0f;
            for (int i = lane; i < Cin; i += 32) acc = fmaf(wr[i], sx[b * Cin + i], acc);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lane == 0) {
                const float v = apply_act_precise(acc + (p.bias ? p.bias[o] : 0.0f), p.act);
                if (p.y) p.y[static_cast<size_t>(b) * p.Cout + o] = v;
                if (p.y_c8) {
                    const int c = o / p.hw, px = o - c * p.hw;
                    p.y_c8[((static_cast<size_t>(b) * (p.cp >> 3) + (c >> 3)) * p.hw + px) * 8 + (c & 7)] =
                        __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// f32 CUDA-core fused conv on NCHW / OIHW (cross-check + tiny layers)
// one thread = one conv-resolution pixel x CO_T consecutive conv output channels
// ---------------------------------------------------------------------------------------------
constexpr int CO_T = 4;

__global__ void conv_f32_kernel(const float* __restrict__ x, int B, int Cin, int H, int W, const float* __restrict__ w,
                                const float* __restrict__ bias, int Cout, int k, int s, int act,
                                const float* __restrict__ resid, const float* __restrict__ g1p,
                                const float* __restrict__ beta, int ldg, float* __restrict__ out_pre,
                                float* __restrict__ out_aff) {
    const int n_conv = Cout * s * s;
    const int co_blocks = (n_conv + CO_T - 1) / CO_T;
    const size_t hw = static_cast<size_t>(H) * W;
    const size_t total = static_cast<size_t>(B) * co_blocks * hw;
    const int pad = (k - 1) / 2;
    const int Ho = H * s, Wo = W * s;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t p = idx % hw;
        size_t r = idx / hw;
        const int cb = r % co_blocks;
        const int b  = r / co_blocks;
        const int h = p / W, ww = p - static_cast<size_t>(h) * W;
        float acc[CO_T];
#pragma unroll
        for (int t = 0; t < CO_T; ++t) {
            const int o = cb * CO_T + t;
            acc[t] = (bias != nullptr && o < n_conv) ? bias[o] : 0.0f;
        }
        for (int ci = 0; ci < Cin; ++ci) {
            const float* xp = x + (static_cast<size_t>(b) * Cin + ci) * hw;
            for (int kr = 0; kr < k; ++kr) {
                const int hi = h + kr - pad;
                if (hi < 0 || hi >= H) continue;
                for (int ks = 0; ks < k; ++ks) {
                    const int wi = ww + ks - pad;
                    if (wi < 0 || wi >= W) continue;
                    const float xv = xp[static_cast<size_t>(hi) * W + wi];
#pragma unroll
                    for (int t = 0; t < CO_T; ++t) {
                        const int o = cb * CO_T + t;
                        if (o < n_conv) acc[t] = fmaf(w[((static_cast<size_t>(o) * Cin + ci) * k + kr) * k + ks], xv, acc[t]);
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < CO_T; ++t) {
            const int o = cb * CO_T + t;
            if (o >= n_conv) continue;
            const int c = o / (s * s), sub = o - c * s * s;
            const int i = sub / s, j = sub - i * s;
            const size_t off = ((static_cast<size_t>(b) * Cout + c) * Ho + (h * s + i)) * Wo + (ww * s + j);
            float v = apply_act_precise(acc[t], act);
            if (resid != nullptr) v += resid[off];
            if (out_pre != nullptr) out_pre[off] = v;
            if (out_aff != nullptr) out_aff[off] = v * g1p[static_cast<size_t>(b) * ldg + c] + beta[static_cast<size_t>(b) * ldg + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 1x1 head conv to <= 4 channels + OutImg (NeRV_Boost / ENeRV_Boost head_layer, model_nerv.py:41,56-57): HBM-bound
// CUDA-core kernel.  Through the tensor-core path this launch is a chain of tiny N = 16 UMMAs bound by per-tile
// latency (NeRV-S 720p: 43 us, E-NeRV-M 1080p: 88 us); here one thread owns one pixel, reads its Cin_p/8 16-byte
// chunks (coalesced across the pixels of a warp), multiplies by the f32 weights held in shared memory and writes the
// NCHW f32 pixels - the read of the last activation map is all that is left.
// ---------------------------------------------------------------------------------------------
constexpr int HEAD1_MAX_COUT = 4;

__global__ void head1x1_kernel(const uint4* __restrict__ x, int B, int Cin, int cin_p, size_t hw, const float* __restrict__ w,
                               const float* __restrict__ bias, int Cout, int act, float* __restrict__ out) {
    extern __shared__ __align__(16) float sw[];   // [Cout][cin_p] (zero beyond Cin) | bias[Cout]
    for (int i = threadIdx.x; i < Cout * cin_p; i += blockDim.x) {
        const int c = i / cin_p, k = i - c * cin_p;
        sw[i] = (k < Cin) ? w[c * Cin + k] : 0.0f;
    }
    if (threadIdx.x < Cout) sw[Cout * cin_p + threadIdx.x] = bias ? bias[threadIdx.x] : 0.0f;
    __syncthreads();
    const int groups = cin_p >> 3;
    const size_t total = static_cast<size_t>(B) * hw;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t b = idx / hw, p = idx - b * hw;
        float acc[HEAD1_MAX_COUT];
#pragma unroll
        for (int c = 0; c < HEAD1_MAX_COUT; ++c) acc[c] = (c < Cout) ? sw[Cout * cin_p + c] : 0.0f;
        for (int g = 0; g < groups; ++g) {
            const uint4 u = x[(b * groups + g) * hw + p];
            const float2 a0 = unpack_h2(u.x), a1 = unpack_h2(u.y), a2 = unpack_h2(u.z), a3 = unpack_h2(u.w);
            const float v[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
            for (int c = 0; c < HEAD1_MAX_COUT; ++c) {
                if (c < Cout) {
                    // two 16-byte broadcast loads per (channel, group) instead of eight scalar ones: the kernel was bound
                    // by shared-memory instruction issue (ncu: issue slots 66 % busy at 40 % of the DRAM peak)
                    const float4 w0 = *reinterpret_cast<const float4*>(sw + c * cin_p + g * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(sw + c * cin_p + g * 8 + 4);
                    acc[c] = fmaf(v[0], w0.x, acc[c]); acc[c] = fmaf(v[1], w0.y, acc[c]);
                    acc[c] = fmaf(v[2], w0.z, acc[c]); acc[c] = fmaf(v[3], w0.w, acc[c]);
                    acc[c] = fmaf(v[4], w1.x, acc[c]); acc[c] = fmaf(v[5], w1.y, acc[c]);
                    acc[c] = fmaf(v[6], w1.z, acc[c]); acc[c] = fmaf(v[7], w1.w, acc[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < HEAD1_MAX_COUT; ++c)
            if (c < Cout) out[(b * Cout + c) * hw + p] = apply_act(acc[c], act);
    }
}

static int grid_for(size_t total, int block) {
    size_t g = (total + block - 1) / block;
    const size_t cap = 148 * 32;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace bnerv

using namespace bnerv;

extern "C" int bnerv_pack_conv_weight(const float* w_oihw, const float* bias, int Cout, int Cin, int k, int s,
                                      void* w_packed, float* bias_packed, void* stream) {
    if (!w_oihw || !w_packed || !bias_packed) return set_error(BNERV_E_BADARG, "pack_conv_weight: null pointer");
    if (Cout <= 0 || Cin <= 0 || s <= 0) return set_error(BNERV_E_BADARG, "pack_conv_weight: non-positive size");
    if (k != 1 && k != 3) return set_error(BNERV_E_UNSUPPORTED, "pack_conv_weight: kernel size %d (only 1 and 3)", k);
    const int cout_p = round_up(Cout, 16), cin_p = round_up(Cin, 16);
    const size_t total = bnerv_packed_weight_numel(Cout, Cin, k, s);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    pack_weight_kernel<<<grid_for(total, 256), 256, 0, st>>>(w_oihw, Cout, Cin, k, s, cout_p, cin_p, static_cast<__half*>(w_packed));
    int rc = check_launch("pack_weight_kernel");
    if (rc) return rc;
    pack_bias_kernel<<<grid_for(static_cast<size_t>(s) * s * cout_p, 256), 256, 0, st>>>(bias, Cout, s, cout_p, bias_packed);
    return check_launch("pack_bias_kernel");
}

template <typename CodeT>
static int pack_q(const void* wq, const float* w_scale, int w_pc, const void* bq, const float* b_scale, int b_pc, int Cout,
                  int Cin, int k, int s, void* w_packed, float* bias_packed, cudaStream_t st) {
    const int cout_p = round_up(Cout, 16), cin_p = round_up(Cin, 16);
    const size_t total = bnerv_packed_weight_numel(Cout, Cin, k, s);
    pack_weight_q_kernel<CodeT><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const CodeT*>(wq), w_scale, w_pc, Cout, Cin, k, s,
                                                                      cout_p, cin_p, static_cast<__half*>(w_packed));
    int rc = check_launch("pack_weight_q_kernel");
    if (rc) return rc;
    pack_bias_q_kernel<CodeT><<<grid_for(static_cast<size_t>(s) * s * cout_p, 256), 256, 0, st>>>(
        static_cast<const CodeT*>(bq), b_scale, b_pc, Cout, s, cout_p, bias_packed);
    return check_launch("pack_bias_q_kernel");
}

extern "C" int bnerv_pack_conv_weight_q(const void* w_codes, const float* w_scale, int w_scale_per_channel,
                                        const void* b_codes, const float* b_scale, int b_scale_per_channel, int code_bytes,
                                        int Cout, int Cin, int k, int s, void* w_packed, float* bias_packed, void* stream) {
    if (!w_codes || !w_scale || !w_packed || !bias_packed) return set_error(BNERV_E_BADARG, "pack_conv_weight_q: null pointer");
    if (b_codes && !b_scale) return set_error(BNERV_E_BADARG, "pack_conv_weight_q: bias codes without a bias scale");
    if (Cout <= 0 || Cin <= 0 || s <= 0) return set_error(BNERV_E_BADARG, "pack_conv_weight_q: non-positive size");
    if (k != 1 && k != 3) return set_error(BNERV_E_UNSUPPORTED, "pack_conv_weight_q: kernel size %d (only 1 and 3)", k);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (code_bytes) {
        case 1: return pack_q<int8_t>(w_codes, w_scale, w_scale_per_channel, b_codes, b_scale, b_scale_per_channel, Cout, Cin, k, s, w_packed, bias_packed, st);
        case 2: return pack_q<int16_t>(w_codes, w_scale, w_scale_per_channel, b_codes, b_scale, b_scale_per_channel, Cout, Cin, k, s, w_packed, bias_packed, st);
        case 4: return pack_q<int32_t>(w_codes, w_scale, w_scale_per_channel, b_codes, b_scale, b_scale_per_channel, Cout, Cin, k, s, w_packed, bias_packed, st);
        default: return set_error(BNERV_E_UNSUPPORTED, "pack_conv_weight_q: code_bytes %d (1, 2 or 4)", code_bytes);
    }
}

extern "C" int bnerv_pack_head_weight(const float* w_oihw, int Cout, int Cin, void* w_head_packed, void* stream) {
    if (!w_oihw || !w_head_packed) return set_error(BNERV_E_BADARG, "pack_head_weight: null pointer");
    if (Cout <= 0 || Cin <= 0) return set_error(BNERV_E_BADARG, "pack_head_weight: non-positive size");
    if (Cout > 3) return set_error(BNERV_E_UNSUPPORTED, "pack_head_weight: Cout = %d (at most 3)", Cout);
    const int cin_p = round_up(Cin, 16);
    pack_head_weight_kernel<<<grid_for(static_cast<size_t>(cin_p) * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w_oihw, Cout, Cin, cin_p, static_cast<__half*>(w_head_packed));
    return check_launch("pack_head_weight_kernel");
}

extern "C" int bnerv_head_conv1(const void* x, int B, int Cin, int H, int W, const float* w_oihw, const float* bias, int Cout,
                                int act, float* out_nchw, void* stream) {
    if (!x || !w_oihw || !out_nchw) return set_error(BNERV_E_BADARG, "head_conv1: null pointer");
    if (B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || Cout <= 0) return set_error(BNERV_E_BADARG, "head_conv1: non-positive size");
    if (Cout > HEAD1_MAX_COUT) return set_error(BNERV_E_UNSUPPORTED, "head_conv1: Cout = %d (at most %d)", Cout, HEAD1_MAX_COUT);
    if (act < BNERV_ACT_NONE || act > BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "head_conv1: act %d", act);
    const int cin_p = round_up(Cin, 16);
    const size_t smem = (static_cast<size_t>(Cout) * cin_p + Cout) * sizeof(float);
    if (smem > 48 * 1024) return set_error(BNERV_E_UNSUPPORTED, "head_conv1: Cin = %d too wide", Cin);
    const size_t hw = static_cast<size_t>(H) * W;
    head1x1_kernel<<<grid_for(static_cast<size_t>(B) * hw, 256), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(x), B, Cin, cin_p, hw, w_oihw, bias, Cout, act, out_nchw);
    return check_launch("head1x1_kernel");
}

extern "C" int bnerv_nchw_to_c8(const float* x, int B, int C, int H, int W, void* y_c8, void* stream) {
    if (!x || !y_c8) return set_error(BNERV_E_BADARG, "nchw_to_c8: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "nchw_to_c8: non-positive size");
    const int cp = round_up(C, 16);
    const size_t total = static_cast<size_t>(B) * (cp / 8) * H * W;
    nchw_to_c8_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, B, C, H, W, cp, static_cast<__half*>(y_c8));
    return check_launch("nchw_to_c8_kernel");
}

extern "C" int bnerv_c8_to_nchw(const void* x_c8, int B, int C, int H, int W, float* y, void* stream) {
    if (!x_c8 || !y) return set_error(BNERV_E_BADARG, "c8_to_nchw: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "c8_to_nchw: non-positive size");
    const int cp = round_up(C, 16);
    const size_t total = static_cast<size_t>(B) * (cp / 8) * H * W;
    c8_to_nchw_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(x_c8), B, C, H, W, cp, y);
    return check_launch("c8_to_nchw_kernel");
}

extern "C" int bnerv_pixel_shuffle(const float* x, int B, int C, int H, int W, int s, float* y, void* stream) {
    if (!x || !y) return set_error(BNERV_E_BADARG, "pixel_shuffle: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || s <= 0) return set_error(BNERV_E_BADARG, "pixel_shuffle: non-positive size");
    const size_t total = static_cast<size_t>(B) * C * H * W * s * s;
    pixel_shuffle_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, B, C, H, W, s, y);
    return check_launch("pixel_shuffle_kernel");
}

extern "C" int bnerv_sft_affine(const bnerv_sft_layer* layers_dev, int n_layers, const float* e, int B, int ch_t, void* stream) {
    if (!layers_dev || !e) return set_error(BNERV_E_BADARG, "sft_affine: null pointer");
    if (n_layers <= 0 || B <= 0 || ch_t <= 0) return set_error(BNERV_E_BADARG, "sft_affine: non-positive size");
    if (ch_t > 1024) return set_error(BNERV_E_UNSUPPORTED, "sft_affine: ch_t %d > 1024", ch_t);
    dim3 grid(n_layers, B);
    sft_affine_kernel<<<grid, 128, 3 * ch_t * sizeof(float), static_cast<cudaStream_t>(stream)>>>(layers_dev, e, ch_t);
    return check_launch("sft_affine_kernel");
}

extern "C" int bnerv_linear_act(const float* x, int B, int Cin, const float* w, const float* bias, int Cout, int act,
                                float* y, void* stream) {
    if (!x || !w || !y) return set_error(BNERV_E_BADARG, "linear_act: null pointer");
    if (B <= 0 || Cin <= 0 || Cout <= 0) return set_error(BNERV_E_BADARG, "linear_act: non-positive size");
    const size_t smem = static_cast<size_t>(B) * Cin * sizeof(float);
    if (smem > 48 * 1024) return set_error(BNERV_E_UNSUPPORTED, "linear_act: B*Cin = %d floats exceeds 48 KB of shared memory", B * Cin);
    const int warps = 8;
    int grid = (Cout + warps - 1) / warps;
    if (grid > 148 * 8) grid = 148 * 8;
    linear_act_kernel<<<grid, warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(x, B, Cin, w, bias, Cout, act, y);
    return check_launch("linear_act_kernel");
}

static int linear_pair_launch(bool pe, const bnerv_linear_problem* probs, int B, const float* t, const float* bases, int levels,
                              void* stream) {
    if (!probs) return set_error(BNERV_E_BADARG, "linear_pair: null problems");
    if (B <= 0) return set_error(BNERV_E_BADARG, "linear_pair: non-positive batch");
    LinearProblem p[2];
    int blocks[2], cin_max = 0;
    for (int k = 0; k < 2; ++k) {
        const bnerv_linear_problem& q = probs[k];
        if (!q.w || (!q.y && !q.y_c8) || (!pe && !q.x)) return set_error(BNERV_E_BADARG, "linear_pair: null pointer in problem %d", k);
        if (q.Cin <= 0 || q.Cout <= 0) return set_error(BNERV_E_BADARG, "linear_pair: non-positive size in problem %d", k);
        if (pe && q.Cin != 2 * levels) return set_error(BNERV_E_BADARG, "linear_pair: Cin = %d but the position encoding has %d entries", q.Cin, 2 * levels);
        if (q.act < BNERV_ACT_NONE || q.act > BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "linear_pair: act %d", q.act);
        if (q.y_c8 && (q.hw <= 0 || q.Cout % q.hw != 0)) return set_error(BNERV_E_BADARG, "linear_pair: C8 output needs Cout = C * hw");
        p[k].x = q.x; p[k].w = q.w; p[k].bias = q.bias; p[k].y = q.y; p[k].y_c8 = static_cast<__half*>(q.y_c8);
        p[k].Cin = q.Cin; p[k].Cout = q.Cout; p[k].act = q.act; p[k].hw = q.y_c8 ? q.hw : 1;
        p[k].cp = q.y_c8 ? round_up(q.Cout / q.hw, 16) : 16;
        blocks[k] = std::min((q.Cout + 7) / 8, 148 * 8);
        cin_max = std::max(cin_max, q.Cin);
    }
    const size_t smem = static_cast<size_t>(B) * cin_max * sizeof(float);
    if (smem > 48 * 1024) return set_error(BNERV_E_UNSUPPORTED, "linear_pair: B*Cin = %d floats exceeds 48 KB of shared memory", B * cin_max);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pe) linear_pair_kernel<true><<<blocks[0] + blocks[1], 256, smem, st>>>(p[0], p[1], B, t, bases, levels, blocks[0]);
    else    linear_pair_kernel<false><<<blocks[0] + blocks[1], 256, smem, st>>>(p[0], p[1], B, nullptr, nullptr, 0, blocks[0]);
    return check_launch("linear_pair_kernel");
}

extern "C" int bnerv_pe_linear_pair(const float* t, int B, const float* bases, int levels, const bnerv_linear_problem* probs,
                                    void* stream) {
    if (!t || !bases || levels <= 0) return set_error(BNERV_E_BADARG, "pe_linear_pair: null / empty position encoding");
    return linear_pair_launch(true, probs, B, t, bases, levels, stream);
}

extern "C" int bnerv_linear_pair(const bnerv_linear_problem* probs, int B, void* stream) {
    return linear_pair_launch(false, probs, B, nullptr, nullptr, 0, stream);
}

extern "C" int bnerv_conv_fused_f32(const float* x, int B, int Cin, int H, int W, const float* w, const float* bias,
                                    int Cout, int k, int s, int act, const float* resid, const float* g1p,
                                    const float* beta, int ldg, float* out_pre, float* out_aff, void* stream) {
    if (!x || !w) return set_error(BNERV_E_BADARG, "conv_fused_f32: null operand");
    if (B <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || s <= 0) return set_error(BNERV_E_BADARG, "conv_fused_f32: non-positive size");
    if (k != 1 && k != 3) return set_error(BNERV_E_UNSUPPORTED, "conv_fused_f32: kernel size %d (only 1 and 3)", k);
    if ((g1p == nullptr) != (beta == nullptr) || (g1p != nullptr) != (out_aff != nullptr))
        return set_error(BNERV_E_BADARG, "conv_fused_f32: g1p, beta and out_aff go together");
    if (!out_pre && !out_aff) return set_error(BNERV_E_BADARG, "conv_fused_f32: no output");
    if (act < BNERV_ACT_NONE || act > BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "conv_fused_f32: act %d", act);
    const size_t total = static_cast<size_t>(B) * ((Cout * s * s + CO_T - 1) / CO_T) * H * W;
    conv_f32_kernel<<<grid_for(total, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x, B, Cin, H, W, w, bias, Cout, k, s, act, resid, g1p, beta, ldg, out_pre, out_aff);
    return check_launch("conv_f32_kernel");
}
