// Backward-pass kernels around the tensor-core convs (HBM-bound element-wise work + per-channel reductions):
// the transposes of OutImg / SFTLayer / ResBlock_SFT / NeRVBlock element-wise steps (model_blocks.py:57-63,
// 74-105, 34-46) as torch.autograd would apply them for train_nerv_all.py:342-348.
//
// Gradient maps are C8 f16 like the activations, multiplied by one power-of-two loss scale S chosen on the
// device from max|dL/dz_head| (f16 has 5 exponent bits; gradients of a mean over 6e6 pixels are ~1e-7).  Every
// f32 reduction below is therefore S * (true gradient); the caller multiplies by scale[1] = 1/S.
#include "common.cuh"

namespace bnerv {

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    const float2 a0 = unpack_h2(u.x), a1 = unpack_h2(u.y), a2 = unpack_h2(u.z), a3 = unpack_h2(u.w);
    v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y; v[4] = a2.x; v[5] = a2.y; v[6] = a3.x; v[7] = a3.y;
}
__device__ __forceinline__ uint4 pack8f(const float (&v)[8]) {
    uint4 o;
    o.x = pack_h2_sat(v[0], v[1]); o.y = pack_h2_sat(v[2], v[3]);
    o.z = pack_h2_sat(v[4], v[5]); o.w = pack_h2_sat(v[6], v[7]);
    return o;
}

// Block-wide sum of N_ACC per-thread values, result atomically added to dst[i] by one thread per value.
template <int N_ACC>
__device__ __forceinline__ void block_reduce_atomic(float (&acc)[N_ACC], float* const (&dst)[N_ACC]) {
    __shared__ float red[N_ACC][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N_ACC; ++i) {
        float v = acc[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[i][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < N_ACC) {
        float v = 0.0f;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
        if (dst[threadIdx.x] != nullptr) atomicAdd(dst[threadIdx.x], v);
    }
}

constexpr int BW_THREADS = 256;     // 8 warps (block_reduce_atomic assumes <= 8)

// Gradient-range monitor.  Gradient maps are f16 behind one loss scale picked at the head, and every conversion saturates
// (cvt.satfinite): a gradient that grows past 65504/S deeper in the cascade would be clipped silently.  The two element-wise
// kernels every block's gradients pass through (inputs = the dgrad conv outputs, outputs = the next dgrad's input) OR bits into
// status[0] when they see |value| >= 65504 (bit 0: saturated, i.e. clipped here or by the conv that produced it) or a
// non-finite value (bit 1).  The pointer is set per process by bnerv_bwd_set_status (NULL: monitoring off).
// status[1] (float bits): the largest |scaled gradient| these kernels saw since the last bnerv_head_bwd - the feedback of the loss-scale
// controller (scale_ctrl_kernel): gradients of the low-resolution stages GROW relative to the head's while a model trains (15M
// HNeRV: from 1.6x to 12 000x the head's maximum within 40 Adam steps, tools/grad_range_probe.py), so a scale fixed from the head's
// maximum alone ends in saturation.  status[2] (float bits): the controller's current target for S * max|dL/dz_head|.
static int* g_bwd_status = nullptr;

__device__ __forceinline__ void bwd_flag(int* status, float amax) {
    if (status == nullptr) return;
    float wmax = amax;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
    if ((threadIdx.x & 31) == 0 && wmax > 0.0f && wmax <= 3.0e38f) atomicMax(status + 1, __float_as_int(wmax));
    if (!(amax < 65504.0f)) {
        const int bit = (amax <= 3.0e38f) ? 1 : 2;            // NaN / inf fail this comparison as well
        const int any = __reduce_or_sync(__activemask(), bit);
        if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1)) atomicOr(status, any);
    }
}

// One thread, before every head_bwd_kernel: target <- target adjusted by the previous backward's largest scaled gradient so that it
// stays within [2^10, 2^14] (two octaves of margin below 65504 for step-to-step growth; the high-resolution maps, ~1e-4 of the
// largest, remain f16 normals), never above 8 (13 bits of headroom when nothing is known yet).
__global__ void scale_ctrl_kernel(int* status) {
    float target = __int_as_float(status[2]);
    if (!(target > 0.0f)) target = 8.0f;
    const float last = __int_as_float(status[1]);
    if (last > 16384.0f) target *= exp2f(-ceilf(log2f(last / 4096.0f)));
    else if (last > 0.0f && last < 1024.0f && target < 8.0f) target = fminf(8.0f, target * 2.0f);
    target = fmaxf(target, 1.0e-30f);
    status[2] = __float_as_int(target);
    status[1] = 0;
}

// ---------------------------------------------------------------------------------------------
// head: dz = S * dimg * d/dz(0.5 tanh z + 0.5) = S * dimg * 2 img (1 - img)
// ---------------------------------------------------------------------------------------------
__global__ void head_absmax_kernel(const float* __restrict__ dimg, const float* __restrict__ img, size_t n, float* amax) {
    float m = 0.0f;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float o = img[i];
        m = fmaxf(m, fabsf(dimg[i] * 2.0f * o * (1.0f - o)));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));   // non-negative floats order like ints
}

__device__ __forceinline__ float scale_from_amax(float amax, float target) {
    // largest power of two S with S * amax <= 8: 13 bits of headroom to the f16 maximum for gradients that GROW on their way
    // down the cascade (at full size they do: with S * amax <= 64 the 15M HNeRV saturated within 300 Adam steps - the
    // gradient-range monitor, bnerv_bwd_set_status, caught it), 27 bits above the smallest subnormal for those that shrink
    if (!(amax > 0.0f) || !isfinite(amax)) return 1.0f;
    float e = floorf(log2f(target / amax));
    e = fminf(fmaxf(e, -100.0f), 100.0f);
    return exp2f(e);
}

__global__ void head_bwd_kernel(const float* __restrict__ dimg, const float* __restrict__ img, int B, int C, int H, int W,
                                const float* __restrict__ amax, float* __restrict__ scale, __half* __restrict__ dz, const int* status) {
    const float S = scale_from_amax(*amax, status ? __int_as_float(status[2]) : 8.0f);
    if (blockIdx.x == 0 && threadIdx.x == 0) { scale[0] = S; scale[1] = 1.0f / S; }
    const int cp = (C + 15) / 16 * 16;
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
            v[k] = 0.0f;
            if (c < C) {
                const size_t off = (static_cast<size_t>(b) * C + c) * hw + p;
                const float o = img[off];
                v[k] = S * dimg[off] * 2.0f * o * (1.0f - o);
            }
        }
        reinterpret_cast<uint4*>(dz)[idx] = pack8f(v);
    }
}

// ---------------------------------------------------------------------------------------------
// per-channel sums of a C8 map: out[(per_b ? b : 0)][c] += sum_{(b,)h,w} x[b,c,h,w]
// grid (chunks, Cp/8, B)
// ---------------------------------------------------------------------------------------------
__global__ void channel_sum_kernel(const __half* __restrict__ x, int cp, size_t hw, int per_b, float* __restrict__ out) {
    const int g = blockIdx.y, b = blockIdx.z;
    const uint4* plane = reinterpret_cast<const uint4*>(x) + (static_cast<size_t>(b) * (cp >> 3) + g) * hw;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < hw; p += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float v[8];
        unpack8(plane[p], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
    float* o = out + (per_b ? static_cast<size_t>(b) * cp : 0) + g * 8;
    float* const dst[8] = {o, o + 1, o + 2, o + 3, o + 4, o + 5, o + 6, o + 7};
    block_reduce_atomic<8>(acc, dst);
}

// ---------------------------------------------------------------------------------------------
// ResBlock middle (model_blocks.py:86-87 transposed).  Forward: v = gelu(c0); w = v*g1p + beta1.
//   given dw = dL/dw (output of conv1's dgrad):  dc0 = dw * g1p * gelu'(c0)
//   dG1[b][c] += sum dw*v ; dB1[b][c] += sum dw ; db0[c] += sum dc0
// ---------------------------------------------------------------------------------------------
__global__ void mid_bwd_kernel(const __half* __restrict__ dw, const __half* __restrict__ v, const __half* __restrict__ dact,
                               const float* __restrict__ g1p, int cp, size_t hw, __half* __restrict__ dc0,
                               float* __restrict__ dG, float* __restrict__ dB, float* __restrict__ dbias, int* status) {
    float amax = 0.0f;
    const int g = blockIdx.y, b = blockIdx.z;
    const size_t base = (static_cast<size_t>(b) * (cp >> 3) + g) * hw;
    float gg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) gg[k] = g1p[static_cast<size_t>(b) * cp + g * 8 + k];
    float acc[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) acc[k] = 0.0f;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;        // two pixels per iteration, see front_bwd_kernel
    for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < hw; p += 2 * stride) {
        const size_t p2 = p + stride;
        const bool two = p2 < hw;
        const size_t q = two ? p2 : p;
        const uint4 ua0 = reinterpret_cast<const uint4*>(dw)[base + p], uv0 = reinterpret_cast<const uint4*>(v)[base + p];
        const uint4 ud0 = reinterpret_cast<const uint4*>(dact)[base + p];
        const uint4 ua1 = reinterpret_cast<const uint4*>(dw)[base + q], uv1 = reinterpret_cast<const uint4*>(v)[base + q];
        const uint4 ud1 = reinterpret_cast<const uint4*>(dact)[base + q];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            if (it == 1 && !two) break;
            float a[8], vv[8], d[8], o[8];
            unpack8(it ? ua1 : ua0, a);
            unpack8(it ? uv1 : uv0, vv);
            unpack8(it ? ud1 : ud0, d);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                o[k] = a[k] * gg[k] * d[k];
                acc[k] += a[k] * vv[k];
                acc[8 + k] += a[k];
                acc[16 + k] += o[k];
                amax = fmaxf(amax, fmaxf(fabsf(a[k]), fabsf(o[k])));
                if (o[k] != o[k]) amax = __int_as_float(0x7fc00000);
            }
            reinterpret_cast<uint4*>(dc0)[base + (it ? p2 : p)] = pack8f(o);
        }
    }
    float* const pG = dG + static_cast<size_t>(b) * cp + g * 8;
    float* const pB = dB + static_cast<size_t>(b) * cp + g * 8;
    float* const pb = dbias + g * 8;
    float* const dst[24] = {pG, pG + 1, pG + 2, pG + 3, pG + 4, pG + 5, pG + 6, pG + 7,
                            pB, pB + 1, pB + 2, pB + 3, pB + 4, pB + 5, pB + 6, pB + 7,
                            pb, pb + 1, pb + 2, pb + 3, pb + 4, pb + 5, pb + 6, pb + 7};
    block_reduce_atomic<24>(acc, dst);
    bwd_flag(status, amax);
}

// ---------------------------------------------------------------------------------------------
// Block front (model_blocks.py:37,85,89 transposed).  Forward: x0 = act(y); u = x0*g0p + beta0; out = x0 + conv1(..).
//   given dout = dL/dout and du = dL/du (output of conv0's dgrad):
//   dx0 = dout + du*g0p ; dy = dx0 * act'(y) ; dG0[b][c] += sum du*x0 ; dB0[b][c] += sum du ; db1[c] += sum dout
// ---------------------------------------------------------------------------------------------
__global__ void front_bwd_kernel(const __half* __restrict__ du, const __half* __restrict__ dout, const __half* __restrict__ x0,
                                 const __half* __restrict__ dact, const float* __restrict__ g0p, int cp, size_t hw,
                                 __half* __restrict__ dy, float* __restrict__ dG, float* __restrict__ dB,
                                 float* __restrict__ dbias1, float* __restrict__ dbias_up, int* status) {
    float amax = 0.0f;
    const int g = blockIdx.y, b = blockIdx.z;
    const size_t base = (static_cast<size_t>(b) * (cp >> 3) + g) * hw;
    float gg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) gg[k] = g0p[static_cast<size_t>(b) * cp + g * 8 + k];
    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.0f;
    // two pixels per iteration: all eight 16-byte loads are in flight before the first is consumed (the one-pixel loop
    // left ~50 KB per SM in flight, about the HBM latency-bandwidth product: 60 % of the copy bandwidth)
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < hw; p += 2 * stride) {
        const size_t p2 = p + stride;
        const bool two = p2 < hw;
        const size_t q = two ? p2 : p;
        const uint4 ua0 = reinterpret_cast<const uint4*>(du)[base + p], ue0 = reinterpret_cast<const uint4*>(dout)[base + p];
        const uint4 ux0 = reinterpret_cast<const uint4*>(x0)[base + p], ud0 = reinterpret_cast<const uint4*>(dact)[base + p];
        const uint4 ua1 = reinterpret_cast<const uint4*>(du)[base + q], ue1 = reinterpret_cast<const uint4*>(dout)[base + q];
        const uint4 ux1 = reinterpret_cast<const uint4*>(x0)[base + q], ud1 = reinterpret_cast<const uint4*>(dact)[base + q];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            if (it == 1 && !two) break;
            float a[8], e[8], xx[8], d[8], o[8];
            unpack8(it ? ua1 : ua0, a);
            unpack8(it ? ue1 : ue0, e);
            unpack8(it ? ux1 : ux0, xx);
            unpack8(it ? ud1 : ud0, d);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                o[k] = (e[k] + a[k] * gg[k]) * d[k];
                acc[k] += a[k] * xx[k];
                acc[8 + k] += a[k];
                acc[16 + k] += e[k];
                acc[24 + k] += o[k];
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(a[k]), fabsf(e[k])), fabsf(o[k])));
                if (o[k] != o[k]) amax = __int_as_float(0x7fc00000);
            }
            reinterpret_cast<uint4*>(dy)[base + (it ? p2 : p)] = pack8f(o);
        }
    }
    float* const pG = dG + static_cast<size_t>(b) * cp + g * 8;
    float* const pB = dB + static_cast<size_t>(b) * cp + g * 8;
    float* const pb = dbias1 + g * 8;
    float* const pu = dbias_up ? dbias_up + g * 8 : nullptr;       // s == 1 up-conv: its bias gradient is the channel sum of dy
    float* const dst[32] = {pG, pG + 1, pG + 2, pG + 3, pG + 4, pG + 5, pG + 6, pG + 7,
                            pB, pB + 1, pB + 2, pB + 3, pB + 4, pB + 5, pB + 6, pB + 7,
                            pb, pb + 1, pb + 2, pb + 3, pb + 4, pb + 5, pb + 6, pb + 7,
                            pu, pu ? pu + 1 : pu, pu ? pu + 2 : pu, pu ? pu + 3 : pu, pu ? pu + 4 : pu, pu ? pu + 5 : pu,
                            pu ? pu + 6 : pu, pu ? pu + 7 : pu};
    block_reduce_atomic<32>(acc, dst);
    bwd_flag(status, amax);
}

// ---------------------------------------------------------------------------------------------
// PixelShuffle transposed on C8 maps: src [B][Cp/8][H*s][W*s][8] -> dst [B][s*s*Cp/8][H][W][8] with the
// un-shuffled channel order m = (i*s + j)*Cp + c  (the K order bnerv_pack_conv_weight_dgrad and bnerv_conv_wgrad use)
// ---------------------------------------------------------------------------------------------
__global__ void unshuffle_c8_kernel(const uint4* __restrict__ src, int B, int cp, int H, int W, int s, uint4* __restrict__ dst) {
    const int g8 = cp >> 3;
    const size_t total = static_cast<size_t>(B) * s * s * g8 * H * W;
    const int Wo = W * s;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int w = idx % W;
        size_t r = idx / W;
        const int h = r % H;
        r /= H;
        const int gd = r % (s * s * g8);
        const int b = r / (s * s * g8);
        const int ij = gd / g8, g = gd - ij * g8;
        const int i = ij / s, j = ij - i * s;
        dst[idx] = src[((static_cast<size_t>(b) * g8 + g) * (static_cast<size_t>(H) * s) + (h * s + i)) * Wo + (w * s + j)];
    }
}

// Same permutation with a thread per LOW-resolution pixel: the s x s source chunks of one pixel are s runs of s*16
// contiguous bytes (coalesced across neighbouring pixels), every destination plane is written densely, and the channel
// sums of the un-shuffled map (= the up-conv's bias gradient) come out of the same pass.  grid (chunks, Cp/8, B).
template <int S>
__global__ void unshuffle_sum_kernel(const uint4* __restrict__ src, int cp, int H, int W, uint4* __restrict__ dst,
                                     float* __restrict__ sums) {
    const int g = blockIdx.y, b = blockIdx.z, g8 = cp >> 3;
    const size_t hw = static_cast<size_t>(H) * W;
    const int Wo = W * S;
    const uint4* sp = src + (static_cast<size_t>(b) * g8 + g) * hw * S * S;
    float acc[S * S * 8];
#pragma unroll
    for (int k = 0; k < S * S * 8; ++k) acc[k] = 0.0f;
    for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < hw; p += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int h = static_cast<int>(p / W), w = static_cast<int>(p - static_cast<size_t>(h) * W);
#pragma unroll
        for (int i = 0; i < S; ++i) {
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const uint4 v = sp[static_cast<size_t>(h * S + i) * Wo + (w * S + j)];
                dst[(static_cast<size_t>(b) * S * S * g8 + (i * S + j) * g8 + g) * hw + p] = v;
                float f[8];
                unpack8(v, f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[(i * S + j) * 8 + k] += f[k];
            }
        }
    }
    float* dptr[S * S * 8];
#pragma unroll
    for (int ij = 0; ij < S * S; ++ij)
#pragma unroll
        for (int k = 0; k < 8; ++k) dptr[ij * 8 + k] = sums + ij * cp + g * 8 + k;
    float* const (&dref)[S * S * 8] = dptr;
    block_reduce_atomic<S * S * 8>(acc, dref);
}

// packed weight of the dgrad conv (see bnerv_pack_conv_weight_dgrad)
__global__ void pack_weight_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int s, int cout_p, int cin_p,
                                         __half* __restrict__ wp) {
    const int kp = s * s * cout_p;           // K of the dgrad conv
    const int taps = k * k;
    const size_t total = static_cast<size_t>(taps) * kp * cin_p;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int kk = idx & 7;
        size_t r = idx >> 3;
        const int n = r % cin_p;             // output row of the dgrad conv = input channel of the forward conv
        r /= cin_p;
        const int kg = r % (kp >> 3);
        const int tap = r / (kp >> 3);
        const int km = kg * 8 + kk;          // un-shuffled gradient channel
        const int ij = km / cout_p, c = km - ij * cout_p;
        float v = 0.0f;
        if (c < Cout && n < Cin) {
            const int o = c * s * s + ij;
            v = w[(static_cast<size_t>(o) * Cin + n) * taps + (taps - 1 - tap)];      // taps flipped in both directions
        }
        wp[idx] = __float2half_rn(v);
    }
}

// un-shuffled channel sums [s*s*Cout_p] (scaled) -> bias gradient [Cout*s*s] in the reference order
__global__ void bias_finalize_kernel(const float* __restrict__ acc, int Cout, int s, int cout_p, const float* __restrict__ inv_scale,
                                     int accumulate, float* __restrict__ grad) {
    const float sc = inv_scale ? *inv_scale : 1.0f;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < Cout * s * s; o += gridDim.x * blockDim.x) {
        const int c = o / (s * s), ij = o - c * s * s;
        const float v = acc[ij * cout_p + c] * sc;
        grad[o] = accumulate ? grad[o] + v : v;
    }
}

// ---------------------------------------------------------------------------------------------
// per-frame error sums for the metrics / pixel losses (hnerv_utils.py:338-341,400-403): deterministic two-stage
// reduction (fixed block order, f64 partials), so PSNR does not depend on atomics ordering
// ---------------------------------------------------------------------------------------------
constexpr int FM_BLOCKS = 296;      // partial sums per frame (2 per SM)

__global__ void frame_err_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, double* __restrict__ part) {
    const float* pa = a + static_cast<size_t>(blockIdx.y) * n;
    const float* pb = b + static_cast<size_t>(blockIdx.y) * n;
    float sq = 0.0f, ab = 0.0f;
    double dsq = 0.0, dab = 0.0;
    int cnt = 0;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float d = pa[i] - pb[i];
        sq = fmaf(d, d, sq);
        ab += fabsf(d);
        if (++cnt == 64) { dsq += sq; dab += ab; sq = ab = 0.0f; cnt = 0; }      // bound the f32 running sums
    }
    dsq += sq; dab += ab;
    __shared__ double red[2][BW_THREADS];
    red[0][threadIdx.x] = dsq; red[1][threadIdx.x] = dab;
    __syncthreads();
    for (int off = BW_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) { red[0][threadIdx.x] += red[0][threadIdx.x + off]; red[1][threadIdx.x] += red[1][threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 2 + 0] = red[0][0];
        part[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 2 + 1] = red[1][0];
    }
}

__global__ void frame_err_final_kernel(const double* __restrict__ part, int nblk, size_t n, float* __restrict__ out) {
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    double sq = 0.0, ab = 0.0;
    for (int i = 0; i < nblk; ++i) { sq += part[(static_cast<size_t>(b) * nblk + i) * 2]; ab += part[(static_cast<size_t>(b) * nblk + i) * 2 + 1]; }
    const double mse = sq / static_cast<double>(n), mae = ab / static_cast<double>(n);
    out[b * 3 + 0] = static_cast<float>(mse);
    out[b * 3 + 1] = static_cast<float>(mae);
    out[b * 3 + 2] = -10.0f * log10f(static_cast<float>(mse) + 1e-9f);      // psnr_fn_single, hnerv_utils.py:400-403 (f32 like torch)
}

static int grid_1d(size_t total, int block) {
    size_t g = (total + block - 1) / block;
    const size_t cap = 148 * 16;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}
// grid of the plane-reduction kernels: enough pixel chunks to fill the machine, never more than the plane has work for
static dim3 grid_planes(int B, int cp, size_t hw) {
    const int planes = B * (cp >> 3);
    size_t chunks = (148 * 8 + planes - 1) / planes;
    const size_t max_chunks = (hw + BW_THREADS * 4 - 1) / (BW_THREADS * 4);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    return dim3(static_cast<unsigned>(chunks), static_cast<unsigned>(cp >> 3), static_cast<unsigned>(B));
}

}  // namespace bnerv

using namespace bnerv;

extern "C" int bnerv_head_bwd(const float* dimg, const float* img, int B, int C, int H, int W, float* amax_scratch,
                              float* scale, void* dz_c8, void* stream) {
    if (!dimg || !img || !amax_scratch || !scale || !dz_c8) return set_error(BNERV_E_BADARG, "head_bwd: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "head_bwd: non-positive size");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(amax_scratch, 0, sizeof(float), st);
    if (e != cudaSuccess) return set_error(static_cast<int>(e), "head_bwd memset: %s", cudaGetErrorString(e));
    const size_t n = static_cast<size_t>(B) * C * H * W;
    head_absmax_kernel<<<grid_1d(n, 256), 256, 0, st>>>(dimg, img, n, amax_scratch);
    int rc = check_launch("head_absmax_kernel");
    if (rc) return rc;
    if (g_bwd_status != nullptr) {                 // loss-scale controller: feedback from the previous backward's gradient range
        scale_ctrl_kernel<<<1, 1, 0, st>>>(g_bwd_status);
        if ((rc = check_launch("scale_ctrl_kernel"))) return rc;
    }
    const size_t total = static_cast<size_t>(B) * (round_up(C, 16) / 8) * H * W;
    head_bwd_kernel<<<grid_1d(total, 256), 256, 0, st>>>(dimg, img, B, C, H, W, amax_scratch, scale, static_cast<__half*>(dz_c8),
                                                          g_bwd_status);
    return check_launch("head_bwd_kernel");
}

extern "C" int bnerv_channel_sum(const void* x_c8, int B, int Cp, int H, int W, int per_b, float* out, void* stream) {
    if (!x_c8 || !out) return set_error(BNERV_E_BADARG, "channel_sum: null pointer");
    if (B <= 0 || Cp <= 0 || (Cp & 7) || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "channel_sum: bad size");
    const size_t hw = static_cast<size_t>(H) * W;
    channel_sum_kernel<<<grid_planes(B, Cp, hw), BW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(x_c8), Cp, hw, per_b, out);
    return check_launch("channel_sum_kernel");
}

extern "C" int bnerv_bwd_set_status(int* status) {
    bnerv::g_bwd_status = status;
    return 0;
}

extern "C" int bnerv_resblock_mid_bwd(const void* dw, const void* v, const void* dact, const float* g1p, int B, int C, int H,
                                      int W, void* dc0, float* dG, float* dB, float* dbias0, void* stream) {
    if (!dw || !v || !dact || !g1p || !dc0 || !dG || !dB || !dbias0) return set_error(BNERV_E_BADARG, "resblock_mid_bwd: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "resblock_mid_bwd: non-positive size");
    const int cp = round_up(C, 16);
    const size_t hw = static_cast<size_t>(H) * W;
    mid_bwd_kernel<<<grid_planes(B, cp, hw), BW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(dw), static_cast<const __half*>(v), static_cast<const __half*>(dact), g1p, cp, hw,
        static_cast<__half*>(dc0), dG, dB, dbias0, g_bwd_status);
    return check_launch("mid_bwd_kernel");
}

extern "C" int bnerv_block_front_bwd(const void* du, const void* dout, const void* x0, const void* dact, const float* g0p,
                                     int B, int C, int H, int W, void* dy, float* dG, float* dB, float* dbias1,
                                     float* dbias_up, void* stream) {
    if (!du || !dout || !x0 || !dact || !g0p || !dy || !dG || !dB || !dbias1) return set_error(BNERV_E_BADARG, "block_front_bwd: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "block_front_bwd: non-positive size");
    const int cp = round_up(C, 16);
    const size_t hw = static_cast<size_t>(H) * W;
    front_bwd_kernel<<<grid_planes(B, cp, hw), BW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(du), static_cast<const __half*>(dout), static_cast<const __half*>(x0),
        static_cast<const __half*>(dact), g0p, cp, hw, static_cast<__half*>(dy), dG, dB, dbias1, dbias_up, g_bwd_status);
    return check_launch("front_bwd_kernel");
}

extern "C" int bnerv_unshuffle_c8(const void* src, int B, int C, int H, int W, int s, void* dst, float* sums, void* stream) {
    if (!src || !dst) return set_error(BNERV_E_BADARG, "unshuffle_c8: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || s <= 0) return set_error(BNERV_E_BADARG, "unshuffle_c8: non-positive size");
    const int cp = round_up(C, 16);
    if (sums != nullptr) {
        const size_t hw = static_cast<size_t>(H) * W;
        const dim3 grid = grid_planes(B, cp, hw);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (s == 2) {
            unshuffle_sum_kernel<2><<<grid, BW_THREADS, 0, st>>>(static_cast<const uint4*>(src), cp, H, W, static_cast<uint4*>(dst), sums);
            return check_launch("unshuffle_sum_kernel<2>");
        }
        if (s == 3) {
            unshuffle_sum_kernel<3><<<grid, BW_THREADS, 0, st>>>(static_cast<const uint4*>(src), cp, H, W, static_cast<uint4*>(dst), sums);
            return check_launch("unshuffle_sum_kernel<3>");
        }
        return set_error(BNERV_E_UNSUPPORTED, "unshuffle_c8: fused channel sums only for s = 2, 3 (got %d); pass sums = NULL and use bnerv_channel_sum", s);
    }
    const size_t total = static_cast<size_t>(B) * s * s * (cp / 8) * H * W;
    unshuffle_c8_kernel<<<grid_1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(src), B, cp, H, W, s, static_cast<uint4*>(dst));
    return check_launch("unshuffle_c8_kernel");
}

extern "C" int bnerv_pack_conv_weight_dgrad(const float* w_oihw, int Cout, int Cin, int k, int s, void* w_packed, void* stream) {
    if (!w_oihw || !w_packed) return set_error(BNERV_E_BADARG, "pack_conv_weight_dgrad: null pointer");
    if (Cout <= 0 || Cin <= 0 || s <= 0) return set_error(BNERV_E_BADARG, "pack_conv_weight_dgrad: non-positive size");
    if (k != 1 && k != 3) return set_error(BNERV_E_UNSUPPORTED, "pack_conv_weight_dgrad: kernel size %d (only 1 and 3)", k);
    const int cout_p = round_up(Cout, 16), cin_p = round_up(Cin, 16);
    const size_t total = static_cast<size_t>(k) * k * s * s * cout_p * cin_p;
    pack_weight_dgrad_kernel<<<grid_1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w_oihw, Cout, Cin, k, s, cout_p, cin_p, static_cast<__half*>(w_packed));
    return check_launch("pack_weight_dgrad_kernel");
}

extern "C" int bnerv_bias_finalize(const float* acc, int Cout, int s, const float* inv_scale, int accumulate, float* grad,
                                   void* stream) {
    if (!acc || !grad) return set_error(BNERV_E_BADARG, "bias_finalize: null pointer");
    if (Cout <= 0 || s <= 0) return set_error(BNERV_E_BADARG, "bias_finalize: non-positive size");
    bias_finalize_kernel<<<grid_1d(static_cast<size_t>(Cout) * s * s, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        acc, Cout, s, round_up(Cout, 16), inv_scale, accumulate, grad);
    return check_launch("bias_finalize_kernel");
}

extern "C" size_t bnerv_frame_metrics_scratch_doubles(int B) { return B > 0 ? static_cast<size_t>(B) * FM_BLOCKS * 2 : 0; }

extern "C" int bnerv_frame_metrics(const float* img, const float* gt, int B, size_t n_per_frame, double* scratch, float* out,
                                   void* stream) {
    if (!img || !gt || !scratch || !out) return set_error(BNERV_E_BADARG, "frame_metrics: null pointer");
    if (B <= 0 || n_per_frame == 0) return set_error(BNERV_E_BADARG, "frame_metrics: non-positive size");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    frame_err_partial_kernel<<<dim3(FM_BLOCKS, B), BW_THREADS, 0, st>>>(img, gt, n_per_frame, scratch);
    int rc = check_launch("frame_err_partial_kernel");
    if (rc) return rc;
    frame_err_final_kernel<<<B, 32, 0, st>>>(scratch, FM_BLOCKS, n_per_frame, out);
    return check_launch("frame_err_final_kernel");
}
