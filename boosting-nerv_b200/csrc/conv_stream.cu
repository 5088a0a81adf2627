// One 3x3 conv (stride 1, no PixelShuffle) with Cin_p == Cout_p in {32, 48} in the ROW-STREAMING form of block_stream32.cu - the
// pipeline of that kernel cut after its first conv - with the full epilogue of bnerv_conv_fused: bias, activation, residual,
// TAT affine, one or two C8 outputs.  For the mid-width layers of E-NeRV-Boost M (43 channels at 540p: conv0 / conv1 of its
// ResBlock_SFT, the 43 -> 43 up-conv), where the generic conv_tc_kernel is bound by the instruction count of its epilogue
// (16 M instructions, 37 us for 100 MB of traffic) and a two-conv streaming kernel does not fit tensor memory (an A row of
// K = 48 is 72 columns).
//
//     TMA warp      : input rows -> shared-memory ring
//     front WGs     : 2; build A(h): lane = pixel column, 3 shifted copies x (CG/2) K steps x 8 TMEM columns
//     MMA warp      : per output row 3 ring rows x 3 shifts x CG/2 K steps MMAs (M = 128, N = 8*CG, K = 16), A from tensor memory
//     epilogue WGs  : 4; D(h) -> bias, act, (+ residual), stores of `pre` and / or `aff = pre*g + beta` to global memory
// Tensor memory: A ring 4 x 12*CG columns + accumulator ring 4 x 8*CG columns (CG = 6: 288 + 192).  "Full" is signalled per
// consumer warpgroup, "empty" per accumulator slot (see block_stream32.cu).
//
// Arithmetic = bnerv_conv_fused's (K steps outermost, taps r*3+sx inside; same epilogue functions): bit-identical results.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

namespace bnerv {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();          // conv_tc.cu

constexpr int CS_WG_F = 2;                    // front warpgroups (rows i = p mod 2: each sees every use of its input / A slots)
constexpr int CS_WG_E = 4;                    // epilogue warpgroups
constexpr int CS_WARP_E = 4 * CS_WG_F, CS_WARP_MMA = CS_WARP_E + 4 * CS_WG_E;     // front | epilogue | MMA issue | TMA
constexpr int CS_THREADS = (CS_WARP_MMA + 2) * 32;
constexpr int CS_NA = 4, CS_ND = 4, CS_NI = 6;
constexpr int CS_VALID = 126;                 // valid output columns per strip: lanes [0, 126) (lane m reads row pixels m .. m+2 of 128)

template <int CG> struct CsCfg {
    static constexpr int NW = 8 * CG;                      // output channels (padded) = accumulator columns
    static constexpr int KS = CG / 2;                      // K steps of 16 channels
    static constexpr int ACOLS = 3 * KS * 8;               // TMEM columns of one A row
    static constexpr int D0 = CS_NA * ACOLS;               // accumulator ring follows the A ring
    static constexpr int ROW_B = CG * 128 * 16;            // one input row in shared memory: [CG groups][128 px][16 B]
    static constexpr int W_B = 9 * CG * NW * 16;           // weights: [tap][CG groups][NW rows][16 B]
    static_assert(D0 + CS_ND * NW <= 512, "TMEM columns");
};

template <int CG> struct CsSmem {
    uint8_t w[CsCfg<CG>::W_B];
    uint8_t in_ring[CS_NI][CsCfg<CG>::ROW_B];
    float bias[8 * CG], g1p[8 * CG], beta[8 * CG];
    uint64_t in_full[CS_NI], in_empty[CS_NI];
    uint64_t a_full[CS_NA], a_empty[CS_NA];
    uint64_t d_full[CS_WG_E], d_empty[CS_ND];
    uint32_t tmem_slot, pad;
};

struct CsArgs {
    int B, H, W, C;
    int act;
    int strips, segs, seg_rows;
    const __half* w;
    const float *bias, *g1p, *beta;      // g1p / beta: null = no affine output
    const __half* resid;                 // null = no residual
    __half* out_pre;                     // null = not stored
    __half* out_aff;
};

__device__ __forceinline__ void cs_tmem_st8(uint32_t taddr, const uint4& lo, const uint4& hi) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
}
__device__ __forceinline__ void cs_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint4 cs_pack8(const float2* x) {
    uint4 o;
    o.x = pack_h2_satfinite(x[0]); o.y = pack_h2_satfinite(x[1]);
    o.z = pack_h2_satfinite(x[2]); o.w = pack_h2_satfinite(x[3]);
    return o;
}
template <int ACT>
__device__ __forceinline__ float2 cs_act2(float2 x, int act) {
    if (ACT >= 0) return act2<ACT>(x);
    switch (act) {
        case BNERV_ACT_SIN:    return sin2(x);
        case BNERV_ACT_GELU:   return gelu2(x);
        case BNERV_ACT_RELU:   return act2<BNERV_ACT_RELU>(x);
        case BNERV_ACT_TANH01: return tanh01_2(x);
        default:               return x;
    }
}

// NPAIR: channel pairs that carry data (whole dead pairs are set to the 0 they would compute: zero weight rows, zero bias,
// activations that map 0 to 0, zero residual); 4*CG = all.
template <int CG, int ACT, int NPAIR>
__global__ void __launch_bounds__(CS_THREADS, 1)
conv_stream_kernel(const __grid_constant__ CUtensorMap tmIn, const CsArgs a) {
    using Cfg = CsCfg<CG>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    CsSmem<CG>& sm = *reinterpret_cast<CsSmem<CG>*>(smem_raw);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int q = warp & 3;
    const int m = q * 32 + lane;                      // TMEM lane == column of the strip

    const int seg = blockIdx.x % a.segs;
    const int strip = (blockIdx.x / a.segs) % a.strips;
    const int fb = blockIdx.x / (a.segs * a.strips);
    const int y0 = seg * a.seg_rows;
    const int y1 = (y0 + a.seg_rows < a.H) ? y0 + a.seg_rows : a.H;
    const int rows = y1 - y0;
    const int sx0 = strip * CS_VALID;                 // image column of lane 0
    const int col = sx0 + m;
    const int n_in = rows + 2;                        // input rows y0-1 .. y1
    const int in_row0 = y0 - 1, in_x0 = sx0 - 1;      // pixel 0 of a shared-memory row = image column sx0 - 1

    if (threadIdx.x == 0) {
        for (int i = 0; i < CS_NI; ++i) { mbar_init(smem_u32(&sm.in_full[i]), 1); mbar_init(smem_u32(&sm.in_empty[i]), 4); }
        for (int i = 0; i < CS_NA; ++i) { mbar_init(smem_u32(&sm.a_full[i]), 4); mbar_init(smem_u32(&sm.a_empty[i]), 1); }
        for (int i = 0; i < CS_ND; ++i) mbar_init(smem_u32(&sm.d_empty[i]), 4);
        for (int i = 0; i < CS_WG_E; ++i) mbar_init(smem_u32(&sm.d_full[i]), 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmIn);
    }
    if (warp == CS_WARP_MMA) tmem_alloc(smem_u32(&sm.tmem_slot), 512);
    {   // weights: the packed global form [tap][CG groups][NW rows][8 halves] is the shared-memory form
        const uint4* s0 = reinterpret_cast<const uint4*>(a.w);
        uint4* d0 = reinterpret_cast<uint4*>(sm.w);
        for (int i = threadIdx.x; i < Cfg::W_B / 16; i += CS_THREADS) d0[i] = __ldg(s0 + i);
    }
    if (threadIdx.x < Cfg::NW) sm.bias[threadIdx.x] = __ldg(a.bias + threadIdx.x);
    pdl_wait();                       // TAT tables / activations come from earlier kernels
    pdl_launch_dependents();
    if (threadIdx.x < Cfg::NW) {
        sm.g1p[threadIdx.x] = a.g1p ? __ldg(a.g1p + fb * Cfg::NW + threadIdx.x) : 0.0f;
        sm.beta[threadIdx.x] = a.beta ? __ldg(a.beta + fb * Cfg::NW + threadIdx.x) : 0.0f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_slot;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    if (warp == CS_WARP_MMA + 1) {
        // =============================== TMA producer: input rows -> ring ===============================
        for (int i = 0; i < n_in; ++i) {
            const int slot = i % CS_NI;
            mbar_wait(smem_u32(&sm.in_empty[slot]), ((i / CS_NI) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(smem_u32(&sm.in_full[slot]), Cfg::ROW_B);
                tma_load_3d(smem_u32(sm.in_ring[slot]), &tmIn, smem_u32(&sm.in_full[slot]), 2 * in_x0, in_row0 + i, fb * CG);
            }
            __syncwarp();
        }
    } else if (warp == CS_WARP_MMA) {
        // =============================== MMA issuer ===============================
        const uint32_t idesc = umma_idesc_f16(128, Cfg::NW);
        const uint64_t b_d = umma_desc_hi_noswz(static_cast<uint32_t>(Cfg::NW) * 16u, 128u);
        const uint32_t b_hi = static_cast<uint32_t>(b_d >> 32);
        const uint32_t b_lo = static_cast<uint32_t>(b_d) | ((smem_u32(sm.w) & 0x3FFFFu) >> 4);
        int a_waited = 0;
        for (int j = 0; j < rows; ++j) {
            while (a_waited <= j + 2) {                             // A rows j, j+1, j+2 (image rows h-1, h, h+1)
                mbar_wait(smem_u32(&sm.a_full[a_waited % CS_NA]), (a_waited / CS_NA) & 1);
                ++a_waited;
            }
            const int ds = j % CS_ND;
            mbar_wait(smem_u32(&sm.d_empty[ds]), ((j / CS_ND) & 1) ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + Cfg::D0 + ds * Cfg::NW;
                // accumulation order of conv_tc_kernel: K steps outermost, the nine taps (r*3 + sx) inside a K step
#pragma unroll
                for (int ks = 0; ks < Cfg::KS; ++ks) {              // one K step of B = 2 channel groups x NW rows x 16 B
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const uint32_t at = tmem_base + ((j + r) % CS_NA) * Cfg::ACOLS;
#pragma unroll
                        for (int sx = 0; sx < 3; ++sx)
                            cs_umma_ts(d, at + (sx * Cfg::KS + ks) * 8, b_lo + ((r * 3 + sx) * CG + 2 * ks) * Cfg::NW, b_hi, idesc,
                                       (ks | r | sx) ? 1u : 0u);
                    }
                }
                umma_commit(smem_u32(&sm.d_full[j % CS_WG_E]));
                umma_commit(smem_u32(&sm.a_empty[j % CS_NA]));       // A row j has had its last reader
            }
            __syncwarp();
        }
    } else if (warp < CS_WARP_E) {
        // =============================== front warpgroups: input rows -> A rows ===============================
        for (int i = warp >> 2; i < n_in; i += CS_WG_F) {
            const int slot = i % CS_NI;
            mbar_wait(smem_u32(&sm.in_full[slot]), (i / CS_NI) & 1);
            const uint8_t* row = sm.in_ring[slot];
            const int as = i % CS_NA;
            const uint32_t t = lane_base + as * Cfg::ACOLS;
            auto load = [&](int sx, uint4 (&g)[CG]) {
                int px = m + sx;
                px = px > 127 ? 127 : px;                            // lanes 126, 127: not valid lanes, any finite data
#pragma unroll
                for (int c = 0; c < CG; ++c) g[c] = *reinterpret_cast<const uint4*>(row + static_cast<size_t>(c * 128 + px) * 16);
            };
            auto store = [&](int sx, const uint4 (&g)[CG]) {
#pragma unroll
                for (int ks = 0; ks < Cfg::KS; ++ks) cs_tmem_st8(t + (sx * Cfg::KS + ks) * 8, g[2 * ks], g[2 * ks + 1]);
            };
            uint4 g[CG], gn[CG];                                     // one shift in flight ahead of the stores
            load(0, g);
            mbar_wait(smem_u32(&sm.a_empty[as]), ((i / CS_NA) & 1) ^ 1);
            tc_fence_after();
            load(1, gn);
            store(0, g);
            load(2, g);
            store(1, gn);
            store(2, g);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&sm.a_full[as]));
                mbar_arrive(smem_u32(&sm.in_empty[slot]));
            }
        }
    } else if (warp < CS_WARP_MMA) {
        // =============================== epilogue warpgroups ===============================
        const int par = (warp - CS_WARP_E) >> 2;
        const size_t plane = static_cast<size_t>(a.H) * a.W * 8;
        const bool lane_valid = (m < CS_VALID) && (col < a.W);
        int it = 0;
        for (int k = par; k < rows; k += CS_WG_E, ++it) {
            const int h = y0 + k;
            const int ds = k % CS_ND;
            const size_t goff = ((static_cast<size_t>(fb) * CG) * a.H + h) * static_cast<size_t>(a.W) * 8 + static_cast<size_t>(col) * 8;
            uint4 r[CG];
#pragma unroll
            for (int c = 0; c < CG; ++c) {
                r[c] = make_uint4(0, 0, 0, 0);
                if (a.resid && lane_valid && 4 * c < NPAIR) r[c] = __ldg(reinterpret_cast<const uint4*>(a.resid + goff + c * plane));
            }
            mbar_wait(smem_u32(&sm.d_full[par]), it & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < CG / 2; ++half) {
                uint32_t v[16];
                tmem_ld16(lane_base + Cfg::D0 + ds * Cfg::NW + half * 16, v);
                tmem_ld_wait();
                if (half == CG / 2 - 1) {                            // all of this row's accumulator has been read
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&sm.d_empty[ds]));
                }
                float2 x[8];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float4 bs = *reinterpret_cast<const float4*>(sm.bias + 16 * half + 4 * p);
                    x[2 * p]     = add2(make_float2(__uint_as_float(v[4 * p]),     __uint_as_float(v[4 * p + 1])), make_float2(bs.x, bs.y));
                    x[2 * p + 1] = add2(make_float2(__uint_as_float(v[4 * p + 2]), __uint_as_float(v[4 * p + 3])), make_float2(bs.z, bs.w));
                }
#pragma unroll
                for (int p = 0; p < 8; ++p)
                    x[p] = (half * 8 + p < NPAIR) ? cs_act2<ACT>(x[p], a.act) : make_float2(0.0f, 0.0f);
                if (a.resid) {
                    const uint4 r0 = r[2 * half], r1 = r[2 * half + 1];
                    x[0] = add2(x[0], unpack_h2(r0.x)); x[1] = add2(x[1], unpack_h2(r0.y));
                    x[2] = add2(x[2], unpack_h2(r0.z)); x[3] = add2(x[3], unpack_h2(r0.w));
                    x[4] = add2(x[4], unpack_h2(r1.x)); x[5] = add2(x[5], unpack_h2(r1.y));
                    x[6] = add2(x[6], unpack_h2(r1.z)); x[7] = add2(x[7], unpack_h2(r1.w));
                }
                if (!lane_valid) continue;
                const size_t o = goff + (2 * half) * plane;
                if (a.out_pre) {
                    *reinterpret_cast<uint4*>(a.out_pre + o) = cs_pack8(x);
                    *reinterpret_cast<uint4*>(a.out_pre + o + plane) = cs_pack8(x + 4);
                }
                if (a.out_aff) {
                    float2 y[8];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float4 gg = *reinterpret_cast<const float4*>(sm.g1p + 16 * half + 4 * p);
                        const float4 ee = *reinterpret_cast<const float4*>(sm.beta + 16 * half + 4 * p);
                        y[2 * p]     = fma2(x[2 * p],     make_float2(gg.x, gg.y), make_float2(ee.x, ee.y));
                        y[2 * p + 1] = fma2(x[2 * p + 1], make_float2(gg.z, gg.w), make_float2(ee.z, ee.w));
                    }
                    *reinterpret_cast<uint4*>(a.out_aff + o) = cs_pack8(y);
                    *reinterpret_cast<uint4*>(a.out_aff + o + plane) = cs_pack8(y + 4);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CS_WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int g_cs_sms = 0;

template <int CG>
static int conv_stream_launch(const void* x, CsArgs& a, cudaStream_t stream) {
    if (g_cs_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_cs_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_cs_sms <= 0) g_cs_sms = 148;
    }
    a.strips = (a.W + CS_VALID - 1) / CS_VALID;
    {   // a segment of R rows costs ~R + 6 row times (2 halo rows + pipeline fill): minimise waves x (R + 6)
        int best_segs = 1;
        long long best_cost = -1;
        const int max_segs = (a.H + 7) / 8;
        for (int sg = 1; sg <= max_segs; ++sg) {
            const int r = (a.H + sg - 1) / sg;
            const int real = (a.H + r - 1) / r;
            const long long ctas = 1LL * a.B * a.strips * real;
            const long long waves = (ctas + g_cs_sms - 1) / g_cs_sms;
            const long long cost = waves * (r + 6);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_segs = sg; }
        }
        a.seg_rows = (a.H + best_segs - 1) / best_segs;
        a.segs = (a.H + a.seg_rows - 1) / a.seg_rows;
    }
    const long long grid = 1LL * a.B * a.strips * a.segs;
    if (grid > 0x3fffffffLL) return set_error(BNERV_E_UNSUPPORTED, "conv_stream: too many CTAs");

    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return set_error(BNERV_E_NODRIVER, "cuTensorMapEncodeTiled entry point not available");
    CUtensorMap tm;
    cuuint64_t dims[3] = {2ull * a.W, static_cast<cuuint64_t>(a.H), 1ull * CG * a.B};
    cuuint64_t strides[2] = {16ull * a.W, 16ull * a.W * a.H};
    cuuint32_t box[3] = {256, 1, CG};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(static_cast<int>(r), "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));

    using KernelFn = void (*)(const CUtensorMap, const CsArgs);
    constexpr int ALL = 4 * CG;
    const int pairs = (a.act == BNERV_ACT_TANH01) ? ALL : (a.C + 1) / 2;
    const bool trim = pairs <= ALL - 2;                   // at least two dead pairs: the NPAIR = ALL - 2 instantiations
    KernelFn fn;
    int slot;
    if (a.act == BNERV_ACT_GELU)      { fn = trim ? conv_stream_kernel<CG, BNERV_ACT_GELU, ALL - 2> : conv_stream_kernel<CG, BNERV_ACT_GELU, ALL>; slot = 0 + trim; }
    else if (a.act == BNERV_ACT_SIN)  { fn = trim ? conv_stream_kernel<CG, BNERV_ACT_SIN, ALL - 2> : conv_stream_kernel<CG, BNERV_ACT_SIN, ALL>; slot = 2 + trim; }
    else if (a.act == BNERV_ACT_NONE) { fn = trim ? conv_stream_kernel<CG, BNERV_ACT_NONE, ALL - 2> : conv_stream_kernel<CG, BNERV_ACT_NONE, ALL>; slot = 4 + trim; }
    else                              { fn = conv_stream_kernel<CG, -1, ALL>; slot = 6; }
    static bool attr_set[7][32] = {};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    const size_t smem = sizeof(CsSmem<CG>) + 1024;
    if (!attr_set[slot][cur_dev & 31]) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
        attr_set[slot][cur_dev & 31] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(CS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("BNERV_NO_PDL") != nullptr;
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, tm, a);
    if (e != cudaSuccess) {
        count_launch();
        return set_error(static_cast<int>(e), "conv_stream_kernel launch: %s", cudaGetErrorString(e));
    }
    return check_launch("conv_stream_kernel");
}

}  // namespace bnerv

using namespace bnerv;

extern "C" int bnerv_conv_stream(const void* x, int B, int Cin, int H, int W, const void* w_packed, const float* bias_packed, int Cout,
                                 int act, const void* resid, const float* g1p, const float* beta, void* out_pre, void* out_aff,
                                 void* stream) {
    if (!x || !w_packed || !bias_packed) return set_error(BNERV_E_BADARG, "conv_stream: null operand");
    if (B <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "conv_stream: non-positive size");
    if ((g1p == nullptr) != (beta == nullptr) || (g1p != nullptr) != (out_aff != nullptr))
        return set_error(BNERV_E_BADARG, "conv_stream: g1p, beta and out_aff go together");
    if (!out_pre && !out_aff) return set_error(BNERV_E_BADARG, "conv_stream: no output");
    if (act < BNERV_ACT_NONE || act > BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "conv_stream: act %d", act);
    const int cin_p = round_up(Cin, 16), cout_p = round_up(Cout, 16);
    if (cin_p != cout_p || (cout_p != 32 && cout_p != 48))
        return set_error(BNERV_E_UNSUPPORTED, "conv_stream: Cin = %d, Cout = %d (equal padded widths of 32 or 48 channels)", Cin, Cout);
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(resid) |
                               reinterpret_cast<uintptr_t>(out_pre) | reinterpret_cast<uintptr_t>(out_aff);
    if (align_or & 15) return set_error(BNERV_E_BADARG, "conv_stream: pointers must be 16-byte aligned");
    CsArgs a{};
    a.B = B; a.H = H; a.W = W; a.C = Cout; a.act = act;
    a.w = static_cast<const __half*>(w_packed); a.bias = bias_packed; a.g1p = g1p; a.beta = beta;
    a.resid = static_cast<const __half*>(resid);
    a.out_pre = static_cast<__half*>(out_pre);
    a.out_aff = static_cast<__half*>(out_aff);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return cout_p == 32 ? conv_stream_launch<4>(x, a, st) : conv_stream_launch<6>(x, a, st);
}
