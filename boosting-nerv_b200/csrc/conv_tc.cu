// Fused implicit-GEMM conv for sm_100a: TMA-staged halo tiles -> tcgen05.mma (f16 x f16 -> f32 in TMEM)
// -> in-register epilogue (bias, sin/GELU, residual, TAT affine, PixelShuffle addressing) -> 16-byte
// vector stores.  Replaces CustomConv2d.forward + PixelShuffle + Sin/GELU + SFTLayer affine + residual
// (lib/quant_ops.py:39-41, model_blocks.py:37,86-89,105,204,217) with one launch.
//
// GEMM view (per image):  D[pixels, n'] = sum_{tap, c} X[pixel + tap, c] * Wp[tap][c][n']
//   M: a CTA works on a 16-row x 16-col pixel super-tile = MT(2) UMMA tiles of 16 rows x 8 cols (M=128)
//   N: N_ACC <= 128 packed output rows per CTA (n' = PixelShuffle-major, see bnerv_b200.h)
//   K: 9 taps x Cin_p channels, consumed 16 channels (one UMMA K step) per pipeline stage.
//
// Shared-memory operand layout is the UMMA "no-swizzle, K-major" canonical form: a core matrix is
// 8 rows x 16 B = 128 contiguous bytes.  The activation layout in HBM ([Cp/8][H][W][8] f16) makes one
// TMA box {18 px, 18 rows, 2 channel groups} land as [group][row][px][16 B]: 8 neighbouring pixels of
// one image row ARE a core matrix, so the A operand of tap (r,s) is simply the same halo tile read
// at start address + (r*18 + s)*16 B with SBO = one halo row (288 B).  The halo tile is fetched once
// and reused by all 9 taps (no im2col materialisation, no 9x re-fetch).
//
// Warp roles (320 threads): warp 0 = TMA producer (1 lane), warp 1 = TMEM owner + MMA issuer (1 lane),
// warps 2..9 = epilogue (two groups of 4 warps, one group per UMMA tile; warp%4 selects its TMEM lane
// quarter).  Two TMEM accumulator buffers (2 x 2 x 128 columns = all 512) let the epilogue of tile i
// overlap the MMAs of tile i+1.  Persistent grid: one CTA per SM, static round-robin tile order.
#include <cuda.h>
#include "common.cuh"

namespace bnerv {

constexpr int MT          = 2;                    // UMMA tiles (128 px each) per CTA super-tile
constexpr int TILE_H      = 16;                   // pixel rows per super-tile
constexpr int TILE_W      = 8 * MT;               // pixel cols per super-tile
constexpr int HALO_W      = TILE_W + 2;           // 18
constexpr int HALO_H      = TILE_H + 2;           // 18
constexpr int A_GROUP_B   = HALO_H * HALO_W * 16; // bytes of one 8-channel group of the halo tile (5184)
constexpr int A_STAGE_B   = 2 * A_GROUP_B;        // one K step = 16 channels = 2 groups (10368, 128-aligned)
constexpr int ACC_COLS    = 128;                  // TMEM columns reserved per accumulator
constexpr int TMEM_COLS   = 512;
constexpr int N_EPI_WARPS = 8 * MT;            // per UMMA tile: 4 lane quarters x 2 column parities
constexpr int N_THREADS   = 64 + 32 * N_EPI_WARPS;
constexpr int MAX_STAGES  = 8;
constexpr int SMEM_LIMIT  = 227 * 1024;
constexpr int BAR_BYTES   = (2 * MAX_STAGES + 4) * 8 + 16;   // mbarriers + TMEM base slot (16-byte multiple)
constexpr int CST_N       = 128;                  // per-tile epilogue constants: [2 buffers][bias|g1p|beta][CST_N] f32
constexpr int CST_BYTES   = 2 * 3 * CST_N * 4;

struct ConvTcArgs {
    int B, H, W;            // conv-resolution geometry (input == pre-shuffle output)
    int cin_groups;         // Cin_p / 8
    int ksteps;             // Cin_p / 16
    int taps;               // 1 or 9
    int n_total;            // s*s*Cout_p
    int n_acc;              // packed rows per CTA (multiple of 16, <= 128)
    int n_tiles;            // ceil(n_total / n_acc)
    int cout, cout_p;       // real / padded output channels
    int s;                  // PixelShuffle factor
    int act;
    int flags;              // F_* epilogue features (used by the generic instantiation)
    int tiles_x, tiles_y;
    int total_tiles;
    int stages;
    int b_stage_bytes;      // taps * 2 * n_acc * 16
    const float* bias;      // [n_total]
    const float* g1p;       // [B][cout_p] or null
    const float* beta;      // [B][cout_p] or null
    const __half* resid;    // C8 at output resolution or null
    __half* out_pre;        // C8 or null
    __half* out_aff;        // C8 or null
    float* out_nchw;        // NCHW f32 or null
};

struct TileCoord { int n0, b, h0, w0; };

__device__ __forceinline__ TileCoord decode_tile(const ConvTcArgs& a, int tile) {
    TileCoord t;
    int nt   = tile % a.n_tiles;
    int rest = tile / a.n_tiles;
    int tx   = rest % a.tiles_x;
    rest /= a.tiles_x;
    int ty = rest % a.tiles_y;
    t.b    = rest / a.tiles_y;
    t.n0   = nt * a.n_acc;
    t.h0   = ty * TILE_H;
    t.w0   = tx * TILE_W;
    return t;
}

// epilogue feature flags (compile-time in the specialised instantiations, run-time in the generic one)
constexpr int F_RESID = 1, F_AFF = 2, F_PRE = 4, F_NCHW = 8, F_SHUF = 16;

template <int ACT>
__device__ __forceinline__ float2 act2_rt(float2 x, int act) {
    if (ACT >= 0) return act2<ACT>(x);
    switch (act) {
        case BNERV_ACT_SIN:    return sin2(x);
        case BNERV_ACT_GELU:   return gelu2(x);
        case BNERV_ACT_RELU:   return act2<BNERV_ACT_RELU>(x);
        case BNERV_ACT_TANH01: return tanh01_2(x);
        default:               return x;
    }
}

// One 8-channel group of one pixel: bias + activation (+ residual) (+ affine) and the stores.
// cb points at this tile's constants in shared memory: [0..127] bias, [128..255] g1p, [256..383] beta, indexed by
// the packed row relative to the tile (col = g16*16 + hh*8); all lanes read the same address (broadcast).
template <int ACT, int FLAGS>
__device__ __forceinline__ void epilogue_chunk(const ConvTcArgs& a, int flags, const uint32_t* v, const float* cb, int col,
                                               int cc, int b, size_t off, bool valid, const uint4& rr, int ho, int wo,
                                               int Ho, int Wo) {
    const float4 b0 = *reinterpret_cast<const float4*>(cb + col);
    const float4 b1 = *reinterpret_cast<const float4*>(cb + col + 4);
    float2 x[4];
    x[0] = add2(make_float2(__uint_as_float(v[0]), __uint_as_float(v[1])), make_float2(b0.x, b0.y));
    x[1] = add2(make_float2(__uint_as_float(v[2]), __uint_as_float(v[3])), make_float2(b0.z, b0.w));
    x[2] = add2(make_float2(__uint_as_float(v[4]), __uint_as_float(v[5])), make_float2(b1.x, b1.y));
    x[3] = add2(make_float2(__uint_as_float(v[6]), __uint_as_float(v[7])), make_float2(b1.z, b1.w));
#pragma unroll
    for (int p = 0; p < 4; ++p) x[p] = act2_rt<ACT>(x[p], a.act);
    if (flags & F_RESID) {
        x[0] = add2(x[0], unpack_h2(rr.x)); x[1] = add2(x[1], unpack_h2(rr.y));
        x[2] = add2(x[2], unpack_h2(rr.z)); x[3] = add2(x[3], unpack_h2(rr.w));
    }
    if (!valid) return;
    if (flags & F_PRE) {
        uint4 o;
        o.x = pack_h2_satfinite(x[0]); o.y = pack_h2_satfinite(x[1]);
        o.z = pack_h2_satfinite(x[2]); o.w = pack_h2_satfinite(x[3]);
        *reinterpret_cast<uint4*>(a.out_pre + off) = o;
    }
    if (flags & F_NCHW) {
        const float xs[8] = {x[0].x, x[0].y, x[1].x, x[1].y, x[2].x, x[2].y, x[3].x, x[3].y};
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (cc + k < a.cout) a.out_nchw[(static_cast<size_t>(b * a.cout + cc + k) * Ho + ho) * Wo + wo] = xs[k];
    }
    if (flags & F_AFF) {
        const float4 g0 = *reinterpret_cast<const float4*>(cb + CST_N + col);
        const float4 g1 = *reinterpret_cast<const float4*>(cb + CST_N + col + 4);
        const float4 e0 = *reinterpret_cast<const float4*>(cb + 2 * CST_N + col);
        const float4 e1 = *reinterpret_cast<const float4*>(cb + 2 * CST_N + col + 4);
        uint4 o;
        o.x = pack_h2_satfinite(fma2(x[0], make_float2(g0.x, g0.y), make_float2(e0.x, e0.y)));
        o.y = pack_h2_satfinite(fma2(x[1], make_float2(g0.z, g0.w), make_float2(e0.z, e0.w)));
        o.z = pack_h2_satfinite(fma2(x[2], make_float2(g1.x, g1.y), make_float2(e1.x, e1.y)));
        o.w = pack_h2_satfinite(fma2(x[3], make_float2(g1.z, g1.w), make_float2(e1.z, e1.w)));
        *reinterpret_cast<uint4*>(a.out_aff + off) = o;
    }
}

struct Pipe {
    uint32_t smem_base;      // shared-window address of the stage ring
    uint32_t full, empty;    // shared-window addresses of full_bar[0] / empty_bar[0]
    uint32_t tfull, tempty;  // ... of tfull_bar[0] / tempty_bar[0]
    uint32_t tmem_base;
    int stage_bytes;
};

// ===================== TMA producer (warp 0, converged; one elected lane issues) =====================
__device__ __forceinline__ void producer_role(const ConvTcArgs& a, const Pipe& p, const CUtensorMap* tmA, const CUtensorMap* tmB) {
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = A_STAGE_B + a.b_stage_bytes;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(a, tile);
        for (int kc = 0; kc < a.ksteps; ++kc) {
            mbar_wait(p.empty + stage * 8, phase ^ 1);
            if (elect_one()) {
                const uint32_t fb = p.full + stage * 8;
                const uint32_t sa = p.smem_base + stage * p.stage_bytes;
                mbar_expect_tx(fb, tx_bytes);
                // activations viewed as u64 elements: 2 per pixel-group -> x coordinate = 2*w
                tma_load_3d(sa, tmA, fb, 2 * (t.w0 - 1), t.h0 - 1, t.b * a.cin_groups + 2 * kc);
                tma_load_3d(sa + A_STAGE_B, tmB, fb, 2 * t.n0, 2 * kc, 0);
            }
            __syncwarp();
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
    }
}

// ===================== MMA issuer (warp 1, converged; one elected lane issues) =====================
// Everything the issue needs is warp-uniform and the taps are unrolled with constant descriptor offsets, so one
// K step (TAPS x MT UMMAs) is a straight run of UTCHMMA separated by a few uniform adds.
template <int TAPS>
__device__ __forceinline__ void mma_role(const ConvTcArgs& a, const Pipe& p) {
    int stage = 0;
    uint32_t phase = 0;
    int abuf = 0;
    uint32_t aphase = 0;
    const uint32_t idesc   = umma_idesc_f16_m128(a.n_acc);
    const uint32_t b_tap16 = 2u * a.n_acc;                      // one tap's [2 groups][n_acc][16 B] slab, in 16-byte units
    const uint64_t a_hi    = umma_desc_hi_noswz(A_GROUP_B, HALO_W * 16);
    const uint64_t b_hi    = umma_desc_hi_noswz(a.n_acc * 16u, 128u);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        mbar_wait(p.tempty + abuf * 8, aphase ^ 1);
        tc_fence_after();
        const uint32_t d0 = p.tmem_base + abuf * (MT * ACC_COLS);
        for (int kc = 0; kc < a.ksteps; ++kc) {
            mbar_wait(p.full + stage * 8, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa16 = (p.smem_base + stage * p.stage_bytes) >> 4;
                const uint32_t sb16 = sa16 + (A_STAGE_B >> 4);
#pragma unroll
                for (int tp = 0; tp < TAPS; ++tp) {
                    const int tap = (TAPS == 1) ? 4 : tp;       // 1x1 conv: the single tap reads the halo tile's centre
                    const int r = tap / 3, sx = tap % 3;
                    const uint64_t bdesc = b_hi | static_cast<uint64_t>(sb16 + tp * b_tap16);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint64_t adesc = a_hi | static_cast<uint64_t>(sa16 + (r * HALO_W + sx + mt * 8));
                        umma_f16(d0 + mt * ACC_COLS, adesc, bdesc, idesc, (tp > 0) ? 1u : (kc > 0 ? 1u : 0u));
                    }
                }
                umma_commit(p.empty + stage * 8);
                if (kc == a.ksteps - 1) umma_commit(p.tfull + abuf * 8);
            }
            __syncwarp();
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        abuf ^= 1;
        if (abuf == 0) aphase ^= 1;
    }
}

template <int ACT, int FLAGS>
__global__ void __launch_bounds__(N_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int flags = (FLAGS >= 0) ? FLAGS : a.flags;

    const int stage_bytes = A_STAGE_B + a.b_stage_bytes;
    uint8_t* bar_base     = smem + static_cast<size_t>(a.stages) * stage_bytes;
    uint64_t* full_bar    = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar   = full_bar + MAX_STAGES;
    uint64_t* tfull_bar   = empty_bar + MAX_STAGES;   // [2]
    uint64_t* tempty_bar  = tfull_bar + 2;            // [2]
    uint32_t* tmem_slot   = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float*    cst         = reinterpret_cast<float*>(bar_base + BAR_BYTES);   // [2][3][CST_N]

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&tfull_bar[i]), 1);
            mbar_init(smem_u32(&tempty_bar[i]), N_EPI_WARPS);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    Pipe p;
    p.smem_base = smem_u32(smem);
    p.full = smem_u32(full_bar);     p.empty = smem_u32(empty_bar);
    p.tfull = smem_u32(tfull_bar);   p.tempty = smem_u32(tempty_bar);
    p.tmem_base = *tmem_slot;
    p.stage_bytes = stage_bytes;

    if (warp == 0) {
        producer_role(a, p, &tmA, &tmB);
    } else if (warp == 1) {
        if (a.taps == 9) mma_role<9>(a, p); else mma_role<1>(a, p);
    } else {
        // ===================== epilogue: 16 warps =====================
        // warp -> (TMEM lane quarter q = warp%4 [hardware rule], UMMA tile mt, column parity half):
        // a warp owns the 16-column groups g16 = half, half+2, half+4, half+6 of its tile.
        const int e    = warp - 2;
        const int q    = warp & 3;
        const int mt   = (e >> 2) >> 1;
        const int half = (e >> 2) & 1;
        const int m    = q * 32 + lane;                // row of the UMMA tile == pixel
        const int et   = threadIdx.x - 64;             // 0 .. 32*N_EPI_WARPS-1
        int abuf = 0;
        uint32_t aphase = 0;
        const int s  = a.s;
        const int Ho = a.H * s, Wo = a.W * s;
        const int cout_groups = a.cout_p >> 3;
        const size_t plane = static_cast<size_t>(Ho) * Wo * 8;      // halves per 8-channel plane
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
            const TileCoord t = decode_tile(a, tile);
            const int h = t.h0 + (m >> 3);
            const int w = t.w0 + mt * 8 + (m & 7);
            const bool valid = (h < a.H) && (w < a.W);
            const size_t base_b = static_cast<size_t>(t.b) * cout_groups * plane;

            // Stage this tile's per-row constants (bias, TAT scale+1, TAT shift) in shared memory: one global load
            // per constant, issued before the accumulator wait; the chunks then read them as LDS broadcasts.
            float* cb = cst + abuf * (3 * CST_N);
            if (et < a.n_acc) {
                const int nn = t.n0 + et;
                float bv = 0.0f, gv = 0.0f, ev = 0.0f;
                if (nn < a.n_total) {
                    bv = __ldg(a.bias + nn);
                    if (flags & F_AFF) {
                        const int cc = (flags & F_SHUF) ? nn % a.cout_p : nn;
                        gv = __ldg(a.g1p + static_cast<size_t>(t.b) * a.cout_p + cc);
                        ev = __ldg(a.beta + static_cast<size_t>(t.b) * a.cout_p + cc);
                    }
                }
                cb[et] = bv; cb[CST_N + et] = gv; cb[2 * CST_N + et] = ev;
            }

            // per 16-column group owned by this warp: packed row, channel, output pixel
            int  nn16[4], cc16[4], ho16[4], wo16[4];
            bool act16[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int g16 = half + 2 * j;
                nn16[j]  = t.n0 + g16 * 16;
                act16[j] = (g16 * 16 < a.n_acc) && (nn16[j] < a.n_total);
                int sub = 0, cc = nn16[j];
                if (flags & F_SHUF) { sub = nn16[j] / a.cout_p; cc = nn16[j] - sub * a.cout_p; }
                const int i = sub / s, jj = sub - i * s;
                cc16[j] = cc;
                ho16[j] = h * s + i;
                wo16[j] = w * s + jj;
            }

            // Residual prefetch: issued before waiting for the accumulator so the HBM latency hides
            // behind the MMAs of this tile.
            uint4 rres[4][2];
            if (flags & F_RESID) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        rres[j][hh] = make_uint4(0, 0, 0, 0);
                        if (act16[j] && valid) {
                            const size_t off = base_b + static_cast<size_t>((cc16[j] >> 3) + hh) * plane +
                                               (static_cast<size_t>(ho16[j]) * Wo + wo16[j]) * 8;
                            rres[j][hh] = __ldg(reinterpret_cast<const uint4*>(a.resid + off));
                        }
                    }
            }

            // constants visible to all epilogue warps; also orders this tile's writes to cb after every warp's
            // reads of the same buffer two tiles ago
            named_bar_sync(1, 32 * N_EPI_WARPS);
            mbar_wait(p.tfull + abuf * 8, aphase);
            tc_fence_after();
            const uint32_t taddr = p.tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (abuf * MT + mt) * ACC_COLS;

            uint32_t v[2][16];
            if (act16[0]) tmem_ld16(taddr + half * 16, v[0]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (act16[j]) {                                           // CTA-uniform
                    tmem_ld_wait();
                    if (j + 1 < 4 && act16[(j + 1) & 3]) tmem_ld16(taddr + (half + 2 * (j + 1)) * 16, v[(j + 1) & 1]);
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const size_t off = base_b + static_cast<size_t>((cc16[j] >> 3) + hh) * plane +
                                           (static_cast<size_t>(ho16[j]) * Wo + wo16[j]) * 8;
                        epilogue_chunk<ACT, FLAGS>(a, flags, &v[j & 1][hh * 8], cb, (half + 2 * j) * 16 + hh * 8,
                                                   cc16[j] + hh * 8, t.b, off, valid, rres[j][hh], ho16[j], wo16[j], Ho, Wo);
                    }
                }
            }
            // all TMEM reads of this warp for this buffer are complete -> hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p.tempty + abuf * 8);
            abuf ^= 1;
            if (abuf == 0) aphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(p.tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

static int make_map_u64_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                           uint64_t stride1_b, uint64_t stride2_b, uint32_t b0, uint32_t b1, uint32_t b2) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return set_error(BNERV_E_NODRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3]    = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_b, stride2_b};
    cuuint32_t box[3]     = {b0, b1, b2};
    cuuint32_t estr[3]    = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(static_cast<int>(r), "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
    return 0;
}

static int g_num_sms = 0;

int choose_n_acc(int n_total) {
    // fewest N tiles with N_ACC <= 128, then the smallest multiple of 16 that covers them evenly
    int tiles = (n_total + 127) / 128;
    int per   = (n_total + tiles - 1) / tiles;
    return round_up(per, 16);
}

}  // namespace bnerv

using namespace bnerv;

extern "C" int bnerv_conv_fused(const void* x, int B, int Cin, int H, int W, const void* w_packed,
                                const float* bias_packed, int Cout, int k, int s, int act, const void* resid,
                                const float* g1p, const float* beta, void* out_pre, void* out_aff, float* out_nchw,
                                void* stream) {
    if (!x || !w_packed || !bias_packed) return set_error(BNERV_E_BADARG, "conv_fused: null operand");
    if (B <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || s <= 0) return set_error(BNERV_E_BADARG, "conv_fused: non-positive size");
    if (k != 1 && k != 3) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: kernel size %d (only 1 and 3)", k);
    if ((g1p == nullptr) != (beta == nullptr)) return set_error(BNERV_E_BADARG, "conv_fused: g1p and beta must both be given");
    if ((g1p != nullptr) != (out_aff != nullptr)) return set_error(BNERV_E_BADARG, "conv_fused: out_aff requires g1p/beta and vice versa");
    if (!out_pre && !out_aff && !out_nchw) return set_error(BNERV_E_BADARG, "conv_fused: no output");
    if (act < BNERV_ACT_NONE || act > BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: act %d", act);
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_packed) |
                               reinterpret_cast<uintptr_t>(bias_packed) | reinterpret_cast<uintptr_t>(resid) |
                               reinterpret_cast<uintptr_t>(g1p) | reinterpret_cast<uintptr_t>(beta) |
                               reinterpret_cast<uintptr_t>(out_pre) | reinterpret_cast<uintptr_t>(out_aff);
    if (align_or & 15) return set_error(BNERV_E_BADARG, "conv_fused: pointers must be 16-byte aligned");

    const int cin_p = round_up(Cin, 16), cout_p = round_up(Cout, 16);
    ConvTcArgs a{};
    a.B = B; a.H = H; a.W = W;
    a.cin_groups = cin_p / 8;
    a.ksteps     = cin_p / 16;
    a.taps       = k * k;
    a.n_total    = s * s * cout_p;
    a.n_acc      = choose_n_acc(a.n_total);
    a.n_tiles    = (a.n_total + a.n_acc - 1) / a.n_acc;
    a.cout = Cout; a.cout_p = cout_p; a.s = s; a.act = act;
    a.tiles_x = (W + TILE_W - 1) / TILE_W;
    a.tiles_y = (H + TILE_H - 1) / TILE_H;
    const long long total = 1LL * a.n_tiles * B * a.tiles_x * a.tiles_y;
    if (total > 0x7fffffffLL) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: too many tiles");
    a.total_tiles   = static_cast<int>(total);
    a.b_stage_bytes = a.taps * 2 * a.n_acc * 16;
    const int stage_bytes = A_STAGE_B + a.b_stage_bytes;
    const int bar_bytes   = BAR_BYTES + CST_BYTES;
    int stages = (SMEM_LIMIT - bar_bytes - 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: stage of %d bytes does not fit twice", stage_bytes);
    a.stages = stages;
    a.bias = bias_packed; a.g1p = g1p; a.beta = beta;
    a.resid = static_cast<const __half*>(resid);
    a.out_pre = static_cast<__half*>(out_pre);
    a.out_aff = static_cast<__half*>(out_aff);
    a.out_nchw = out_nchw;
    a.flags = (resid ? F_RESID : 0) | (g1p ? F_AFF : 0) | (out_pre ? F_PRE : 0) | (out_nchw ? F_NCHW : 0) | (s > 1 ? F_SHUF : 0);

    CUtensorMap tmA, tmB;
    // activations as u64: dims {2W, H, B*cin_groups}; strides {W*16, H*W*16} bytes
    int rc = make_map_u64_3d(&tmA, x, 2ull * W, H, 1ull * B * a.cin_groups, 16ull * W, 16ull * W * H,
                             2 * HALO_W, HALO_H, 2);
    if (rc) return rc;
    // packed weights as u64: dims {2*Np, Kp/8, taps}; strides {Np*16, Np*16*Kp/8}
    rc = make_map_u64_3d(&tmB, w_packed, 2ull * a.n_total, a.cin_groups, a.taps, 16ull * a.n_total,
                         16ull * a.n_total * a.cin_groups, 2 * a.n_acc, 2, a.taps);
    if (rc) return rc;

    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    const size_t smem_bytes = static_cast<size_t>(stages) * stage_bytes + bar_bytes;
    const int grid = a.total_tiles < g_num_sms ? a.total_tiles : g_num_sms;
    // specialised epilogues for the five launch shapes of the decoder cascade, generic otherwise
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const ConvTcArgs);
    KernelFn fn = conv_tc_kernel<-1, -1>;
    int slot = 0;
    const int fl = a.flags;
#define BNERV_PICK(ID, ACT_, FL_)                                              \
    if (act == (ACT_) && fl == (FL_)) { fn = conv_tc_kernel<(ACT_), (FL_)>; slot = (ID); }
    BNERV_PICK(1, BNERV_ACT_SIN, F_AFF | F_PRE)                 // up-conv 1x1 / s=1 (+sin, x0 and u)
    BNERV_PICK(2, BNERV_ACT_SIN, F_AFF | F_PRE | F_SHUF)        // up-conv + PixelShuffle
    BNERV_PICK(3, BNERV_ACT_GELU, F_AFF)                        // conv0 + GELU + TAT affine
    BNERV_PICK(4, BNERV_ACT_NONE, F_RESID | F_PRE)              // conv1 + residual
    BNERV_PICK(5, BNERV_ACT_NONE, F_PRE | F_SHUF)               // E-NeRV stage-0 up-conv
    BNERV_PICK(6, BNERV_ACT_TANH01, F_NCHW)                     // head conv -> image
#undef BNERV_PICK
    static bool smem_set[8] = {false, false, false, false, false, false, false, false};
    if (!smem_set[slot]) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
        smem_set[slot] = true;
    }
    fn<<<grid, N_THREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, a);
    return check_launch("conv_tc_kernel");
}
