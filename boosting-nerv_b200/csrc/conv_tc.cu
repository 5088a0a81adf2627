// Fused implicit-GEMM conv for sm_100a: weight-stationary CTA pairs.
//   TMA-staged halo tiles -> tcgen05.mma.cta_group::2 (f16 x f16 -> f32 in TMEM, M = 256 over two SMs)
//   -> in-register epilogue (bias, sin/GELU, residual, TAT affine, PixelShuffle addressing) -> 16-byte stores.
// Replaces CustomConv2d.forward + PixelShuffle + Sin/GELU + SFTLayer affine + residual
// (lib/quant_ops.py:39-41, model_blocks.py:37,86-89,105,204,217) with one launch.
//
// GEMM view (per image):  D[pixels, n'] = sum_{tap, c} X[pixel + tap, c] * Wp[tap][c][n']
//   M: each CTA of a pair owns a 16-row x 8*MT-col pixel super-tile = MT UMMA row blocks of 128 pixels; one UMMA
//      (M = 256) covers row block mt of BOTH CTAs.
//   N: n_acc <= 256/MT packed output rows (n' = PixelShuffle-major, see bnerv_b200.h) per pass ("n-tile").
//   K: 9 taps x Cin_p channels, 16 channels (one UMMA K step) per pipeline stage.
//
// Why pairs + resident weights (profiles/r01_v3_*): with one CTA per tile every UMMA reads A (4 KB) and B
// (N*32 B) from shared memory -> ~134 B/clk at N = 112, above the ~128 B/clk an SM delivers, and the weight
// slab is re-streamed from L2 for every tile.  In a pair each SM reads its own A and only HALF of B, and each
// CTA keeps its half of the n-tile's weights ([K steps][taps][2 groups][n_acc/2][16 B]) resident in shared
// memory for the whole pass, so only activations stream (10 KB per K step).
//
// Shared-memory operand layout is the UMMA "no-swizzle, K-major" canonical form: a core matrix is
// 8 rows x 16 B = 128 contiguous bytes.  The activation layout in HBM ([Cp/8][H][W][8] f16) makes one
// TMA box {HALO_W px, 18 rows, 2 channel groups} land as [group][row][px][16 B]: 8 neighbouring pixels of
// one image row ARE a core matrix, so the A operand of tap (r,s) is simply the same halo tile read
// at start address + (r*HALO_W + s)*16 B with SBO = one halo row.  The halo tile is fetched once
// and reused by all 9 taps (no im2col materialisation, no 9x re-fetch).
//
// Warp roles (576 threads per CTA): warp 0 = TMA producer, warp 1 = TMEM owner + (leader CTA only) MMA issuer,
// warps 2..17 = epilogue (TMEM lane quarter = warp%4).  Two TMEM accumulator buffers (2 x 256 columns) let the
// epilogue of tile i overlap the MMAs of tile i+1.  Persistent grid: 74 pairs, static round-robin tile order.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace bnerv {

constexpr int TILE_H      = 16;                   // pixel rows per super-tile
constexpr int HALO_H      = TILE_H + 2;
constexpr int TMEM_COLS   = 512;
constexpr int BUF_COLS    = 256;                  // TMEM columns per accumulator buffer
constexpr int N_EPI_WARPS = 16;
constexpr int N_THREADS   = 64 + 32 * N_EPI_WARPS;
constexpr int MAX_STAGES  = 12;
constexpr int MIN_STAGES  = 3;
constexpr int SMEM_LIMIT  = 227 * 1024;
constexpr int N_BARS      = 2 * MAX_STAGES + 6;   // full[], empty[], tfull[2], tempty[2], wfull, wempty
constexpr int BAR_BYTES   = N_BARS * 8 + 16;      // + TMEM base slot
constexpr int CST_N       = 256;                  // per-tile epilogue constants, see Cst
constexpr int MAX_GROUPS  = CST_N / 16;

template <int MT_>
struct Geo {
    static constexpr int MT         = MT_;
    static constexpr int TILE_W     = 8 * MT;
    static constexpr int HALO_W     = TILE_W + 2;
    static constexpr int A_GROUP_B  = HALO_H * HALO_W * 16;  // bytes of one 8-channel group of the halo tile
    static constexpr int A_STAGE_B  = 2 * A_GROUP_B;         // one K step = 16 channels = 2 groups (128-byte multiple)
    static constexpr int ACC_STRIDE = BUF_COLS / MT;         // TMEM columns between the MT accumulators of a buffer
    static constexpr int CS         = 4 / MT;                // epilogue warps sharing one (lane quarter, row block)
};

struct ChunkInfo {          // 8 consecutive packed rows of the current n-tile = 8 channels of one output position
    long long goff;         // output offset (halves) relative to the pixel's (i=j=0) position in channel group 0
    int cc;                 // first output channel
    int ij;                 // PixelShuffle sub-position: i | (j << 16)
};
struct Cst {                // staged per tile by the epilogue warps, read back as LDS broadcasts
    float bias[CST_N], g1p[CST_N], beta[CST_N];
    ChunkInfo chk[2 * MAX_GROUPS];
};

struct ConvTcArgs {
    int B, H, W;            // conv-resolution geometry (input == pre-shuffle output)
    int cin_groups;         // Cin_p / 8
    int ksteps;             // Cin_p / 16
    int taps;               // 1 or 9
    int n_total;            // s*s*Cout_p
    int n_acc;              // packed rows per n-tile (multiple of 16, <= 256/MT)
    int n_half;             // n_acc / 2 : rows of B each CTA of the pair keeps
    int n_tiles;            // ceil(n_total / n_acc)
    int cout, cout_p;       // real / padded output channels
    int s;                  // PixelShuffle factor
    int act;
    int flags;              // F_* epilogue features (used by the generic instantiation)
    int tiles_x, tiles_y;   // super-tiles per image
    int pairs_x;            // ceil(tiles_x / 2)
    int pair_tiles;         // B * tiles_y * pairs_x : work items per n-tile
    int work_total;         // n_tiles * pair_tiles, n-tile major; pair p owns the contiguous range [p, p+1) * total / n_pairs
    int n_pairs;            // CTA pairs in the grid
    int stages;
    int b_kstep_bytes;      // taps * 2 * n_half * 16 : one K step of the resident weight half
    int b_bytes;            // ksteps * b_kstep_bytes (0 in streaming mode)
    int b_stream;           // 1: weights do not stay resident; every pipeline stage carries its K step of B next to A
    int stage_bytes;        // A stage (+ one K step of the weight half in streaming mode)
    const float* bias;      // [n_total]
    const float* g1p;       // [B][cout_p] or null
    const float* beta;      // [B][cout_p] or null
    const __half* resid;    // C8 at output resolution or null
    __half* out_pre;        // C8 or null
    __half* out_aff;        // C8 or null
    float* out_nchw;        // NCHW f32 or null
    __half* out_deriv;      // C8 or null: act'(pre-activation) (training forward)
    long long split_stride; // halves between the hi and the lo block of a split map: (Cout_p/8) * Ho * Wo * 8
};

struct TileCoord { int b, h0, w0; };

template <int MT>
__device__ __forceinline__ TileCoord decode_tile(const ConvTcArgs& a, int pt, int rank) {
    TileCoord t;
    const int px = pt % a.pairs_x;
    int rest = pt / a.pairs_x;
    const int ty = rest % a.tiles_y;
    t.b  = rest / a.tiles_y;
    t.h0 = ty * TILE_H;
    t.w0 = (2 * px + rank) * Geo<MT>::TILE_W;      // may lie entirely outside the image (odd tiles_x): TMA zero-fills
    return t;
}

// Work split: the (n-tile major) item list is cut into one contiguous range per pair, so a pair switches its resident
// weights at most (range / pair_tiles + 1) times and pairs ~n_pairs/n_tiles apart walk the same pixel tiles at about
// the same time (the activation tile is then served from L2 to all of them).
struct WorkRange { int begin, end; };
__device__ __forceinline__ WorkRange work_range(const ConvTcArgs& a, uint32_t pair) {
    WorkRange r;
    r.begin = static_cast<int>(static_cast<long long>(a.work_total) * pair / a.n_pairs);
    r.end   = static_cast<int>(static_cast<long long>(a.work_total) * (pair + 1) / a.n_pairs);
    return r;
}

// epilogue feature flags (compile-time in the specialised instantiations, run-time in the generic one)
constexpr int F_RESID = 1, F_AFF = 2, F_PRE = 4, F_NCHW = 8, F_SHUF = 16;
constexpr int F_WIDE = 32;   // s == 2 row packing: the two chunks of a group are 32 contiguous output bytes
constexpr int F_DERIV = 64;  // also write act'(pre-activation) (training forward)
constexpr int F_HEAD = 128;  // 3x3 conv to <= 3 channels as ONE 1x1 contraction to 9*Cout columns + shift-sum (see mma_role_head)
// Split ("precise") maps, generic instantiation only: a value v is kept as the f16 pair hi = f16(v), lo = f16(v - hi) (~22
// significant bits) and a map is laid out [hi | lo | hi] over 3*Cout_p channels, so that the NEXT conv - an ordinary conv over
// 3*Cout_p input channels with the weight blocks [W_hi ; W_hi ; W_lo] - accumulates hi*W_hi + lo*W_hi + hi*W_lo in f32.
constexpr int F_SPLIT_OUT = 256;   // out_pre / out_aff are split maps
constexpr int F_SPLIT_RES = 512;   // resid is a split map (hi + lo are both added)

// Head mode geometry: P[q][(tap, c)] for the 18 x 18 halo pixels q of a 16 x 16 output tile, staged in shared memory
constexpr int HEAD_N     = 32;                    // UMMA N (27 used for Cout = 3)
constexpr int HEAD_PSTR  = 27;                    // floats per halo pixel (odd: conflict-free across a pixel row)
constexpr int HEAD_PBUF_FLOATS = 18 * 18 * HEAD_PSTR;
constexpr int HEAD_SMEM_BYTES  = 2 * HEAD_PBUF_FLOATS * 4;

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

// 256-bit store (STG.E.ENL2.256): both horizontal PixelShuffle neighbours of a pixel's 8 channels in one request.
__device__ __forceinline__ void st_global_32B(__half* p, const uint4& lo, const uint4& hi) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
}

__device__ __forceinline__ uint4 pack8(const float2* x) {
    uint4 o;
    o.x = pack_h2_satfinite(x[0]); o.y = pack_h2_satfinite(x[1]);
    o.z = pack_h2_satfinite(x[2]); o.w = pack_h2_satfinite(x[3]);
    return o;
}

// lo half of the split representation of 8 values whose hi half is `hi`
__device__ __forceinline__ uint4 pack8_lo(const float2* x, const uint4& hi) {
    float2 d[4];
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float2 f = unpack_h2(h[p]);
        d[p] = make_float2(x[p].x - f.x, x[p].y - f.y);
    }
    return pack8(d);
}

// the lo and second hi block of a split output map (the first hi block is stored by the regular path)
__device__ __forceinline__ void store_split_tail(__half* out, size_t off0, size_t off1, long long ss, bool wide,
                                                 const float2* x, const uint4& o0, const uint4& o1) {
    const uint4 l0 = pack8_lo(x, o0), l1 = pack8_lo(x + 4, o1);
    if (wide) {
        st_global_32B(out + off0 + ss, l0, l1);
        st_global_32B(out + off0 + 2 * ss, o0, o1);
    } else {
        *reinterpret_cast<uint4*>(out + off0 + ss) = l0;
        *reinterpret_cast<uint4*>(out + off1 + ss) = l1;
        *reinterpret_cast<uint4*>(out + off0 + 2 * ss) = o0;
        *reinterpret_cast<uint4*>(out + off1 + 2 * ss) = o1;
    }
}

// Epilogue of one 16-column accumulator group of one pixel = two chunks of 8 channels (chunk hh = columns hh*8..+7):
// bias + activation (+ residual) (+ TAT affine) and the stores.  Constants are LDS broadcasts from the tile's Cst;
// the bias is fetched by the caller BEFORE the accumulator wait, scale/shift are requested right after the bias add
// so their latency hides behind the activation math.
struct GroupAddr { size_t off[2]; int cc[2], ho[2], wo[2]; };

template <int ACT, int FLAGS>
__device__ __forceinline__ void epilogue_group(const ConvTcArgs& a, int flags, const uint32_t* v, const float4* bs,
                                               const Cst* cb, int col, int b, const GroupAddr& ga, bool valid,
                                               const uint4* rr, int Ho, int Wo, long long res_adj = 0) {
    float2 x[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        x[2 * p]     = add2(make_float2(__uint_as_float(v[4 * p]),     __uint_as_float(v[4 * p + 1])), make_float2(bs[p].x, bs[p].y));
        x[2 * p + 1] = add2(make_float2(__uint_as_float(v[4 * p + 2]), __uint_as_float(v[4 * p + 3])), make_float2(bs[p].z, bs[p].w));
    }
    float4 gg[4], ee[4];
    if (flags & F_AFF) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            gg[p] = *reinterpret_cast<const float4*>(cb->g1p + col + 4 * p);
            ee[p] = *reinterpret_cast<const float4*>(cb->beta + col + 4 * p);
        }
    }
    float2 dv[8];
    if (flags & F_DERIV) {
#pragma unroll
        for (int p = 0; p < 8; ++p) x[p] = act2_with_deriv(x[p], (ACT >= 0) ? ACT : a.act, dv[p]);
    } else {
#pragma unroll
        for (int p = 0; p < 8; ++p) x[p] = act2_rt<ACT>(x[p], a.act);
    }
    if (flags & F_RESID) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            x[4 * hh + 0] = add2(x[4 * hh + 0], unpack_h2(rr[hh].x)); x[4 * hh + 1] = add2(x[4 * hh + 1], unpack_h2(rr[hh].y));
            x[4 * hh + 2] = add2(x[4 * hh + 2], unpack_h2(rr[hh].z)); x[4 * hh + 3] = add2(x[4 * hh + 3], unpack_h2(rr[hh].w));
        }
        if ((flags & F_SPLIT_RES) && valid) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint4 rl = __ldg(reinterpret_cast<const uint4*>(a.resid + ga.off[hh] + res_adj + a.split_stride));
                x[4 * hh + 0] = add2(x[4 * hh + 0], unpack_h2(rl.x)); x[4 * hh + 1] = add2(x[4 * hh + 1], unpack_h2(rl.y));
                x[4 * hh + 2] = add2(x[4 * hh + 2], unpack_h2(rl.z)); x[4 * hh + 3] = add2(x[4 * hh + 3], unpack_h2(rl.w));
            }
        }
    }
    if (!valid) return;
    if (flags & F_DERIV) {
        const uint4 o0 = pack8(dv), o1 = pack8(dv + 4);
        if (flags & F_WIDE) {
            st_global_32B(a.out_deriv + ga.off[0], o0, o1);
        } else {
            *reinterpret_cast<uint4*>(a.out_deriv + ga.off[0]) = o0;
            *reinterpret_cast<uint4*>(a.out_deriv + ga.off[1]) = o1;
        }
    }
    if (flags & F_PRE) {
        const uint4 o0 = pack8(x), o1 = pack8(x + 4);
        if (flags & F_WIDE) {
            st_global_32B(a.out_pre + ga.off[0], o0, o1);
        } else {
            *reinterpret_cast<uint4*>(a.out_pre + ga.off[0]) = o0;
            *reinterpret_cast<uint4*>(a.out_pre + ga.off[1]) = o1;
        }
        if (flags & F_SPLIT_OUT) store_split_tail(a.out_pre, ga.off[0], ga.off[1], a.split_stride, (flags & F_WIDE) != 0, x, o0, o1);
    }
    if (flags & F_NCHW) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const float xs[8] = {x[4 * hh].x, x[4 * hh].y, x[4 * hh + 1].x, x[4 * hh + 1].y,
                                 x[4 * hh + 2].x, x[4 * hh + 2].y, x[4 * hh + 3].x, x[4 * hh + 3].y};
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (ga.cc[hh] + k < a.cout)
                    a.out_nchw[(static_cast<size_t>(b * a.cout + ga.cc[hh] + k) * Ho + ga.ho[hh]) * Wo + ga.wo[hh]] = xs[k];
        }
    }
    if (flags & F_AFF) {
        float2 y[8];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            y[2 * p]     = fma2(x[2 * p],     make_float2(gg[p].x, gg[p].y), make_float2(ee[p].x, ee[p].y));
            y[2 * p + 1] = fma2(x[2 * p + 1], make_float2(gg[p].z, gg[p].w), make_float2(ee[p].z, ee[p].w));
        }
        const uint4 o0 = pack8(y), o1 = pack8(y + 4);
        if (flags & F_WIDE) {
            st_global_32B(a.out_aff + ga.off[0], o0, o1);
        } else {
            *reinterpret_cast<uint4*>(a.out_aff + ga.off[0]) = o0;
            *reinterpret_cast<uint4*>(a.out_aff + ga.off[1]) = o1;
        }
        if (flags & F_SPLIT_OUT) store_split_tail(a.out_aff, ga.off[0], ga.off[1], a.split_stride, (flags & F_WIDE) != 0, y, o0, o1);
    }
}

// Same for ONE chunk of 8 channels (narrow layers: n_acc == 16, where the two chunks of the only accumulator group go to
// different warps so that all 16 epilogue warps work instead of 8).
template <int ACT, int FLAGS>
__device__ __forceinline__ void epilogue_chunk(const ConvTcArgs& a, int flags, const uint32_t* v, const Cst* cb, int col, int b,
                                               size_t off, int cc, int ho, int wo, bool valid, const uint4& rr, int Ho, int Wo,
                                               long long res_adj = 0) {
    const float4 b0 = *reinterpret_cast<const float4*>(cb->bias + col), b1 = *reinterpret_cast<const float4*>(cb->bias + col + 4);
    float2 x[4];
    x[0] = add2(make_float2(__uint_as_float(v[0]), __uint_as_float(v[1])), make_float2(b0.x, b0.y));
    x[1] = add2(make_float2(__uint_as_float(v[2]), __uint_as_float(v[3])), make_float2(b0.z, b0.w));
    x[2] = add2(make_float2(__uint_as_float(v[4]), __uint_as_float(v[5])), make_float2(b1.x, b1.y));
    x[3] = add2(make_float2(__uint_as_float(v[6]), __uint_as_float(v[7])), make_float2(b1.z, b1.w));
    float4 gg[2], ee[2];
    if (flags & F_AFF) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            gg[p] = *reinterpret_cast<const float4*>(cb->g1p + col + 4 * p);
            ee[p] = *reinterpret_cast<const float4*>(cb->beta + col + 4 * p);
        }
    }
    float2 dv[4];
    if (flags & F_DERIV) {
#pragma unroll
        for (int p = 0; p < 4; ++p) x[p] = act2_with_deriv(x[p], (ACT >= 0) ? ACT : a.act, dv[p]);
    } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) x[p] = act2_rt<ACT>(x[p], a.act);
    }
    if (flags & F_RESID) {
        x[0] = add2(x[0], unpack_h2(rr.x)); x[1] = add2(x[1], unpack_h2(rr.y));
        x[2] = add2(x[2], unpack_h2(rr.z)); x[3] = add2(x[3], unpack_h2(rr.w));
        if ((flags & F_SPLIT_RES) && valid) {
            const uint4 rl = __ldg(reinterpret_cast<const uint4*>(a.resid + off + res_adj + a.split_stride));
            x[0] = add2(x[0], unpack_h2(rl.x)); x[1] = add2(x[1], unpack_h2(rl.y));
            x[2] = add2(x[2], unpack_h2(rl.z)); x[3] = add2(x[3], unpack_h2(rl.w));
        }
    }
    if (!valid) return;
    if (flags & F_DERIV) *reinterpret_cast<uint4*>(a.out_deriv + off) = pack8(dv);
    if (flags & F_PRE) {
        const uint4 o = pack8(x);
        *reinterpret_cast<uint4*>(a.out_pre + off) = o;
        if (flags & F_SPLIT_OUT) {
            *reinterpret_cast<uint4*>(a.out_pre + off + a.split_stride) = pack8_lo(x, o);
            *reinterpret_cast<uint4*>(a.out_pre + off + 2 * a.split_stride) = o;
        }
    }
    if (flags & F_NCHW) {
        const float xs[8] = {x[0].x, x[0].y, x[1].x, x[1].y, x[2].x, x[2].y, x[3].x, x[3].y};
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (cc + k < a.cout) a.out_nchw[(static_cast<size_t>(b * a.cout + cc + k) * Ho + ho) * Wo + wo] = xs[k];
    }
    if (flags & F_AFF) {
        float2 y[4];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            y[2 * p]     = fma2(x[2 * p],     make_float2(gg[p].x, gg[p].y), make_float2(ee[p].x, ee[p].y));
            y[2 * p + 1] = fma2(x[2 * p + 1], make_float2(gg[p].z, gg[p].w), make_float2(ee[p].z, ee[p].w));
        }
        const uint4 o = pack8(y);
        *reinterpret_cast<uint4*>(a.out_aff + off) = o;
        if (flags & F_SPLIT_OUT) {
            *reinterpret_cast<uint4*>(a.out_aff + off + a.split_stride) = pack8_lo(y, o);
            *reinterpret_cast<uint4*>(a.out_aff + off + 2 * a.split_stride) = o;
        }
    }
}

struct Pipe {
    uint32_t w_base;         // shared-window address of the resident weight half
    uint32_t a_base;         // ... of the activation stage ring
    uint32_t full, empty;    // ... of full_bar[0] / empty_bar[0]
    uint32_t tfull, tempty;  // ... of tfull_bar[0] / tempty_bar[0]
    uint32_t wfull, wempty;
    uint32_t tmem_base;
    uint32_t rank;           // CTA rank in the pair (0 = leader)
    uint32_t pair;           // pair index in the grid
};

// ===================== TMA producer (warp 0 of both CTAs, converged; one elected lane issues) =====================
// Every load of either CTA credits its bytes to the LEADER's barrier (the leader's MMA thread is the only consumer);
// the leader alone arms the barrier with the pair's total byte count.
template <int MT>
__device__ __forceinline__ void producer_role(const ConvTcArgs& a, const Pipe& p, const CUtensorMap* tmA, const CUtensorMap* tmB) {
    using G = Geo<MT>;
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t full_leader  = map_to_cta(p.full, 0);
    const uint32_t wfull_leader = map_to_cta(p.wfull, 0);
    const WorkRange wr = work_range(a, p.pair);
    int wc = 0;                                   // weight loads issued so far
    if (wr.begin >= wr.end) { pdl_wait(); pdl_launch_dependents(); }
    for (int it = wr.begin; it < wr.end; ++it) {
        const int nt = it / a.pair_tiles, pt = it - nt * a.pair_tiles;
        if (a.b_stream) {
            if (it == wr.begin) { pdl_wait(); pdl_launch_dependents(); }
        } else if (it == wr.begin || pt == 0) {
            // the previous n-tile's MMAs (which read the resident weights of both CTAs) must have completed
            if (wc > 0) mbar_wait(p.wempty, (wc - 1) & 1);
            if (elect_one()) {
                if (p.rank == 0) mbar_expect_tx(p.wfull, 2u * a.b_bytes);
                const int row0 = nt * a.n_acc + p.rank * a.n_half;
                for (int kc = 0; kc < a.ksteps; ++kc)
                    tma_load_3d_pair(p.w_base + kc * a.b_kstep_bytes, tmB, wfull_leader, 2 * row0, 2 * kc, 0);
            }
            __syncwarp();
            ++wc;
            if (it == wr.begin) {
                // The first weight set was requested above, overlapping the previous kernel's tail (programmatic
                // dependent launch; packed weights are never written by a kernel that triggers early).  Activations
                // are the previous kernel's output: wait for it to complete before the first halo-tile load.
                pdl_wait();
                pdl_launch_dependents();
            }
        }
        {
            const TileCoord t = decode_tile<MT>(a, pt, p.rank);
            for (int kc = 0; kc < a.ksteps; ++kc) {
                mbar_wait(p.empty + stage * 8, phase ^ 1);
                if (elect_one()) {
                    if (p.rank == 0) mbar_expect_tx(p.full + stage * 8, 2u * a.stage_bytes);
                    // activations viewed as u64 elements: 2 per pixel-group -> x coordinate = 2*w
                    tma_load_3d_pair(p.a_base + stage * a.stage_bytes, tmA, full_leader + stage * 8,
                                     2 * (t.w0 - 1), t.h0 - 1, t.b * a.cin_groups + 2 * kc);
                    if (a.b_stream)      // this K step of this CTA's half of the n-tile's weights rides in the same stage
                        tma_load_3d_pair(p.a_base + stage * a.stage_bytes + G::A_STAGE_B, tmB, full_leader + stage * 8,
                                         2 * (nt * a.n_acc + static_cast<int>(p.rank) * a.n_half), 2 * kc, 0);
                }
                __syncwarp();
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    }
}

// ===================== MMA issuer (warp 1 of the leader CTA, converged; one elected lane issues) =====================
// Everything the issue needs is warp-uniform and the taps are unrolled with constant descriptor offsets, so one
// K step (TAPS x MT UMMAs of M = 256) is a straight run of UTCHMMA separated by a few uniform adds.
template <int MT, int TAPS>
__device__ __forceinline__ void mma_role(const ConvTcArgs& a, const Pipe& p) {
    using G = Geo<MT>;
    int stage = 0;
    uint32_t phase = 0;
    int abuf = 0;
    uint32_t aphase = 0;
    const uint32_t idesc   = umma_idesc_f16(256, a.n_acc);
    const uint32_t b_tap16 = 2u * a.n_half;                     // one tap's [2 groups][n_half][16 B] slab, in 16-byte units
    const uint32_t b_ks16  = static_cast<uint32_t>(a.b_kstep_bytes) >> 4;
    const uint64_t a_hi    = umma_desc_hi_noswz(G::A_GROUP_B, G::HALO_W * 16);
    const uint64_t b_hi    = umma_desc_hi_noswz(a.n_half * 16u, 128u);
    const uint32_t wb16    = (p.w_base & 0x3FFFFu) >> 4;
    const uint32_t ab16    = (p.a_base & 0x3FFFFu) >> 4;
    const WorkRange wr = work_range(a, p.pair);
    int wc = 0;                                   // weight sets consumed so far
    for (int it = wr.begin; it < wr.end; ++it) {
        const int pt = it % a.pair_tiles;
        if (!a.b_stream && (it == wr.begin || pt == 0)) {
            if (wc > 0) {                         // every MMA that reads the old weights has been issued: release them
                if (elect_one()) umma_commit_pair(p.wempty);
                __syncwarp();
            }
            mbar_wait(p.wfull, wc & 1);
            tc_fence_after();
            ++wc;
        }
        {
            mbar_wait(p.tempty + abuf * 8, aphase ^ 1);
            tc_fence_after();
            const uint32_t d0 = p.tmem_base + abuf * BUF_COLS;
            for (int kc = 0; kc < a.ksteps; ++kc) {
                mbar_wait(p.full + stage * 8, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa16 = ab16 + stage * (static_cast<uint32_t>(a.stage_bytes) >> 4);
                    const uint32_t sb16 = a.b_stream ? sa16 + (G::A_STAGE_B >> 4) : wb16 + kc * b_ks16;
#pragma unroll
                    for (int tp = 0; tp < TAPS; ++tp) {
                        const int tap = (TAPS == 1) ? 4 : tp;   // 1x1 conv: the single tap reads the halo tile's centre
                        const int r = tap / 3, sx = tap % 3;
                        const uint64_t bdesc = b_hi | static_cast<uint64_t>(sb16 + tp * b_tap16);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint64_t adesc = a_hi | static_cast<uint64_t>(sa16 + (r * G::HALO_W + sx + mt * 8));
                            umma_f16_pair(d0 + mt * G::ACC_STRIDE, adesc, bdesc, idesc, (tp > 0) ? 1u : (kc > 0 ? 1u : 0u));
                        }
                    }
                    umma_commit_pair(p.empty + stage * 8);
                    if (kc == a.ksteps - 1) umma_commit_pair(p.tfull + abuf * 8);
                }
                __syncwarp();
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
            abuf ^= 1;
            if (abuf == 0) aphase ^= 1;
        }
    }
}

// ===================== MMA issuer, head mode =====================
// out[p, c] = sum_tap sum_k W[c][k][tap] X[p + tap, k]  is evaluated as  P[q, (tap, c)] = sum_k X[q, k] W[c][k][tap]
// for every HALO pixel q (one 1x1 contraction, N = 9*Cout <= 32) followed by out[p, c] = sum_tap P[p + tap, (tap, c)]
// in the epilogue: 6 UMMAs (M = 256, N = 32) per K step cover the 18 x 18 halo with 16-row x 8-px blocks at row
// offsets {0, 2} and px offsets {0, 8, 10} - against 18 UMMAs (9 taps x 2 blocks, N = 16) of the generic path, whose
// cost is the A-operand read, not the math.
template <int MT>
__device__ __forceinline__ void mma_role_head(const ConvTcArgs& a, const Pipe& p) {
    using G = Geo<MT>;
    int stage = 0;
    uint32_t phase = 0;
    int abuf = 0;
    uint32_t aphase = 0;
    const uint32_t idesc   = umma_idesc_f16(256, HEAD_N);
    const uint32_t b_ks16  = static_cast<uint32_t>(a.b_kstep_bytes) >> 4;
    const uint64_t a_hi    = umma_desc_hi_noswz(G::A_GROUP_B, G::HALO_W * 16);
    const uint64_t b_hi    = umma_desc_hi_noswz(a.n_half * 16u, 128u);
    const uint32_t wb16    = (p.w_base & 0x3FFFFu) >> 4;
    const uint32_t ab16    = (p.a_base & 0x3FFFFu) >> 4;
    const WorkRange wr = work_range(a, p.pair);
    if (wr.begin < wr.end) {
        mbar_wait(p.wfull, 0);                    // one n-tile: the weights are loaded once
        tc_fence_after();
    }
    for (int it = wr.begin; it < wr.end; ++it) {
        mbar_wait(p.tempty + abuf * 8, aphase ^ 1);
        tc_fence_after();
        const uint32_t d0 = p.tmem_base + abuf * BUF_COLS;
        for (int kc = 0; kc < a.ksteps; ++kc) {
            mbar_wait(p.full + stage * 8, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa16 = ab16 + stage * (static_cast<uint32_t>(a.stage_bytes) >> 4);
                const uint64_t bdesc = b_hi | static_cast<uint64_t>(wb16 + kc * b_ks16);
#pragma unroll
                for (int blk = 0; blk < 6; ++blk) {
                    const int ro = (blk / 3) * 2, po = (blk % 3 == 0) ? 0 : (blk % 3 == 1 ? 8 : 10);
                    const uint64_t adesc = a_hi | static_cast<uint64_t>(sa16 + ro * G::HALO_W + po);
                    umma_f16_pair(d0 + blk * HEAD_N, adesc, bdesc, idesc, kc > 0 ? 1u : 0u);
                }
                umma_commit_pair(p.empty + stage * 8);
                if (kc == a.ksteps - 1) umma_commit_pair(p.tfull + abuf * 8);
            }
            __syncwarp();
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        abuf ^= 1;
        if (abuf == 0) aphase ^= 1;
    }
}

// ===================== epilogue, head mode (all 16 epilogue warps) =====================
template <int MT, int ACT>
__device__ __forceinline__ void epilogue_head(const ConvTcArgs& a, const Pipe& p, float* pbuf, int warp, int lane) {
    const int e   = warp - 2;
    const int q   = warp & 3;
    const int sub = e >> 2;                         // 0..3: blocks {sub, sub + 4}
    const int m   = q * 32 + lane;                  // row of a 16-row x 8-px P block
    const int et  = threadIdx.x - 64;
    const uint32_t tempty_leader = map_to_cta(p.tempty, 0);
    const WorkRange wr = work_range(a, p.pair);
    int abuf = 0;
    uint32_t aphase = 0;
    float bias_c[2] = {0.0f, 0.0f};                 // this thread's output channels: c = (et >> 8) + 2*i
    for (int i = 0; i < 2; ++i) {
        const int c = (et >> 8) + 2 * i;
        if (c < a.cout) bias_c[i] = __ldg(a.bias + c);
    }
    pdl_wait();
    for (int it = wr.begin; it < wr.end; ++it) {
        const TileCoord t = decode_tile<MT>(a, it, p.rank);      // n_tiles == 1: item == pixel pair-tile
        float* pb = pbuf + abuf * HEAD_PBUF_FLOATS;
        mbar_wait(p.tfull + abuf * 8, aphase);
        tc_fence_after();
        // phase A: P blocks TMEM -> shared memory (only the pixels a block is the first to cover)
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            const int blk = sub + 4 * rep;
            if (blk < 6) {                           // warp-uniform
                const int ro = (blk / 3) * 2, pj = blk % 3, po = (pj == 0) ? 0 : (pj == 1 ? 8 : 10);
                const uint32_t taddr = p.tmem_base + (static_cast<uint32_t>(q * 32) << 16) + abuf * BUF_COLS + blk * HEAD_N;
                uint32_t v0[16], v1[16];
                tmem_ld16(taddr, v0);
                tmem_ld16(taddr + 16, v1);
                tmem_ld_wait();
                const bool fresh = (ro == 0 || (m >> 3) >= 14) && (pj != 2 || (m & 7) >= 6);
                if (fresh) {
                    float* dst = pb + ((ro + (m >> 3)) * 18 + po + (m & 7)) * HEAD_PSTR;
#pragma unroll
                    for (int k = 0; k < 16; ++k) dst[k] = __uint_as_float(v0[k]);
#pragma unroll
                    for (int k = 0; k < HEAD_PSTR - 16; ++k) dst[16 + k] = __uint_as_float(v1[k]);
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + abuf * 8);       // TMEM buffer free as soon as it is copied out
        named_bar_sync(1, 32 * N_EPI_WARPS);
        // phase B: shift-sum of the 9 taps, bias, activation, NCHW f32 store
        {
            const int pidx = et & 255, hh = pidx >> 4, ww = pidx & 15;
            const int h = t.h0 + hh, w = t.w0 + ww;
            const bool valid = (h < a.H) && (w < a.W);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c = (et >> 8) + 2 * i;
                if (c < a.cout) {
                    float acc = bias_c[i];
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
                        acc += pb[((hh + tap / 3) * 18 + ww + tap % 3) * HEAD_PSTR + tap * a.cout + c];
                    const float2 o = act2_rt<ACT>(make_float2(acc, 0.0f), a.act);
                    if (valid) a.out_nchw[(static_cast<size_t>(t.b * a.cout + c) * a.H + h) * a.W + w] = o.x;
                }
            }
        }
        abuf ^= 1;
        if (abuf == 0) aphase ^= 1;
    }
}

template <int MT, int ACT, int FLAGS, bool NARROW>
__global__ void __launch_bounds__(N_THREADS, 1)     // 18 warps = 5 on two of the four SM sub-partitions (16 K registers each): <= 96 registers/thread
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcArgs a) {
    using G = Geo<MT>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int flags = (FLAGS >= 0) ? FLAGS : a.flags;

    uint8_t* a_ring       = smem + a.b_bytes;
    uint8_t* bar_base     = a_ring + static_cast<size_t>(a.stages) * a.stage_bytes;
    uint64_t* full_bar    = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar   = full_bar + MAX_STAGES;
    uint64_t* tfull_bar   = empty_bar + MAX_STAGES;   // [2]
    uint64_t* tempty_bar  = tfull_bar + 2;            // [2]
    uint64_t* wfull_bar   = tempty_bar + 2;
    uint64_t* wempty_bar  = wfull_bar + 1;
    uint32_t* tmem_slot   = reinterpret_cast<uint32_t*>(wempty_bar + 1);
    Cst*      cst         = reinterpret_cast<Cst*>(bar_base + BAR_BYTES);     // [2]

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&tfull_bar[i]), 1);
            mbar_init(smem_u32(&tempty_bar[i]), 2 * N_EPI_WARPS);      // the epilogue warps of BOTH CTAs (leader's copy is used)
        }
        mbar_init(smem_u32(wfull_bar), 1);
        mbar_init(smem_u32(wempty_bar), 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();            // barriers of both CTAs initialised before any remote arrive / TMA credit
    tc_fence_after();

    Pipe p;
    p.w_base = smem_u32(smem);
    p.a_base = smem_u32(a_ring);
    p.full = smem_u32(full_bar);     p.empty = smem_u32(empty_bar);
    p.tfull = smem_u32(tfull_bar);   p.tempty = smem_u32(tempty_bar);
    p.wfull = smem_u32(wfull_bar);   p.wempty = smem_u32(wempty_bar);
    p.tmem_base = *tmem_slot;
    // clusters of (2,1,1) are consecutive blocks in x: taking rank / pair index from blockIdx (not from the %cluster_ctarank /
    // %clusterid.x special registers) lets the compiler keep all tile arithmetic derived from them - work ranges, the
    // div/mod of the tile decode - on the uniform datapath instead of replicating ~150 integer instructions per tile in
    // every thread (half of a narrow layer's epilogue instruction stream)
    p.rank = blockIdx.x & 1u;
    p.pair = blockIdx.x >> 1;

    if (warp == 0) {
        producer_role<MT>(a, p, &tmA, &tmB);
    } else if (warp == 1) {
        if (p.rank == 0) {
            if (flags & F_HEAD) mma_role_head<MT>(a, p);
            else if (a.taps == 9) mma_role<MT, 9>(a, p);
            else mma_role<MT, 1>(a, p);
        }
    } else {
      if (flags & F_HEAD) {
        epilogue_head<MT, ACT>(a, p, reinterpret_cast<float*>(cst + 2), warp, lane);
      } else {
        // ===================== epilogue: 16 warps per CTA =====================
        // warp -> (TMEM lane quarter q = warp%4 [hardware rule], row block mt, column slot cs of CS):
        // a warp owns the 16-column groups g16 = cs, cs+CS, cs+2CS, cs+3CS of its row block.
        const int e    = warp - 2;
        const int q    = warp & 3;
        const int sub  = e >> 2;                        // 0..3
        const int mt   = (MT == 2) ? (sub >> 1) : 0;
        const int cs   = (MT == 2) ? (sub & 1) : sub;
        const int m    = q * 32 + lane;                 // row of the UMMA row block == pixel
        const int et   = threadIdx.x - 64;              // 0 .. 32*N_EPI_WARPS-1
        int abuf = 0;
        uint32_t aphase = 0;
        const int s  = a.s;
        const int Ho = a.H * s, Wo = a.W * s;
        const int cout_groups = a.cout_p >> 3;
        const size_t plane = static_cast<size_t>(Ho) * Wo * 8;      // halves per 8-channel plane
        const uint32_t tempty_leader = map_to_cta(p.tempty, 0);
        const WorkRange wr = work_range(a, p.pair);
        int cst_key0 = -1, cst_key1 = -1;        // (n-tile, batch) the two constant buffers currently hold
        pdl_wait();            // residual / TAT tables come from earlier kernels; our stores must not overtake their readers
        for (int it = wr.begin; it < wr.end; ++it) {
            const int nt = it / a.pair_tiles, pt = it - nt * a.pair_tiles;
            const int n0 = nt * a.n_acc;
            {
                const TileCoord t = decode_tile<MT>(a, pt, p.rank);
                const int h = t.h0 + (m >> 3);
                const int w = t.w0 + mt * 8 + (m & 7);
                const bool valid = (h < a.H) && (w < a.W);
                const size_t base_b = static_cast<size_t>(t.b) * cout_groups * plane * ((flags & F_SPLIT_OUT) ? 3 : 1);
                // a residual map whose batch stride differs from the output's (split vs plain)
                const long long res_adj = static_cast<long long>(t.b) * cout_groups * static_cast<long long>(plane) *
                                          (((flags & F_SPLIT_RES) ? 3 : 1) - ((flags & F_SPLIT_OUT) ? 3 : 1));
                const size_t pix = (static_cast<size_t>(h) * s * Wo + static_cast<size_t>(w) * s) * 8;

                // Stage the per-row constants (bias, TAT scale+1, TAT shift) and per-16-column-group addressing in
                // shared memory.  They depend on (n-tile, batch index) only, so a buffer is re-staged only when that
                // key changes: narrow layers (one K step per tile) would otherwise pay a global-load latency and a
                // 512-thread barrier per tile - 2/3 of their epilogue time.
                Cst* cb = cst + abuf;
                const int ckey = nt * a.B + t.b;
                const bool restage = ((abuf ? cst_key1 : cst_key0) != ckey);     // uniform over the epilogue warps
                if (restage) {
                  if (abuf) cst_key1 = ckey; else cst_key0 = ckey;
                  named_bar_sync(1, 32 * N_EPI_WARPS);                     // every warp is done reading this buffer (any earlier tile)
                  if (et < a.n_acc) {
                    const int nn = n0 + et;
                    float bv = 0.0f, gv = 0.0f, ev = 0.0f;
                    if (nn < a.n_total) {
                        bv = __ldg(a.bias + nn);
                        if (flags & F_AFF) {
                            int cc = nn, i, j;
                            if (flags & F_SHUF) packed_row_to_cij(nn, s, a.cout_p, cc, i, j);
                            gv = __ldg(a.g1p + static_cast<size_t>(t.b) * a.cout_p + cc);
                            ev = __ldg(a.beta + static_cast<size_t>(t.b) * a.cout_p + cc);
                        }
                    }
                    cb->bias[et] = bv; cb->g1p[et] = gv; cb->beta[et] = ev;
                  } else if (et >= 256 && et < 256 + 2 * MAX_GROUPS) {
                    const int ch = et - 256;                    // chunk = 8 packed rows
                    const int nn = n0 + ch * 8;
                    int cc = nn, i = 0, j = 0;
                    if (flags & F_SHUF) packed_row_to_cij(nn, s, a.cout_p, cc, i, j);
                    ChunkInfo ci;
                    ci.goff = static_cast<long long>(cc >> 3) * static_cast<long long>(plane) + (static_cast<long long>(i) * Wo + j) * 8;
                    ci.cc = cc;
                    ci.ij = i | (j << 16);
                    cb->chk[ch] = ci;
                  }
                  named_bar_sync(1, 32 * N_EPI_WARPS);                     // constants visible to all epilogue warps
                }

              if (NARROW) {                 // compile-time: MT == 2 && n_acc == 16 (launch_conv)
                // ---- narrow layers: one 16-column group; warp (mt, cs) takes its chunk cs (8 channels) ----
                size_t off;
                int cc, ho, wo;
                if (flags & F_SHUF) {
                    const ChunkInfo ci = cb->chk[(flags & F_WIDE) ? 0 : cs];
                    off = base_b + pix + static_cast<size_t>(ci.goff) + ((flags & F_WIDE) ? 8 * cs : 0);
                    cc = ci.cc;
                    ho = h * s + (ci.ij & 0xffff);
                    wo = w * s + (ci.ij >> 16) + ((flags & F_WIDE) ? cs : 0);
                } else {
                    cc = n0 + cs * 8;
                    off = base_b + pix + static_cast<size_t>(cc >> 3) * plane;
                    ho = h;
                    wo = w;
                }
                uint4 rr = make_uint4(0, 0, 0, 0);
                if ((flags & F_RESID) && valid) rr = __ldg(reinterpret_cast<const uint4*>(a.resid + off + res_adj));
                mbar_wait(p.tfull + abuf * 8, aphase);
                tc_fence_after();
                uint32_t v8[8];
                tmem_ld8(p.tmem_base + (static_cast<uint32_t>(q * 32) << 16) + abuf * BUF_COLS + mt * G::ACC_STRIDE + cs * 8, v8);
                tmem_ld_wait();
                epilogue_chunk<ACT, FLAGS>(a, flags, v8, cb, cs * 8, t.b, off, cc, ho, wo, valid, rr, Ho, Wo, res_adj);
              } else {
                bool act16[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int g16 = cs + G::CS * j;
                    act16[j] = (g16 * 16 < a.n_acc) && (n0 + g16 * 16 < a.n_total);
                }

                // output addressing of one of this thread's groups: arithmetic for plain convs, the staged table for
                // PixelShuffle (recomputed where needed instead of kept live across the accumulator wait)
                auto group_addr = [&](int g16) {
                    GroupAddr ga;
                    if (flags & F_SHUF) {
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const ChunkInfo ci = cb->chk[2 * g16 + ((flags & F_WIDE) ? 0 : hh)];
                            ga.off[hh] = base_b + pix + static_cast<size_t>(ci.goff) + ((flags & F_WIDE) ? 8 * hh : 0);
                            ga.cc[hh] = ci.cc;
                            ga.ho[hh] = h * s + (ci.ij & 0xffff);
                            ga.wo[hh] = w * s + (ci.ij >> 16) + ((flags & F_WIDE) ? hh : 0);
                        }
                    } else {
                        const int cc = n0 + g16 * 16;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            ga.off[hh] = base_b + pix + static_cast<size_t>((cc >> 3) + hh) * plane;
                            ga.cc[hh] = cc + 8 * hh;
                            ga.ho[hh] = h;
                            ga.wo[hh] = w;
                        }
                    }
                    return ga;
                };

                // Residual prefetch: issued before waiting for the accumulator so the HBM latency hides
                // behind the MMAs of this tile.
                uint4 rres[4][2];
                if (flags & F_RESID) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const GroupAddr ga = group_addr(cs + G::CS * j);
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            rres[j][hh] = make_uint4(0, 0, 0, 0);
                            if (act16[j] && valid) rres[j][hh] = __ldg(reinterpret_cast<const uint4*>(a.resid + ga.off[hh] + res_adj));
                        }
                    }
                }

                mbar_wait(p.tfull + abuf * 8, aphase);
                tc_fence_after();
                const uint32_t taddr = p.tmem_base + (static_cast<uint32_t>(q * 32) << 16) + abuf * BUF_COLS + mt * G::ACC_STRIDE;

                uint32_t v[2][16];
                if (act16[0]) tmem_ld16(taddr + cs * 16, v[0]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (act16[j]) {                                           // CTA-uniform
                        const int col = (cs + G::CS * j) * 16;
                        float4 bs[4];
#pragma unroll
                        for (int p4 = 0; p4 < 4; ++p4) bs[p4] = *reinterpret_cast<const float4*>(cb->bias + col + 4 * p4);
                        tmem_ld_wait();
                        if (j + 1 < 4 && act16[(j + 1) & 3]) tmem_ld16(taddr + (cs + G::CS * (j + 1)) * 16, v[(j + 1) & 1]);
                        const GroupAddr ga = group_addr(cs + G::CS * j);
                        epilogue_group<ACT, FLAGS>(a, flags, v[j & 1], bs, cb, col, t.b, ga, valid, rres[j], Ho, Wo, res_adj);
                    }
                }
              }
                // all TMEM reads of this warp for this buffer are complete -> hand it back to the leader's MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader + abuf * 8);
                abuf ^= 1;
                if (abuf == 0) aphase ^= 1;
            }
        }
      }
    }

    // both CTAs done (the leader's MMAs write the peer's TMEM; the peer's loads credit the leader's barriers)
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(p.tmem_base, TMEM_COLS);
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

PFN_encodeTiled get_encode_tiled() { return get_encode(); }     // shared with conv_wgrad.cu

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

static int a_stage_bytes(int mt) { return mt == 2 ? Geo<2>::A_STAGE_B : Geo<1>::A_STAGE_B; }
static int fixed_smem_bytes() { return BAR_BYTES + 2 * static_cast<int>(sizeof(Cst)); }

// Rows per n-tile.  The pair keeps taps*Kp*n_acc*2 bytes of weights resident (half per CTA) next to at least
// MIN_STAGES activation stages.  Among the sizes that fit, minimise (padded N) x (shared-memory operand-bandwidth
// penalty of a narrow N: one UMMA reads 32*(128 + n/2) B per SM in 128*n/256 cycles); ties go to the larger tile.
int choose_n_acc(int n_total, int ksteps, int taps) {
    int best = 0;
    double best_cost = 0.0;
    for (int n = 16; n <= 256; n += 16) {
        const int mt = n <= 128 ? 2 : 1;
        const long long b_half = 1LL * ksteps * taps * 2 * (n / 2) * 16;
        if (b_half + 1LL * MIN_STAGES * a_stage_bytes(mt) + fixed_smem_bytes() > SMEM_LIMIT) break;
        const int tiles = (n_total + n - 1) / n;
        const double bw = 8192.0 * (1.0 / n + 1.0 / 256.0) / 110.0;
        const double cost = static_cast<double>(tiles) * n * (bw > 1.0 ? bw : 1.0);
        if (best == 0 || cost <= best_cost) { best = n; best_cost = cost; }
        if (n >= n_total) break;
    }
    return best;
}

// Streaming mode (weights too wide to stay resident next to a useful n-tile, e.g. the dgrad of a PixelShuffle
// up-conv whose K is s*s*Cout): every stage carries A plus one K step of the weight half; at least 4 stages.
int choose_n_acc_stream(int n_total, int taps) {
    int best = 0;
    double best_cost = 0.0;
    for (int n = 16; n <= 256; n += 16) {
        const int mt = n <= 128 ? 2 : 1;
        const long long stage = a_stage_bytes(mt) + 1LL * taps * 2 * (n / 2) * 16;
        if (4 * stage + fixed_smem_bytes() > SMEM_LIMIT) break;
        const int tiles = (n_total + n - 1) / n;
        const double bw = 8192.0 * (1.0 / n + 1.0 / 256.0) / 110.0;
        const double cost = static_cast<double>(tiles) * n * (bw > 1.0 ? bw : 1.0);
        if (best == 0 || cost <= best_cost) { best = n; best_cost = cost; }
        if (n >= n_total) break;
    }
    return best;
}

template <int MT, bool NARROW>
static int launch_conv_n(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvTcArgs& a, int act, size_t smem_bytes,
                         cudaStream_t stream) {
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const ConvTcArgs);
    // specialised epilogues for the launch shapes of the decoder cascade, generic otherwise
    KernelFn fn = conv_tc_kernel<MT, -1, -1, NARROW>;
    int slot = 0;
    const int fl = a.flags;
#define BNERV_PICK(ID, ACT_, FL_)                                              \
    if (act == (ACT_) && fl == (FL_)) { fn = conv_tc_kernel<MT, (ACT_), (FL_), NARROW>; slot = (ID); }
    BNERV_PICK(1, BNERV_ACT_SIN, F_AFF | F_PRE)                 // up-conv 1x1 / s=1 (+sin, x0 and u)
    BNERV_PICK(2, BNERV_ACT_SIN, F_AFF | F_PRE | F_SHUF)        // up-conv + PixelShuffle (s = 3, 5)
    BNERV_PICK(7, BNERV_ACT_SIN, F_AFF | F_PRE | F_SHUF | F_WIDE)   // up-conv + PixelShuffle(2), 32-byte stores
    BNERV_PICK(3, BNERV_ACT_GELU, F_AFF)                        // conv0 + GELU + TAT affine
    BNERV_PICK(4, BNERV_ACT_NONE, F_RESID | F_PRE)              // conv1 + residual
    BNERV_PICK(5, BNERV_ACT_NONE, F_PRE | F_SHUF)               // E-NeRV stage-0 up-conv
    BNERV_PICK(6, BNERV_ACT_TANH01, F_NCHW)                     // head conv -> image
    // training forward: the same launches also write act'(pre-activation)
    BNERV_PICK(8,  BNERV_ACT_SIN, F_AFF | F_PRE | F_DERIV)
    BNERV_PICK(9,  BNERV_ACT_SIN, F_AFF | F_PRE | F_SHUF | F_DERIV)
    BNERV_PICK(10, BNERV_ACT_SIN, F_AFF | F_PRE | F_SHUF | F_WIDE | F_DERIV)
    BNERV_PICK(11, BNERV_ACT_GELU, F_AFF | F_PRE | F_DERIV)
    BNERV_PICK(12, BNERV_ACT_NONE, F_PRE)                       // dgrad
    BNERV_PICK(13, BNERV_ACT_TANH01, F_NCHW | F_HEAD)           // 3x3 head conv, 1x1-contraction + shift-sum form
#undef BNERV_PICK
    static bool smem_set_dev[16][32] = {};          // function attributes are per device (context)
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& smem_set_flag = smem_set_dev[slot][cur_dev & 31];
    if (!smem_set_flag) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
        smem_set_flag = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * a.n_pairs);
    cfg.blockDim = dim3(N_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue + weight fetch overlap the previous kernel's tail
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("BNERV_NO_PDL") != nullptr;      // debugging switch
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 1 : 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, tmA, tmB, a);
    if (e != cudaSuccess) {
        count_launch();
        return set_error(static_cast<int>(e), "conv_tc_kernel launch: %s", cudaGetErrorString(e));
    }
    return check_launch("conv_tc_kernel");
}

// NARROW (one 16-column accumulator group, MT = 2): the two 8-channel chunks go to different epilogue warps
template <int MT>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvTcArgs& a, int act, size_t smem_bytes,
                       cudaStream_t stream) {
    if (MT == 2 && a.n_acc == 16 && !(a.flags & F_HEAD)) return launch_conv_n<2, true>(tmA, tmB, a, act, smem_bytes, stream);
    return launch_conv_n<MT, false>(tmA, tmB, a, act, smem_bytes, stream);
}

}  // namespace bnerv

using namespace bnerv;

extern "C" int bnerv_conv_fused(const void* x, int B, int Cin, int H, int W, const void* w_packed,
                                const float* bias_packed, int Cout, int k, int s, int act, const void* resid,
                                const float* g1p, const float* beta, void* out_pre, void* out_aff, float* out_nchw,
                                void* stream) {
    return bnerv_conv_fused_ex(x, B, Cin, H, W, w_packed, bias_packed, Cout, k, s, act, resid, g1p, beta, out_pre, out_aff,
                               out_nchw, nullptr, stream);
}

static int conv_fused_impl(const void* x, int B, int Cin, int H, int W, const void* w_packed,
                           const float* bias_packed, int Cout, int k, int s, int act, const void* resid,
                           const float* g1p, const float* beta, void* out_pre, void* out_aff, float* out_nchw,
                           void* out_deriv, int split, void* stream);

extern "C" int bnerv_conv_fused_ex(const void* x, int B, int Cin, int H, int W, const void* w_packed,
                                   const float* bias_packed, int Cout, int k, int s, int act, const void* resid,
                                   const float* g1p, const float* beta, void* out_pre, void* out_aff, float* out_nchw,
                                   void* out_deriv, void* stream) {
    return conv_fused_impl(x, B, Cin, H, W, w_packed, bias_packed, Cout, k, s, act, resid, g1p, beta, out_pre, out_aff, out_nchw,
                           out_deriv, 0, stream);
}

extern "C" int bnerv_conv_fused_split(const void* x, int B, int Cin, int H, int W, const void* w_packed,
                                      const float* bias_packed, int Cout, int k, int s, int act, const void* resid,
                                      const float* g1p, const float* beta, void* out_pre, void* out_aff, float* out_nchw,
                                      int split, void* stream) {
    if (split & ~3) return set_error(BNERV_E_BADARG, "conv_fused_split: split = %d (bit 0: split outputs, bit 1: split residual)", split);
    if ((split & 2) && !resid) return set_error(BNERV_E_BADARG, "conv_fused_split: split residual without a residual");
    if ((split & 1) && !out_pre && !out_aff) return set_error(BNERV_E_BADARG, "conv_fused_split: split output without a C8 output");
    return conv_fused_impl(x, B, Cin, H, W, w_packed, bias_packed, Cout, k, s, act, resid, g1p, beta, out_pre, out_aff, out_nchw,
                           nullptr, split, stream);
}

static int conv_fused_impl(const void* x, int B, int Cin, int H, int W, const void* w_packed,
                           const float* bias_packed, int Cout, int k, int s, int act, const void* resid,
                           const float* g1p, const float* beta, void* out_pre, void* out_aff, float* out_nchw,
                           void* out_deriv, int split, void* stream) {
    if (!x || !w_packed || !bias_packed) return set_error(BNERV_E_BADARG, "conv_fused: null operand");
    if (B <= 0 || Cin <= 0 || Cout <= 0 || H <= 0 || W <= 0 || s <= 0) return set_error(BNERV_E_BADARG, "conv_fused: non-positive size");
    if (k != 1 && k != 3) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: kernel size %d (only 1 and 3)", k);
    if ((g1p == nullptr) != (beta == nullptr)) return set_error(BNERV_E_BADARG, "conv_fused: g1p and beta must both be given");
    if ((g1p != nullptr) != (out_aff != nullptr)) return set_error(BNERV_E_BADARG, "conv_fused: out_aff requires g1p/beta and vice versa");
    if (!out_pre && !out_aff && !out_nchw) return set_error(BNERV_E_BADARG, "conv_fused: no output");
    if (out_deriv && resid) return set_error(BNERV_E_BADARG, "conv_fused: out_deriv is the derivative of the activation; not defined with a residual");
    if (act < BNERV_ACT_NONE || act > BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: act %d", act);
    if (s > 0xffff) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: PixelShuffle factor %d", s);
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_packed) |
                               reinterpret_cast<uintptr_t>(bias_packed) | reinterpret_cast<uintptr_t>(resid) |
                               reinterpret_cast<uintptr_t>(g1p) | reinterpret_cast<uintptr_t>(beta) |
                               reinterpret_cast<uintptr_t>(out_pre) | reinterpret_cast<uintptr_t>(out_aff) |
                               reinterpret_cast<uintptr_t>(out_deriv);
    if (align_or & 15) return set_error(BNERV_E_BADARG, "conv_fused: pointers must be 16-byte aligned");

    const int cin_p = round_up(Cin, 16), cout_p = round_up(Cout, 16);
    ConvTcArgs a{};
    a.B = B; a.H = H; a.W = W;
    a.cin_groups = cin_p / 8;
    a.ksteps     = cin_p / 16;
    a.taps       = k * k;
    a.n_total    = s * s * cout_p;
    a.n_acc      = choose_n_acc(a.n_total, a.ksteps, a.taps);
    static const bool force_stream = getenv("BNERV_FORCE_STREAM") != nullptr;      // testing switch
    // resident tiles narrower than 64 rows pay more in padded columns / narrow-N shared-memory bandwidth than the
    // weights' L2 re-reads cost; between 64 and 96 rows the two modes measured within noise of each other in-model
    a.b_stream   = (force_stream || a.n_acc == 0 || (a.n_acc < 64 && a.n_acc < a.n_total)) ? 1 : 0;
    if (a.b_stream) a.n_acc = choose_n_acc_stream(a.n_total, a.taps);
    if (a.n_acc == 0) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: no tile configuration fits (Cin = %d, k = %d)", Cin, k);
    a.n_half     = a.n_acc / 2;
    a.n_tiles    = (a.n_total + a.n_acc - 1) / a.n_acc;
    a.cout = Cout; a.cout_p = cout_p; a.s = s; a.act = act;
    const int mt = a.n_acc <= 128 ? 2 : 1;
    const int tile_w = 8 * mt;
    a.tiles_x = (W + tile_w - 1) / tile_w;
    a.tiles_y = (H + TILE_H - 1) / TILE_H;
    a.pairs_x = (a.tiles_x + 1) / 2;
    const long long pair_tiles = 1LL * B * a.tiles_y * a.pairs_x;
    if (pair_tiles > 0x3fffffffLL) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: too many tiles");
    a.pair_tiles    = static_cast<int>(pair_tiles);
    if (pair_tiles * a.n_tiles > 0x3fffffffLL) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: too many tiles");
    a.work_total    = a.pair_tiles * a.n_tiles;
    a.b_kstep_bytes = a.taps * 2 * a.n_half * 16;
    a.b_bytes       = a.b_stream ? 0 : a.ksteps * a.b_kstep_bytes;
    const int a_stage = a_stage_bytes(mt) + (a.b_stream ? a.b_kstep_bytes : 0);
    a.stage_bytes   = a_stage;
    int stages = (SMEM_LIMIT - fixed_smem_bytes() - a.b_bytes) / a_stage;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < MIN_STAGES) return set_error(BNERV_E_UNSUPPORTED, "conv_fused: %d activation stages do not fit", MIN_STAGES);
    a.stages = stages;
    a.bias = bias_packed; a.g1p = g1p; a.beta = beta;
    a.resid = static_cast<const __half*>(resid);
    a.out_pre = static_cast<__half*>(out_pre);
    a.out_aff = static_cast<__half*>(out_aff);
    a.out_nchw = out_nchw;
    a.out_deriv = static_cast<__half*>(out_deriv);
    a.flags = (out_deriv ? F_DERIV : 0) | (resid ? F_RESID : 0) | (g1p ? F_AFF : 0) | (out_pre ? F_PRE : 0) | (out_nchw ? F_NCHW : 0) | (s > 1 ? F_SHUF : 0) | (s == 2 ? F_WIDE : 0) |
              ((split & 1) ? F_SPLIT_OUT : 0) | ((split & 2) ? F_SPLIT_RES : 0);
    a.split_stride = static_cast<long long>(cout_p / 8) * (1LL * H * s) * (1LL * W * s) * 8;

    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    const int max_pairs = g_num_sms / 2;
    a.n_pairs = a.work_total < max_pairs ? a.work_total : max_pairs;

    CUtensorMap tmA, tmB;
    // activations as u64: dims {2W, H, B*cin_groups}; strides {W*16, H*W*16} bytes
    int rc = make_map_u64_3d(&tmA, x, 2ull * W, H, 1ull * B * a.cin_groups, 16ull * W, 16ull * W * H,
                             2 * (tile_w + 2), HALO_H, 2);
    if (rc) return rc;
    // packed weights as u64: dims {2*Np, Kp/8, taps}; strides {Np*16, Np*16*Kp/8}; one box = one K step of one CTA's half
    rc = make_map_u64_3d(&tmB, w_packed, 2ull * a.n_total, a.cin_groups, a.taps, 16ull * a.n_total,
                         16ull * a.n_total * a.cin_groups, 2 * a.n_half, 2, a.taps);
    if (rc) return rc;

    const size_t smem_bytes = static_cast<size_t>(a.b_bytes) + static_cast<size_t>(stages) * a_stage + fixed_smem_bytes();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return mt == 2 ? launch_conv<2>(tmA, tmB, a, act, smem_bytes, st) : launch_conv<1>(tmA, tmB, a, act, smem_bytes, st);
}


// ---------------------------------------------------------------------------------------------
// head mode entry point
// ---------------------------------------------------------------------------------------------
extern "C" int bnerv_head_conv3(const void* x, int B, int Cin, int H, int W, const void* w_head_packed, const float* bias,
                                int Cout, int act, float* out_nchw, void* stream) {
    if (!x || !w_head_packed || !bias || !out_nchw) return set_error(BNERV_E_BADARG, "head_conv3: null pointer");
    if (B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || Cout <= 0) return set_error(BNERV_E_BADARG, "head_conv3: non-positive size");
    if (9 * Cout > HEAD_PSTR) return set_error(BNERV_E_UNSUPPORTED, "head_conv3: Cout = %d (at most 3)", Cout);
    if (act < BNERV_ACT_NONE || act > BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "head_conv3: act %d", act);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_head_packed)) & 15)
        return set_error(BNERV_E_BADARG, "head_conv3: pointers must be 16-byte aligned");
    const int cin_p = round_up(Cin, 16);
    ConvTcArgs a{};
    a.B = B; a.H = H; a.W = W;
    a.cin_groups = cin_p / 8;
    a.ksteps = cin_p / 16;
    a.taps = 1;                                   // the packed weight is one [Kp][32] slab
    a.n_total = HEAD_N; a.n_acc = HEAD_N; a.n_half = HEAD_N / 2; a.n_tiles = 1;
    a.cout = Cout; a.cout_p = 16; a.s = 1; a.act = act;
    a.flags = F_NCHW | F_HEAD;
    const int tile_w = 16;
    a.tiles_x = (W + tile_w - 1) / tile_w;
    a.tiles_y = (H + TILE_H - 1) / TILE_H;
    a.pairs_x = (a.tiles_x + 1) / 2;
    const long long pair_tiles = 1LL * B * a.tiles_y * a.pairs_x;
    if (pair_tiles > 0x3fffffffLL) return set_error(BNERV_E_UNSUPPORTED, "head_conv3: too many tiles");
    a.pair_tiles = static_cast<int>(pair_tiles);
    a.work_total = a.pair_tiles;
    a.b_stream = 0;
    a.b_kstep_bytes = 2 * a.n_half * 16;
    a.b_bytes = a.ksteps * a.b_kstep_bytes;
    a.stage_bytes = a_stage_bytes(2);
    int stages = (SMEM_LIMIT - fixed_smem_bytes() - HEAD_SMEM_BYTES - a.b_bytes) / a.stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < MIN_STAGES) return set_error(BNERV_E_UNSUPPORTED, "head_conv3: Cin = %d too wide", Cin);
    a.stages = stages;
    a.bias = bias;
    a.out_nchw = out_nchw;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    const int max_pairs = g_num_sms / 2;
    a.n_pairs = a.work_total < max_pairs ? a.work_total : max_pairs;
    CUtensorMap tmA, tmB;
    int rc = make_map_u64_3d(&tmA, x, 2ull * W, H, 1ull * B * a.cin_groups, 16ull * W, 16ull * W * H, 2 * (tile_w + 2), HALO_H, 2);
    if (rc) return rc;
    rc = make_map_u64_3d(&tmB, w_head_packed, 2ull * HEAD_N, a.cin_groups, 1, 16ull * HEAD_N, 16ull * HEAD_N * a.cin_groups,
                         2 * a.n_half, 2, 1);
    if (rc) return rc;
    const size_t smem_bytes = static_cast<size_t>(a.b_bytes) + static_cast<size_t>(stages) * a.stage_bytes + fixed_smem_bytes() +
                              HEAD_SMEM_BYTES;
    return launch_conv<2>(tmA, tmB, a, act, smem_bytes, static_cast<cudaStream_t>(stream));
}


// ---------------------------------------------------------------------------------------------
// One NeRVBlock = three fused-conv launches chained by programmatic dependent launch
// ---------------------------------------------------------------------------------------------
extern "C" int bnerv_nerv_block_fwd(const void* x, int B, int Cin, int H, int W, const void* w_up, const float* b_up, int k_up,
                                    int s, int act_up, const void* w_c0, const float* b_c0, const void* w_c1, const float* b_c1,
                                    int C, int act_inner, const float* g0p, const float* beta0, const float* g1p,
                                    const float* beta1, void* x0, void* u, void* wmap, void* out, void* stream) {
    if (!x0 || !u || !wmap || !out) return set_error(BNERV_E_BADARG, "nerv_block_fwd: null workspace / output");
    if (!g0p || !beta0 || !g1p || !beta1) return set_error(BNERV_E_BADARG, "nerv_block_fwd: the four TAT tables are required");
    // x0 = act(PS_s(conv_k(x))) and u = x0*g0p + beta0          model_blocks.py:37,216-217 + :105
    int rc = bnerv_conv_fused(x, B, Cin, H, W, w_up, b_up, C, k_up, s, act_up, nullptr, g0p, beta0, x0, u, nullptr, stream);
    if (rc) return rc;
    const int Ho = H * s, Wo = W * s;
    // w = act_inner(conv3(u))*g1p + beta1                          model_blocks.py:86-87
    rc = bnerv_conv_fused(u, B, C, Ho, Wo, w_c0, b_c0, C, 3, 1, act_inner, nullptr, g1p, beta1, nullptr, wmap, nullptr, stream);
    if (rc) return rc;
    // out = x0 + conv3(w)                                          model_blocks.py:88-89
    return bnerv_conv_fused(wmap, B, C, Ho, Wo, w_c1, b_c1, C, 3, 1, BNERV_ACT_NONE, x0, nullptr, nullptr, out, nullptr, nullptr, stream);
}
