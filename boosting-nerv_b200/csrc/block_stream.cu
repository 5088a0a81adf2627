// One kernel per NeRVBlock for the 12..16-channel stages, ROW-STREAMING form with the MMA A operand in TENSOR MEMORY.
// Replaces NeRVBlock.forward (model_blocks.py:34-46): UpConv 3x3 (:213-220, no PixelShuffle in this form) + Sin (:129-134)
// + ResBlock_SFT (:83-89) with its two SFTLayer affines (:101-105).
//
// Why: at N = 16 a tcgen05.mma whose A operand comes from shared memory costs ~50 cycles (4 KB of A per MMA at 128 B/clk),
// the same MMA with A in TMEM costs 17.6 (tools/ts_probe.cu, profiles/r02_ts_probe_umma_ts_vs_ss.txt): the region-tiled
// form (block_fused.cu) keeps the tensor pipe busy with operand reads for ~65 % of a 720p block.  Here the tap shift of a
// 3x3 conv is split: the HORIZONTAL shift goes into the operand - for every image row the threads that own the pixels write
// a "row-im2col" A[lane = pixel w][(sx, c)] = in[h][w + sx - 1][c] into TMEM (3 x 8 columns for 16 channels) - and the
// VERTICAL shift is a choice of A row:   D(h) = sum_r A(h + r - 1) . W[r]   (3 x 3 MMAs of K = 16, N = 16).
// Each A row is written once and read by the MMAs of three output rows.
//
// A CTA owns a strip of 128 columns (122 valid: the three chained 3x3 convs lose 2 + 4 lanes) and a segment of rows, and
// streams rows top to bottom through a three-stage pipeline; different rows are in different stages at the same time:
//     TMA warp   : input rows -> shared-memory ring
//     WG front   : builds A_up(h) from the input ring;  epilogue of D_up(h): x0 = sin(.), u = x0*g0p + beta0 -> u row to
//                  shared memory (neighbour exchange), x0 row to the residual ring, builds A_c0(h)
//     WG middle  : epilogue of D_c0(h): w = gelu(.)*g1p + beta1 -> builds A_c1(h)
//     WG back    : epilogue of D_c1(h): out = . + x0 -> global
//     MMA warp   : one elected lane issues the 9 MMAs of every (stage, row) as its three A rows become ready
// Rings in TMEM: 4 A rows x 24 columns and 4 accumulators x 16 columns per stage = 480 of 512 columns (powers of two:
// slot = row & 3, no integer division in the loops); every hand-off is
// an mbarrier (A full / A free, D full / D free, input full / free, x0 full / free).
// Out-of-image pixels of u / w are written as exact zeros (the reference pads AFTER the affine, model_blocks.py:105,:86).
//
// Arithmetic = the three-launch path's, operation by operation (tap order r*3+sx inside one K step, same epilogue
// functions, f16 rounding of x0 / u / w at the same places): results are expected to be bit-identical to
// bnerv_nerv_block_fwd; tests/test_gpu_block_fused.py checks exactly that.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

namespace bnerv {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();          // conv_tc.cu

// Six epilogue warpgroups, split by the per-row cost of the stages: 3 front (A_up build + sin epilogue + A_c0 build), 2 middle
// (GELU epilogue + A_c1 build), 1 back (residual + store).  Warpgroup p of a stage with n warpgroups takes rows j = p (mod n).
constexpr int BS_WG_F = 3, BS_WG_M = 2, BS_WG_B = 1;
constexpr int BS_WGS = 3;                     // exchange-row buffers are sized for the largest count
constexpr int BS_MMA_WARP = (BS_WG_F + BS_WG_M + BS_WG_B) * 4;   // warps 0-23 epilogues, 24-26 MMA issue (one per stage), 27 TMA
constexpr int BS_THREADS = (BS_MMA_WARP + 4) * 32;
constexpr int BS_NA = 4;                      // A-row ring slots per stage (TMEM)
constexpr int BS_NIK = 2;                     // input rows in flight per front warpgroup
constexpr int BS_NI = BS_WG_F * BS_NIK;       // input-row slots (shared memory): row r -> slot (r % 3) * 2 + (r / 3) % 2, use (r / 3) / 2.
// Every front warpgroup owns its slots, so it waits on EVERY use of them in order.  (A shared ring taken in turns is not safe
// here: TMA loads complete out of order, a warpgroup could test a slot whose previous load - consumed by another warpgroup -
// has not landed yet, and the parity test of an mbarrier cannot tell "one phase behind" from "done".)
constexpr int BS_NX = 24;                     // x0-row ring slots (shared memory); the writer leads the reader by <= 18 rows
constexpr int BS_VALID = 122;                 // valid output columns per strip: lanes [2, 124)
constexpr int BS_ACOLS = 24;                  // TMEM columns of one A row: 3 horizontal taps x 16 channels x f16
constexpr int BS_A_COL0 = 0;                  // A rings: stage S at S * NA * 24
constexpr int BS_D_COL0 = 3 * BS_NA * BS_ACOLS;            // accumulator rings follow the A rings
// Accumulator rings.  s = 1: 4 slots x 16 columns per stage (288 .. 480).  PixelShuffle(2) form (S2): the up-conv runs on
// INPUT rows with N = 4 sub-positions x 16 channels = 64 columns, 2 slots; conv0 4 x 16; conv1 2 x 16 (288 .. 512).
template <bool S2> struct BsRing {
    static constexpr int ND0 = S2 ? 2 : 4, ND1 = 4, ND2 = S2 ? 2 : 4;
    static constexpr int DW0 = S2 ? 64 : 16;
    static constexpr int D0 = BS_D_COL0, D1 = D0 + ND0 * DW0, D2 = D1 + ND1 * 16;
    static_assert(D2 + ND2 * 16 <= 512, "TMEM columns");
    __host__ __device__ static constexpr int nd(int S) { return S == 0 ? ND0 : (S == 1 ? ND1 : ND2); }
    __host__ __device__ static constexpr int dcol(int S) { return S == 0 ? D0 : (S == 1 ? D1 : D2); }
    __host__ __device__ static constexpr int dw(int S) { return S == 0 ? DW0 : 16; }
};
constexpr int BS_ROW_B = 2 * 128 * 16;        // one input row in shared memory: [2 groups][128 px][16 B]
constexpr int BS_XROW_B = 2 * 130 * 16;       // one exchange row: [2 groups][130 px][16 B] (px 0 and 129 stay zero)

struct BsCst { float b_up[64], b_c0[16], b_c1[16], g0p[16], beta0[16], g1p[16], beta1[16]; float head_w[4][16], head_b[4]; };

struct BsBars {
    uint64_t in_full[BS_NI], in_empty[BS_NI];
    uint64_t a_full[3][BS_NA], a_empty[3][BS_NA];
    uint64_t d_full[3][4], d_empty[3][4];
    uint64_t tok[BS_WG_F];       // S2 form: "front warpgroup p has built its pair of A_c0 rows" (orders the builds, see the kernel)
    uint32_t tmem_slot, pad;
};

struct BsSmem {
    uint8_t w_up[9 * 2 * 64 * 16];            // up-conv weights: [tap][2 groups][N = 16 or 64 rows][16 B]
    uint8_t w_c[2][9 * 2 * 16 * 16];          // conv0 / conv1 weights: [tap][2 groups][16 rows][16 B]
    uint8_t in_ring[BS_NI][BS_ROW_B];
    uint8_t u_ring[BS_WG_F][2][2][BS_XROW_B]; // per front warpgroup: double-buffered exchange rows (two rows per step in the S2 form)
    uint8_t w_ring[BS_WG_M][2][BS_XROW_B];
    uint8_t x0_ring[BS_NX][BS_ROW_B];
    BsCst cst;
    BsBars bars;
};

struct BsArgs {
    int B, H, W, C;              // H, W: the block's OUTPUT resolution (= input resolution x s)
    int s;                       // PixelShuffle factor of the up-conv: 1 or 2
    int has_up;                  // 0: the TMA input is u (conv0's input), the residual x0 is read from `resid`
    int act_up, act_inner;
    int strips, segs, seg_rows;
    const __half *w_up, *w_c0, *w_c1;
    const float *b_up, *b_c0, *b_c1, *g0p, *beta0, *g1p, *beta1;
    const __half* resid;
    __half* out;
    long long* dbg;              // bring-up: CTA 0 records clock64 stamps [role < 9][iteration < 32][4] (bnerv_debug_set_buffer)
    // fused 1x1 head conv + OutImg (HEAD instantiations): img = act(head_w . f16(out) + head_b), NCHW f32; `out` is not stored
    const float *head_w, *head_b;
    float* img;
    int head_cout, head_act;
};

#define BS_STAMP(role, it, k) do { if (DBG && a.dbg && blockIdx.x == 0 && (threadIdx.x & 127) == 0 && (it) < 32) \
        a.dbg[((role) * 32 + (it)) * 4 + (k)] = clock64(); } while (0)
#define BS_STAMP1(role, it, k) do { if (DBG && a.dbg && blockIdx.x == 0 && lane == 0 && (it) < 32) \
        a.dbg[((role) * 32 + (it)) * 4 + (k)] = clock64(); } while (0)

extern long long* g_bf_dbg;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& lo, const uint4& hi) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]; B descriptor given as (lo, hi) halves
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ uint4 bs_pack8(const float2* x) {
    uint4 o;
    o.x = pack_h2_satfinite(x[0]); o.y = pack_h2_satfinite(x[1]);
    o.z = pack_h2_satfinite(x[2]); o.w = pack_h2_satfinite(x[3]);
    return o;
}
template <int ACT>
__device__ __forceinline__ float2 bs_act2(float2 x, int act) {
    if (ACT >= 0) return act2<ACT>(x);
    switch (act) {
        case BNERV_ACT_SIN:    return sin2(x);
        case BNERV_ACT_GELU:   return gelu2(x);
        case BNERV_ACT_RELU:   return act2<BNERV_ACT_RELU>(x);
        case BNERV_ACT_TANH01: return tanh01_2(x);
        default:               return x;
    }
}
__device__ __forceinline__ void bs_bias16(const uint32_t* v, const float* bias, float2* x) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float4 bs = *reinterpret_cast<const float4*>(bias + 4 * p);
        x[2 * p]     = add2(make_float2(__uint_as_float(v[4 * p]),     __uint_as_float(v[4 * p + 1])), make_float2(bs.x, bs.y));
        x[2 * p + 1] = add2(make_float2(__uint_as_float(v[4 * p + 2]), __uint_as_float(v[4 * p + 3])), make_float2(bs.z, bs.w));
    }
}
__device__ __forceinline__ void bs_affine8(const float2* x, const float* g, const float* e, float2* y) {
    const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
    const float4 e0 = *reinterpret_cast<const float4*>(e), e1 = *reinterpret_cast<const float4*>(e + 4);
    y[0] = fma2(x[0], make_float2(g0.x, g0.y), make_float2(e0.x, e0.y));
    y[1] = fma2(x[1], make_float2(g0.z, g0.w), make_float2(e0.z, e0.w));
    y[2] = fma2(x[2], make_float2(g1.x, g1.y), make_float2(e1.x, e1.y));
    y[3] = fma2(x[3], make_float2(g1.z, g1.w), make_float2(e1.z, e1.w));
}

// packed global weights [tap][2 groups][16 rows][8 halves] are already the shared-memory form for one K step
__device__ __forceinline__ void bs_stage_weights(uint8_t* dst, const __half* src, int n_rows) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < 9 * 2 * n_rows; i += BS_THREADS) d4[i] = __ldg(s4 + i);
}

// NP: float2 channel pairs that carry data (6 when C <= 12: the 4 pad channels are exact zeros through sin / GELU / ReLU and
// are not evaluated); DBG: records BS_STAMPs.
// S2: the up-conv carries PixelShuffle(2) - it runs on input rows (half resolution), see the front warpgroups.
template <int ACT_UP, int ACT_IN, int NP, bool DBG, bool S2, bool HEAD = false>
__global__ void __launch_bounds__(BS_THREADS, 1)
block_stream_kernel(const __grid_constant__ CUtensorMap tmIn, const BsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    BsSmem& sm = *reinterpret_cast<BsSmem*>(smem_raw);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int wg = warp >> 2;                         // 0-2 front, 3-4 middle, 5 back, 6 = MMA / TMA warps
    const int stage_of_wg = wg < BS_WG_F ? 0 : (wg < BS_WG_F + BS_WG_M ? 1 : 2);
    const int par = stage_of_wg == 0 ? wg : (stage_of_wg == 1 ? wg - BS_WG_F : wg - BS_WG_F - BS_WG_M);   // row residue this warpgroup handles
    // A warpgroup may only wait on ring slots whose EVERY use it sees or whose previous use is implied complete by its own
    // previous row (mbarrier waits test a phase parity: "one use behind" and "done" look alike): the row stride of a stage's
    // warpgroups must not exceed its ring depth.  S2 leaves only 2 accumulator slots for the 64-column up-conv rows, fewer than
    // the three front warpgroups - so there the "full" signal of the up-conv goes to a barrier per CONSUMER WARPGROUP (which sees
    // every phase of it; row j+3 cannot complete before every warp has released row j, the slot chain j -> j+2 and the in-order
    // MMA issue see to that), while "empty" stays per slot for its single waiter, the MMA issuer.
    constexpr int NFRONT = BS_WG_F;
    const int nwg = stage_of_wg == 0 ? NFRONT : (stage_of_wg == 1 ? BS_WG_M : BS_WG_B);
    const int q = warp & 3;
    const int m = q * 32 + lane;                      // TMEM lane == column of the strip

    // CTA -> (frame, strip, row segment)
    const int seg = blockIdx.x % a.segs;
    const int strip = (blockIdx.x / a.segs) % a.strips;
    const int fb = blockIdx.x / (a.segs * a.strips);
    const int y0 = seg * a.seg_rows;
    const int y1 = (y0 + a.seg_rows < a.H) ? y0 + a.seg_rows : a.H;
    const int rows = y1 - y0;                         // >= 1 by construction of the grid
    const int sx0 = strip * BS_VALID - 2;             // image column of lane 0
    const int col = sx0 + m;
    const bool col_in = (col >= 0) && (col < a.W);

    using R = BsRing<S2>;
    const int S0 = a.has_up ? 0 : 1;                  // first stage that runs
    // rows per stage: up y0-2 .. y1+1, c0 y0-1 .. y1, c1 y0 .. y1-1;  A rows per stage: two more than its output rows.
    // S2: the up stage counts INPUT rows: row hi yields output rows 2hi, 2hi+1; y0 is even, so the u rows y0-2 .. y1+1
    // (A_c0 rows 0 .. rows+3) come from input rows (y0-2)/2 + jin, jin < n_jin.
    const int n_ac0 = rows + 4;
    const int n_jin = (n_ac0 + 1) / 2;
    const int n_out[3] = {S2 ? n_jin : rows + 4, rows + 2, rows};
    const int n_in = a.has_up ? n_out[0] + 2 : rows + 4;            // TMA rows: A_up rows (x) or A_c0 rows (u)
    const int in_row0 = a.has_up ? (S2 ? (y0 - 2) / 2 - 1 : y0 - 3) : y0 - 2;
    const int in_x0 = S2 ? sx0 / 2 - 1 : sx0 - 1;     // image column (input resolution) of pixel 0 of an input row in shared memory

    if (threadIdx.x == 0) {
        for (int i = 0; i < BS_NI; ++i) { mbar_init(smem_u32(&sm.bars.in_full[i]), 1); mbar_init(smem_u32(&sm.bars.in_empty[i]), 4); }
        for (int s = 0; s < 3; ++s) {
            for (int i = 0; i < BS_NA; ++i) { mbar_init(smem_u32(&sm.bars.a_full[s][i]), 4); mbar_init(smem_u32(&sm.bars.a_empty[s][i]), 1); }
            for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&sm.bars.d_full[s][i]), 1); mbar_init(smem_u32(&sm.bars.d_empty[s][i]), 4); }
        }
        for (int i = 0; i < BS_WG_F; ++i) mbar_init(smem_u32(&sm.bars.tok[i]), 4);
        fence_mbar_init();
        tma_prefetch_desc(&tmIn);
    }
    if (warp == BS_MMA_WARP) tmem_alloc(smem_u32(&sm.bars.tmem_slot), 512);
    if (a.has_up) bs_stage_weights(sm.w_up, a.w_up, R::DW0);
    bs_stage_weights(sm.w_c[0], a.w_c0, 16);
    bs_stage_weights(sm.w_c[1], a.w_c1, 16);
    if (threadIdx.x < 64) {
        const int i = threadIdx.x;
        sm.cst.b_up[i] = (a.has_up && i < R::DW0) ? __ldg(a.b_up + i) : 0.0f;
        if (i < 16) {
            sm.cst.b_c0[i] = __ldg(a.b_c0 + i);
            sm.cst.b_c1[i] = __ldg(a.b_c1 + i);
        }
    }
    if (HEAD && threadIdx.x >= 64 && threadIdx.x < 128) {        // head weights [Cout][C] -> [4][16], zero beyond (Cout, C)
        const int c = (threadIdx.x - 64) >> 4, k = threadIdx.x & 15;
        sm.cst.head_w[c][k] = (c < a.head_cout && k < a.C) ? __ldg(a.head_w + c * a.C + k) : 0.0f;
        if (k == 0) sm.cst.head_b[c] = (c < a.head_cout && a.head_b) ? __ldg(a.head_b + c) : 0.0f;
    }
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(sm.u_ring) / 16); i += BS_THREADS)      // exchange rows: the edge pixels stay zero
        reinterpret_cast<uint4*>(sm.u_ring)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(sm.w_ring) / 16); i += BS_THREADS)
        reinterpret_cast<uint4*>(sm.w_ring)[i] = make_uint4(0, 0, 0, 0);
    pdl_wait();                       // TAT tables / activations come from earlier kernels
    pdl_launch_dependents();
    if (threadIdx.x < 16) {
        const int i = threadIdx.x;
        sm.cst.g0p[i] = a.has_up ? __ldg(a.g0p + fb * 16 + i) : 0.0f;
        sm.cst.beta0[i] = a.has_up ? __ldg(a.beta0 + fb * 16 + i) : 0.0f;
        sm.cst.g1p[i] = __ldg(a.g1p + fb * 16 + i);
        sm.cst.beta1[i] = __ldg(a.beta1 + fb * 16 + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.bars.tmem_slot;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    // builds A row `iA` of stage S from a [2 groups][pitch px][16 B] shared-memory row: lane m <- pixels m+px0 .. m+px0+2
    // (lane_mask: 127 = every lane its own pixel; 63 = lanes 64..127 repeat lanes 0..63, the S2 up-conv rows)
    auto build_a = [&](int S, int iA, const uint8_t* row, int pitch, int px0, int px_max, int lane_mask = 127) {
        uint4 g0[3], g1[3];
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) {
            int px = (m & lane_mask) + px0 + sx;
            px = px > px_max ? px_max : px;                  // lanes >= 126 of an input row: not valid lanes, any finite data
            g0[sx] = *reinterpret_cast<const uint4*>(row + static_cast<size_t>(px) * 16);
            g1[sx] = *reinterpret_cast<const uint4*>(row + static_cast<size_t>(pitch + px) * 16);
        }
        const int as = iA % BS_NA;
        mbar_wait(smem_u32(&sm.bars.a_empty[S][as]), ((iA / BS_NA) & 1) ^ 1);
        tc_fence_after();
        const uint32_t t = lane_base + BS_A_COL0 + (S * BS_NA + as) * BS_ACOLS;
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) tmem_st8(t + sx * 8, g0[sx], g1[sx]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm.bars.a_full[S][as]));
    };

    if (warp == BS_MMA_WARP + 3) {
        // =============================== TMA producer: input rows -> ring ===============================
        for (int i = 0; i < n_in; ++i) {
            const int owner = i % NFRONT, seq = i / NFRONT;                // row i belongs to front warpgroup `owner`, its seq-th row
            const int slot = owner * BS_NIK + (seq % BS_NIK);
            mbar_wait(smem_u32(&sm.bars.in_empty[slot]), ((seq / BS_NIK) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(smem_u32(&sm.bars.in_full[slot]), BS_ROW_B);
                tma_load_3d(smem_u32(sm.in_ring[slot]), &tmIn, smem_u32(&sm.bars.in_full[slot]), 2 * in_x0, in_row0 + i, fb * 2);
            }
            __syncwarp();
        }
    } else if (warp >= BS_MMA_WARP) {
        // =============================== MMA issuers: warp BS_MMA_WARP + S issues stage S ===============================
        // (one issuer per stage: a single warp walking all three stages spends ~90 cycles in each of its two mbarrier waits
        //  per (stage, row) even when they are already complete - 1400 cycles per row, which was the whole row period)
        const int S = warp - BS_MMA_WARP;
        if (S >= S0) {
            const int N = (S == 0) ? R::DW0 : 16;                       // accumulator columns of one row of this stage
            const int nd = R::nd(S);
            const uint32_t idesc = umma_idesc_f16(128, N);
            const uint64_t b_d = umma_desc_hi_noswz(static_cast<uint32_t>(N) * 16u, 128u);
            const uint32_t b_hi = static_cast<uint32_t>(b_d >> 32);
            const uint32_t b_lo = static_cast<uint32_t>(b_d) | ((smem_u32(S == 0 ? sm.w_up : sm.w_c[S - 1]) & 0x3FFFFu) >> 4);
            const uint32_t tap16 = static_cast<uint32_t>(2 * N);       // one tap of B in 16-byte units
            const uint32_t d_base = tmem_base + R::dcol(S);
            const int n = n_out[S];
            int a_waited = 0;
            for (int j = 0; j < n; ++j) {
                BS_STAMP1(6 + S, j, 0);
                while (a_waited <= j + 2) {                             // A rows j, j+1, j+2 (image rows h-1, h, h+1)
                    mbar_wait(smem_u32(&sm.bars.a_full[S][a_waited % BS_NA]), (a_waited / BS_NA) & 1);
                    ++a_waited;
                }
                BS_STAMP1(6 + S, j, 1);
                const int ds = j % nd;
                mbar_wait(smem_u32(&sm.bars.d_empty[S][ds]), ((j / nd) & 1) ^ 1);
                tc_fence_after();
                BS_STAMP1(6 + S, j, 2);
                if (elect_one()) {
                    const uint32_t d = d_base + ds * N;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const uint32_t at = tmem_base + BS_A_COL0 + (S * BS_NA + (j + r) % BS_NA) * BS_ACOLS;
#pragma unroll
                        for (int sx = 0; sx < 3; ++sx)
                            umma_f16_ts(d, at + sx * 8, b_lo + (r * 3 + sx) * tap16, b_hi, idesc, (r | sx) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(&sm.bars.d_full[S][(S2 && S == 0) ? j % NFRONT : ds]));
                    umma_commit(smem_u32(&sm.bars.a_empty[S][j % BS_NA]));      // A row j has had its last reader
                }
                __syncwarp();
                BS_STAMP1(6 + S, j, 3);
            }
        }
    } else if (stage_of_wg == 0) {
        // =============================== front warpgroups (rows j = par mod nwg) ===============================
        if (!a.has_up) {
            // residual form: the input rows ARE u -> A_c0 rows
            int seq = 0;
            for (int i = par; i < n_in; i += nwg, ++seq) {
                const int slot = par * BS_NIK + (seq % BS_NIK);
                mbar_wait(smem_u32(&sm.bars.in_full[slot]), (seq / BS_NIK) & 1);
                build_a(1, i, sm.in_ring[slot], 128, 0, 127);
                if (lane == 0) mbar_arrive(smem_u32(&sm.bars.in_empty[slot]));
            }
        } else if (S2) {
            // ---------- PixelShuffle(2) up-conv: one step = one INPUT row hi = (y0-2)/2 + jin -> output rows 2hi, 2hi+1 ----------
            // D_up(jin): lane l = input column sx0/2 + l (l < 64 carry data), 64 columns = packed rows n' = ((i*2 + G)*2 + j)*8 + c%8
            // (bnerv_pack_conv_weight, s = 2): 16-column group i*2+G holds output row 2hi+i, channel group G, both output
            // columns 2*(sx0/2 + l) + j.  The sin epilogue therefore runs on lanes 0..63 (warps 0, 1 of the warpgroup); all
            // four warps build the two A_c0 rows afterwards.
            // The 64 input columns of the strip are REPEATED in lanes 64..127 of the A_up rows, so the accumulator rows hold every
            // value twice and all four warps of the warpgroup share the epilogue: lanes 0..63 take output row 2hi (groups 0, 1),
            // lanes 64..127 output row 2hi + 1 (groups 2, 3).  (The MMA is M = 128 either way.)
            auto build_up = [&](int i) {
                const int seq = i / NFRONT;
                const int slot = par * BS_NIK + (seq % BS_NIK);
                mbar_wait(smem_u32(&sm.bars.in_full[slot]), (seq / BS_NIK) & 1);
                build_a(0, i, sm.in_ring[slot], 128, 0, 127, 63);
                if (lane == 0) mbar_arrive(smem_u32(&sm.bars.in_empty[slot]));
            };
            if (par < n_in) build_up(par);
            const int lc = m & 63, i_row = m >> 6;                       // input column of this lane, output-row parity it handles
            int it = 0;
            for (int j = par; j < n_out[0]; j += nwg, ++it) {
                if (j + nwg < n_in) build_up(j + nwg);
                const int ds = j % R::ND0;
                mbar_wait(smem_u32(&sm.bars.d_full[0][par]), it & 1);     // this warpgroup's own "full" barrier (see NFRONT)
                tc_fence_after();
                uint8_t* ubuf = sm.u_ring[par][it & 1][0];               // two exchange rows: + i * BS_XROW_B
                {
#pragma unroll 1
                    for (int G = 0; G < 2; ++G) {
                        const int i = i_row, g = i_row * 2 + G;          // g = i*2 + G
                        uint32_t v[16];
                        tmem_ld16(lane_base + R::D0 + ds * 64 + g * 16, v);
                        tmem_ld_wait();
                        float2 x[8];
                        bs_bias16(v, sm.cst.b_up + g * 16, x);
                        // x[0..3]: 8 channels of output column j = 0, x[4..7]: of j = 1; channel group G: pairs 2, 3 are pad when NP == 6
#pragma unroll
                        for (int p = 0; p < 8; ++p)
                            x[p] = (G == 0 || NP == 8 || (p & 3) < 2) ? bs_act2<ACT_UP>(x[p], a.act_up) : make_float2(0.0f, 0.0f);
                        const int ia = 2 * j + i;                        // A_c0 row = output row y0 - 2 + ia
                        const int h = y0 - 2 + ia;
                        const int k = ia - 2;
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            const int lane_o = 2 * lc + jj;              // output lane of the strip
                            const int ocol = sx0 + lane_o;
                            const bool inside = (ocol >= 0) && (ocol < a.W) && (h >= 0) && (h < a.H);
                            float2 y[4];
                            bs_affine8(x + 4 * jj, sm.cst.g0p + 8 * G, sm.cst.beta0 + 8 * G, y);
                            uint4 uo = bs_pack8(y);
                            if (!inside) uo = make_uint4(0, 0, 0, 0);
                            *reinterpret_cast<uint4*>(ubuf + i * BS_XROW_B + static_cast<size_t>(G * 130 + lane_o + 1) * 16) = uo;
                            if (k >= 0 && k < rows)
                                *reinterpret_cast<uint4*>(sm.x0_ring[k % BS_NX] + static_cast<size_t>(G * 128 + lane_o) * 16) = bs_pack8(x + 4 * jj);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sm.bars.d_empty[0][ds]));
                named_bar_sync(1 + par, 128);                            // both u rows are complete
                // A_c0 rows are built IN ROW ORDER across the three warpgroups (a token from the warpgroup that owns input row
                // j - 1).  Without it this warpgroup's previous a_empty check (row 2j - 5) does not cover conv0's MMA of row
                // 2j - 8, the parity test of row 2j's slot could alias and overwrite an A row that MMA still reads.
                if (j > 0) {
                    const int pred = (par + NFRONT - 1) % NFRONT;
                    const int pit = (par > 0) ? it : it - 1;
                    mbar_wait(smem_u32(&sm.bars.tok[pred]), pit & 1);
                }
#pragma unroll 1
                for (int i = 0; i < 2; ++i)
                    if (2 * j + i < n_ac0) build_a(1, 2 * j + i, ubuf + i * BS_XROW_B, 130, 0, 129);
                if (lane == 0) mbar_arrive(smem_u32(&sm.bars.tok[par]));
            }
        } else {
            auto build_up = [&](int i) {                   // i = par (mod 3): this warpgroup's (i / 3)-th input row
                const int seq = i / NFRONT;
                const int slot = par * BS_NIK + (seq % BS_NIK);
                mbar_wait(smem_u32(&sm.bars.in_full[slot]), (seq / BS_NIK) & 1);
                build_a(0, i, sm.in_ring[slot], 128, 0, 127);
                if (lane == 0) mbar_arrive(smem_u32(&sm.bars.in_empty[slot]));
            };
            // this warpgroup's A_up rows run one of its rows ahead of its epilogue (MMA_up(j) reads A rows j, j+1, j+2; the ring
            // of 5 must not be overrun: a build that waits for a free slot would hold back the epilogue behind it)
            if (par < n_in) build_up(par);
            int it = 0;
            for (int j = par; j < n_out[0]; j += nwg, ++it) {
                BS_STAMP(wg, it, 0);
                if (j + nwg < n_in) build_up(j + nwg);
                const int h = y0 - 2 + j;
                const int ds = j % R::ND0;
                BS_STAMP(wg, it, 1);
                mbar_wait(smem_u32(&sm.bars.d_full[0][ds]), (j / R::ND0) & 1);
                BS_STAMP(wg, it, 2);
                tc_fence_after();
                uint32_t v[16];
                tmem_ld16(lane_base + R::D0 + ds * 16, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sm.bars.d_empty[0][ds]));
                float2 x[8];
                bs_bias16(v, sm.cst.b_up, x);
#pragma unroll
                for (int p = 0; p < 8; ++p) x[p] = (p < NP) ? bs_act2<ACT_UP>(x[p], a.act_up) : make_float2(0.0f, 0.0f);
                const bool inside = col_in && (h >= 0) && (h < a.H);
                uint4 u0, u1;
                {
                    float2 y[4];
                    bs_affine8(x, sm.cst.g0p, sm.cst.beta0, y);
                    u0 = bs_pack8(y);
                    bs_affine8(x + 4, sm.cst.g0p + 8, sm.cst.beta0 + 8, y);
                    u1 = bs_pack8(y);
                    if (!inside) { u0 = make_uint4(0, 0, 0, 0); u1 = u0; }
                }
                uint8_t* urow = sm.u_ring[par][it & 1][0];
                *reinterpret_cast<uint4*>(urow + static_cast<size_t>(m + 1) * 16) = u0;
                *reinterpret_cast<uint4*>(urow + static_cast<size_t>(130 + m + 1) * 16) = u1;
                const int k = j - 2;                                     // this row's index in the conv1 / output sequence
                if (k >= 0 && k < rows) {
                    // x0 ring, no barriers of its own.  Free slot: this row runs at most NA + ND + NA + ND = 16 (+2) rows ahead of
                    // the back warpgroup (the A_c0 / D_c0 / A_c1 / D_c1 rings are bounded), the ring holds 32.  Visibility: the
                    // write precedes this thread's a_full arrive (release), and D_c1(k) - which the reader waits for - is
                    // downstream of that arrive through the conv0 / conv1 MMAs of rows k+1, k+2.
                    const int xs = k % BS_NX;
                    *reinterpret_cast<uint4*>(sm.x0_ring[xs] + static_cast<size_t>(m) * 16) = bs_pack8(x);
                    *reinterpret_cast<uint4*>(sm.x0_ring[xs] + static_cast<size_t>(128 + m) * 16) = bs_pack8(x + 4);
                }
                BS_STAMP(wg, it, 3);
                named_bar_sync(1 + par, 128);                            // the u row is complete
                build_a(1, j, urow, 130, 0, 129);
                // the next iteration but one rewrites this exchange row: every thread has passed the next barrier by then
            }
        }
    } else if (stage_of_wg == 1) {
        // =============================== middle warpgroups: conv0 epilogue -> A_c1 ===============================
        int it = 0;
        for (int k = par; k < n_out[1]; k += nwg, ++it) {
            const int h = y0 - 1 + k;
            const int ds = k % R::ND1;
            BS_STAMP(wg, it, 0);
            mbar_wait(smem_u32(&sm.bars.d_full[1][ds]), (k / R::ND1) & 1);
            BS_STAMP(wg, it, 1);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16(lane_base + R::D1 + ds * 16, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.bars.d_empty[1][ds]));
            float2 x[8];
            bs_bias16(v, sm.cst.b_c0, x);
#pragma unroll
            for (int p = 0; p < 8; ++p) x[p] = (p < NP) ? bs_act2<ACT_IN>(x[p], a.act_inner) : make_float2(0.0f, 0.0f);
            const bool inside = col_in && (h >= 0) && (h < a.H);
            uint4 w0, w1;
            {
                float2 y[4];
                bs_affine8(x, sm.cst.g1p, sm.cst.beta1, y);
                w0 = bs_pack8(y);
                bs_affine8(x + 4, sm.cst.g1p + 8, sm.cst.beta1 + 8, y);
                w1 = bs_pack8(y);
                if (!inside) { w0 = make_uint4(0, 0, 0, 0); w1 = w0; }
            }
            uint8_t* wrow = sm.w_ring[par][it & 1];
            *reinterpret_cast<uint4*>(wrow + static_cast<size_t>(m + 1) * 16) = w0;
            *reinterpret_cast<uint4*>(wrow + static_cast<size_t>(130 + m + 1) * 16) = w1;
            BS_STAMP(wg, it, 2);
            named_bar_sync(1 + BS_WG_F + par, 128);
            build_a(2, k, wrow, 130, 0, 129);
            BS_STAMP(wg, it, 3);
        }
    } else {
        // =============================== back warpgroups: conv1 epilogue + residual -> global ===============================
        const size_t plane = static_cast<size_t>(a.H) * a.W * 8;
        const bool lane_valid = (m >= 2) && (m < 2 + BS_VALID) && col_in;
        for (int k = par; k < rows; k += nwg) {
            const int h = y0 + k;
            const size_t goff = ((static_cast<size_t>(fb) * 2) * a.H + h) * static_cast<size_t>(a.W) * 8 + static_cast<size_t>(col) * 8;
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
            if (!a.has_up && lane_valid) {
                r0 = __ldg(reinterpret_cast<const uint4*>(a.resid + goff));
                r1 = __ldg(reinterpret_cast<const uint4*>(a.resid + goff + plane));
            }
            const int ds = k % R::ND2;
            BS_STAMP(wg, k, 0);
            mbar_wait(smem_u32(&sm.bars.d_full[2][ds]), (k / R::ND2) & 1);
            BS_STAMP(wg, k, 1);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16(lane_base + R::D2 + ds * 16, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.bars.d_empty[2][ds]));
            if (a.has_up) {
                const int xs = k % BS_NX;
                r0 = *reinterpret_cast<const uint4*>(sm.x0_ring[xs] + static_cast<size_t>(m) * 16);
                r1 = *reinterpret_cast<const uint4*>(sm.x0_ring[xs] + static_cast<size_t>(128 + m) * 16);
            }
            float2 x[8];
            bs_bias16(v, sm.cst.b_c1, x);
            x[0] = add2(x[0], unpack_h2(r0.x)); x[1] = add2(x[1], unpack_h2(r0.y));
            x[2] = add2(x[2], unpack_h2(r0.z)); x[3] = add2(x[3], unpack_h2(r0.w));
            x[4] = add2(x[4], unpack_h2(r1.x)); x[5] = add2(x[5], unpack_h2(r1.y));
            x[6] = add2(x[6], unpack_h2(r1.z)); x[7] = add2(x[7], unpack_h2(r1.w));
            if (!HEAD) {
                if (lane_valid) {
                    *reinterpret_cast<uint4*>(a.out + goff) = bs_pack8(x);
                    *reinterpret_cast<uint4*>(a.out + goff + plane) = bs_pack8(x + 4);
                }
            } else if (lane_valid) {
                // bnerv_head_conv1's arithmetic on the f16 values the map would have held (bias, then one FMA per channel in
                // channel order): the image is bit-identical to the separate head launch
                const uint4 o0 = bs_pack8(x), o1 = bs_pack8(x + 4);
                const uint32_t ow[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
                float hacc[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) hacc[c] = sm.cst.head_b[c];
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    if (p < NP) {
                        const float2 v = unpack_h2(ow[p]);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            hacc[c] = fmaf(v.x, sm.cst.head_w[c][2 * p], hacc[c]);
                            hacc[c] = fmaf(v.y, sm.cst.head_w[c][2 * p + 1], hacc[c]);
                        }
                    }
                }
                const size_t hw = static_cast<size_t>(a.H) * a.W;
                const size_t pix = static_cast<size_t>(h) * a.W + col;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < a.head_cout) a.img[(static_cast<size_t>(fb) * a.head_cout + c) * hw + pix] = apply_act(hacc[c], a.head_act);
            }
            BS_STAMP(wg, k, 3);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == BS_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int g_bs_sms = 0;

static int bs_launch(const void* x_in, BsArgs& a, cudaStream_t stream) {
    if (g_bs_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_bs_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_bs_sms <= 0) g_bs_sms = 148;
    }
    a.dbg = g_bf_dbg;
    a.strips = (a.W + BS_VALID - 1) / BS_VALID;
    // One CTA per SM at a time (it owns all of TMEM).  A segment of R rows costs ~R + 12 row times (4 halo rows + pipeline
    // fill/drain): choose the number of segments per strip that minimises (waves over the SMs) x (R + 12).
    static const int forced_rows = getenv("BNERV_BS_ROWS") ? atoi(getenv("BNERV_BS_ROWS")) : 0;
    {
        int best_segs = 1;
        long long best_cost = -1;
        const int max_segs = (a.H + 7) / 8;
        for (int sg = 1; sg <= max_segs; ++sg) {
            const int r = (a.H + sg - 1) / sg;
            const int real = (a.H + r - 1) / r;
            const long long ctas = 1LL * a.B * a.strips * real;
            const long long waves = (ctas + g_bs_sms - 1) / g_bs_sms;
            const long long cost = waves * (r + 12);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_segs = sg; }
        }
        int seg_rows = (a.H + best_segs - 1) / best_segs;
        if (forced_rows > 0) seg_rows = forced_rows;
        if (a.has_up && a.s == 2) seg_rows += seg_rows & 1;             // segments start on even output rows
        a.seg_rows = seg_rows;
        a.segs = (a.H + seg_rows - 1) / seg_rows;
    }
    const long long grid = 1LL * a.B * a.strips * a.segs;
    if (grid > 0x3fffffffLL) return set_error(BNERV_E_UNSUPPORTED, "block_stream: too many CTAs");

    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return set_error(BNERV_E_NODRIVER, "cuTensorMapEncodeTiled entry point not available");
    CUtensorMap tm;
    const int inH = (a.has_up && a.s == 2) ? a.H / 2 : a.H, inW = (a.has_up && a.s == 2) ? a.W / 2 : a.W;
    cuuint64_t dims[3] = {2ull * inW, static_cast<cuuint64_t>(inH), 2ull * a.B};
    cuuint64_t strides[2] = {16ull * inW, 16ull * inW * inH};
    cuuint32_t box[3] = {256, 1, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(x_in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(static_cast<int>(r), "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));

    using KernelFn = void (*)(const CUtensorMap, const BsArgs);
    // pad channels may be skipped when the activations map 0 to 0 (every block activation but OutImg's tanh01)
    const bool np6 = a.C <= 12 && a.act_up != BNERV_ACT_TANH01 && a.act_inner != BNERV_ACT_TANH01;
    const bool spec = (!a.has_up || a.act_up == BNERV_ACT_SIN) && a.act_inner == BNERV_ACT_GELU;
    const bool s2 = a.has_up && a.s == 2;
    KernelFn fn;
    int slot;
    if (a.dbg && !s2 && !a.img) { fn = block_stream_kernel<BNERV_ACT_SIN, BNERV_ACT_GELU, 6, true, false>; slot = 8; }     // bring-up stamps: the 12-channel sin/GELU form
    else if (spec && np6) { fn = s2 ? block_stream_kernel<BNERV_ACT_SIN, BNERV_ACT_GELU, 6, false, true> : block_stream_kernel<BNERV_ACT_SIN, BNERV_ACT_GELU, 6, false, false>; slot = 0 + s2; }
    else if (spec) { fn = s2 ? block_stream_kernel<BNERV_ACT_SIN, BNERV_ACT_GELU, 8, false, true> : block_stream_kernel<BNERV_ACT_SIN, BNERV_ACT_GELU, 8, false, false>; slot = 2 + s2; }
    else if (np6) { fn = s2 ? block_stream_kernel<-1, -1, 6, false, true> : block_stream_kernel<-1, -1, 6, false, false>; slot = 4 + s2; }
    else { fn = s2 ? block_stream_kernel<-1, -1, 8, false, true> : block_stream_kernel<-1, -1, 8, false, false>; slot = 6 + s2; }
    if (a.img) {          // + 1x1 head conv + OutImg in the back warpgroup: the sin / GELU, s = 1 forms (the last block of NeRV-Boost)
        if (!spec || s2) return set_error(BNERV_E_UNSUPPORTED, "block_stream_head: only the sin / GELU, s = 1 form");
        if (np6) { fn = block_stream_kernel<BNERV_ACT_SIN, BNERV_ACT_GELU, 6, false, false, true>; slot = 9; }
        else     { fn = block_stream_kernel<BNERV_ACT_SIN, BNERV_ACT_GELU, 8, false, false, true>; slot = 10; }
    }
    static bool attr_set[11][32] = {};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    const size_t smem = sizeof(BsSmem) + 1024;
    if (!attr_set[slot][cur_dev & 31]) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
        attr_set[slot][cur_dev & 31] = true;
    }
    static const bool verbose = getenv("BNERV_BF_VERBOSE") != nullptr;
    if (verbose)
        fprintf(stderr, "block_stream: has_up=%d %dx%d B=%d -> %d strips x %d segments of %d rows = %lld CTAs, smem %zu B\n", a.has_up,
                a.H, a.W, a.B, a.strips, a.segs, a.seg_rows, grid, smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(BS_THREADS);
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
        return set_error(static_cast<int>(e), "block_stream_kernel launch: %s", cudaGetErrorString(e));
    }
    return check_launch("block_stream_kernel");
}

// 17..32 channels: block_stream32.cu
int resblock_stream32(const void* u, const void* x0, int B, int C, int H, int W, const void* w_c0, const float* b_c0,
                      const void* w_c1, const float* b_c1, int act_inner, const float* g1p, const float* beta1, void* out,
                      const float* head_w, const float* head_b, int head_cout, int head_act, float* img, cudaStream_t stream);

int upconv_stream32(const void* x, int B, int C, int H, int W, const void* w_up, const float* b_up, int act_up, const float* g0p,
                    const float* beta0, void* x0, void* u, cudaStream_t stream);

}  // namespace bnerv

using namespace bnerv;

extern "C" int bnerv_upconv_stream(const void* x, int B, int Cin, int H, int W, const void* w_up, const float* b_up, int C, int act_up,
                                   const float* g0p, const float* beta0, void* x0, void* u, void* stream) {
    if (!x || !w_up || !b_up || !g0p || !beta0 || !x0 || !u) return set_error(BNERV_E_BADARG, "upconv_stream: null pointer");
    if (B <= 0 || Cin <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "upconv_stream: non-positive size");
    if (Cin <= 16 || Cin > 32 || C <= 16 || C > 32) return set_error(BNERV_E_UNSUPPORTED, "upconv_stream: Cin = %d, C = %d (17..32 channels each)", Cin, C);
    if (act_up < BNERV_ACT_NONE || act_up >= BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "upconv_stream: activation code");
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_up) | reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(u);
    if (align_or & 15) return set_error(BNERV_E_BADARG, "upconv_stream: pointers must be 16-byte aligned");
    return upconv_stream32(x, B, C, H, W, w_up, b_up, act_up, g0p, beta0, x0, u, static_cast<cudaStream_t>(stream));
}

extern "C" int bnerv_nerv_block_stream(const void* x, int B, int Cin, int H, int W, const void* w_up, const float* b_up, int k_up,
                                       int s, int act_up, const void* w_c0, const float* b_c0, const void* w_c1, const float* b_c1,
                                       int C, int act_inner, const float* g0p, const float* beta0, const float* g1p,
                                       const float* beta1, void* out, void* stream) {
    if (!x || !w_up || !b_up || !w_c0 || !b_c0 || !w_c1 || !b_c1 || !out) return set_error(BNERV_E_BADARG, "nerv_block_stream: null pointer");
    if (!g0p || !beta0 || !g1p || !beta1) return set_error(BNERV_E_BADARG, "nerv_block_stream: the four TAT tables are required");
    if (B <= 0 || Cin <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "nerv_block_stream: non-positive size");
    if (k_up != 3 || (s != 1 && s != 2)) return set_error(BNERV_E_UNSUPPORTED, "nerv_block_stream: up-conv k = %d, s = %d (only k = 3, s = 1 or 2)", k_up, s);
    if (C > 16 || Cin > 16) return set_error(BNERV_E_UNSUPPORTED, "nerv_block_stream: C = %d / Cin = %d (at most 16 channels)", C, Cin);
    if (act_up < BNERV_ACT_NONE || act_up > BNERV_ACT_TANH01 || act_inner < BNERV_ACT_NONE || act_inner > BNERV_ACT_TANH01)
        return set_error(BNERV_E_UNSUPPORTED, "nerv_block_stream: activation code");
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_up) | reinterpret_cast<uintptr_t>(w_c0) |
                               reinterpret_cast<uintptr_t>(w_c1) | reinterpret_cast<uintptr_t>(out);
    if (align_or & 15) return set_error(BNERV_E_BADARG, "nerv_block_stream: pointers must be 16-byte aligned");
    BsArgs a{};
    a.B = B; a.H = H * s; a.W = W * s; a.C = C; a.s = s; a.has_up = 1; a.act_up = act_up; a.act_inner = act_inner;
    a.w_up = static_cast<const __half*>(w_up); a.w_c0 = static_cast<const __half*>(w_c0); a.w_c1 = static_cast<const __half*>(w_c1);
    a.b_up = b_up; a.b_c0 = b_c0; a.b_c1 = b_c1;
    a.g0p = g0p; a.beta0 = beta0; a.g1p = g1p; a.beta1 = beta1;
    a.out = static_cast<__half*>(out);
    return bs_launch(x, a, static_cast<cudaStream_t>(stream));
}

extern "C" int bnerv_resblock_stream(const void* u, const void* x0, int B, int C, int H, int W, const void* w_c0, const float* b_c0,
                                     const void* w_c1, const float* b_c1, int act_inner, const float* g1p, const float* beta1,
                                     void* out, void* stream) {
    if (!u || !x0 || !w_c0 || !b_c0 || !w_c1 || !b_c1 || !g1p || !beta1 || !out) return set_error(BNERV_E_BADARG, "resblock_stream: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "resblock_stream: non-positive size");
    if (C > 32) return set_error(BNERV_E_UNSUPPORTED, "resblock_stream: C = %d (at most 32 channels)", C);
    if (act_inner < BNERV_ACT_NONE || act_inner > BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "resblock_stream: activation code");
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(w_c0) |
                               reinterpret_cast<uintptr_t>(w_c1) | reinterpret_cast<uintptr_t>(out);
    if (align_or & 15) return set_error(BNERV_E_BADARG, "resblock_stream: pointers must be 16-byte aligned");
    if (C > 16)
        return resblock_stream32(u, x0, B, C, H, W, w_c0, b_c0, w_c1, b_c1, act_inner, g1p, beta1, out, nullptr, nullptr, 0, 0, nullptr,
                                 static_cast<cudaStream_t>(stream));
    BsArgs a{};
    a.B = B; a.H = H; a.W = W; a.C = C; a.s = 1; a.has_up = 0; a.act_up = BNERV_ACT_NONE; a.act_inner = act_inner;
    a.w_c0 = static_cast<const __half*>(w_c0); a.w_c1 = static_cast<const __half*>(w_c1);
    a.b_c0 = b_c0; a.b_c1 = b_c1;
    a.g1p = g1p; a.beta1 = beta1;
    a.resid = static_cast<const __half*>(x0);
    a.out = static_cast<__half*>(out);
    return bs_launch(u, a, static_cast<cudaStream_t>(stream));
}

extern "C" int bnerv_resblock_stream_head(const void* u, const void* x0, int B, int C, int H, int W, const void* w_c0, const float* b_c0,
                                          const void* w_c1, const float* b_c1, int act_inner, const float* g1p, const float* beta1,
                                          const float* head_w, const float* head_b, int head_cout, int head_act, float* img,
                                          void* stream) {
    if (!u || !x0 || !w_c0 || !b_c0 || !w_c1 || !b_c1 || !g1p || !beta1 || !head_w || !img) return set_error(BNERV_E_BADARG, "resblock_stream_head: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "resblock_stream_head: non-positive size");
    if (C <= 16 || C > 32) return set_error(BNERV_E_UNSUPPORTED, "resblock_stream_head: C = %d (17..32 channels)", C);
    if (head_cout < 1 || head_cout > 4) return set_error(BNERV_E_UNSUPPORTED, "resblock_stream_head: %d head channels (1..4)", head_cout);
    if (act_inner < BNERV_ACT_NONE || act_inner > BNERV_ACT_TANH01 || head_act < BNERV_ACT_NONE || head_act > BNERV_ACT_TANH01)
        return set_error(BNERV_E_UNSUPPORTED, "resblock_stream_head: activation code");
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(w_c0) |
                               reinterpret_cast<uintptr_t>(w_c1);
    if (align_or & 15) return set_error(BNERV_E_BADARG, "resblock_stream_head: pointers must be 16-byte aligned");
    return resblock_stream32(u, x0, B, C, H, W, w_c0, b_c0, w_c1, b_c1, act_inner, g1p, beta1, nullptr, head_w, head_b, head_cout,
                             head_act, img, static_cast<cudaStream_t>(stream));
}

extern "C" int bnerv_nerv_block_stream_head(const void* x, int B, int Cin, int H, int W, const void* w_up, const float* b_up,
                                            const void* w_c0, const float* b_c0, const void* w_c1, const float* b_c1, int C,
                                            const float* g0p, const float* beta0, const float* g1p, const float* beta1,
                                            const float* head_w, const float* head_b, int head_cout, int head_act, float* img,
                                            void* stream) {
    if (!x || !w_up || !b_up || !w_c0 || !b_c0 || !w_c1 || !b_c1 || !head_w || !img) return set_error(BNERV_E_BADARG, "nerv_block_stream_head: null pointer");
    if (!g0p || !beta0 || !g1p || !beta1) return set_error(BNERV_E_BADARG, "nerv_block_stream_head: the four TAT tables are required");
    if (B <= 0 || Cin <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(BNERV_E_BADARG, "nerv_block_stream_head: non-positive size");
    if (C > 16 || Cin > 16) return set_error(BNERV_E_UNSUPPORTED, "nerv_block_stream_head: C = %d / Cin = %d (at most 16 channels)", C, Cin);
    if (head_cout < 1 || head_cout > 4) return set_error(BNERV_E_UNSUPPORTED, "nerv_block_stream_head: %d head channels (1..4)", head_cout);
    if (head_act < BNERV_ACT_NONE || head_act > BNERV_ACT_TANH01) return set_error(BNERV_E_UNSUPPORTED, "nerv_block_stream_head: activation code");
    const uintptr_t align_or = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_up) | reinterpret_cast<uintptr_t>(w_c0) |
                               reinterpret_cast<uintptr_t>(w_c1);
    if (align_or & 15) return set_error(BNERV_E_BADARG, "nerv_block_stream_head: pointers must be 16-byte aligned");
    BsArgs a{};
    a.B = B; a.H = H; a.W = W; a.C = C; a.s = 1; a.has_up = 1; a.act_up = BNERV_ACT_SIN; a.act_inner = BNERV_ACT_GELU;
    a.w_up = static_cast<const __half*>(w_up); a.w_c0 = static_cast<const __half*>(w_c0); a.w_c1 = static_cast<const __half*>(w_c1);
    a.b_up = b_up; a.b_c0 = b_c0; a.b_c1 = b_c1;
    a.g0p = g0p; a.beta0 = beta0; a.g1p = g1p; a.beta1 = beta1;
    a.head_w = head_w; a.head_b = head_b; a.head_cout = head_cout; a.head_act = head_act; a.img = img;
    return bs_launch(x, a, static_cast<cudaStream_t>(stream));
}
