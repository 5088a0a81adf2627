// Weight gradient of the fused conv for sm_100a:  dW[tap][m][c] = sum_{b,h,w} dY[b,m,h,w] * X[b,c,h+r-pad,w+s-pad]
// (the transpose of CustomConv2d.forward, lib/quant_ops.py:39-41, as torch.autograd computes it for
// train_nerv_all.py:342-348).  A GEMM whose K dimension is the PIXELS:
//     D_tap[m, c] (+)= A[m, p] * B_tap[p, c],   A = dY (M = gradient channels), B = X shifted by the tap, K = 16 pixels
// Both operands are the C8 f16 maps exactly as they lie in HBM: a TMA box lands in shared memory as
// [channel group][row][px][8 ch] and 8 neighbouring pixels x 8 channels (128 contiguous bytes) ARE one core matrix of
// the UMMA "MN-major, no swizzle" canonical layout (SBO = channel-group stride, LBO = 8-pixel stride), so no
// transposition pass exists anywhere: the tap shift is a start-address offset into the X halo tile.
//
// Work decomposition: job = (block of 128 gradient channels, chunk of Nc input channels, kernel row r); a job keeps its
// 3 taps x Nc f32 accumulators in TMEM (<= 480 of 512 columns) while its CTAs sweep disjoint ranges of pixel tiles
// (8 rows x 16 px; one UMMA M=128,N=Nc,K=16 per tile row and tap), then adds the partial sums into the f32
// accumulation buffer with vector reductions (red.global.add.v4.f32).  Jobs are equal-cost, CTA c takes job c % J.
//
// Narrow layers (3*Cin_p <= 256, i.e. Cin <= 80: every NeRV-S / E-NeRV-M stage above 135p): an N = Cin_p UMMA costs
// the same ~45 issue cycles as a wide one, so the three column taps of a kernel row are STACKED IN N: the X tile is
// loaded three times at px offsets -1, 0, +1 into consecutive channel-group planes ("replicas"), which keeps the
// MN-major group stride uniform, and one UMMA (N = 3*Cin_p) does what three did.  When 9*Cin_p <= 512 TMEM columns
// (Cin <= 48) all three kernel rows live in one job, so dY is fetched once instead of three times.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace bnerv {

constexpr int WG_THREADS = 192;
constexpr int WG_TILE_W  = 16;
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_SMEM_LIMIT = 227 * 1024;
constexpr int WG_BAR_BYTES  = (2 * WG_MAX_STAGES + 2) * 8 + 16;

struct WgradArgs {
    int B, H, W;
    int taps;            // 1 or 9
    int pad;             // 0 or 1
    int T;               // taps per job (1 or 3)
    int R;               // pixel rows per tile
    int m_groups;        // M_p / 8
    int m_p;             // padded gradient channels
    int c_groups;        // Cin_p / 8
    int cin_p;
    int nc;              // input channels per chunk (multiple of 16)
    int m_blocks, c_chunks, r_jobs;
    int jobs;            // m_blocks * c_chunks * r_jobs
    int tiles_x, tiles_y, tiles;
    int stages;
    int dy_bytes, x_bytes, stage_bytes;
    int swap_lbo_sbo;    // debugging switch (BNERV_WGRAD_SWAP)
    int stack;           // 1: column taps stacked in N through replicated X planes
    int rows_per_job;    // kernel rows handled inside one job (3 in stacked mode when 9*Cin_p <= 512, else 1)
    int reps;            // X replicas per stage (stacked: 3 * rows_per_job)
    int rep_bytes;       // bytes of one replica
    int n_mma;           // UMMA N
    int tap_cols;        // accumulator columns per tap
    int swapped;         // 1: operands exchanged (A = X, B = dY shifted), results written transposed with mirrored taps
    int out_mp, out_cinp;   // the caller's M_p / Cin_p (layout of acc) in swapped mode
    float* acc;          // [taps][M_p][Cin_p] f32, accumulated into
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MN-major, no swizzle: LBO = byte distance between core matrices adjacent in K (8 pixels), SBO = between core
// matrices adjacent in M/N (8 channels).  (cute::UMMA::make_umma_desc<Major::MN>, INTERLEAVE row.)
__device__ __forceinline__ uint64_t desc_hi_mn(uint32_t k_stride, uint32_t mn_stride, int swap) {
    return swap ? umma_desc_hi_noswz(mn_stride, k_stride) : umma_desc_hi_noswz(k_stride, mn_stride);
}

struct WgJob { int m_blk, c_chunk, r, t0, t1; bool any; };

__device__ __forceinline__ WgJob wg_job(const WgradArgs& a, int jj) {
    WgJob j;
    const int job = jj % a.jobs;
    int split = 0, splits = 1;
    if (static_cast<int>(gridDim.x) >= a.jobs) {
        split  = jj / a.jobs;
        splits = (static_cast<int>(gridDim.x) - 1 - job) / a.jobs + 1;
    }
    j.r = job % a.r_jobs;
    const int rest = job / a.r_jobs;
    j.c_chunk = rest % a.c_chunks;
    j.m_blk   = rest / a.c_chunks;
    j.t0 = static_cast<int>(static_cast<long long>(a.tiles) * split / splits);
    j.t1 = static_cast<int>(static_cast<long long>(a.tiles) * (split + 1) / splits);
    j.any = j.t1 > j.t0;
    return j;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    uint8_t*  bar_base  = smem + static_cast<size_t>(a.stages) * a.stage_bytes;
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + WG_MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        mbar_init(smem_u32(tfull_bar), 1);
        mbar_init(smem_u32(tempty_bar), 4);
        fence_mbar_init();
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t full = smem_u32(full_bar), empty = smem_u32(empty_bar);
    const uint32_t tfull = smem_u32(tfull_bar), tempty = smem_u32(tempty_bar);
    const int jj_end = (a.jobs > static_cast<int>(gridDim.x)) ? a.jobs : static_cast<int>(gridDim.x);
    const int tiles_img = a.tiles_x * a.tiles_y;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int jj = blockIdx.x; jj < jj_end; jj += gridDim.x) {
            const WgJob j = wg_job(a, jj);
            for (int t = j.t0; t < j.t1; ++t) {
                const int b = t / tiles_img, ti = t - b * tiles_img;
                const int ty = ti / a.tiles_x, tx = ti - ty * a.tiles_x;
                const int h0 = ty * a.R, w0 = tx * WG_TILE_W;
                mbar_wait(empty + stage * 8, phase ^ 1);
                if (elect_one()) {
                    const uint32_t dst = s_base + stage * a.stage_bytes;
                    mbar_expect_tx(full + stage * 8, static_cast<uint32_t>(a.dy_bytes + a.x_bytes));
                    tma_load_3d(dst, &tmDY, full + stage * 8, 2 * w0, h0, b * a.m_groups + j.m_blk * 16);
                    if (a.stack) {
                        for (int rp = 0; rp < a.reps; ++rp) {
                            const int rr = (a.rows_per_job == 3) ? rp / 3 : j.r, sx = rp % 3;
                            tma_load_3d(dst + a.dy_bytes + rp * a.rep_bytes, &tmX, full + stage * 8, 2 * (w0 + sx - 1), h0 + rr - 1,
                                        b * a.c_groups);
                        }
                    } else {
                        tma_load_3d(dst + a.dy_bytes, &tmX, full + stage * 8, 2 * (w0 - a.pad), h0 + j.r - a.pad,
                                    b * a.c_groups + j.c_chunk * (a.nc >> 3));
                    }
                }
                __syncwarp();
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0;
        uint32_t phase = 0, jphase = 0;
        const uint32_t idesc = umma_idesc_f16(128, a.n_mma) | (1u << 15) | (1u << 16);   // A and B MN-major
        const int xw = a.stack ? WG_TILE_W : WG_TILE_W + 2 * a.pad;                       // X tile row length (px)
        const int n_grp = a.stack ? a.rows_per_job : a.T;                                 // UMMAs per tile row
        const uint32_t grp16 = a.stack ? 3u * (static_cast<uint32_t>(a.rep_bytes) >> 4) : 1u;   // B start step between them (16 B units)
        const uint64_t a_hi = desc_hi_mn(128u, static_cast<uint32_t>(a.R) * WG_TILE_W * 16u, a.swap_lbo_sbo);
        const uint64_t b_hi = desc_hi_mn(128u, static_cast<uint32_t>(a.R) * xw * 16u, a.swap_lbo_sbo);
        for (int jj = blockIdx.x; jj < jj_end; jj += gridDim.x) {
            const WgJob j = wg_job(a, jj);
            if (!j.any) continue;
            mbar_wait(tempty, jphase ^ 1);          // the epilogue has drained the previous job's accumulators
            tc_fence_after();
            for (int t = j.t0; t < j.t1; ++t) {
                mbar_wait(full + stage * 8, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sdy16 = ((s_base + stage * a.stage_bytes) & 0x3FFFFu) >> 4;
                    const uint32_t sx16  = sdy16 + (static_cast<uint32_t>(a.dy_bytes) >> 4);
                    for (int y = 0; y < a.R; ++y) {
                        const uint64_t adesc = a_hi | static_cast<uint64_t>(sdy16 + y * WG_TILE_W);
                        for (int s = 0; s < n_grp; ++s) {
                            const uint64_t bdesc = b_hi | static_cast<uint64_t>(sx16 + y * xw + s * grp16);
                            umma_f16(tmem_base + s * a.n_mma, adesc, bdesc, idesc, (t > j.t0 || y > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(empty + stage * 8);
                    if (t == j.t1 - 1) umma_commit(tfull);
                }
                __syncwarp();
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
            jphase ^= 1;
        }
    } else {
        // ===================== epilogue: TMEM -> f32 reductions into acc =====================
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int m = q * 32 + lane;
        uint32_t jphase = 0;
        for (int jj = blockIdx.x; jj < jj_end; jj += gridDim.x) {
            const WgJob j = wg_job(a, jj);
            if (!j.any) continue;
            mbar_wait(tfull, jphase);
            tc_fence_after();
            const int mch = j.m_blk * 128 + m;
            const int c0 = j.c_chunk * a.nc;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
            const int n_taps_job = a.stack ? 3 * a.rows_per_job : a.T;      // tap blocks of tap_cols columns each
            for (int s = 0; s < n_taps_job; ++s) {
                const int tap = (a.taps == 1) ? 0 : ((a.stack && a.rows_per_job == 3) ? s : j.r * 3 + s);
                float* row = a.acc + (static_cast<size_t>(tap) * a.m_p + mch) * a.cin_p + c0;
                for (int g = 0; g < a.tap_cols; g += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + s * a.tap_cols + g, v);    // warp-collective: every lane takes part
                    tmem_ld_wait();
                    if (a.swapped) {
                        // this thread's row is an X channel, its columns are gradient channels of tap' = 8 - tap
                        if (mch < a.out_cinp && g < a.out_mp) {
                            float* col = a.acc + (static_cast<size_t>(a.taps - 1 - tap) * a.out_mp + g) * a.out_cinp + mch;
#pragma unroll
                            for (int k = 0; k < 16; ++k) atomicAdd(col + static_cast<size_t>(k) * a.out_cinp, __uint_as_float(v[k]));
                        }
                    } else if (mch < a.m_p && c0 + g < a.cin_p) {
#pragma unroll
                        for (int k = 0; k < 16; k += 4)
                            red_add_v4(row + g + k, __uint_as_float(v[k]), __uint_as_float(v[k + 1]),
                                       __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
            jphase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// finalize: acc [taps][M_p][Cin_p] (scaled by the loss scale) -> grad_oihw [Cout*s*s][Cin][k][k] (+=)
// gradient channel m = (i*s + j)*Cout_p + c of the un-shuffled map  <->  reference conv channel c*s*s + i*s + j
// ---------------------------------------------------------------------------------------------
__global__ void wgrad_finalize_kernel(const float* __restrict__ acc, int Cout, int Cin, int k, int s, int cout_p, int cin_p,
                                      const float* __restrict__ inv_scale, int accumulate, float* __restrict__ grad) {
    const int taps = k * k;
    const int m_p = s * s * cout_p;
    const size_t total = static_cast<size_t>(Cout) * s * s * Cin * taps;
    const float sc = inv_scale ? *inv_scale : 1.0f;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int tap = idx % taps;
        size_t r = idx / taps;
        const int ci = r % Cin;
        const int o = r / Cin;
        const int c = o / (s * s), ij = o - c * s * s;
        const int m = ij * cout_p + c;
        const float v = acc[(static_cast<size_t>(tap) * m_p + m) * cin_p + ci] * sc;
        grad[idx] = accumulate ? grad[idx] + v : v;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();      // conv_tc.cu

static int make_map3(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                     uint32_t b0, uint32_t b1, uint32_t b2) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return set_error(BNERV_E_NODRIVER, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1, s2};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(static_cast<int>(r), "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
    return 0;
}

}  // namespace bnerv

using namespace bnerv;

extern "C" size_t bnerv_wgrad_acc_numel(int M_p, int Cin, int k) {
    if (M_p <= 0 || Cin <= 0 || k <= 0) return 0;
    return static_cast<size_t>(k) * k * M_p * round_up(Cin, 16);
}

extern "C" int bnerv_conv_wgrad(const void* x, const void* dy, int B, int Cin, int H, int W, int M_p, int k, float* acc,
                                void* stream) {
    if (!x || !dy || !acc) return set_error(BNERV_E_BADARG, "conv_wgrad: null pointer");
    if (B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || M_p <= 0) return set_error(BNERV_E_BADARG, "conv_wgrad: non-positive size");
    if (M_p % 16) return set_error(BNERV_E_BADARG, "conv_wgrad: M_p must be a multiple of 16 (C8 padded channels)");
    if (k != 1 && k != 3) return set_error(BNERV_E_UNSUPPORTED, "conv_wgrad: kernel size %d (only 1 and 3)", k);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(acc)) & 15)
        return set_error(BNERV_E_BADARG, "conv_wgrad: pointers must be 16-byte aligned");
    WgradArgs a{};
    a.B = B; a.H = H; a.W = W;
    a.taps = k * k;
    a.pad = (k - 1) / 2;
    a.T = k;                                  // one kernel row of taps per job
    a.r_jobs = k;
    a.m_p = M_p; a.m_groups = M_p / 8;
    a.cin_p = round_up(Cin, 16); a.c_groups = a.cin_p / 8;
    // A conv to very few channels (the 3x3 head: M_p = 16) would spend a full M = 128, N = Cin_p UMMA per tap on 16 useful
    // rows.  Exchange the operands instead: D'[c, m] of tap' = sum_p X[p, c] dY[p + tap', m] equals the wanted
    // D[m, c] of tap = 8 - tap' (substitute q = p + tap'), the narrow dY becomes the tap-stacked N operand
    // (N = 3 * 16 per kernel row, all rows in one job) and the epilogue writes the transpose with mirrored taps.
    static const bool no_swap = getenv("BNERV_WGRAD_NO_SWAP") != nullptr;        // A/B switch
    a.out_mp = a.m_p; a.out_cinp = a.cin_p;
    a.swapped = (k == 3 && M_p <= 16 && a.cin_p >= 64 && !no_swap) ? 1 : 0;
    if (a.swapped) {
        const void* tmp = x; x = dy; dy = tmp;
        a.m_p = a.out_cinp; a.m_groups = a.m_p / 8;
        a.cin_p = a.out_mp; a.c_groups = a.cin_p / 8;
        M_p = a.m_p;
    }
    const int nc_max = 160;                   // T * nc <= 512 TMEM columns with T = 3
    a.c_chunks = (a.cin_p + nc_max - 1) / nc_max;
    a.nc = round_up((a.cin_p + a.c_chunks - 1) / a.c_chunks, 16);
    a.m_blocks = (M_p + 127) / 128;
    static const bool no_stack = getenv("BNERV_WGRAD_NO_STACK") != nullptr;      // A/B switch
    a.stack = (k == 3 && 3 * a.cin_p <= 256 && !no_stack) ? 1 : 0;
    a.rows_per_job = 1;
    a.n_mma = a.nc;
    a.tap_cols = a.nc;
    a.reps = 1;
    if (a.stack) {
        a.c_chunks = 1;
        a.nc = a.cin_p;
        a.rows_per_job = (9 * a.cin_p <= 512) ? 3 : 1;
        a.r_jobs = (a.rows_per_job == 3) ? 1 : 3;
        a.reps = 3 * a.rows_per_job;
        a.n_mma = 3 * a.cin_p;
        a.tap_cols = a.cin_p;
    }
    a.jobs = a.m_blocks * a.c_chunks * a.r_jobs;
    const int xw = a.stack ? WG_TILE_W : WG_TILE_W + 2 * a.pad;
    // rows per tile: 8 when at least 3 stages fit, else 4 (measured: 4-row tiles with twice the stages are 8 % slower at
    // 112 channels - more TMA boxes and commits per byte)
    a.R = 8;
    for (;;) {
        a.dy_bytes = 16 * a.R * WG_TILE_W * 16;
        a.rep_bytes = (a.nc / 8) * a.R * xw * 16;
        a.x_bytes = a.reps * a.rep_bytes;
        a.stage_bytes = a.dy_bytes + a.x_bytes;
        a.stages = (WG_SMEM_LIMIT - WG_BAR_BYTES) / a.stage_bytes;
        if (a.stages >= 3 || a.R == 4) break;
        a.R = 4;
    }
    if (a.stages > WG_MAX_STAGES) a.stages = WG_MAX_STAGES;
    if (a.stages < 2) return set_error(BNERV_E_UNSUPPORTED, "conv_wgrad: pipeline stages do not fit");
    a.tiles_x = (W + WG_TILE_W - 1) / WG_TILE_W;
    a.tiles_y = (H + a.R - 1) / a.R;
    const long long tiles = 1LL * B * a.tiles_x * a.tiles_y;
    if (tiles > 0x3fffffffLL) return set_error(BNERV_E_UNSUPPORTED, "conv_wgrad: too many tiles");
    a.tiles = static_cast<int>(tiles);
    a.acc = acc;
    static const bool swap = getenv("BNERV_WGRAD_SWAP") != nullptr;
    a.swap_lbo_sbo = swap ? 1 : 0;

    CUtensorMap tmDY, tmX;
    int rc = make_map3(&tmDY, dy, 2ull * W, H, 1ull * B * a.m_groups, 16ull * W, 16ull * W * H, 2 * WG_TILE_W, a.R, 16);
    if (rc) return rc;
    rc = make_map3(&tmX, x, 2ull * W, H, 1ull * B * a.c_groups, 16ull * W, 16ull * W * H, 2 * xw, a.R, a.nc / 8);
    if (rc) return rc;

    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    // every job gets the same number of CTAs; never more CTAs per job than it has tiles
    int per_job = num_sms / a.jobs;
    if (per_job > a.tiles) per_job = a.tiles;
    int grid = per_job >= 1 ? per_job * a.jobs : num_sms;
    const size_t smem_bytes = static_cast<size_t>(a.stages) * a.stage_bytes + WG_BAR_BYTES;
    static bool attr_set_dev[32] = {};              // function attributes are per device (context)
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& attr_set = attr_set_dev[cur_dev & 31];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_LIMIT);
        if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
        attr_set = true;
    }
    conv_wgrad_kernel<<<grid, WG_THREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmDY, tmX, a);
    return check_launch("conv_wgrad_kernel");
}

extern "C" int bnerv_wgrad_finalize(const float* acc, int Cout, int Cin, int k, int s, const float* inv_scale, int accumulate,
                                    float* grad_oihw, void* stream) {
    if (!acc || !grad_oihw) return set_error(BNERV_E_BADARG, "wgrad_finalize: null pointer");
    if (Cout <= 0 || Cin <= 0 || s <= 0 || (k != 1 && k != 3)) return set_error(BNERV_E_BADARG, "wgrad_finalize: bad size");
    const size_t total = static_cast<size_t>(Cout) * s * s * Cin * k * k;
    size_t g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    wgrad_finalize_kernel<<<static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        acc, Cout, Cin, k, s, round_up(Cout, 16), round_up(Cin, 16), inv_scale, accumulate, grad_oihw);
    return check_launch("wgrad_finalize_kernel");
}
