/*
 * bnerv_b200.h — C-ABI of the B200-native Boosting-NeRV conditional-decoder hot path.
 *
 * Every entry point takes plain device pointers, integer shapes and a CUDA stream handle
 * (passed as void* so the header needs no CUDA include), allocates nothing, and returns
 *      0   on success,
 *     <0   BNERV_E_* : the arguments are outside what the kernel supports (nothing launched),
 *     >0   a cudaError_t / CUresult value from the launch.
 * bnerv_last_error() returns a thread-local, human-readable description of the last non-zero return.
 *
 * The reference (Xinjie-Q/Boosting-NeRV) is pure Python/PyTorch and has no FFI; the functions below
 * replace what its modules do through torch ops.  Citations are `file:line` in the reference tree.
 *
 * Data layouts
 *   NCHW f32 : the reference's layout, [B][C][H][W] float (model boundary only).
 *   C8  f16  : this library's activation layout between kernels, [B][Cp/8][H][W][8] __half with
 *              Cp = round_up(C, 16).  Channels >= C hold exact zeros.  One 16-byte group holds 8
 *              consecutive channels of one pixel, so a pixel row of one group is a dense run in HBM
 *              (TMA-friendly) and 8 neighbouring pixels of one group form one UMMA core matrix.
 *   packed conv weight : [taps][Kp/8][Np][8] __half (taps = k*k, Kp = round_up(Cin,16),
 *              Np = s*s*round_up(Cout,16)); row n' holds reference output channel c*s*s + i*s + j,
 *              i.e. PixelShuffle (model_blocks.py:204,217) is folded into the row order:
 *              n' = (i*s + j)*Cout_p + c, except s == 2 where n' = ((i*Cout_p/8 + c/8)*2 + j)*8 + c%8
 *              (both horizontal neighbours of a pixel in one 16-row group -> 32-byte stores).
 *              The layout is private to bnerv_pack_conv_weight / bnerv_conv_fused.  Padding is zero.
 */
#ifndef BNERV_B200_H_
#define BNERV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNERV_ABI_VERSION 4

/* error codes (negative) */
#define BNERV_E_BADARG      (-1)   /* null pointer / non-positive size / misaligned pointer          */
#define BNERV_E_UNSUPPORTED (-2)   /* shape or option outside what the kernels implement             */
#define BNERV_E_NODRIVER    (-3)   /* CUDA driver entry point (cuTensorMapEncodeTiled) not available  */

/* activation codes (model_blocks.py:136-158 — only the ones the shipped scripts select) */
#define BNERV_ACT_NONE   0
#define BNERV_ACT_SIN    1   /* Sin,  model_blocks.py:129-134                      */
#define BNERV_ACT_GELU   2   /* nn.GELU() exact-erf, model_blocks.py:146           */
#define BNERV_ACT_RELU   3   /* nn.ReLU, used inside SFTLayer, model_blocks.py:99   */
#define BNERV_ACT_TANH01 4   /* OutImg 'tanh': tanh(x)*0.5+0.5, model_blocks.py:61 */

int         bnerv_abi_version(void);
const char* bnerv_last_error(void);
/* Number of kernel launches issued by this library in the calling process since load (monotonic). */
uint64_t    bnerv_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Weight packing.  Replaces nothing in the reference (it feeds F.conv2d OIHW weights directly,
 * lib/quant_ops.py:39-41) — this is the "effective weight" ingest: pass `dequant_w ?? weight`.
 *   w_oihw  : [Cout*s*s][Cin][k][k] f32     bias : [Cout*s*s] f32 or NULL (treated as zeros)
 *   w_packed: [k*k][Kp/8][Np][8] f16        bias_packed : [Np] f32 (n' order)
 * k in {1,3}; s >= 1 is the PixelShuffle factor folded into the row order (s = 1: plain conv).
 * ---------------------------------------------------------------------------------------------- */
int bnerv_pack_conv_weight(const float* w_oihw, const float* bias, int Cout, int Cin, int k, int s,
                           void* w_packed, float* bias_packed, void* stream);

/* Quantised-weight ingest: the compression path's effective weight is
 * dequant_w = round(w / scale) * scale (Scale_T, lib/transform_ops.py:239-251; consumed at lib/quant_ops.py:40).
 * Given the integer codes round(w/scale) and the scale(s) this packs (float)code * scale - the same f32 product torch
 * forms - so the result is bit-identical to bnerv_pack_conv_weight(dequant_w, dequant_b) without materialising them.
 *   w_codes : [Cout*s*s][Cin][k][k] signed integers of `code_bytes` (1, 2 or 4) bytes (Scale_T does not clamp, so
 *             8-bit quantisation can exceed int8; the caller picks the narrowest type that holds its codes)
 *   w_scale : 1 float, or [Cout*s*s] when w_scale_per_channel;  b_codes/b_scale : same for the bias, b_codes may be NULL */
int bnerv_pack_conv_weight_q(const void* w_codes, const float* w_scale, int w_scale_per_channel,
                             const void* b_codes, const float* b_scale, int b_scale_per_channel, int code_bytes,
                             int Cout, int Cin, int k, int s, void* w_packed, float* bias_packed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused conv (the hot op).  One launch computes, for a stride-1 'same' conv with k in {1,3}:
 *     y   = conv_k(x; W, b)                        CustomConv2d.forward, lib/quant_ops.py:39-41
 *     y   = PixelShuffle_s(y)                      UpConv, model_blocks.py:213-220
 *     x0  = act(y)                                 NeRVBlock.forward, model_blocks.py:37
 *     x0 += resid                                  ResBlock_SFT.forward, model_blocks.py:89
 *     u   = x0 * g1p + beta                        SFTLayer.forward, model_blocks.py:105 (g1p = scale+1)
 * and writes x0 and/or u (C8 f16), or an NCHW f32 image for the head (model_blocks.py:57-63).
 *   x        : C8 f16 [B][Cin_p/8][H][W][8]
 *   w_packed, bias_packed : from bnerv_pack_conv_weight (same k, s)
 *   resid    : C8 f16 at the OUTPUT resolution [B][Cout_p/8][H*s][W*s][8], or NULL
 *   g1p,beta : f32 [B][Cout_p] (from bnerv_sft_affine), or both NULL for no affine
 *   out_pre  : C8 f16, receives x0 (may be NULL when out_aff is given)
 *   out_aff  : C8 f16, receives u  (NULL when g1p is NULL)
 *   out_nchw : f32 [B][Cout][H*s][W*s], receives x0 (after act) in the reference layout, or NULL
 * Computation: f16 operands, f32 accumulation (tcgen05.mma.cta_group::2 kind::f16, accumulators in TMEM);
 * each CTA pair keeps its weight tile resident in shared memory; when Cin is too wide for a useful tile to stay
 * resident the weights stream through the stage ring with the activations instead.  The launch uses programmatic dependent launch:
 * w_packed / bias_packed must not be written by the kernel enqueued immediately before this call unless
 * that kernel is one of this library's (none of which trigger early completion while writing them).
 * ---------------------------------------------------------------------------------------------- */
int bnerv_conv_fused(const void* x, int B, int Cin, int H, int W,
                     const void* w_packed, const float* bias_packed, int Cout, int k, int s,
                     int act, const void* resid, const float* g1p, const float* beta,
                     void* out_pre, void* out_aff, float* out_nchw, void* stream);

/* bnerv_conv_fused plus one more optional output for the TRAINING forward:
 *   out_deriv : C8 f16 at the output resolution, receives act'(y) evaluated at the pre-activation (cos(y) for
 *               SIN, Phi(y) + y*phi(y) for GELU, 1 for NONE) - the map the backward pass multiplies by, so neither the
 *               pre-activation nor a second transcendental pass is needed later.  Must be NULL when resid is given. */
int bnerv_conv_fused_ex(const void* x, int B, int Cin, int H, int W,
                        const void* w_packed, const float* bias_packed, int Cout, int k, int s,
                        int act, const void* resid, const float* g1p, const float* beta,
                        void* out_pre, void* out_aff, float* out_nchw, void* out_deriv, void* stream);

/* The higher-precision ("split") form of bnerv_conv_fused, selectable per layer: f16 operands carry 11 significant bits, which
 * bounds a deep cascade of 1000+-term contractions at ~2e-3 of the f32 result on trained weights ("precision").  A
 * split map keeps every value v as the f16 pair hi = f16(v), lo = f16(v - hi) (~22 bits) laid out as ONE C8 map over 3*Cp
 * channels, [hi | lo | hi].  The conv that consumes it is an ordinary bnerv_conv_fused over Cin = 3*Cp input channels whose
 * packed weight holds the input-channel blocks [W_hi ; W_hi ; W_lo] (W_hi = f16(W), W_lo = f16(W - W_hi)): the f32 accumulator
 * receives hi*W_hi + lo*W_hi + hi*W_lo, i.e. the product to ~2^-21 instead of 2^-11 - three times the tensor work, same kernel.
 *   split bit 0 : out_pre / out_aff are written as split maps [B][3*Cout_p/8][Ho][Wo][8]
 *   split bit 1 : resid is a split map; hi + lo are both added
 * x / Cin / w_packed are used as given (a split input is simply a map with 3*Cp channels).  Replaces the same reference lines
 * as bnerv_conv_fused. */
int bnerv_conv_fused_split(const void* x, int B, int Cin, int H, int W,
                           const void* w_packed, const float* bias_packed, int Cout, int k, int s,
                           int act, const void* resid, const float* g1p, const float* beta,
                           void* out_pre, void* out_aff, float* out_nchw, int split, void* stream);

/* One NeRVBlock (model_blocks.py:34-46 with the ResBlock_SFT of :74-89) = the three fused-conv launches above issued
 * back to back on `stream` and chained by programmatic dependent launch:
 *     x0 = act_up(PS_s(conv_k(x; w_up)))         u   = x0*g0p + beta0
 *     w  = act_inner(conv3(u; w_c0))*g1p + beta1 out = x0 + conv3(w; w_c1)
 * x: C8 f16 [B][Cin_p/8][H][W][8]; w_*: packed by bnerv_pack_conv_weight (w_up with k_up and s, the others k = 3, s = 1);
 * g0p/beta0/g1p/beta1: f32 [B][C_p] from bnerv_sft_affine; x0, u, wmap (workspaces the caller may reuse afterwards; x0
 * is also the block's pre-residual activation) and out: C8 f16 [B][C_p/8][H*s][W*s][8]. */
int bnerv_nerv_block_fwd(const void* x, int B, int Cin, int H, int W, const void* w_up, const float* b_up, int k_up,
                         int s, int act_up, const void* w_c0, const float* b_c0, const void* w_c1, const float* b_c1,
                         int C, int act_inner, const float* g0p, const float* beta0, const float* g1p,
                         const float* beta1, void* x0, void* u, void* wmap, void* out, void* stream);

/* ONE kernel per NeRVBlock for the narrow stages (C <= 48 channels; north_star: "each NeRVBlock is a single fused kernel that
 * stages the small per-stage feature map in shared memory"): same contract and the same arithmetic, operation by operation,
 * as bnerv_nerv_block_fwd (bit-identical results), but x0, u and w never leave the SM - a CTA keeps the weights of all three
 * convs resident and walks regions of the output map; the region's tile is rewritten in place (input -> u -> w) between the
 * three tcgen05 stages, x0 waits in shared memory for the residual.  One read of x, one write of out.
 *   Supported: k_up = 3, s in {1, 2}, round_up(C,16) <= 48, round_up(Cin,16) <= 64, s*s*round_up(C,16) <= 256; anything else
 *   returns BNERV_E_UNSUPPORTED (nothing launched) and the caller uses bnerv_nerv_block_fwd. */
int bnerv_nerv_block_fused(const void* x, int B, int Cin, int H, int W, const void* w_up, const float* b_up, int k_up,
                           int s, int act_up, const void* w_c0, const float* b_c0, const void* w_c1, const float* b_c1,
                           int C, int act_inner, const float* g0p, const float* beta0, const float* g1p,
                           const float* beta1, void* out, void* stream);
/* The ResBlock_SFT half alone (model_blocks.py:83-89) for blocks whose up-conv is too wide to fuse (PixelShuffle 3 / 5, wide
 * inputs): u = x0*g0p + beta0 and x0 come from a bnerv_conv_fused launch;  out = x0 + conv3(act_inner(conv3(u))*g1p + beta1).
 * u, x0, out: C8 f16 [B][C_p/8][H][W][8]; round_up(C,16) <= 48. */
int bnerv_resblock_fused(const void* u, const void* x0, int B, int C, int H, int W, const void* w_c0, const float* b_c0,
                         const void* w_c1, const float* b_c1, int act_inner, const float* g1p, const float* beta1,
                         void* out, void* stream);

/* The same two contracts in ROW-STREAMING form for stages of at most 16 channels (all 12-channel stages of NeRV-Boost XS / S):
 * a CTA streams the rows of a 128-column strip through up-conv -> conv0 -> conv1, and the A operand of every tcgen05.mma
 * comes from TENSOR MEMORY - the threads that own a row's pixels write its "row-im2col" (3 horizontal taps x 16 channels)
 * into TMEM once, the vertical taps select among the A rows of a ring - instead of being re-read from shared memory nine
 * times (17.6 vs 50.4 cycles per N = 16 MMA, profiles/r02_ts_probe_umma_ts_vs_ss.txt).  Bit-identical to
 * bnerv_nerv_block_fwd.  Supported: k_up = 3, s = 1 or 2 (PixelShuffle inside the kernel), C <= 16, Cin <= 16;
 * bnerv_resblock_stream: C <= 32 (17..32 channels run the K = 32 form of block_stream32.cu: A rows of two K steps, the
 * residual rows prefetched from global memory; E-NeRV-Boost M's 1080p stages).  Anything else returns BNERV_E_UNSUPPORTED
 * with nothing launched. */
int bnerv_nerv_block_stream(const void* x, int B, int Cin, int H, int W, const void* w_up, const float* b_up, int k_up,
                            int s, int act_up, const void* w_c0, const float* b_c0, const void* w_c1, const float* b_c1,
                            int C, int act_inner, const float* g0p, const float* beta0, const float* g1p,
                            const float* beta1, void* out, void* stream);
int bnerv_resblock_stream(const void* u, const void* x0, int B, int C, int H, int W, const void* w_c0, const float* b_c0,
                          const void* w_c1, const float* b_c1, int act_inner, const float* g1p, const float* beta1,
                          void* out, void* stream);

/* The up-conv half of a 17..32-channel NeRVBlock in the same row-streaming form (k = 3, no PixelShuffle, 17..32 input channels):
 * x0 = act_up(conv3(x) + b_up), u = x0*g0p + beta0 - what bnerv_conv_fused does for this layer (bit-identical), at the MMA-issue
 * floor instead of the generic epilogue's instruction count.  x, x0, u: C8 f16 maps of 32 padded channels. */
int bnerv_upconv_stream(const void* x, int B, int Cin, int H, int W, const void* w_up, const float* b_up, int C, int act_up,
                        const float* g0p, const float* beta0, void* x0, void* u, void* stream);

/* bnerv_conv_fused for one shape class - 3x3, stride 1, no PixelShuffle, equal padded widths of 32 or 48 channels (Cin, Cout in
 * 17..32 or 33..48) - in the row-streaming form: A rows in tensor memory, one launch = one conv with the same epilogue
 * (bias, activation, residual, TAT affine, out_pre and / or out_aff) and bit-identical results.  For E-NeRV-Boost M's 43-channel
 * 540p layers, where bnerv_conv_fused is bound by its epilogue's instruction count.  Other shapes: BNERV_E_UNSUPPORTED. */
int bnerv_conv_stream(const void* x, int B, int Cin, int H, int W, const void* w_packed, const float* bias_packed, int Cout,
                      int act, const void* resid, const float* g1p, const float* beta, void* out_pre, void* out_aff, void* stream);

/* bnerv_resblock_stream (17..32 channels) with the model's 1x1 head conv and OutImg (model_enerv.py:311-313 / model_nerv.py:56-57,
 * model_blocks.py:57-63) folded into the last warpgroup: img[b][c][h][w] = act(head_b[c] + sum_k head_w[c][k] * f16(out[k])), f32
 * NCHW, head_w = the raw f32 [head_cout][C] weights, head_cout <= 4.  The block output is not stored (it would be written once
 * and read once by bnerv_head_conv1); the image is bit-identical to that two-launch sequence.  act_inner must be GELU. */
int bnerv_resblock_stream_head(const void* u, const void* x0, int B, int C, int H, int W, const void* w_c0, const float* b_c0,
                               const void* w_c1, const float* b_c1, int act_inner, const float* g1p, const float* beta1,
                               const float* head_w, const float* head_b, int head_cout, int head_act, float* img, void* stream);

/* The same for a whole 12..16-channel NeRVBlock (3x3 up-conv without PixelShuffle, sin, GELU: the last block of NeRV-Boost,
 * model_nerv.py:53-57): bnerv_nerv_block_stream + the 1x1 head conv + OutImg, one kernel from the block input to the image.
 * Bit-identical to the two launches; measured 2 % SLOWER in a NeRV-S frame (the one back warpgroup becomes the critical stage of
 * the 720p block), so the engine keeps the separate head launch for <= 16 channels. */
int bnerv_nerv_block_stream_head(const void* x, int B, int Cin, int H, int W, const void* w_up, const float* b_up,
                                 const void* w_c0, const float* b_c0, const void* w_c1, const float* b_c1, int C,
                                 const float* g0p, const float* beta0, const float* g1p, const float* beta1,
                                 const float* head_w, const float* head_b, int head_cout, int head_act, float* img, void* stream);

/* Bring-up instrumentation for the fused-block kernel: device buffer of n_ctas*4*12 int64 receiving clock64 phase stamps of
 * the first 4 regions of the first n_ctas CTAs of subsequent launches (slot 11 = SM id); NULL switches it off. */
int bnerv_debug_set_buffer(void* buf, int n_ctas);

/* The 3x3 head conv to <= 3 channels (HNeRV_Boost.head_layer, model_hnerv.py:214,273) + OutImg (model_blocks.py:57-63)
 * in its own form: out[p,c] = act(b[c] + sum_tap P[p+tap][(tap,c)]) with P = X . Wp ONE 1x1 tensor-core contraction to
 * 9*Cout (<= 27) columns over the halo tile, summed over the taps from shared memory.  Same result as
 * bnerv_conv_fused(k = 3, out_nchw) with a third of the UMMAs (the N = 16 launch is bound by its A-operand reads).
 *   w_head_packed : bnerv_pack_head_weight output, [Kp/8][32][8] f16 (32*Kp halves); bias : the raw f32 [Cout]. */
int bnerv_pack_head_weight(const float* w_oihw, int Cout, int Cin, void* w_head_packed, void* stream);
int bnerv_head_conv3(const void* x, int B, int Cin, int H, int W, const void* w_head_packed, const float* bias,
                     int Cout, int act, float* out_nchw, void* stream);

/* The 1x1 head conv to <= 4 channels + OutImg (NeRV_Boost / ENeRV_Boost head_layer, model_nerv.py:41,56-57) as an
 * HBM-bound CUDA-core kernel: x C8 f16, w_oihw the RAW f32 [Cout][Cin] weights (no packing, no f16 rounding of the
 * weights), bias f32 [Cout] or NULL, out NCHW f32.  Replaces a chain of latency-bound N = 16 UMMAs. */
int bnerv_head_conv1(const void* x, int B, int Cin, int H, int W, const float* w_oihw, const float* bias, int Cout,
                     int act, float* out_nchw, void* stream);

/* Same contract and operand layouts as bnerv_conv_fused, computed by an f32 CUDA-core kernel on the
 * reference's own layouts (NCHW f32 activations, OIHW f32 weights) — the exact-arithmetic path used
 * for tiny layers and as the on-device cross-check of the tensor-core kernel.
 *   x: [B][Cin][H][W] f32; w: [Cout*s*s][Cin][k][k] f32; bias: [Cout*s*s] or NULL;
 *   resid/out_pre/out_aff: [B][Cout][H*s][W*s] f32; g1p/beta: [B][ldg] f32 with row stride `ldg`. */
int bnerv_conv_fused_f32(const float* x, int B, int Cin, int H, int W,
                         const float* w, const float* bias, int Cout, int k, int s,
                         int act, const float* resid, const float* g1p, const float* beta, int ldg,
                         float* out_pre, float* out_aff, void* stream);

/* ------------------------------------------------------------------------------------------------
 * TAT affine parameters for one SFTLayer (model_blocks.py:92-105):
 *     g1p  = Ws1 · relu(Ws0 · e + bs0) + bs1 + 1 ;   beta = Wh1 · relu(Wh0 · e + bh0) + bh1
 *   e: [B][ch_t] f32; Ws0/Wh0: [ch_t][ch_t]; Ws1/Wh1: [C][ch_t]; biases f32.
 *   g1p, beta: [B][Cp] f32 with Cp = round_up(C,16); entries >= C are written as 0.
 * `n_layers` SFT layers are evaluated by ONE launch: every pointer argument is a device array of
 * n_layers pointers / ints (built once per model by the caller).
 * ---------------------------------------------------------------------------------------------- */
typedef struct bnerv_sft_layer {
    const float *ws0, *bs0, *ws1, *bs1;   /* scale branch  SFT_scale_conv0/1 */
    const float *wh0, *bh0, *wh1, *bh1;   /* shift branch  SFT_shift_conv0/1 */
    float *g1p, *beta;                    /* outputs [B][Cp]                  */
    int C, Cp;
} bnerv_sft_layer;
int bnerv_sft_affine(const bnerv_sft_layer* layers_dev, int n_layers, const float* e, int B, int ch_t,
                     void* stream);

/* 1x1-conv / linear layer on a [B][Cin] vector with activation: y = act(W x + b)
 * (NeRV_MLP, model_blocks.py:66-71).  W: [Cout][Cin] f32 (a 1x1 OIHW weight is the same memory). */
int bnerv_linear_act(const float* x, int B, int Cin, const float* w, const float* bias, int Cout,
                     int act, float* y, void* stream);

/* Two independent layers in ONE launch - the stem of a frame is two small MLPs (NeRV_Boost: stem and stem_t,
 * model_nerv.py:47-52), and at batch 1 each of their layers is launch-latency, not work.
 *   bnerv_pe_linear_pair : both problems read x = cat(sin(t*bases), cos(t*bases)) (PositionEncoding.forward,
 *                          model_blocks.py:120-126: f32 product, sinf / cosf as torch's CUDA kernels evaluate them), built
 *                          in shared memory; Cin of both must be 2*levels, `x` is ignored.
 *   bnerv_linear_pair    : x given per problem.
 * y (f32 [B][Cout]) and/or y_c8: the output as the C8 f16 map [B][Cp/8][hw][8] of its .view(B, Cout/hw, h, w) (the cascade
 * input, model_nerv.py:50; padding channels are NOT written - zero the buffer once).  Per-output arithmetic is
 * bnerv_linear_act's, so results are bit-identical to it. */
typedef struct bnerv_linear_problem {
    const float* x;       /* [B][Cin] */
    const float* w;       /* [Cout][Cin] */
    const float* bias;    /* [Cout] or NULL */
    float* y;             /* [B][Cout] or NULL */
    void* y_c8;           /* C8 f16 map or NULL */
    int32_t Cin, Cout, act, hw;
} bnerv_linear_problem;
int bnerv_pe_linear_pair(const float* t, int B, const float* bases, int levels, const bnerv_linear_problem* probs, void* stream);
int bnerv_linear_pair(const bnerv_linear_problem* probs, int B, void* stream);

/* Layout conversion at the model boundary. */
int bnerv_nchw_to_c8(const float* x, int B, int C, int H, int W, void* y_c8, void* stream);
int bnerv_c8_to_nchw(const void* x_c8, int B, int C, int H, int W, float* y, void* stream);

/* nn.PixelShuffle(s) on NCHW f32 (model_blocks.py:204,217): a pure index permutation, bit-exact.
 *   x: [B][C*s*s][H][W] -> y: [B][C][H*s][W*s] */
int bnerv_pixel_shuffle(const float* x, int B, int C, int H, int W, int s, float* y, void* stream);

/* ================================================================================================
 * Backward of the cascade: what torch.autograd computes for the reference's
 * loss.backward() (train_nerv_all.py:342-348) through CustomConv2d / PixelShuffle / Sin / GELU / SFTLayer /
 * OutImg, as native kernels on the same C8 f16 maps.
 *
 * Gradient maps are C8 f16 multiplied by ONE power-of-two loss scale S per backward pass, chosen on the device
 * by bnerv_head_bwd (scale[0] = S, scale[1] = 1/S, no host synchronisation); every f32 reduction below is
 * S * (true gradient) and is un-scaled by the *_finalize calls / by the caller with scale[1].
 * "Un-shuffled" gradient map of a PixelShuffle(s) up-conv: C8 f16 [B][s*s*Cout_p/8][H][W][8] whose channel
 * m = (i*s + j)*Cout_p + c is reference conv channel c*s*s + i*s + j (s = 1: the plain C8 map).
 * ================================================================================================ */

/* OutImg 'tanh' transposed (model_blocks.py:61): dz = S * dimg * 2*img*(1 - img), written as a C8 f16 map.
 *   dimg, img : NCHW f32 [B][C][H][W] (img = the forward output); amax_scratch : 1 float of device scratch;
 *   scale : 2 floats (device), receives {S, 1/S}; dz_c8 : C8 f16 [B][Cp/8][H][W][8]. */
int bnerv_head_bwd(const float* dimg, const float* img, int B, int C, int H, int W, float* amax_scratch,
                   float* scale, void* dz_c8, void* stream);

/* dgrad: dx = conv_fused(x = un-shuffled dy, w = this packing, Cin' = s*s*Cout_p, Cout' = Cin, same k, s' = 1,
 * act NONE, zero bias).  Packs W[o][ci][r][q] (OIHW f32) transposed and tap-flipped into the layout
 * bnerv_conv_fused reads; w_packed holds bnerv_packed_weight_numel(Cin, s*s*Cout_p, k, 1) halves. */
int bnerv_pack_conv_weight_dgrad(const float* w_oihw, int Cout, int Cin, int k, int s, void* w_packed, void* stream);

/* wgrad: acc[tap][m][c] += sum_{b,h,w} dy[b,m,h,w] * x[b,c,h+r-pad,w+q-pad]   (tcgen05, K = pixels)
 *   x  : C8 f16 [B][Cin_p/8][H][W][8], the conv's forward input;  dy : un-shuffled C8 f16 gradient with M_p channels;
 *   acc: f32 [k*k][M_p][Cin_p], caller-zeroed, accumulated with vector reductions (bnerv_wgrad_acc_numel floats). */
int bnerv_conv_wgrad(const void* x, const void* dy, int B, int Cin, int H, int W, int M_p, int k, float* acc,
                     void* stream);
size_t bnerv_wgrad_acc_numel(int M_p, int Cin, int k);
/* acc (M_p = s*s*Cout_p) -> reference layout: grad_oihw[Cout*s*s][Cin][k][k] (=|+=) acc * (*inv_scale). */
int bnerv_wgrad_finalize(const float* acc, int Cout, int Cin, int k, int s, const float* inv_scale, int accumulate,
                         float* grad_oihw, void* stream);
/* un-shuffled channel sums [s*s*Cout_p] -> grad[Cout*s*s] (=|+=) acc * (*inv_scale) in the reference order. */
int bnerv_bias_finalize(const float* acc, int Cout, int s, const float* inv_scale, int accumulate, float* grad,
                        void* stream);

/* out[(per_b ? b : 0)][c] += sum_{h,w(,b)} x[b,c,h,w] over a C8 f16 map with Cp (multiple of 8) channels. */
int bnerv_channel_sum(const void* x_c8, int B, int Cp, int H, int W, int per_b, float* out, void* stream);

/* Gradient-range monitor of the native backward.  Gradient maps are f16 behind ONE power-of-two loss scale chosen at the head
 * (bnerv_head_bwd) and every conversion saturates, so a gradient that outgrows 65504 / S deeper in the cascade would be clipped
 * silently.  bnerv_resblock_mid_bwd / bnerv_block_front_bwd - which every block's gradients pass through - OR into
 * status[0]: bit 0 when they read or produce |value| >= 65504 (saturated), bit 1 for a non-finite value.  They also keep
 * status[1] = the largest |scaled gradient| of the step (float bits), which bnerv_head_bwd feeds back into the NEXT step's loss
 * scale: status[2] (float bits, 0 = default 8) is the target for S * max|dL/dz_head|, lowered when the deepest maps exceed 2^14
 * and raised again (up to 8) when they fall below 2^10 - the gradients of the low-resolution stages grow by orders of magnitude
 * relative to the head's while a model trains.  `status` = FOUR device ints the caller zeroes once and reads when it chooses (no
 * synchronisation here); NULL switches monitor and controller off (fixed target 8).  Process-wide; kernels captured in a CUDA
 * graph keep the pointer they were launched with. */
int bnerv_bwd_set_status(int* status);

/* ResBlock_SFT middle transposed (model_blocks.py:86-87; forward v = gelu(c0), w = v*g1p + beta1):
 *   dc0 = dw * g1p * dact;  dG[b][c] += sum dw*v;  dB[b][c] += sum dw;  dbias0[c] += sum dc0
 *   dw, v, dact (= gelu'(c0) from bnerv_conv_fused_ex), dc0 : C8 f16 [B][Cp/8][H][W][8]; g1p, dG, dB : f32 [B][Cp];
 *   dbias0 : f32 [Cp].  The three reductions accumulate (caller zeroes). */
int bnerv_resblock_mid_bwd(const void* dw, const void* v, const void* dact, const float* g1p, int B, int C, int H,
                           int W, void* dc0, float* dG, float* dB, float* dbias0, void* stream);

/* NeRVBlock front transposed (model_blocks.py:37,85,89; forward x0 = act(y), u = x0*g0p + beta0, out = x0 + conv1(..)):
 *   dy = (dout + du*g0p) * dact;  dG[b][c] += sum du*x0;  dB[b][c] += sum du;  dbias1[c] += sum dout
 *   (dout is also dL/d conv1-output, hence conv1's bias gradient).  dy is at the block's output resolution;
 *   bnerv_unshuffle_c8 turns it into the up-conv's gradient map when s > 1.  dbias_up (f32 [Cp], may be NULL):
 *   += sum dy, the up-conv's bias gradient when it has no PixelShuffle (s = 1). */
int bnerv_block_front_bwd(const void* du, const void* dout, const void* x0, const void* dact, const float* g0p,
                          int B, int C, int H, int W, void* dy, float* dG, float* dB, float* dbias1, float* dbias_up,
                          void* stream);

/* PixelShuffle(s) transposed on C8 maps: src [B][Cp/8][H*s][W*s][8] -> dst [B][s*s*Cp/8][H][W][8] (un-shuffled order).
 * sums (f32 [s*s*Cp], may be NULL; s = 2, 3 only): += channel sums of dst = the up-conv's un-shuffled bias gradient. */
int bnerv_unshuffle_c8(const void* src, int B, int C, int H, int W, int s, void* dst, float* sums, void* stream);

/* Per-frame error metrics on the device (hnerv_utils.py:338-341 'L2'/'L1' pixel losses and
 * psnr_fn_single :400-403) without the reference's per-step .cpu() synchronisation:
 *   out[b] = { mean((img-gt)^2), mean(|img-gt|), -10*log10(mse + 1e-9) }   (3 floats per frame)
 * img, gt : f32 [B][n_per_frame]; scratch : bnerv_frame_metrics_scratch_doubles(B) doubles.  Deterministic
 * (fixed-order two-stage reduction with f64 partial sums). */
int bnerv_frame_metrics(const float* img, const float* gt, int B, size_t n_per_frame, double* scratch, float* out,
                        void* stream);
size_t bnerv_frame_metrics_scratch_doubles(int B);

/* SSIM statistics of one pyramid level and their gradient: the device side of the `ssim` /
 * `ms_ssim` terms of hnerv_utils.loss_fn (:338-395), which the reference takes from pytorch_msssim==0.2.1 (not vendored;
 * algorithm restated in oracle/msssim_oracle.py, parity unpinned).  11-tap Gaussian window (sigma 1.5), 'valid' extent.
 *   x, y  : f32 [planes][H][W] (planes = batch * channels), H, W > 10
 *   stats : f64 [planes][2], caller-zeroed; += {sum ssim_map, sum cs_map} over the (H-10) x (W-10) valid pixels
 *   gw    : f32 [planes][2] = dL/d{mean ssim, mean cs} / ((H-10)*(W-10));  scratch: bnerv_ssim_scratch_floats floats
 *   dx    : f32 [planes][H][W] (=|+=) dL/dx  (y is treated as a constant, hnerv_utils.py:336) */
int bnerv_ssim_stats(const float* x, const float* y, int planes, int H, int W, float C1, float C2, double* stats,
                     void* stream);
int bnerv_ssim_grad(const float* x, const float* y, int planes, int H, int W, float C1, float C2, const float* gw,
                    float* scratch, int accumulate, float* dx, void* stream);
size_t bnerv_ssim_scratch_floats(int planes, int H, int W);

/* ------------------------------------------------------------------------------------------------
 * Post-training quantisation + Huffman bit accounting (the on-disk format either side of the
 * decoder).  Replaces quant_tensor (hnerv_utils.py:101-134) as called for every decoder tensor by quant_model
 * (train_nerv_all.py:620-641) and for the frame embeddings (:542), and the statistics half of the Huffman stage
 * (train_nerv_all.py:581-607: .tolist() of every code, np.unique, dahuffman's HuffmanCodec.from_data).
 *
 * quant_tensor tries one (min, scale) pair for the whole tensor (kept in f32) and, for every axis whose reduced table is
 * < 2 % of the tensor (extent > 50), one pair per slice along that axis (stored as f16), and keeps the candidate with the
 * smallest mean |t - dequant| (the first one on ties).  bnerv_ptq_plan_tensor (host only, no CUDA call) lists the
 * candidates of a shape and where their tables go; bnerv_ptq_quant_tensor runs them all and selects on the device:
 *   t       : f32, contiguous, `shape[0..ndim)`, ndim <= 4, fewer than 2^31 elements, finite values
 *   quant   : u8 [numel]  codes of the best candidate        new_t : f32 [numel] its reconstruction, or NULL
 *   tables  : f32 [plan.table_floats]; candidate c owns [table_offset[c], +groups[c]) = min, the next groups[c] = scale
 *             (per-axis values already rounded to f16, held in f32); group index = flat index of the keepdim table
 *   err     : f64 [BNERV_PTQ_MAX_CAND] mean |t - new_t| per candidate;  best : i32 [1] winning candidate index
 *   scratch : f64 [plan.scratch_doubles]
 * Every arithmetic step is the reference's in round-to-nearest f32 (true division as torch's CPU kernels do; torch's CUDA
 * kernel multiplies by 1/(2^bits-1) when dividing by a scalar, which can move a whole-tensor scale by 1 ulp), so codes,
 * tables and reconstruction are bit-identical to the CPU reference; the candidate errors are summed in f64 in a fixed
 * order and compared after rounding to f32 (the reference: f32 pairwise means), which can only matter for near-ties between
 * candidates.  Preconditions - NOT checked, and where behaviour deliberately differs from the reference: inputs must be finite
 * (CUDA fminf / fmaxf drop NaN where torch propagates it), and a group whose range is zero - a constant tensor, or a per-axis
 * slice whose f16 scale rounds to 0 - gets code 0 and new_t = min here, where the reference divides by zero and stores NaN. */
#define BNERV_PTQ_MAX_CAND 5
typedef struct bnerv_ptq_plan {
    int32_t n_cand;                           /* 1 + number of eligible axes                                  */
    int32_t axis[BNERV_PTQ_MAX_CAND];         /* -1: whole tensor; else the axis min/max reduce over          */
    int64_t groups[BNERV_PTQ_MAX_CAND];       /* entries of the candidate's min table (= of its scale table)  */
    int64_t table_offset[BNERV_PTQ_MAX_CAND]; /* float offset of [min table | scale table] inside `tables`    */
    int64_t table_floats;                     /* floats the caller provides for `tables`                      */
    int64_t scratch_doubles;                  /* doubles the caller provides for `scratch`                    */
} bnerv_ptq_plan;
int bnerv_ptq_plan_tensor(const int64_t* shape, int ndim, bnerv_ptq_plan* plan);
int bnerv_ptq_quant_tensor(const float* t, const int64_t* shape, int ndim, int bits, uint8_t* quant, float* new_t,
                           float* tables, double* err, int32_t* best, double* scratch, void* stream);
/* Multi-tensor form: ALL tensors of a model in five launches (min/max, tables, candidate errors, selection, codes) over a
 * descriptor table instead of ~13 launches per tensor - the loop of train_nerv_all.py:630-636 as one call.  Per tensor the
 * arithmetic, the partition of the error sums and therefore every output bit are those of bnerv_ptq_quant_tensor (which is this
 * call with one job).  Pointers are device buffers laid out as described above; `tables_f16` (optional) receives the same
 * tables as f16 values - exact for the per-axis candidates, whose entries are f16 values already - so the caller can hand out
 * the stored form without a conversion launch per tensor.  `scratch`: 256-byte aligned device buffer of
 * bnerv_ptq_quant_tensors_scratch_bytes(jobs, n_jobs) bytes (host-only helper, 0 on a bad job list). */
typedef struct bnerv_ptq_job {
    const float* t;
    int64_t shape[BNERV_PTQ_MAX_CAND - 1];
    int32_t ndim;
    int32_t reserved;
    uint8_t* quant;
    float* new_t;          /* may be NULL */
    float* tables;         /* f32 [plan.table_floats] */
    void* tables_f16;      /* f16 [plan.table_floats] or NULL */
    double* err;           /* f64 [BNERV_PTQ_MAX_CAND] */
    int32_t* best;         /* i32 [1] */
} bnerv_ptq_job;
size_t bnerv_ptq_quant_tensors_scratch_bytes(const bnerv_ptq_job* jobs, int n_jobs);
int bnerv_ptq_quant_tensors(const bnerv_ptq_job* jobs, int n_jobs, int bits, void* scratch, size_t scratch_bytes, void* stream);
/* Decode side of the same format: out = min + scale * quant in f32 from the stored u8 codes and the winning candidate's
 * tables (axis = -1: one f32 pair, tables_f16 = 0; axis >= 0: f16 keepdim tables over that axis, tables_f16 = 1) -
 * bit-identical to the `new_t` bnerv_ptq_quant_tensor produced, i.e. to what quant_model loads into the quantised model
 * (train_nerv_all.py:634-638).  (hnerv_utils.dequant_tensor, :185-188, is unused by the reference and evaluates the same
 * expression in f16 when the tables are f16.) */
int bnerv_ptq_dequant_tensor(const uint8_t* quant, const int64_t* shape, int ndim, int axis, const void* tmin,
                             const void* scale, int tables_f16, float* out, void* stream);
/* counts256[v] += number of codes equal to v (u64 [256], caller-zeroed before the first tensor of a model). */
int bnerv_histogram_u8(const uint8_t* codes, size_t n, uint64_t* counts256, void* stream);
/* HOST function (no CUDA call): Huffman code length in bits of every symbol with a non-zero count, 0 for the others,
 * as dahuffman==0.4.1's HuffmanCodec.from_data assigns them - an EOF leaf of frequency 1 joins the alphabet and equal
 * frequencies are ordered by (first leaf's symbol), EOF first.  The package is not vendored by the reference: its
 * published algorithm is restated (parity unpinned for the tie-breaking: every Huffman tree has the same total cost
 * INCLUDING the EOF leaf, but which equally frequent symbols end up beside that leaf decides a few bits of the sum over
 * the real symbols the reference reports). */
int bnerv_huffman_code_lengths(const uint64_t* counts, int n_symbols, int32_t* lengths);

/* ------------------------------------------------------------------------------------------------
 * ConvNeXt encoder of HNeRV_Boost, forward only (encoder half): one stage of
 * ConvNeXt.forward (model_blocks.py:314-320) per call -
 *     [LayerNorm channels_first] -> Conv2d(Cin, Cout, kernel s, stride s) -> [LayerNorm channels_first (stage 0)]
 *     -> n_blocks x Block.forward (model_blocks.py:246-260): x + gamma * pwconv2(GELU(pwconv1(LayerNorm(dwconv7x7(x)))))
 * as HNeRV_Boost.forward_encoder (model_hnerv.py:230-234) runs it per frame in evaluate() and when a sequence's
 * embeddings are extracted.  f32 CUDA-core arithmetic with exact erf (the embedding feeds the whole decoder; gate 1e-5
 * against the reference's f32 result); all LayerNorms use eps = 1e-6 like the reference.  Training keeps the torch module.
 *   weights : the module's own f32 parameter storage, no packing (down_w [Cout][Cin][s][s], dw_w [C][1][7][7],
 *             pw1_w [4C][C], pw2_w [C][4C], gamma [C]; ln_in_* / ln_out_* NULL where the stage has no such LayerNorm)
 *   x       : the frame, NCHW f32 (x_is_nchw = 1, stage 0) or the previous stage's channels-last output [B][Hin][Win][Cin]
 *   y_nhwc  : [B][Hin/s][Win/s][Cout] f32 (channels-last; bnerv_nhwc_to_nchw converts the last stage's output)
 *   work    : f32 [bnerv_convnext_stage_work_floats(B, Hin, Win, s, Cout)] */
typedef struct bnerv_convnext_block {
    const float *dw_w, *dw_b, *ln_w, *ln_b, *pw1_w, *pw1_b, *pw2_w, *pw2_b, *gamma;
} bnerv_convnext_block;
typedef struct bnerv_convnext_stage {
    const float *ln_in_w, *ln_in_b, *down_w, *down_b, *ln_out_w, *ln_out_b;
    const bnerv_convnext_block* blocks;       /* HOST array of n_blocks entries (device pointers inside) */
    int32_t n_blocks, Cin, Cout, s;
} bnerv_convnext_stage;
int    bnerv_convnext_stage_fwd(const bnerv_convnext_stage* stage, const float* x, int x_is_nchw, int B, int Hin, int Win,
                                float* y_nhwc, float* work, void* stream);
size_t bnerv_convnext_stage_work_floats(int B, int Hin, int Win, int s, int Cout);
int    bnerv_nhwc_to_nchw(const float* x_nhwc, int B, int H, int W, int C, float* y_nchw, void* stream);

/* Sizes (in elements) of the buffers the caller must provide. */
size_t bnerv_c8_numel(int B, int C, int H, int W);                 /* __half elements            */
size_t bnerv_packed_weight_numel(int Cout, int Cin, int k, int s); /* __half elements            */
size_t bnerv_packed_bias_numel(int Cout, int s);                   /* float elements (= Np)      */

#ifdef __cplusplus
}
#endif
#endif /* BNERV_B200_H_ */
