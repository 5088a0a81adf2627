// Shared device helpers for the bnerv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/bnerv_b200.h"

namespace bnerv {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (definitions in capi.cu)
// ---------------------------------------------------------------------------------------------
int  set_error(int code, const char* fmt, ...);
int  check_launch(const char* what);      // cudaGetLastError -> return code (+ bumps launch counter)
void count_launch();

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Packed conv-output row n' -> (output channel c of the shuffled map, PixelShuffle sub-position i, j); the
// reference conv channel is c*s*s + i*s + j (model_blocks.py:204,217).  Every 8 consecutive rows share (i, j).
//   s == 2 : n' = ((i*Cout_p/8 + c/8)*2 + j)*8 + c%8  -- the two horizontal neighbours of one pixel's 8 channels are
//            adjacent rows of ONE 16-column accumulator group, so a thread stores 32 contiguous bytes (a full
//            sector) instead of two half sectors written by different n-tile passes (2x DRAM writes + fill reads).
//   else   : n' = (i*s + j)*Cout_p + c
__host__ __device__ __forceinline__ void packed_row_to_cij(int n, int s, int cout_p, int& c, int& i, int& j) {
    if (s == 2) {
        const int u = n >> 4, r = n & 15, g8 = cout_p >> 3;
        j = r >> 3;
        i = u / g8;
        c = (u - i * g8) * 8 + (r & 7);
    } else {
        const int sub = n / cout_p;
        c = n - sub * cout_p;
        i = sub / s;
        j = sub - i * s;
    }
}

// ---------------------------------------------------------------------------------------------
// epilogue math.  All in f32; accuracy targets are << the 1e-3 parity budget.
// ---------------------------------------------------------------------------------------------
// sin with explicit two-constant Cody-Waite reduction to [-pi, pi] followed by MUFU.SIN
// (abs error 2^-21.4 on that interval).  Pre-activations are O(1..1e3); k stays exact in f32.
__device__ __forceinline__ float fast_sin(float x) {
    const float inv2pi = 0.15915494309189535f;
    const float c_hi   = 6.2831854820251465f;        // float(2*pi)
    const float c_lo   = -1.7484555314695172e-07f;   // 2*pi - c_hi
    float k = rintf(x * inv2pi);
    float r = fmaf(-k, c_hi, x);
    r = fmaf(-k, c_lo, r);
    return __sinf(r);
}

// erf via Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7), branch-free: 2 MUFU + ~10 FMA.
__device__ __forceinline__ float fast_erf(float x) {
    float ax = fabsf(x);
    float t  = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
    float p  = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float e = __expf(-ax * ax);
    float r = fmaf(-p, e, 1.0f);
    return copysignf(r, x);
}

__device__ __forceinline__ float gelu_erf(float x) {       // nn.GELU() default (exact erf form)
    return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752f));
}

__device__ __forceinline__ float tanh01(float x) {         // OutImg 'tanh': tanh(x)*0.5+0.5 == sigmoid(2x)
    float xc = fminf(fmaxf(x, -15.0f), 15.0f);
    float e  = __expf(-2.0f * xc);
    return __fdividef(1.0f, 1.0f + e);
}

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case BNERV_ACT_SIN:    return fast_sin(x);
        case BNERV_ACT_GELU:   return gelu_erf(x);
        case BNERV_ACT_RELU:   return fmaxf(x, 0.0f);
        case BNERV_ACT_TANH01: return tanh01(x);
        default:               return x;
    }
}

// precise variants for the f32 cross-check kernel (libdevice sinf/erff/tanhf)
__device__ __forceinline__ float apply_act_precise(float x, int act) {
    switch (act) {
        case BNERV_ACT_SIN:    return sinf(x);
        case BNERV_ACT_GELU:   return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
        case BNERV_ACT_RELU:   return fmaxf(x, 0.0f);
        case BNERV_ACT_TANH01: return tanhf(x) * 0.5f + 0.5f;
        default:               return x;
    }
}

// saturating f32x2 -> f16x2 (never produces inf from a finite f32)
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
    lo = fminf(fmaxf(lo, -65504.0f), 65504.0f);
    hi = fminf(fmaxf(hi, -65504.0f), 65504.0f);
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    return __half22float2(h);
}

// ---------------------------------------------------------------------------------------------
// packed f32x2 epilogue math (sm_100 FFMA2/FADD2/FMUL2): two channels per instruction on the FMA pipe.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

__device__ __forceinline__ float mufu_sin(float x) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ float2 sin2(float2 x) {
    // k = round(x / 2pi) by the 1.5*2^23 trick (|x| < 2^22 * 2pi), r = x - k*2pi in two pieces, MUFU.SIN
    const float magic = 12582912.0f;
    float2 t = fma2(x, f2(0.15915494309189535f), f2(magic));
    float2 k = add2(t, f2(-magic));
    float2 r = fma2(k, f2(-6.2831854820251465f), x);
    r = fma2(k, f2(1.7484555314695172e-07f), r);
    return make_float2(mufu_sin(r.x), mufu_sin(r.y));
}

__device__ __forceinline__ float2 gelu2(float2 x) {
    // 0.5 x (1 + erf(x/sqrt2)), erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7); coefficients negated so
    // that erf(|z|) = 1 + pn(t) * exp(-z^2) comes out of one FFMA2.
    float2 z  = mul2(x, f2(0.70710678118654752f));
    float2 az = make_float2(fabsf(z.x), fabsf(z.y));
    float2 d  = fma2(az, f2(0.3275911f), f2(1.0f));
    float2 t  = make_float2(mufu_rcp(d.x), mufu_rcp(d.y));
    float2 pn = fma2(t, f2(-1.061405429f), f2(1.453152027f));
    pn = fma2(pn, t, f2(-1.421413741f));
    pn = fma2(pn, t, f2(0.284496736f));
    pn = fma2(pn, t, f2(-0.254829592f));
    pn = mul2(pn, t);
    float2 q = mul2(mul2(z, z), f2(-1.4426950408889634f));
    float2 e = make_float2(mufu_ex2(q.x), mufu_ex2(q.y));
    float2 r = fma2(pn, e, f2(1.0f));                              // erf(|z|)
    r = make_float2(copysignf(r.x, z.x), copysignf(r.y, z.y));
    float2 h = mul2(x, f2(0.5f));
    return fma2(h, r, h);
}

__device__ __forceinline__ float2 tanh01_2(float2 x) { return make_float2(tanh01(x.x), tanh01(x.y)); }

// Activation AND its derivative at the pre-activation (training forward: the derivative map is what the backward
// pass multiplies by, so neither the pre-activation nor a second transcendental pass is needed later).
__device__ __forceinline__ float mufu_cos(float x) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ float2 sin2_d(float2 x, float2& d) {           // d = cos(x)
    const float magic = 12582912.0f;
    float2 t = fma2(x, f2(0.15915494309189535f), f2(magic));
    float2 k = add2(t, f2(-magic));
    float2 r = fma2(k, f2(-6.2831854820251465f), x);
    r = fma2(k, f2(1.7484555314695172e-07f), r);
    d = make_float2(mufu_cos(r.x), mufu_cos(r.y));
    return make_float2(mufu_sin(r.x), mufu_sin(r.y));
}

__device__ __forceinline__ float2 gelu2_d(float2 x, float2& d) {          // d = Phi(x) + x * phi(x)
    float2 z  = mul2(x, f2(0.70710678118654752f));
    float2 az = make_float2(fabsf(z.x), fabsf(z.y));
    float2 dd = fma2(az, f2(0.3275911f), f2(1.0f));
    float2 t  = make_float2(mufu_rcp(dd.x), mufu_rcp(dd.y));
    float2 pn = fma2(t, f2(-1.061405429f), f2(1.453152027f));
    pn = fma2(pn, t, f2(-1.421413741f));
    pn = fma2(pn, t, f2(0.284496736f));
    pn = fma2(pn, t, f2(-0.254829592f));
    pn = mul2(pn, t);
    float2 q = mul2(mul2(z, z), f2(-1.4426950408889634f));
    float2 e = make_float2(mufu_ex2(q.x), mufu_ex2(q.y));                 // exp(-x^2/2)
    float2 r = fma2(pn, e, f2(1.0f));
    r = make_float2(copysignf(r.x, z.x), copysignf(r.y, z.y));            // erf(x/sqrt2)
    float2 h = mul2(x, f2(0.5f));
    float2 cdf = fma2(r, f2(0.5f), f2(0.5f));
    d = fma2(mul2(x, e), f2(0.3989422804014327f), cdf);
    return fma2(h, r, h);
}

// act: BNERV_ACT_* (run-time); returns act(x), d = act'(x)
__device__ __forceinline__ float2 act2_with_deriv(float2 x, int act, float2& d) {
    switch (act) {
        case BNERV_ACT_SIN:  return sin2_d(x, d);
        case BNERV_ACT_GELU: return gelu2_d(x, d);
        case BNERV_ACT_RELU: d = make_float2(x.x > 0.0f ? 1.0f : 0.0f, x.y > 0.0f ? 1.0f : 0.0f);
                             return make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f));
        case BNERV_ACT_TANH01: { float2 o = tanh01_2(x); d = make_float2(2.0f * o.x * (1.0f - o.x), 2.0f * o.y * (1.0f - o.y)); return o; }
        default: d = f2(1.0f); return x;
    }
}

template <int ACT>
__device__ __forceinline__ float2 act2(float2 x) {
    if (ACT == BNERV_ACT_SIN) return sin2(x);
    if (ACT == BNERV_ACT_GELU) return gelu2(x);
    if (ACT == BNERV_ACT_RELU) return make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f));
    if (ACT == BNERV_ACT_TANH01) return tanh01_2(x);
    return x;
}

// f32x2 -> f16x2 with saturation to +-65504 in one instruction (F2FP.SATFINITE.F16.F32.PACK_AB)
__device__ __forceinline__ uint32_t pack_h2_satfinite(float2 v) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v.y), "f"(v.x));
    return r;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// try_wait with a suspend-time hint: the thread is parked by the hardware until the phase completes or the
// hint (ns) expires, instead of polling every ~25 cycles and stealing issue slots from the working warps.
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as a
// hung GPU.  ~2 s at 2 GHz.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

// One lane of a converged warp (elect.sync): the issuing lane for TMA / tcgen05 instructions.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// named barrier over a subset of the CTA's warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], f16 operands, f32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 16 consecutive f32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 8 consecutive f32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major, no swizzle ("interleave" canonical layout):
//   core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes;
//   LBO = byte distance between core matrices adjacent in K, SBO = between 8-row groups in M/N.
// (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor; version field = 1 on sm_100.)
__device__ __host__ __forceinline__ uint64_t umma_desc_hi_noswz(uint32_t lbo, uint32_t sbo) {
    return (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16)
         | (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32)
         | (1ull << 46);
}
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | umma_desc_hi_noswz(lbo, sbo);
}
// instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n (cute UMMA::InstrDescriptor)
__device__ __host__ __forceinline__ uint32_t umma_idesc_f16_m128(int n) {
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

// Programmatic dependent launch: wait until the preceding kernel in the stream has completed and its writes are
// visible / allow the next kernel's CTAs to start their prologue as this kernel's CTAs retire.
__device__ __forceinline__ void pdl_wait()              { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x()    { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Remote arrive with the default (CTA-scope release) semantics: the consumer only needs the TMEM reads ordered,
// which tcgen05.fence::before_thread_sync provides; a .release.cluster arrive would make every epilogue warp wait
// for its outstanding global stores (MEMBAR.ALL.CTA + ERRBAR, 22 % of epilogue samples in profiles/r01_v4a).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair: data lands in the issuing CTA's shared memory, the transaction
// bytes are credited to `bar_cluster` (a shared::cluster address, normally the pair leader's barrier).
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const void* tmap, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {   // one whole warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 split over the pair (128 rows each, A from each CTA's own shared
// memory) and B's N rows split half/half over the two CTAs' shared memory.  Issued by the leader CTA only.
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all previously issued MMAs completed) on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __host__ __forceinline__ uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace bnerv
