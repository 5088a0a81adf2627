// Post-training quantisation of the decoder's tensors and the symbol statistics of the Huffman stage
// (the on-disk format either side of the decode path).
//
//   bnerv_ptq_quant_tensor   quant_tensor, hnerv_utils.py:101-134: min/max tables for the whole tensor and for every
//                            axis longer than 50, round-to-nearest codes, reconstruction, mean |error| per candidate,
//                            best candidate chosen ON THE DEVICE (no host round trip between the passes)
//   bnerv_histogram_u8       np.unique(quant_v_list, return_counts=True), train_nerv_all.py:592-593, without the
//                            .tolist() of every code
//   bnerv_huffman_code_lengths  HuffmanCodec.from_data(...).get_code_table() bit sizes, train_nerv_all.py:596-599
//                            (dahuffman 0.4.1's heap construction incl. its EOF leaf) - host code, <= 4096 symbols
//
// Arithmetic is the reference's, operation by operation, in round-to-nearest f32 intrinsics that nvcc never contracts
// into FMAs, so codes, tables and reconstructions are bit-identical to the torch CPU result.  All kernels are one-pass
// HBM-bound sweeps; the whole job of a 15 M parameter model is a few hundred launches of a few microseconds.
#include <algorithm>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace bnerv {

static constexpr int PTQ_ERR_BLOCKS = 592;            // 148 SMs x 4 resident blocks of 256 threads
static constexpr long long PTQ_TARGET_THREADS = 148LL * 2048;
static constexpr int PTQ_BLOCK_MODE_BELOW = 1024;     // fewer groups than this: one block per (group, split)

// A candidate's view of the tensor: element (o, a, i) lives at (o*A + a)*inner + i and belongs to group o*inner + i
// (axis candidates: A = t.shape[axis]; whole tensor: outer = inner = 1, A = numel).
struct PtqView {
    unsigned outer, A, inner;
};

struct PtqCandSet {
    int n;
    PtqView view[BNERV_PTQ_MAX_CAND];
    long long table_off[BNERV_PTQ_MAX_CAND];
    long long groups[BNERV_PTQ_MAX_CAND];
};

static int splits_for(long long G, long long A) {
    if (G >= PTQ_BLOCK_MODE_BELOW) return static_cast<int>(std::min<long long>(A, (PTQ_TARGET_THREADS + G - 1) / G));
    const long long want = std::max<long long>(1, (PTQ_ERR_BLOCKS + G - 1) / G);
    return static_cast<int>(std::max<long long>(1, std::min<long long>((A + 2047) / 2048, want)));
}

__device__ __forceinline__ float round_f16(float x) { return __half2float(__float2half_rn(x)); }

// ---- multi-tensor work tables ----------------------------------------------------------------------------------
// One PtqJobDev per tensor, one PtqSeg per (tensor, candidate).  Every pass is ONE launch over all segments: a block finds its
// segment by binary search over the segments' first-block prefix, then does exactly what the single-tensor grid did (same
// splits of the reduced axis, same 592-block strided error sums), so results do not depend on how many tensors share a launch.
struct PtqJobDev {
    const float* t;
    uint8_t* quant;
    float* new_t;
    float* tables;
    __half* tables_f16;
    double* err;
    int* best;
    unsigned n;
    int n_cand;
    int err_blocks;
    int first_seg;
    unsigned apply_first;       // first block of this job in the apply pass
    PtqView view[BNERV_PTQ_MAX_CAND];
    long long table_off[BNERV_PTQ_MAX_CAND];
};

struct PtqSeg {
    int job, cand;
    int S;                      // splits of the reduced axis in the min/max pass
    int block_mode;             // 1: one block per (group, split); 0: one thread per group
    unsigned G;
    unsigned mm_first, table_first, err_first;     // first block of this segment in the three passes
    long long mm_off;           // float2 offset of its [G][S] min/max partials
    long long errp_off;         // double offset of its err_blocks partial sums
};

struct PtqTables {
    const PtqJobDev* jobs;
    const PtqSeg* segs;
    int n_jobs, n_segs;
    float2* mm;
    double* errp;
    float levels;
};

template <unsigned PtqSeg::*FIRST>
__device__ __forceinline__ int ptq_find_seg(const PtqTables& T, unsigned blk) {
    int lo = 0, hi = T.n_segs - 1;
    while (lo < hi) {           // last segment whose first block is <= blk
        const int mid = (lo + hi + 1) >> 1;
        if (T.segs[mid].*FIRST <= blk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- pass 1: min / max over the reduced axis ------------------------------------------------------------------
__global__ void __launch_bounds__(256) ptq_minmax_kernel(PtqTables T) {
    const PtqSeg sg = T.segs[ptq_find_seg<&PtqSeg::mm_first>(T, blockIdx.x)];
    const PtqJobDev& jb = T.jobs[sg.job];
    const PtqView v = jb.view[sg.cand];
    const unsigned local = blockIdx.x - sg.mm_first;
    float2* partial = T.mm + sg.mm_off;
    const int S = sg.S;
    if (!sg.block_mode) {
        const unsigned nbx = (sg.G + 255u) / 256u;
        const unsigned g = (local % nbx) * 256u + threadIdx.x, sp = local / nbx;
        if (g >= sg.G) return;
        const unsigned a0 = static_cast<unsigned>(static_cast<unsigned long long>(v.A) * sp / S);
        const unsigned a1 = static_cast<unsigned>(static_cast<unsigned long long>(v.A) * (sp + 1) / S);
        const unsigned o = g / v.inner, i = g - o * v.inner;
        const float* p = jb.t + (static_cast<size_t>(o) * v.A) * v.inner + i;
        float mn = INFINITY, mx = -INFINITY;
        for (unsigned a = a0; a < a1; ++a) {
            const float x = __ldg(p + static_cast<size_t>(a) * v.inner);
            mn = fminf(mn, x);
            mx = fmaxf(mx, x);
        }
        partial[static_cast<size_t>(g) * S + sp] = make_float2(mn, mx);
        return;
    }
    const unsigned g = local % sg.G, sp = local / sg.G;
    const unsigned a0 = static_cast<unsigned>(static_cast<unsigned long long>(v.A) * sp / S);
    const unsigned a1 = static_cast<unsigned>(static_cast<unsigned long long>(v.A) * (sp + 1) / S);
    const unsigned o = g / v.inner, i = g - o * v.inner;
    const float* p = jb.t + (static_cast<size_t>(o) * v.A) * v.inner + i;
    float mn = INFINITY, mx = -INFINITY;
    for (unsigned a = a0 + threadIdx.x; a < a1; a += blockDim.x) {
        const float x = __ldg(p + static_cast<size_t>(a) * v.inner);
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    __shared__ float smn[8], smx[8];
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned w = 1; w < blockDim.x / 32; ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
        partial[static_cast<size_t>(g) * S + sp] = make_float2(mn, mx);
    }
}

// ---- pass 2: scale = (max - min) / (2^bits - 1) in f32 (hnerv_utils.py:106,111); per-axis tables are stored as f16 (:113)
__global__ void __launch_bounds__(256) ptq_table_kernel(PtqTables T) {
    const PtqSeg sg = T.segs[ptq_find_seg<&PtqSeg::table_first>(T, blockIdx.x)];
    const PtqJobDev& jb = T.jobs[sg.job];
    const unsigned g = (blockIdx.x - sg.table_first) * 256u + threadIdx.x;
    if (g >= sg.G) return;
    const float2* partial = T.mm + sg.mm_off;
    float mn = INFINITY, mx = -INFINITY;
    for (int s = 0; s < sg.S; ++s) {
        const float2 p = partial[static_cast<size_t>(g) * sg.S + s];
        mn = fminf(mn, p.x);
        mx = fmaxf(mx, p.y);
    }
    float scale = __fdiv_rn(__fsub_rn(mx, mn), T.levels);
    if (sg.cand > 0) { mn = round_f16(mn); scale = round_f16(scale); }
    float* table = jb.tables + jb.table_off[sg.cand];
    table[g] = mn;
    table[sg.G + g] = scale;
    if (jb.tables_f16) {
        __half* th = jb.tables_f16 + jb.table_off[sg.cand];
        th[g] = __float2half_rn(mn);
        th[sg.G + g] = __float2half_rn(scale);
    }
}

// quant = clamp(round((t - min) / scale), 0, levels); new_t = min + scale * quant   (hnerv_utils.py:120-121)
__device__ __forceinline__ float ptq_reconstruct(float x, float tmin, float scale, float levels, float& q) {
    q = fminf(fmaxf(rintf(__fdiv_rn(__fsub_rn(x, tmin), scale)), 0.0f), levels);
    return __fadd_rn(tmin, __fmul_rn(scale, q));
}

__device__ __forceinline__ unsigned ptq_group_of(unsigned idx, const PtqView& v) {
    if (v.inner == 1) return idx / v.A;                 // whole tensor (outer 1 -> 0) or last axis
    const unsigned span = v.A * v.inner;
    const unsigned o = idx / span, rem = idx - o * span;
    return o * v.inner + rem % v.inner;
}

// ---- pass 3: mean |t - new_t| of every candidate, stage 1: err_blocks blocks per segment, fixed per-thread order, one f64
// partial per block
__global__ void __launch_bounds__(256) ptq_error_kernel(PtqTables T) {
    const PtqSeg sg = T.segs[ptq_find_seg<&PtqSeg::err_first>(T, blockIdx.x)];
    const PtqJobDev& jb = T.jobs[sg.job];
    const PtqView v = jb.view[sg.cand];
    const unsigned lb = blockIdx.x - sg.err_first, nb = jb.err_blocks;
    const float* table = jb.tables + jb.table_off[sg.cand];
    const float* t = jb.t;
    double acc = 0.0;
    for (unsigned idx = lb * 256u + threadIdx.x; idx < jb.n; idx += nb * 256u) {
        const unsigned g = ptq_group_of(idx, v);
        const float x = __ldg(t + idx);
        float q;
        const float nt = ptq_reconstruct(x, __ldg(table + g), __ldg(table + sg.G + g), T.levels, q);
        acc += static_cast<double>(fabsf(__fsub_rn(x, nt)));
    }
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    __shared__ double sw[8];
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) acc += sw[w];
        T.errp[sg.errp_off + lb] = acc;
    }
}

// ---- pass 4: one block per tensor: sum the partials of each candidate, best = first candidate with the smallest error
// (min(err_t_list) + list.index, hnerv_utils.py:127-128)
__global__ void __launch_bounds__(256) ptq_select_kernel(PtqTables T) {
    const PtqJobDev& jb = T.jobs[blockIdx.x];
    __shared__ double s[256];
    __shared__ double errs[BNERV_PTQ_MAX_CAND];
    for (int c = 0; c < jb.n_cand; ++c) {
        const double* partial = T.errp + T.segs[jb.first_seg + c].errp_off;
        double acc = 0.0;
        for (int i = threadIdx.x; i < jb.err_blocks; i += 256) acc += partial[i];
        s[threadIdx.x] = acc;
        __syncthreads();
        for (int d = 128; d > 0; d >>= 1) {
            if (static_cast<int>(threadIdx.x) < d) s[threadIdx.x] += s[threadIdx.x + d];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            errs[c] = s[0] / static_cast<double>(jb.n);
            jb.err[c] = errs[c];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // the reference compares f32 means (torch .mean() of an f32 tensor): round before comparing, so candidates whose errors
        // agree to f32 precision tie exactly as they do there (first one wins)
        int best = 0;
        for (int c = 1; c < jb.n_cand; ++c)
            if (static_cast<float>(errs[c]) < static_cast<float>(errs[best])) best = c;
        *jb.best = best;
    }
}

// ---- pass 5: the codes and the reconstruction of the winning candidate
__global__ void __launch_bounds__(256) ptq_apply_kernel(PtqTables T) {
    int lo = 0, hi = T.n_jobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (T.jobs[mid].apply_first <= blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const PtqJobDev& jb = T.jobs[lo];
    const int best = *jb.best;
    const PtqView v = jb.view[best];
    const unsigned G = v.outer * v.inner;
    const float* table = jb.tables + jb.table_off[best];
    const unsigned lb = blockIdx.x - jb.apply_first, nb = jb.err_blocks;
    for (unsigned idx = lb * 256u + threadIdx.x; idx < jb.n; idx += nb * 256u) {
        const unsigned g = ptq_group_of(idx, v);
        float q;
        const float nt = ptq_reconstruct(__ldg(jb.t + idx), __ldg(table + g), __ldg(table + G + g), T.levels, q);
        jb.quant[idx] = static_cast<uint8_t>(q);
        if (jb.new_t) jb.new_t[idx] = nt;
    }
}

// new_t = min + scale * quant from the stored form (u8 codes + f32 scalar pair or f16 keepdim tables): the decode side of the
// format, bit-identical to the reconstruction quant_tensor returned (and quant_model loads, train_nerv_all.py:634-638)
template <typename T>
__global__ void __launch_bounds__(256) ptq_dequant_kernel(const uint8_t* __restrict__ quant, unsigned n, PtqView v,
                                                          const T* __restrict__ tmin, const T* __restrict__ scale,
                                                          float* __restrict__ out) {
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const unsigned g = ptq_group_of(idx, v);
        const float m = static_cast<float>(tmin[g]), sc = static_cast<float>(scale[g]);
        out[idx] = __fadd_rn(m, __fmul_rn(sc, static_cast<float>(quant[idx])));
    }
}

// ---- histogram of uint8 codes ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) histogram_u8_kernel(const uint8_t* __restrict__ codes, size_t n,
                                                           unsigned long long* __restrict__ counts) {
    __shared__ unsigned h[8][256];                      // one sub-histogram per warp: peaked code distributions
    for (int k = threadIdx.x; k < 8 * 256; k += 256) (&h[0][0])[k] = 0;
    __syncthreads();
    unsigned* mine = h[threadIdx.x >> 5];
    const size_t misalign = (4 - (reinterpret_cast<uintptr_t>(codes) & 3)) & 3;
    const size_t head = n < misalign ? n : misalign;
    const size_t words = (n - head) / 4;
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x, nthr = static_cast<size_t>(gridDim.x) * blockDim.x;
    const unsigned* body = reinterpret_cast<const unsigned*>(codes + head);
    for (size_t w = tid; w < words; w += nthr) {
        const unsigned x = __ldg(body + w);
        atomicAdd(mine + (x & 255u), 1u);
        atomicAdd(mine + ((x >> 8) & 255u), 1u);
        atomicAdd(mine + ((x >> 16) & 255u), 1u);
        atomicAdd(mine + (x >> 24), 1u);
    }
    if (tid < head) atomicAdd(mine + codes[tid], 1u);
    const size_t tail0 = head + words * 4;
    if (tid < n - tail0) atomicAdd(mine + codes[tail0 + tid], 1u);
    __syncthreads();
    unsigned total = 0;
    for (int w = 0; w < 8; ++w) total += h[w][threadIdx.x];
    if (total) atomicAdd(counts + threadIdx.x, static_cast<unsigned long long>(total));
}

static int ptq_make_plan(const int64_t* shape, int ndim, bnerv_ptq_plan* plan, PtqCandSet* cs, long long* numel) {
    if (!shape) return set_error(BNERV_E_BADARG, "ptq: null shape");
    if (ndim < 0 || ndim > BNERV_PTQ_MAX_CAND - 1) return set_error(BNERV_E_UNSUPPORTED, "ptq: %d dimensions (at most %d)", ndim, BNERV_PTQ_MAX_CAND - 1);
    long long n = 1;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] <= 0) return set_error(BNERV_E_BADARG, "ptq: non-positive extent");
        n *= shape[d];
        if (n >= (1LL << 31)) return set_error(BNERV_E_UNSUPPORTED, "ptq: tensors of 2^31 elements or more");
    }
    bnerv_ptq_plan p{};
    PtqCandSet c{};
    p.axis[0] = -1; p.groups[0] = 1; p.table_offset[0] = 0;
    c.view[0] = PtqView{1u, static_cast<unsigned>(n), 1u};
    int nc = 1;
    long long off = 2, gmax = 1;
    for (int d = 0; d < ndim; ++d) {
        const long long G = n / shape[d];
        // t_min.nelement() / t.nelement() < 0.02 (hnerv_utils.py:110), evaluated like Python: a double division
        if (!(static_cast<double>(G) / static_cast<double>(n) < 0.02)) continue;
        long long inner = 1;
        for (int e = d + 1; e < ndim; ++e) inner *= shape[e];
        p.axis[nc] = d; p.groups[nc] = G; p.table_offset[nc] = off;
        c.view[nc] = PtqView{static_cast<unsigned>(G / inner), static_cast<unsigned>(shape[d]), static_cast<unsigned>(inner)};
        off += 2 * G;
        gmax = std::max(gmax, G);
        ++nc;
    }
    p.n_cand = nc;
    p.table_floats = off;
    // one job's multi-tensor scratch: job + segment tables, [G][S] min/max pairs and 592 partial sums per candidate (+ alignment)
    p.scratch_doubles = 512 + BNERV_PTQ_MAX_CAND * (64 + PTQ_ERR_BLOCKS + gmax + PTQ_TARGET_THREADS + 32);
    c.n = nc;
    for (int k = 0; k < nc; ++k) { c.table_off[k] = p.table_offset[k]; c.groups[k] = p.groups[k]; }
    if (plan) *plan = p;
    if (cs) *cs = c;
    if (numel) *numel = n;
    return 0;
}

}  // namespace bnerv

using namespace bnerv;

extern "C" int bnerv_ptq_plan_tensor(const int64_t* shape, int ndim, bnerv_ptq_plan* plan) {
    if (!plan) return set_error(BNERV_E_BADARG, "ptq_plan_tensor: null plan");
    return ptq_make_plan(shape, ndim, plan, nullptr, nullptr);
}

namespace bnerv {

// splits of the reduced axis for the min/max pass: enough threads to fill the GPU when a tensor is alone, but never fewer
// than ~32 elements per thread (min/max are exact, so the split changes nothing but the scratch size)
static int splits_multi(long long G, long long A) {
    const long long s = splits_for(G, A);
    return static_cast<int>(std::max<long long>(1, std::min<long long>(s, A / 32)));
}

struct PtqHostLayout {
    std::vector<PtqJobDev> jobs;
    std::vector<PtqSeg> segs;
    unsigned mm_blocks = 0, table_blocks = 0, err_blocks = 0, apply_blocks = 0;
    long long mm_floats2 = 0, errp_doubles = 0;
    size_t off_jobs = 0, off_segs = 0, off_mm = 0, off_errp = 0, total = 0;
};

static int ptq_layout(const bnerv_ptq_job* jobs, int n_jobs, PtqHostLayout& L) {
    if (!jobs || n_jobs <= 0) return set_error(BNERV_E_BADARG, "ptq_quant_tensors: no jobs");
    L.jobs.resize(n_jobs);
    for (int j = 0; j < n_jobs; ++j) {
        const bnerv_ptq_job& in = jobs[j];
        bnerv_ptq_plan plan;
        PtqCandSet cs;
        long long n = 0;
        if (int rc = ptq_make_plan(in.shape, in.ndim, &plan, &cs, &n)) return rc;
        PtqJobDev& d = L.jobs[j];
        d.t = in.t; d.quant = in.quant; d.new_t = in.new_t; d.tables = in.tables;
        d.tables_f16 = static_cast<__half*>(in.tables_f16); d.err = in.err; d.best = in.best;
        d.n = static_cast<unsigned>(n);
        d.n_cand = cs.n;
        d.err_blocks = static_cast<int>(std::min<long long>(PTQ_ERR_BLOCKS, (n + 255) / 256));
        d.first_seg = static_cast<int>(L.segs.size());
        d.apply_first = L.apply_blocks;
        L.apply_blocks += d.err_blocks;
        for (int c = 0; c < cs.n; ++c) {
            d.view[c] = cs.view[c];
            d.table_off[c] = cs.table_off[c];
            PtqSeg sg{};
            sg.job = j; sg.cand = c;
            sg.G = static_cast<unsigned>(cs.groups[c]);
            sg.S = splits_multi(cs.groups[c], cs.view[c].A);
            sg.block_mode = cs.groups[c] < PTQ_BLOCK_MODE_BELOW ? 1 : 0;
            sg.mm_first = L.mm_blocks; sg.table_first = L.table_blocks; sg.err_first = L.err_blocks;
            sg.mm_off = L.mm_floats2; sg.errp_off = L.errp_doubles;
            const long long nb_mm = sg.block_mode ? 1LL * sg.G * sg.S : 1LL * ((sg.G + 255) / 256) * sg.S;
            if (L.mm_blocks + nb_mm > 0x7fffffffLL) return set_error(BNERV_E_UNSUPPORTED, "ptq_quant_tensors: too many blocks");
            L.mm_blocks += static_cast<unsigned>(nb_mm);
            L.table_blocks += (sg.G + 255) / 256;
            L.err_blocks += d.err_blocks;
            L.mm_floats2 += 1LL * sg.G * sg.S;
            L.errp_doubles += d.err_blocks;
            L.segs.push_back(sg);
        }
    }
    auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
    L.off_jobs = 0;
    L.off_segs = up(L.jobs.size() * sizeof(PtqJobDev));
    L.off_mm   = L.off_segs + up(L.segs.size() * sizeof(PtqSeg));
    L.off_errp = L.off_mm + up(static_cast<size_t>(L.mm_floats2) * sizeof(float2));
    L.total    = L.off_errp + up(static_cast<size_t>(L.errp_doubles) * sizeof(double));
    return 0;
}

static int ptq_run(const bnerv_ptq_job* jobs, int n_jobs, int bits, void* scratch, size_t scratch_bytes, void* stream) {
    if (bits < 1 || bits > 8) return set_error(BNERV_E_UNSUPPORTED, "ptq_quant_tensors: %d bits (codes are uint8: 1..8)", bits);
    if (!scratch) return set_error(BNERV_E_BADARG, "ptq_quant_tensors: null scratch");
    if (reinterpret_cast<uintptr_t>(scratch) & 255) return set_error(BNERV_E_BADARG, "ptq_quant_tensors: scratch must be 256-byte aligned");
    for (int j = 0; j < n_jobs; ++j)
        if (!jobs[j].t || !jobs[j].quant || !jobs[j].tables || !jobs[j].err || !jobs[j].best)
            return set_error(BNERV_E_BADARG, "ptq_quant_tensors: null pointer in job %d", j);
    PtqHostLayout L;
    if (int rc = ptq_layout(jobs, n_jobs, L)) return rc;
    if (scratch_bytes < L.total) return set_error(BNERV_E_BADARG, "ptq_quant_tensors: scratch of %zu bytes, %zu needed", scratch_bytes, L.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* base = static_cast<uint8_t*>(scratch);
    // one upload of both tables (pageable source: the call returns once it has been staged)
    std::vector<uint8_t> stage(L.off_mm, 0);
    memcpy(stage.data() + L.off_jobs, L.jobs.data(), L.jobs.size() * sizeof(PtqJobDev));
    memcpy(stage.data() + L.off_segs, L.segs.data(), L.segs.size() * sizeof(PtqSeg));
    cudaError_t e = cudaMemcpyAsync(base, stage.data(), stage.size(), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return set_error(static_cast<int>(e), "ptq_quant_tensors: table upload: %s", cudaGetErrorString(e));
    PtqTables T;
    T.jobs = reinterpret_cast<const PtqJobDev*>(base + L.off_jobs);
    T.segs = reinterpret_cast<const PtqSeg*>(base + L.off_segs);
    T.n_jobs = n_jobs; T.n_segs = static_cast<int>(L.segs.size());
    T.mm = reinterpret_cast<float2*>(base + L.off_mm);
    T.errp = reinterpret_cast<double*>(base + L.off_errp);
    T.levels = static_cast<float>((1 << bits) - 1);
    ptq_minmax_kernel<<<L.mm_blocks, 256, 0, st>>>(T);
    if (int rc = check_launch("ptq_minmax_kernel")) return rc;
    ptq_table_kernel<<<L.table_blocks, 256, 0, st>>>(T);
    if (int rc = check_launch("ptq_table_kernel")) return rc;
    ptq_error_kernel<<<L.err_blocks, 256, 0, st>>>(T);
    if (int rc = check_launch("ptq_error_kernel")) return rc;
    ptq_select_kernel<<<n_jobs, 256, 0, st>>>(T);
    if (int rc = check_launch("ptq_select_kernel")) return rc;
    ptq_apply_kernel<<<L.apply_blocks, 256, 0, st>>>(T);
    return check_launch("ptq_apply_kernel");
}

}  // namespace bnerv

extern "C" size_t bnerv_ptq_quant_tensors_scratch_bytes(const bnerv_ptq_job* jobs, int n_jobs) {
    PtqHostLayout L;
    if (ptq_layout(jobs, n_jobs, L)) return 0;
    return L.total;
}

extern "C" int bnerv_ptq_quant_tensors(const bnerv_ptq_job* jobs, int n_jobs, int bits, void* scratch, size_t scratch_bytes,
                                       void* stream) {
    if (!jobs || n_jobs <= 0) return set_error(BNERV_E_BADARG, "ptq_quant_tensors: no jobs");
    return ptq_run(jobs, n_jobs, bits, scratch, scratch_bytes, stream);
}

extern "C" int bnerv_ptq_quant_tensor(const float* t, const int64_t* shape, int ndim, int bits, uint8_t* quant, float* new_t,
                                      float* tables, double* err, int32_t* best, double* scratch, void* stream) {
    if (!t || !quant || !tables || !err || !best || !scratch) return set_error(BNERV_E_BADARG, "ptq_quant_tensor: null pointer");
    if (!shape || ndim < 0 || ndim > BNERV_PTQ_MAX_CAND - 1) return set_error(BNERV_E_UNSUPPORTED, "ptq: %d dimensions (at most %d)", ndim, BNERV_PTQ_MAX_CAND - 1);
    bnerv_ptq_job job{};
    job.t = t; job.ndim = ndim;
    for (int d = 0; d < ndim; ++d) job.shape[d] = shape[d];
    job.quant = quant; job.new_t = new_t; job.tables = tables; job.tables_f16 = nullptr; job.err = err; job.best = best;
    bnerv_ptq_plan plan;
    if (int rc = ptq_make_plan(shape, ndim, &plan, nullptr, nullptr)) return rc;
    // the single-tensor entry point: `scratch` (plan.scratch_doubles doubles) is the multi-tensor scratch of one job
    return ptq_run(&job, 1, bits, scratch, static_cast<size_t>(plan.scratch_doubles) * sizeof(double), stream);
}

extern "C" int bnerv_ptq_dequant_tensor(const uint8_t* quant, const int64_t* shape, int ndim, int axis, const void* tmin,
                                        const void* scale, int tables_f16, float* out, void* stream) {
    if (!quant || !tmin || !scale || !out) return set_error(BNERV_E_BADARG, "ptq_dequant_tensor: null pointer");
    bnerv_ptq_plan plan;
    long long n = 0;
    if (int rc = ptq_make_plan(shape, ndim, &plan, nullptr, &n)) return rc;
    if (axis < -1 || axis >= ndim) return set_error(BNERV_E_BADARG, "ptq_dequant_tensor: axis %d of %d dimensions", axis, ndim);
    PtqView v{1u, static_cast<unsigned>(n), 1u};
    if (axis >= 0) {
        long long inner = 1;
        for (int e = axis + 1; e < ndim; ++e) inner *= shape[e];
        v = PtqView{static_cast<unsigned>(n / shape[axis] / inner), static_cast<unsigned>(shape[axis]), static_cast<unsigned>(inner)};
    }
    const int blocks = static_cast<int>(std::min<long long>(PTQ_ERR_BLOCKS, (n + 255) / 256));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (tables_f16)
        ptq_dequant_kernel<__half><<<blocks, 256, 0, st>>>(quant, static_cast<unsigned>(n), v, static_cast<const __half*>(tmin),
                                                          static_cast<const __half*>(scale), out);
    else
        ptq_dequant_kernel<float><<<blocks, 256, 0, st>>>(quant, static_cast<unsigned>(n), v, static_cast<const float*>(tmin),
                                                         static_cast<const float*>(scale), out);
    return check_launch("ptq_dequant_kernel");
}

extern "C" int bnerv_histogram_u8(const uint8_t* codes, size_t n, uint64_t* counts256, void* stream) {
    if (!codes || !counts256) return set_error(BNERV_E_BADARG, "histogram_u8: null pointer");
    if (n == 0) return 0;
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>(PTQ_ERR_BLOCKS, (n / 4 + 255) / 256 + 1));
    histogram_u8_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(codes, n, reinterpret_cast<unsigned long long*>(counts256));
    return check_launch("histogram_u8_kernel");
}

// dahuffman 0.4.1, HuffmanCodec.from_frequencies: heap of (frequency, [leaves...]) tuples plus one EOF leaf of frequency
// 1; Python orders equal frequencies by the leaf lists, whose first elements always differ (the leaf sets of two live
// nodes are disjoint), so the order key is (frequency, first leaf's symbol) with EOF below every symbol; a merged node
// starts with the leaves of the smaller operand.  A strict total order fixes the pop sequence whatever the heap shape.
extern "C" int bnerv_huffman_code_lengths(const uint64_t* counts, int n_symbols, int32_t* lengths) {
    if (!counts || !lengths) return set_error(BNERV_E_BADARG, "huffman_code_lengths: null pointer");
    if (n_symbols <= 0 || n_symbols > 4096) return set_error(BNERV_E_UNSUPPORTED, "huffman_code_lengths: %d symbols (1..4096)", n_symbols);
    struct Node { uint64_t freq; int first; int left, right; };
    std::vector<Node> nodes;
    std::vector<int> live;
    nodes.push_back(Node{1, -1, -1, -1});               // EOF
    live.push_back(0);
    for (int s = 0; s < n_symbols; ++s) {
        lengths[s] = 0;
        if (counts[s]) { nodes.push_back(Node{counts[s], s, -1, -1}); live.push_back(static_cast<int>(nodes.size()) - 1); }
    }
    if (live.size() < 2) return set_error(BNERV_E_BADARG, "huffman_code_lengths: every count is zero");
    auto less = [&](int a, int b) {
        return nodes[a].freq != nodes[b].freq ? nodes[a].freq < nodes[b].freq : nodes[a].first < nodes[b].first;
    };
    auto pop_min = [&]() {
        size_t k = 0;
        for (size_t i = 1; i < live.size(); ++i) if (less(live[i], live[k])) k = i;
        const int id = live[k];
        live[k] = live.back();
        live.pop_back();
        return id;
    };
    while (live.size() > 1) {
        const int a = pop_min(), b = pop_min();
        nodes.push_back(Node{nodes[a].freq + nodes[b].freq, nodes[a].first, a, b});
        live.push_back(static_cast<int>(nodes.size()) - 1);
    }
    std::vector<std::pair<int, int>> stack{{live[0], 0}};
    while (!stack.empty()) {
        const auto [id, depth] = stack.back();
        stack.pop_back();
        if (nodes[id].left < 0) { if (nodes[id].first >= 0) lengths[nodes[id].first] = depth; continue; }
        stack.push_back({nodes[id].left, depth + 1});
        stack.push_back({nodes[id].right, depth + 1});
    }
    return 0;
}
