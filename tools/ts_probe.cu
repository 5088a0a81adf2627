// Micro-probe (bring-up evidence for csrc/block_stream.cu): tcgen05.mma with the A operand in TENSOR MEMORY ("TS" form)
// against the shared-memory form ("SS") for the narrow-N shapes of the 12-channel NeRV stages.
//   1. correctness of the TS operand layout: A[m][k] (f16) lives in TMEM lane m, 32-bit column k/2, half k%2, written with
//      tcgen05.st.32x32b by the thread that owns lane m; D = A x B^T is compared with a CPU reference;
//   2. cycles per MMA (M = 128, K = 16) for N in {16, 48, 64}: SS reads 4 KB of A from shared memory per MMA, TS does not.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ts_probe tools/ts_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../boosting-nerv_b200/csrc/common.cuh"

namespace bnerv {
int set_error(int code, const char*, ...) { return code; }
int check_launch(const char*) { return 0; }
void count_launch() {}
}
using namespace bnerv;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// smem layouts: A (SS): K-major no-swizzle, [2 k-groups][128 rows][16 B]  (LBO = 2048, SBO = 128)
//               B     : K-major no-swizzle, [2 k-groups][N rows][16 B]    (LBO = N*16, SBO = 128)
// res[0..] : D (128 x N floats) of the TS product; cyc[0] = cycles of `reps` TS MMAs, cyc[1] = of `reps` SS MMAs
__global__ void __launch_bounds__(128, 1) ts_probe_kernel(const __half* A, const __half* B, int N, int reps, float* res, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = threadIdx.x;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 4096;
    // stage A and B in the canonical layouts
    for (int i = threadIdx.x; i < 128 * 2; i += 128) {
        const int row = i % 128, g = i / 128;
        *reinterpret_cast<uint4*>(sA + (g * 128 + row) * 16) = *reinterpret_cast<const uint4*>(A + row * 16 + g * 8);
    }
    for (int i = threadIdx.x; i < N * 2; i += 128) {
        const int row = i % N, g = i / N;
        *reinterpret_cast<uint4*>(sB + (g * N + row) * 16) = *reinterpret_cast<const uint4*>(B + row * 16 + g * 8);
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t lane_base = tb + (static_cast<uint32_t>(warp * 32) << 16);
    // A into TMEM columns [256, 264): lane m holds A[m][0..15] as 8 packed pairs
    {
        uint32_t v[8];
        const uint4 lo = *reinterpret_cast<const uint4*>(A + m * 16), hi = *reinterpret_cast<const uint4*>(A + m * 16 + 8);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        tmem_st8(lane_base + 256, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t bdesc = umma_desc_kmajor_noswz(smem_u32(sB), N * 16, 128);
    const uint64_t adesc = umma_desc_kmajor_noswz(smem_u32(sA), 2048, 128);
    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        umma_f16_ts(tb, tb + 256, bdesc, idesc, 0);                 // TS product -> columns [0, N)
        umma_f16(tb + 64, adesc, bdesc, idesc, 0);                  // SS product -> columns [64, 64 + N)  (N <= 64 here)
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase); phase ^= 1;
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8], w[8];
        tmem_ld8(lane_base + c0, v);
        tmem_ld8(lane_base + 64 + c0, w);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) {
            res[m * N + c0 + j] = __uint_as_float(v[j]);
            res[128 * N + m * N + c0 + j] = __uint_as_float(w[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // timing: `reps` accumulating MMAs of each form issued back to back by one thread, rotating over `chains` independent
    // accumulators (chains = 1: a dependent chain, i.e. the latency of one MMA; chains = 4: the pipe's throughput)
    for (int form = 0; form < 4; ++form) {
        const int chains = (form & 2) ? 4 : 1;
        long long t0 = 0;
        if (__shfl_sync(0xffffffffu, warp, 0) == 1) {   // warp-uniform branch + elect.sync: UTCHMMA issues from the uniform datapath
            if (elect_one()) {
                t0 = clock64();
                for (int i = 0; i < reps; i += 4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t d = tb + 128 + (chains == 4 ? j * 64 : 0);
                        if ((form & 1) == 0) umma_f16_ts(d, tb + 400 + j * 8, bdesc, idesc, 1);
                        else umma_f16(d, adesc, bdesc, idesc, 1);
                    }
                }
                umma_commit(smem_u32(&bar));
            }
            __syncwarp();
        }
        mbar_wait(smem_u32(&bar), phase); phase ^= 1;
        if (t0 != 0) cyc[form] = clock64() - t0;
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
    const int Ns[] = {16, 32, 48, 64};
    std::vector<__half> hA(128 * 16), hB(64 * 16);
    srand(1);
    for (auto& v : hA) v = __float2half((rand() % 2001 - 1000) / 500.0f);
    for (auto& v : hB) v = __float2half((rand() % 2001 - 1000) / 500.0f);
    __half *dA, *dB; float* dres; long long* dcyc;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dres, 2 * 128 * 64 * 4); cudaMalloc(&dcyc, 32);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    for (int N : Ns) {
        const int reps = 2000;
        cudaMemset(dres, 0, 2 * 128 * 64 * 4);
        ts_probe_kernel<<<1, 128, 16384>>>(dA, dB, N, reps, dres, dcyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
        std::vector<float> res(2 * 128 * N);
        long long cyc[4];
        cudaMemcpy(res.data(), dres, res.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(cyc, dcyc, 32, cudaMemcpyDeviceToHost);
        double err_ts = 0, err_ss = 0, ref_max = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int k = 0; k < 16; ++k) acc += double(__half2float(hA[m * 16 + k])) * double(__half2float(hB[n * 16 + k]));
                err_ts = fmax(err_ts, fabs(acc - res[m * N + n]));
                err_ss = fmax(err_ss, fabs(acc - res[128 * N + m * N + n]));
                ref_max = fmax(ref_max, fabs(acc));
            }
        printf("N=%2d: max|D_ts - ref| = %.3e, max|D_ss - ref| = %.3e (max|ref| %.2f); cycles per MMA (M=128, K=16): dependent chain TS %.1f / SS %.1f, 4 independent accumulators TS %.1f / SS %.1f\n",
               N, err_ts, err_ss, ref_max, double(cyc[0]) / reps, double(cyc[1]) / reps, double(cyc[2]) / reps, double(cyc[3]) / reps);
    }
    return 0;
}
