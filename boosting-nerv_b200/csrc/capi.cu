// C-ABI plumbing shared by every entry point: thread-local error text, launch accounting, size helpers.
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include "common.cuh"

namespace bnerv {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char* what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

}  // namespace bnerv

using namespace bnerv;

extern "C" int bnerv_abi_version(void) { return BNERV_ABI_VERSION; }
extern "C" const char* bnerv_last_error(void) { return bnerv::g_err; }
extern "C" uint64_t bnerv_launch_count(void) { return bnerv::g_launches.load(std::memory_order_relaxed); }

extern "C" size_t bnerv_c8_numel(int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return static_cast<size_t>(B) * round_up(C, 16) * H * W;
}
extern "C" size_t bnerv_packed_weight_numel(int Cout, int Cin, int k, int s) {
    if (Cout <= 0 || Cin <= 0 || k <= 0 || s <= 0) return 0;
    return static_cast<size_t>(k) * k * round_up(Cin, 16) * s * s * round_up(Cout, 16);
}
extern "C" size_t bnerv_packed_bias_numel(int Cout, int s) {
    if (Cout <= 0 || s <= 0) return 0;
    return static_cast<size_t>(s) * s * round_up(Cout, 16);
}
