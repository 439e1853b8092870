// runtime.cu — error reporting, launch accounting and ABI version for libia2c_b200.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace ia2c {
namespace {
thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(err));
        return IA2C_ERR_CUDA;
    }
    count_launch(1);
    return IA2C_OK;
}
}  // namespace ia2c

extern "C" const char* ia2c_last_error(void) { return ia2c::g_error; }
extern "C" int ia2c_abi_version(void) { return IA2C_ABI_VERSION; }
extern "C" uint64_t ia2c_launch_count(void) { return ia2c::g_launches.load(std::memory_order_relaxed); }
