// C-ABI plumbing: version, build target, thread-local error text.
#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

namespace gs {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("GSPLAT_B200_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return GS_ERR_CUDA;
}

}  // namespace gs

extern "C" int gs_abi_version(void) { return GS_ABI_VERSION; }

extern "C" const char* gs_last_error_string(void) { return gs::g_error; }

namespace gs { long long launches(); }
extern "C" int64_t gs_kernel_launch_count(void) { return (int64_t)gs::launches(); }

extern "C" int gs_built_for_sm(void) {
#ifdef GS_BUILT_FOR_SM
    return GS_BUILT_FOR_SM;
#else
    return 0;
#endif
}
