// Library plumbing: error string, version, launch counter.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace nppc {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}
}  // namespace nppc

extern "C" {
const char* nppc_last_error(void) { return nppc::g_err; }
const char* nppc_version(void) { return "nppc_b200 0.1 sm_100a"; }
long long nppc_launch_count(void) { return nppc::g_launches.load(); }
void nppc_reset_launch_count(void) { nppc::g_launches.store(0); }
}
