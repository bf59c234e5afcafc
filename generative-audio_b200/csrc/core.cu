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
int sm_count() {   // cached per device (a process may drive several GPUs)
    static int n[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (n[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}
}  // namespace nppc

extern "C" {
const char* nppc_last_error(void) { return nppc::g_err; }
const char* nppc_version(void) { return "nppc_b200 0.1 sm_100a"; }
long long nppc_launch_count(void) { return nppc::g_launches.load(); }
void nppc_reset_launch_count(void) { nppc::g_launches.store(0); }
}
