#include "common.h"
#include <cstring>

namespace gnngls {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int cached_attr(cudaDeviceAttr attr, int *cache /* per device, 64 */) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (cache[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, attr, dev) == cudaSuccess) cache[dev] = v;
    }
    return cache[dev];
}

int device_sm_count() {
    static int cache[64] = {0};
    return cached_attr(cudaDevAttrMultiProcessorCount, cache);
}

int device_max_optin_smem() {
    static int cache[64] = {0};
    return cached_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, cache);
}

}  // namespace gnngls

extern "C" int gnngls_abi_version(void) { return GNNGLS_B200_ABI_VERSION; }
extern "C" const char *gnngls_last_error_string(void) { return gnngls::g_err; }
