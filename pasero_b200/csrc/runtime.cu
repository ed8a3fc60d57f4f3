// pasero_b200 — library-wide plumbing: per-thread error text, device queries, ABI version.
#include "common.cuh"
#include "../../include/pasero_b200.h"

#include <stdarg.h>
#include <stdlib.h>

namespace pb200_rt {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PB200_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

}  // namespace pb200_rt

extern "C" int32_t pb200_abi_version(void) { return PB200_ABI_VERSION; }
extern "C" const char* pb200_last_error_string(void) { return pb200_rt::g_err; }
extern "C" int32_t pb200_num_sms(void) { return pb200_rt::num_sms(); }
