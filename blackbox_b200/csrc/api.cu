// api.cu -- error reporting and version of libbbx.so
#include <stdarg.h>
#include "bbx_common.cuh"

static thread_local char g_err[512] = "";

void bbx_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Multiprocessor count of the calling thread's current device, queried on first use and cached
// per device (cudaDeviceGetAttribute is legal during stream capture).
int bbx_sm_count(void)
{
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev];
    if (n > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
    return n;
}

extern "C" const char *bbx_last_error(void) { return g_err; }
extern "C" int bbx_version(void) { return 100; }
