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

extern "C" const char *bbx_last_error(void) { return g_err; }
extern "C" int bbx_version(void) { return 100; }
