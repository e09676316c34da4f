// api.cu — error reporting and version of the C ABI (include/posecnn_b200.h)
#include <stdarg.h>

#include "common.cuh"

namespace pcnn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace pcnn

extern "C" const char* pcnn_last_error(void) { return pcnn::g_err; }
extern "C" int pcnn_version(void) { return 100; }
