// api.cu — error reporting and version of the C ABI (include/posecnn_b200.h)
#include <stdarg.h>

#include <mutex>
#include <set>
#include <utility>

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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: the opt-in is remembered per
// (kernel, device ordinal), so a process that drives several GPUs opts in on each of them; failures are reported.
int smem_optin(const void* func, int bytes, const char* what)
{
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("%s: cudaGetDevice: %s", what, cudaGetErrorString(e)); return PCNN_E_CUDA; }
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({func, dev})) return PCNN_OK;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("%s: cannot reserve %d B of dynamic shared memory on device %d: %s", what, bytes, dev, cudaGetErrorString(e));
        return PCNN_E_CUDA;
    }
    done.insert({func, dev});
    return PCNN_OK;
}
}  // namespace pcnn

extern "C" const char* pcnn_last_error(void) { return pcnn::g_err; }
extern "C" int pcnn_version(void) { return 100; }
