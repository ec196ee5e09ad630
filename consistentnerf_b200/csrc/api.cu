// Version / error plumbing of the C ABI.
#include "common.cuh"
#include <string.h>

namespace cnerf {
static thread_local char g_err[512] = {0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace cnerf

extern "C" int cnerf_version(void) { return CNERF_VERSION; }

extern "C" int cnerf_last_error(char* buf, int len) {
    int n = (int)strlen(cnerf::g_err);
    if (buf && len > 0) {
        int c = n < len - 1 ? n : len - 1;
        memcpy(buf, cnerf::g_err, c);
        buf[c] = 0;
    }
    return n;
}

extern "C" int cnerf_device_info(int* cc, int* sms) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cnerf::check_cuda(e, "cudaGetDevice");
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return cnerf::check_cuda(e, "cudaGetDeviceProperties");
    if (cc) *cc = p.major * 10 + p.minor;
    if (sms) *sms = p.multiProcessorCount;
    return CNERF_OK;
}
