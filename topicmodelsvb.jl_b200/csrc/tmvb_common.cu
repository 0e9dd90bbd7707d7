// tmvb_common.cu -- error plumbing and host fp64 special functions shared by the model files.
#include <math.h>
#include <stdarg.h>

#include "tmvb_common.cuh"

namespace tmvb {

std::string &last_error()
{
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

// psi(x), x > 0 (host, fp64): recurrence up to x >= 10, then the Bernoulli series.
double h_digamma(double x)
{
    double r = 0.0;
    while (x < 10.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    double t = 1.0 / x, t2 = t * t;
    double s = t2 * (1.0 / 12 - t2 * (1.0 / 120 - t2 * (1.0 / 252 - t2 * (1.0 / 240 - t2 * (1.0 / 132 - t2 * (691.0 / 32760 - t2 * (1.0 / 12)))))));
    return r + log(x) - 0.5 * t - s;
}

// psi'(x), x > 0 (host, fp64)
double h_trigamma(double x)
{
    double r = 0.0;
    while (x < 10.0) {
        r += 1.0 / (x * x);
        x += 1.0;
    }
    double t = 1.0 / x, t2 = t * t;
    double s = t * (1.0 + 0.5 * t + t2 * (1.0 / 6 - t2 * (1.0 / 30 - t2 * (1.0 / 42 - t2 * (1.0 / 30 - t2 * (5.0 / 66 - t2 * (691.0 / 2730 - t2 * (7.0 / 6))))))));
    return r + s;
}

}  // namespace tmvb

extern "C" {

int tmvb_version(void) { return TMVB_VERSION; }

const char *tmvb_last_error(void) { return tmvb::last_error().c_str(); }

int tmvb_device_count(int *count)
{
    TMVB_CHECK_ARG(count != nullptr, "count is NULL");
    *count = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return tmvb::fail((int)e, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    *count = n;
    return 0;
}

int tmvb_alloc_pinned(void **ptr, int64_t bytes)
{
    TMVB_CHECK_ARG(ptr != nullptr && bytes >= 0, "bad pinned allocation request");
    *ptr = nullptr;
    TMVB_CUDA(cudaHostAlloc(ptr, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocPortable | cudaHostAllocMapped));
    return 0;
}

int tmvb_free_pinned(void *ptr)
{
    if (ptr) TMVB_CUDA(cudaFreeHost(ptr));
    return 0;
}

}  // extern "C"
