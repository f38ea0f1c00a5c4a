// lib.cu -- library-level entry points and shared host helpers of libsln_b200.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace sln {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// SM count of the CURRENT device, asked every time (the runtime answers from its own table in well under a microsecond):
// no process-global cache, so a process that switches devices -- or the header's "no global state" -- stays correct
int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

// read on every launch (a getenv is ~100 ns) so that one process can A/B both settings
bool pdl_enabled()
{
    const char *e = getenv("SLN_PDL");
    return !(e && e[0] == '0');
}

}  // namespace sln

extern "C" int sln_version(void) { return 100; }   // 0.1.0

extern "C" const char *sln_last_error_string(void) { return sln::g_err; }

extern "C" int sln_device_info(int *sm_count_out, int *cc, size_t *smem_optin, size_t *l2_bytes)
{
    int dev = 0;
    SLN_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SLN_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sm_count_out) *sm_count_out = p.multiProcessorCount;
    if (cc) *cc = p.major * 10 + p.minor;
    if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
    if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
    return SLN_OK;
}
