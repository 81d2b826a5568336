// api.cu — process-wide state behind the C-ABI (include/repsurf_b200.h): last-error text,
// cached device properties, launch counter.
#include "common.cuh"
#include <stdarg.h>
#include <mutex>

unsigned long long g_rsb_launches = 0;

static thread_local char t_err[512] = "";

void rsb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

RSB_EXPORT const char *rsb_last_error(void) { return t_err; }

int rsb_sm_count()
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;  // B200
    }
    return sms;
}

RSB_EXPORT unsigned long long rsb_launch_count(void) { return g_rsb_launches; }
RSB_EXPORT void rsb_reset_launch_count(void) { g_rsb_launches = 0; }
RSB_EXPORT int rsb_abi_version(void) { return 1; }
