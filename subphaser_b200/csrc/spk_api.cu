// Error string, version and device attributes of libspk.
#include <stdarg.h>
#include <string.h>
#include "spk_common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_spk_launches = 0;

void spk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int spk_num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
        cached_dev = dev;
        cached_sms = sms;
    }
    return cached_sms;
}

extern "C" const char* spk_last_error(void) { return g_err; }
extern "C" int spk_version(void) { return 100; }
extern "C" int spk_sm_count(void) { return spk_num_sms(); }
extern "C" uint64_t spk_launch_count(void) { return g_spk_launches; }
