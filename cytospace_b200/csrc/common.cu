// Library-wide plumbing of libcytospace_b200.so: ABI version, per-thread error
// string, device query.  (include/cytospace_b200.h documents the contract.)
#include "common.h"

#include <cstring>

namespace cyb {

char *error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace cyb

extern "C" int cyb_abi_version(void) { return CYB_ABI_VERSION; }

extern "C" const char *cyb_last_error(void) { return cyb::error_buffer(); }

extern "C" int cyb_device_info(int device, int *sm_count, int *cc_major, int *cc_minor,
                               size_t *total_mem_bytes) {
    cudaDeviceProp prop;
    CYB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem_bytes) *total_mem_bytes = prop.totalGlobalMem;
    return CYB_OK;
}
