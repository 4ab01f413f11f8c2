// Host-side helpers shared by the translation units of libcytospace_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

#include "cytospace_b200.h"

namespace cyb {

// Per-thread error string behind cyb_last_error().
char *error_buffer();
int set_error(int code, const char *fmt, ...);

#define CYB_CUDA_CHECK(expr)                                                         \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess)                                                       \
            return ::cyb::set_error(CYB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,    \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);     \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace cyb
