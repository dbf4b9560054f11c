// Shared host-side helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include "gnngls_b200.h"

namespace gnngls {

// thread-local message behind gnngls_last_error_string()
void set_error(const char *fmt, ...);
int device_sm_count();
int device_max_optin_smem();

#define GNNGLS_REQUIRE(cond, code, ...)                 \
    do {                                                \
        if (!(cond)) {                                  \
            ::gnngls::set_error(__VA_ARGS__);           \
            return (code);                              \
        }                                               \
    } while (0)

#define GNNGLS_CUDA_OK(expr)                                                               \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            ::gnngls::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),   \
                                __FILE__, __LINE__);                                       \
            return GNNGLS_ERR_CUDA;                                                        \
        }                                                                                  \
    } while (0)

#define GNNGLS_LAUNCH_OK(what)                                                             \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            ::gnngls::set_error("launch of %s failed: %s", what, cudaGetErrorString(e__)); \
            return GNNGLS_ERR_CUDA;                                                        \
        }                                                                                  \
    } while (0)

}  // namespace gnngls
