// common.h -- host-side helpers shared by the translation units of libdfb_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/dfb.h"

namespace dfb {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace dfb

#ifndef DFB_REQUIRE
#define DFB_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            dfb::set_error(__VA_ARGS__);  \
            return DFB_ERR_INVALID;       \
        }                                 \
    } while (0)
#endif

#define DFB_CUDA(call)                                   \
    do {                                                 \
        int _r = dfb::check_cuda((call), #call);         \
        if (_r != DFB_OK) return _r;                     \
    } while (0)

#define DFB_LAUNCH_CHECK(name)                                        \
    do {                                                              \
        int _r = dfb::check_cuda(cudaGetLastError(), "launch " name); \
        if (_r != DFB_OK) return _r;                                  \
    } while (0)
