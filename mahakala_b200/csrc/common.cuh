// Shared host-side helpers of libmahakala_b200.so (error reporting, launch geometry).
#pragma once
#include <cuda_runtime.h>
#include "fp64_math.cuh"
#include <cstdint>
#include <cstdio>

namespace mk {

void set_error(const char* fmt, ...);
int sm_count();                 // multiprocessors of the current device (148 on B200)
unsigned int* queue_counter(cudaStream_t stream, int slot);   // zeroed device counter for persistent kernels

#define MK_CUDA_CHECK(expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            mk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define MK_REQUIRE(cond, msg)                                                                      \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            mk::set_error("invalid argument: %s (%s)", msg, #cond);                                \
            return 2;                                                                              \
        }                                                                                          \
    } while (0)

}  // namespace mk
