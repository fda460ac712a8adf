// Stand-in for the reference's include/cuda_toolkit/cuda_macro.h: vector types and the error convention.
// The reference exit(1)s on a CUDA error (cuda_macro.h:21-31); the wrappers in this tree throw gie::Error instead
// (define GIE_COMPAT_EXIT_ON_ERROR to get the reference's behaviour).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector_types.h>
#include <vector_functions.h>
#include "gie_b200.h"

namespace gie {
struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string &what) : std::runtime_error(what), status(st) {}
};
inline void check(int status, const char *where)
{
    if (status == GIE_OK) return;
    std::string msg = std::string(where) + ": gie status " + std::to_string(status) + ": " + gie_last_error();
#ifdef GIE_COMPAT_EXIT_ON_ERROR
    std::fprintf(stderr, "%s\n", msg.c_str());
    std::exit(1);
#else
    throw Error(status, msg);
#endif
}
}  // namespace gie
#define GIE_CHECK(call) ::gie::check((call), #call)

// The device-memory helpers the reference's ROS-typed MapMakers use for their sensor staging buffers
// (src/*_map_maker.cpp: GPU_MALLOC / GPU_MEMCPY_H2D / GPU_FREE) and the node's profiling sync (volumetric_mapper.cpp:186).
#include <cuda_runtime_api.h>
namespace gie {
inline void cuda_ok(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return;
    std::fprintf(stderr, "CUDA error in %s: %s\n", what, cudaGetErrorString(e));
    std::exit(1);
}
}  // namespace gie
#define GPU_MALLOC(ptr, bytes) ::gie::cuda_ok(cudaMalloc((void **)(ptr), (bytes)), "GPU_MALLOC")
#define GPU_FREE(ptr) ::gie::cuda_ok(cudaFree(ptr), "GPU_FREE")
#define GPU_MEMSET(ptr, v, bytes) ::gie::cuda_ok(cudaMemset((ptr), (v), (bytes)), "GPU_MEMSET")
#define GPU_MEMCPY_H2D(d, s, bytes) ::gie::cuda_ok(cudaMemcpy((d), (s), (bytes), cudaMemcpyHostToDevice), "GPU_MEMCPY_H2D")
#define GPU_MEMCPY_D2H(d, s, bytes) ::gie::cuda_ok(cudaMemcpy((d), (s), (bytes), cudaMemcpyDeviceToHost), "GPU_MEMCPY_D2H")
#define GPU_MEMCPY_D2D(d, s, bytes) ::gie::cuda_ok(cudaMemcpy((d), (s), (bytes), cudaMemcpyDeviceToDevice), "GPU_MEMCPY_D2D")
#define GPU_DEV_SYNC() ::gie::cuda_ok(cudaDeviceSynchronize(), "GPU_DEV_SYNC")
#define SENS_FAR_DIST 1000.f
#define GPU_PI_FLOAT 3.1415926f
