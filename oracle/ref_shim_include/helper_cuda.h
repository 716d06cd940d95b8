// helper_cuda.h -- stand-in for the cuda-samples header the reference includes (reference Dockerfile:80-83 fetches
// cuda-samples v10.2; it is not in this image and there is no network).  TEST/BENCH INFRASTRUCTURE ONLY: it exists so
// that oracle/Makefile's `ref_gpu` target can compile the reference's own .cu files where they lie.  The reference uses
// exactly one thing from the real header, the abort-on-error macro.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define checkCudaErrors(call)                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            std::exit(1);                                                                             \
        }                                                                                             \
    } while (0)
