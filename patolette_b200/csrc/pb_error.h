// pb_error.h - CUDA error plumbing shared by host-only and device translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#define PB_CUDA_OK(expr)                                                                    \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            fprintf(stderr, "patolette_b200: CUDA error %s at %s:%d: %s\n",               \
                    cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e));      \
            throw pb_cuda_error(_e);                                                        \
        }                                                                                   \
    } while (0)

struct pb_cuda_error {
    cudaError_t code;
    explicit pb_cuda_error(cudaError_t c) : code(c) {}
};

