// common.cuh -- host-side error checking for the C++ driver and the kind wrappers.
// Same names and behaviour as reference/include/common.cuh:1-23 (USED_DEVICE, FULL_MASK,
// checkCudaErr prints and aborts), plus checkSpmvStatus for the C-ABI return codes, which
// prints and exits the way the reference's library wrappers do
// (reference/include/spmv/cusparse.cuh:13-21, spmv.h:46-47).
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "spmv_b200.h"
#include "timer.hpp"

#define USED_DEVICE 0

#define FULL_MASK 0xffffffff

template <typename T>
void CheckCudaErr(T result, char const *const func, const char *const file, int const line) {
    if (result) {
        fprintf(stderr, "CUDA error at %s:%d code=%d(%s) \"%s\" \n", file, line,
                static_cast<unsigned int>(result), cudaGetErrorName(result), func);
        abort();
    }
}
#define checkCudaErr(val) CheckCudaErr((val), #val, __FILE__, __LINE__)

inline void CheckSpmvStatus(int status, char const *const func, const char *const file,
                            int const line) {
    if (status != SPMVB200_OK) {
        fprintf(stderr, "SpMV error at %s:%d status=%d(%s) \"%s\" %s\n", file, line, status,
                spmvb200_status_string(status), func,
                (status == SPMVB200_ERR_CUDA || status == SPMVB200_ERR_CUSPARSE)
                    ? spmvb200_last_cuda_error()
                    : "");
        exit(EXIT_FAILURE);
    }
}
#define checkSpmvStatus(val) CheckSpmvStatus((val), #val, __FILE__, __LINE__)

// SpmvStream (the stream every kind enqueues on) lives in timer.hpp, beside the Timer that
// records its events on it.
