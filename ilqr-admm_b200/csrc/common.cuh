// Shared host-side helpers of libisls_b200.so (error reporting used by every translation unit).
#pragma once
#include <cuda_runtime.h>

#include <string>

int isls_fail(int code, const std::string &msg);                 // sets isls_last_error_string(), returns code
int isls_cuda_fail(cudaError_t e, const char *what);

#define CK(call)                                                \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return isls_cuda_fail(e__, #call);  \
  } while (0)
