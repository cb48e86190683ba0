// ufair_internal.h -- shared between the translation units of libufair.so (not installed).
#pragma once
#include <cuda_runtime.h>

#include "../../include/ufair.h"

namespace ufair {
int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what);
int validate_desc(const ufair_desc* d, size_t elem);
template <typename Real> int run_device(const ufair_desc* d, cudaStream_t stream);
template <typename Real> int run_stats_pass(const ufair_desc* d, cudaStream_t stream);
}  // namespace ufair
