// idw_kernels_fp32.cuh -- FP32 fast path (placeholder until the kernel lands; fails loudly).
#pragma once
#include "idw_internal.h"
namespace v2r {
int launch_rows_fp32(cudaStream_t, const double *, const double *, const double *, const long long *,
                            const long long *, long long, const v2r_idw_grid &, long long, long long,
                            const double *, int, double *) {
    return set_error(V2R_IDW_EINVAL, "FP32 path is not built in this library version");
}
std::string describe_fp32(const std::vector<Pass> &) { return "fp32/unavailable"; }
}  // namespace v2r
