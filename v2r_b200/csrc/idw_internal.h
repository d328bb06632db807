// idw_internal.h -- declarations shared by the translation units of libv2r_idw.so.
#pragma once
#include <atomic>
#include <cstdarg>
#include <exception>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/v2r_idw.h"

namespace v2r {

// One kernel pass over the points: `count` consecutive exponents of the sweep starting at `first`.
struct Pass {
    int mode;         // Mode
    int first, count; // exponent slice
    int k0, dk;       // ladder classes: h_e = base^(k0 + e*dk)
    double cexp[16];  // general class: c_e = -p_e / 2
};

int set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
const char *last_error();

// Nothing may unwind across the C ABI (cgo would abort the Go process): every entry point that can allocate runs
// its body through this.
template <typename F>
int guarded(F &&body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc &) {
        return set_error(V2R_IDW_ENOMEM, "out of host memory");
    } catch (const std::exception &e) {
        return set_error(V2R_IDW_EINVAL, "internal error: %s", e.what());
    } catch (...) {
        return set_error(V2R_IDW_EINVAL, "internal error");
    }
}

constexpr int kMaxSweep = 4096;  // exponents per call (a sweep runs as ceil(n / 10) passes at most)
int plan_sweep(const double *exps, int n_exps, std::vector<Pass> &passes);
int validate_geometry(long long n, const v2r_idw_grid *g, long long row0, long long nrows, int n_exps);
int validate_launch(const void *d_x, const void *d_y, const void *d_z, long long n, const v2r_idw_grid *g,
                    long long row0, long long nrows, int n_exps, const void *d_out);
int launch_rows_fp64(int device, cudaStream_t stream, const double *d_x, const double *d_y, const double *d_z,
                     const long long *d_kr, const long long *d_kc, long long n, const v2r_idw_grid &g,
                     long long row0, long long nrows, const double *exps, int n_exps, double *d_out);
int launch_rows_fp32(int device, cudaStream_t stream, const double *d_x, const double *d_y, const double *d_z,
                     const long long *d_kr, const long long *d_kc, long long n, const v2r_idw_grid &g,
                     long long row0, long long nrows, const double *exps, int n_exps, double *d_out);
std::string describe_fp32(const std::vector<Pass> &passes);
struct KParams;
struct GenTables;
int device_slots(int device, const void *fn, int threads, int smem, const GenTables **tables);
void plan_grid(KParams &kp, int tile_r, int tile_c, int slots);
int alloc_scratch(cudaStream_t stream, KParams &kp, int nacc, int threads, void **block, unsigned **flag);
void fill_params(KParams &kp, const double *d_x, const double *d_y, const double *d_z, long long n, const v2r_idw_grid &g,
                 long long row0, long long nrows);
int launch_poison(cudaStream_t stream, const double *d_x, const double *d_y, long long n, const v2r_idw_grid &g, long long row0,
                  long long nrows, double *d_out, int ne, unsigned *flag);
double estimate_pair_rate(const double *exps, int n_exps, int precision);
long long launch_count();
void count_launch();
int launch_patch_exact(cudaStream_t stream, const double *d_z, const long long *d_kr, const long long *d_kc, long long n,
                       const v2r_idw_grid &g, long long row0, long long nrows, double *d_out, int n_exps);

}  // namespace v2r
