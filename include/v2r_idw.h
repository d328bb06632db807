/*
 * v2r_idw.h -- C ABI of libv2r_idw.so, the B200 (sm_100a) implementation of Dewberry/v2r's
 * Inverse Distance Weighting gridding hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI seam for IDW (it is 100 % Go); the seam
 * is the one its Go entry points would bind through cgo (see INTEGRATION.md for the shim):
 *
 *   reference interface (file:line, relative to the v2r repo)         replaced / served by
 *   ---------------------------------------------------------------   ---------------------------------
 *   idw.FullSolve hot loop            features/idw/idw.go:34-40       v2r_idw_solve
 *   idw.ChunkSolve / makeGridIDW      features/idw/idw_chunk.go:27-41,59-99
 *                                                                     v2r_idw_solve_rows (+ _begin/_next streaming)
 *   idw.calculateIDW                  features/idw/idw_utils.go:8-27  the CUDA kernels behind both
 *   tools.DistExp / euclidDist        tools/utils.go:156-168          (device arithmetic)
 *   tools.RCToPoint                   tools/utils.go:78-82            (device arithmetic, mul then add, unfused)
 *   exponent sweep loop               cmd/idw.go:153-161              exps[] / n_exps of ONE call
 *   tools.GetDimensions               tools/utils.go:133-137          v2r_idw_get_dimensions
 *   tools.PointToPair                 tools/utils.go:70-72            v2r_idw_point_to_pair
 *   tools.MakeCoordSpace              tools/utils.go:97-119           v2r_idw_make_coord_space (deterministic)
 *   expRange / sweep list             cmd/idw.go:78-90,153            v2r_idw_exp_range
 *
 * Conventions
 *   - Plain pointers and sizes only; no C++/torch types.  All floating point data is IEEE float64,
 *     all indices int64 (Go `int` on amd64).
 *   - Every function returns an int status (0 = V2R_IDW_OK).  The message of the last failure on
 *     the calling thread is available from v2r_idw_last_error().  Nothing aborts or throws.
 *   - The library never keeps a caller pointer after the call returns (cgo rule): _solve* return when
 *     `out` is complete, _stream_begin returns when the point arrays have been read completely.  `out`
 *     buffers are caller-provided; pinned buffers from v2r_idw_alloc_pinned() make the copies asynchronous
 *     inside a call.
 *   - Callable from any OS thread (goroutines migrate): every call sets its device explicitly and
 *     a context serialises its own calls with an internal mutex.
 *   - Raster layout = the reference's flattenGrid (features/idw/idw_utils.go:33-42): row-major
 *     float64, row 0 is y = y_min (south-up geotransform, tools/processing/io_gdal.go:59); a sweep
 *     returns n_exps such rasters back to back, exponent-major.
 *   - Exact-hit rule (features/idw/idw_utils.go:13-16): a cell (r,c) present in the key table
 *     (key_r[i], key_c[i]) takes z[i] verbatim for every exponent.  Keys must be unique, as Go map
 *     keys are; v2r_idw_make_coord_space() produces such a table.  Keys outside the grid never hit.
 *   - There is NO CPU fallback: without a CUDA device v2r_idw_init* fails with V2R_IDW_ECUDA.
 */
#ifndef V2R_IDW_H
#define V2R_IDW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V2R_IDW_ABI_VERSION 2

/* status codes */
#define V2R_IDW_OK        0
#define V2R_IDW_EINVAL    1 /* bad argument */
#define V2R_IDW_ECUDA     2 /* CUDA runtime / driver error, or no device */
#define V2R_IDW_ENCCL     3 /* NCCL error or NCCL not loadable (multi-GPU contexts only) */
#define V2R_IDW_ENOMEM    4 /* host or device allocation failed */
#define V2R_IDW_ESTATE    5 /* call sequence error (e.g. _next without _begin) */

/* precision of the accumulation path */
#define V2R_IDW_FP64 0 /* float64 throughout, like the reference; parity bar 1e-12 relative */
#define V2R_IDW_FP32 1 /* float32 distances/weights on tile-local coordinates, float64 across tiles; ~1e-5 */

typedef struct v2r_idw_ctx v2r_idw_ctx;

/* Grid geometry = the reference's (xInfo, yInfo) pair after GetDimensions. */
typedef struct v2r_idw_grid {
    double x_min, x_step; /* tools.Info{Min,Step} of x; cell column c sits at x_min + c*x_step */
    double y_min, y_step; /* tools.Info{Min,Step} of y; cell row    r sits at y_min + r*y_step */
    int64_t rows, cols;   /* tools.GetDimensions(xInfo, yInfo) */
} v2r_idw_grid;

/* ---- library / device ------------------------------------------------------------------ */
int v2r_idw_abi_version(void);
const char *v2r_idw_last_error(void);
/* Message of the last failed call that took this context, kept with the context: safe to fetch from another OS
 * thread than the one that made the failing call (a goroutine may migrate between two cgo calls). */
const char *v2r_idw_ctx_last_error(v2r_idw_ctx *ctx);
int v2r_idw_device_count(int *count);

/* One process driving GPUs 0..n_gpus-1 (n_gpus in {1,2,4,8}): the shape the Go host needs.
 * n_gpus > 1 loads NCCL (ncclCommInitAll): points are broadcast from device 0 and each device
 * computes a row band.  The bands reach the caller's host raster either straight from every device
 * (one PCIe link per GPU; taken when `out` is page-locked, e.g. from v2r_idw_alloc_pinned) or, for
 * pageable `out` (or V2R_IDW_GATHER=nccl), by an NCCL gather to device 0 and one copy from there. */
int v2r_idw_init(int n_gpus, v2r_idw_ctx **out);
/* One context on one given device: for one-process-per-GPU hosts (torchrun, MPI). */
int v2r_idw_init_device(int device, v2r_idw_ctx **out);
void v2r_idw_destroy(v2r_idw_ctx *ctx);

/* Page-locked host memory so readers can append straight into DMA-able SoA buffers. */
int v2r_idw_alloc_pinned(size_t bytes, void **ptr);
int v2r_idw_free_pinned(void *ptr);

/* ---- host-side helpers that define the inputs of the hot path ------------------------- */
/* tools/utils.go:133-137: cols = int((1 + xMax - xMin)/xStep), rows likewise (Go truncation). */
int v2r_idw_get_dimensions(double x_min, double x_max, double x_step, double y_min, double y_max,
                           double y_step, int64_t *rows, int64_t *cols);
/* tools/utils.go:70-72: r = int((y - yMin)/yStep), c = int((x - xMin)/xStep). */
int v2r_idw_point_to_pair(double x, double y, double x_min, double x_step, double y_min,
                          double y_step, int64_t *r, int64_t *c);
/* cmd/idw.go:153: exp = es; exp < ee; exp += ei accumulated in float64.  Writes at most max_out
 * values, stores the full count in *count.  EINVAL when ei <= 0 or ee <= es (cmd/idw.go:83-88). */
int v2r_idw_exp_range(double es, double ee, double ei, double *out, int64_t max_out, int64_t *count);
/* tools/utils.go:97-119 with the arrival order fixed to input order: colliding points are folded
 * by the pairwise running average (new + old)/2 and keep the X,Y of the last arrival.  Outputs
 * (capacity n each) are in first-arrival order of each key; *n_out = number of distinct keys. */
int v2r_idw_make_coord_space(int64_t n, const double *x, const double *y, const double *z,
                             double x_min, double x_step, double y_min, double y_step,
                             double *out_x, double *out_y, double *out_z, int64_t *out_key_r,
                             int64_t *out_key_c, int64_t *n_out);

/* ---- the hot path, host buffers in / host buffers out --------------------------------- */
/* Whole grid, all exponents, one pass over the points.
 *   x,y,z,key_r,key_c : n de-duplicated points (host memory, pageable or pinned)
 *   exps, n_exps      : the sweep (1 <= n_exps <= 4096; runs as passes of <= 10 exponents); each raster is IDW
 *                       with exponent exps[e]
 *   out               : n_exps * rows * cols float64, host memory                                  */
int v2r_idw_solve(v2r_idw_ctx *ctx, const double *x, const double *y, const double *z,
                  const int64_t *key_r, const int64_t *key_c, int64_t n, const v2r_idw_grid *grid,
                  const double *exps, int n_exps, int precision, double *out);

/* Rows [row0, row0 + nrows) only: out is n_exps * nrows * cols.  This is the chunk-write protocol
 * of ChunkSolve with offsets = (row0, 0) and bufferSize = (nrows, cols). */
int v2r_idw_solve_rows(v2r_idw_ctx *ctx, const double *x, const double *y, const double *z,
                       const int64_t *key_r, const int64_t *key_c, int64_t n, const v2r_idw_grid *grid,
                       int64_t row0, int64_t nrows, const double *exps, int n_exps, int precision,
                       double *out);

/* Streaming variant so the host can overlap its raster writer with the GPU: _begin uploads the
 * points and fixes the sweep; every _next computes the next band of up to band_rows rows for all
 * exponents into `out` (n_exps * band_rows * cols capacity; rasters packed with the band's actual
 * row count) while the following band is already running; *row0 / *nrows describe what was
 * returned, *nrows == 0 means the grid is finished.  _end releases the session. */
int v2r_idw_stream_begin(v2r_idw_ctx *ctx, const double *x, const double *y, const double *z,
                         const int64_t *key_r, const int64_t *key_c, int64_t n, const v2r_idw_grid *grid,
                         const double *exps, int n_exps, int precision, int64_t band_rows);
int v2r_idw_stream_next(v2r_idw_ctx *ctx, double *out, int64_t *row0, int64_t *nrows);
int v2r_idw_stream_end(v2r_idw_ctx *ctx);

/* ---- the hot path on device-resident data (one device, caller's stream) --------------- */
/* For hosts that already own device memory (a per-GPU worker under torchrun/MPI whose points
 * arrived by an NCCL broadcast).  All d_* pointers are device pointers on `device`, 16-byte
 * aligned; `stream` is a cudaStream_t (NULL = default stream).  Asynchronous: returns after the
 * launches are queued.  d_out receives n_exps * nrows * cols float64. */
int v2r_idw_launch_rows(int device, void *stream, const double *d_x, const double *d_y,
                        const double *d_z, const int64_t *d_key_r, const int64_t *d_key_c, int64_t n,
                        const v2r_idw_grid *grid, int64_t row0, int64_t nrows, const double *exps,
                        int n_exps, int precision, double *d_out);

/* ---- measurement ------------------------------------------------------------------------ */
/* Timing of the most recent solve on this context (CUDA events, milliseconds; max over the context's devices):
 * host->device upload + broadcast; span of the kernels (first launch to last completion); copy-out (gather +
 * device->host) that was NOT hidden behind kernels, i.e. last kernel end to last copy end.  Any pointer may be NULL. */
int v2r_idw_last_timing(v2r_idw_ctx *ctx, double *h2d_ms, double *kernel_ms, double *d2h_ms);
/* Number of kernel launches issued by this library since it was loaded (all contexts). */
int64_t v2r_idw_launch_count(void);
/* Describes the kernel variant the dispatcher picks for a sweep: writes a NUL-terminated string
 * such as "fp64/rcp E=1 cells=4x4 k0=1 dk=0" and the modelled FP64-pipe instruction count per
 * cell-point pair. */
int v2r_idw_describe(const double *exps, int n_exps, int precision, char *buf, size_t buf_len,
                     double *fp64_ops_per_pair);
/* Roofline denominators measured on the device: dependent-free DFMA issue rate and MUFU.RCP64H /
 * FP32 MUFU.RCP rates, in G instructions (thread-level) per second. Runs for about `ms` each. */
int v2r_idw_measure_peaks(int device, double ms, double *dfma_ginst_s, double *mufu64_ginst_s,
                          double *mufu32_ginst_s, double *ffma_ginst_s);

/* One microbenchmark at a time.  DFMA_3SRC issues DFMAs with three distinct register-pair sources -- the form
 * the IDW kernels execute -- as opposed to DFMA's one register source + two constants. */
#define V2R_IDW_PEAK_DFMA         0
#define V2R_IDW_PEAK_MUFU_RCP64H  1
#define V2R_IDW_PEAK_MUFU_RCP_F32 2
#define V2R_IDW_PEAK_FFMA         3
#define V2R_IDW_PEAK_DFMA_3SRC    4
#define V2R_IDW_PEAK_DADD         5
int v2r_idw_measure_peak(int device, double ms, int which, double *ginst_s);

#ifdef __cplusplus
}
#endif
#endif /* V2R_IDW_H */
