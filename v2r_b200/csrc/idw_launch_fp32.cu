// idw_launch_fp32.cu -- dispatch of the FP32 fast path (see idw_kernels_fp32.cuh).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "idw_internal.h"
#include "idw_kernels_fp32.cuh"

namespace v2r {

struct FVariant {
    int mode, E;
    bool packed, iso;  // packed: the two-wide p = 2 body; iso: x_step == y_step (no row rescaling)
    int cr, cc, threads, minb;
    void (*fn)(const KParams);
    int smem;
    const char *tag;  // non-null: tuning build (V2R_IDW_TUNE=<tag>)
};

#define V2R_FVARIANT(MODE, E, CR, CC, MINB)                                                           \
    FVariant { MODE, E, false, false, CR, CC, 256, MINB, idw_fp32_kernel<MODE, E, CR, CC, 256, MINB>, \
               smem_bytes(Fp32Body<MODE, E, CR, CC, 256>::BODY_BYTES, 2 * (E) * (CR) * (CC), 256), nullptr }
#define V2R_FPACKED(TAG, CR, CC, ISO, THREADS, MINB)                                                                  \
    FVariant { MODE_RCP, 1, true, ISO, CR, CC, THREADS, MINB, idw_fp32_packed_kernel<CR, CC, ISO, THREADS, MINB>,     \
               smem_bytes(Fp32PackedBodyFor<CR, CC, ISO, THREADS, MINB>::BODY_BYTES, 2 * (CR) * (CC), THREADS), TAG }
// cell blocks sized so that 2 * E * CR * CC FP64 accumulators + as many FP32 partials fit the register budget without
// spilling (two CTAs per SM up to 6 exponents, one beyond)
#define V2R_FROW(MODE)                                                                          \
    V2R_FVARIANT(MODE, 1, 2, 4, 2), V2R_FVARIANT(MODE, 2, 2, 2, 2),                             \
        V2R_FVARIANT(MODE, 3, 2, 2, 2), V2R_FVARIANT(MODE, 4, 1, 2, 2),                         \
        V2R_FVARIANT(MODE, 5, 1, 2, 2), V2R_FVARIANT(MODE, 6, 1, 2, 2),                         \
        V2R_FVARIANT(MODE, 7, 1, 2, 1), V2R_FVARIANT(MODE, 8, 1, 2, 1),                         \
        V2R_FVARIANT(MODE, 9, 1, 2, 1), V2R_FVARIANT(MODE, 10, 1, 2, 1)

static const FVariant kFVariants[] = {
    // tuning builds of the packed p = 2 kernel (cell block, CTAs per SM)
    V2R_FPACKED("f2x8", 2, 8, true, 256, 1),   // (4x8 and 2x16 were measured too, 5 383 / 5 082; they spill and were dropped)
    // p = 2: packed FP32 (FFMA2 / FADD2 / FMUL2), two points per MUFU.RCP, squared row distances staged in shared
    // memory; 4x4 cells measured 5 604 G evals/s on c2 against 5 383 (4x8), 5 379 (2x8) and 5 082 (2x16)
    V2R_FPACKED(nullptr, 4, 4, true, 256, 1), V2R_FPACKED(nullptr, 4, 4, false, 256, 1),
    V2R_FROW(MODE_RCP), V2R_FROW(MODE_RSQ), V2R_FROW(MODE_QRT), V2R_FROW(MODE_GEN),
};

static const FVariant *pick_fvariant(const Pass &ps, bool iso) {
    static const char *tune = getenv("V2R_IDW_TUNE");
    const bool p2 = ps.mode == MODE_RCP && ps.count == 1 && ps.k0 == 1;
    if (tune && *tune && p2 && iso)
        for (const FVariant &v : kFVariants)
            if (v.tag && !strcmp(v.tag, tune)) return &v;
    for (const FVariant &v : kFVariants) {
        if (v.tag) continue;
        if (v.packed) {
            if (p2 && v.iso == iso) return &v;
            continue;
        }
        if (v.mode == ps.mode && v.E == ps.count) return &v;
    }
    return nullptr;
}

int launch_rows_fp32(int device, cudaStream_t stream, const double *d_x, const double *d_y, const double *d_z,
                     const long long *d_kr, const long long *d_kc, long long n, const v2r_idw_grid &g,
                     long long row0, long long nrows, const double *exps, int n_exps, double *d_out) {
    std::vector<Pass> passes;
    int rc = plan_sweep(exps, n_exps, passes);
    if (rc) return rc;
    const long long plane = nrows * g.cols;
    for (const Pass &ps : passes) {
        const FVariant *v = pick_fvariant(ps, g.x_step == g.y_step);
        if (!v) return set_error(V2R_IDW_EINVAL, "no FP32 kernel variant for mode %d E %d", ps.mode, ps.count);
        KParams kp{};
        fill_params(kp, d_x, d_y, d_z, n, g, row0, nrows);
        kp.out = d_out + (long long)ps.first * plane;
        kp.k0 = ps.k0 > 0 ? ps.k0 : 1;
        kp.dk = ps.dk > 0 ? ps.dk : 1;
        for (int i = 0; i < ps.count; ++i) kp.cexp[i] = ps.cexp[i];
        const int slots = device_slots(device, (const void *)v->fn, v->threads, v->smem, nullptr);
        if (!slots) return V2R_IDW_ECUDA;
        plan_grid(kp, (v->threads / 32) * v->cr, 32 * v->cc, slots);
        const long long grid = kp.n_main + kp.n_tail;
        if (grid > 2147483647LL || kp.n_tiles * kp.segs > 2147483647LL)
            return set_error(V2R_IDW_EINVAL, "band of %lld rows x %lld cols exceeds one launch; use smaller bands", nrows, (long long)g.cols);
        void *block = nullptr;
        if (kp.n_tail > 0 && (rc = alloc_scratch(stream, kp, 2 * v->E * v->cr * v->cc, v->threads, &block, nullptr))) return rc;
        v->fn<<<(unsigned)grid, v->threads, v->smem, stream>>>(kp);
        count_launch();
        cudaError_t ce = cudaGetLastError();
        // the packed kernel has no d == 0 semantics of its own (coordinates are rescaled): exact coincidences -> NaN here
        if (ce == cudaSuccess && v->packed) rc = launch_poison(stream, d_x, d_y, n, g, row0, nrows, kp.out, 1, nullptr);
        if (block) cudaFreeAsync(block, stream);
        if (rc) return rc;
        if (ce != cudaSuccess) return set_error(V2R_IDW_ECUDA, "idw fp32 kernel launch: %s", cudaGetErrorString(ce));
    }
    return launch_patch_exact(stream, d_z, d_kr, d_kc, n, g, row0, nrows, d_out, n_exps);
}

std::string describe_fp32(const std::vector<Pass> &passes) {
    std::string s;
    static const char *names[] = {"rcp", "rsqrt", "root4", "general"};
    for (const Pass &ps : passes) {
        const FVariant *v = pick_fvariant(ps, true);
        char tmp[160];
        snprintf(tmp, sizeof tmp, "%sfp32/%s%s E=%d cells=%dx%d k0=%d dk=%d", s.empty() ? "" : " + ", names[ps.mode],
                 v && v->packed ? " (packed f32x2, paired rcp)" : "", ps.count, v ? v->cr : 0, v ? v->cc : 0, ps.k0, ps.dk);
        s += tmp;
    }
    return s;
}

}  // namespace v2r
