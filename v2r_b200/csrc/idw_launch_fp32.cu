// idw_launch_fp32.cu -- dispatch of the FP32 fast path (see idw_kernels_fp32.cuh).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "idw_internal.h"
#include "idw_kernels_fp32.cuh"

namespace v2r {

struct FVariant {
    int mode, E;
    bool paired;
    int cr, cc, threads, minb;
    void (*fn)(const KParams);
};

#define V2R_FVARIANT(MODE, E, CR, CC, PAIRED, MINB) \
    FVariant { MODE, E, PAIRED, CR, CC, 256, MINB, idw_fp32_kernel<MODE, E, CR, CC, PAIRED, 256, MINB> }
#define V2R_FROW(MODE)                                                                                      \
    V2R_FVARIANT(MODE, 1, 4, 4, false, 2), V2R_FVARIANT(MODE, 2, 2, 4, false, 2),                           \
        V2R_FVARIANT(MODE, 3, 2, 2, false, 2), V2R_FVARIANT(MODE, 4, 2, 2, false, 2),                       \
        V2R_FVARIANT(MODE, 5, 1, 2, false, 2), V2R_FVARIANT(MODE, 6, 1, 2, false, 2),                       \
        V2R_FVARIANT(MODE, 7, 1, 2, false, 2), V2R_FVARIANT(MODE, 8, 1, 2, false, 2),                       \
        V2R_FVARIANT(MODE, 9, 1, 2, false, 2), V2R_FVARIANT(MODE, 10, 1, 2, false, 2)

static const FVariant kFVariants[] = {
    V2R_FVARIANT(MODE_RCP, 1, 4, 4, true, 1),  // p = 2: two points per MUFU.RCP; one CTA per SM measured 4.6 % faster than two
    V2R_FVARIANT(MODE_RCP, 1, 4, 4, true, 2),  // (V2R_IDW_FP32_MINB=2 selects the two-CTA build; tuning aid)
    V2R_FROW(MODE_RCP), V2R_FROW(MODE_RSQ), V2R_FROW(MODE_QRT), V2R_FROW(MODE_GEN),
};

static const FVariant *pick_fvariant(const Pass &ps) {
    static const int want_minb = getenv("V2R_IDW_FP32_MINB") ? atoi(getenv("V2R_IDW_FP32_MINB")) : 1;
    for (const FVariant &v : kFVariants) {
        if (v.paired && v.minb != want_minb) continue;
        if (v.mode != ps.mode || v.E != ps.count) continue;
        if (v.paired && !(ps.k0 == 1 && ps.count == 1)) continue;
        return &v;
    }
    return nullptr;
}

int launch_rows_fp32(cudaStream_t stream, const double *d_x, const double *d_y, const double *d_z,
                     const long long *d_kr, const long long *d_kc, long long n, const v2r_idw_grid &g,
                     long long row0, long long nrows, const double *exps, int n_exps, double *d_out) {
    std::vector<Pass> passes;
    int rc = plan_sweep(exps, n_exps, passes);
    if (rc) return rc;
    const long long plane = nrows * g.cols;
    for (const Pass &ps : passes) {
        const FVariant *v = pick_fvariant(ps);
        if (!v) return set_error(V2R_IDW_EINVAL, "no FP32 kernel variant for mode %d E %d", ps.mode, ps.count);
        KParams kp{};
        kp.x = d_x; kp.y = d_y; kp.w = d_z; kp.n = n;
        kp.x_min = g.x_min; kp.x_step = g.x_step; kp.y_min = g.y_min; kp.y_step = g.y_step;
        kp.cols = g.cols; kp.row0 = row0; kp.nrows = nrows;
        kp.out = d_out + (long long)ps.first * plane;
        kp.plane = plane;
        kp.k0 = ps.k0 > 0 ? ps.k0 : 1;
        kp.dk = ps.dk > 0 ? ps.dk : 1;
        for (int i = 0; i < ps.count; ++i) kp.cexp[i] = ps.cexp[i];
        const long long tile_r = (long long)(v->threads / 32) * v->cr, tile_c = 32LL * v->cc;
        const long long gx = (g.cols + tile_c - 1) / tile_c, gy = (nrows + tile_r - 1) / tile_r;
        if (gx > 2147483647LL || gy > 65535LL)
            return set_error(V2R_IDW_EINVAL, "band of %lld rows x %lld cols exceeds one launch; use smaller bands", nrows, (long long)g.cols);
        v->fn<<<dim3((unsigned)gx, (unsigned)gy), v->threads, 0, stream>>>(kp);
        count_launch();
        cudaError_t ce = cudaGetLastError();
        if (ce != cudaSuccess) return set_error(V2R_IDW_ECUDA, "idw fp32 kernel launch: %s", cudaGetErrorString(ce));
    }
    return launch_patch_exact(stream, d_z, d_kr, d_kc, n, g, row0, nrows, d_out, n_exps);
}

std::string describe_fp32(const std::vector<Pass> &passes) {
    std::string s;
    static const char *names[] = {"rcp", "rsqrt", "root4", "general"};
    for (const Pass &ps : passes) {
        const FVariant *v = pick_fvariant(ps);
        char tmp[160];
        snprintf(tmp, sizeof tmp, "%sfp32/%s%s E=%d cells=%dx%d k0=%d dk=%d", s.empty() ? "" : " + ", names[ps.mode],
                 v && v->paired ? " (paired rcp)" : "", ps.count, v ? v->cr : 0, v ? v->cc : 0, ps.k0, ps.dk);
        s += tmp;
    }
    return s;
}

}  // namespace v2r
