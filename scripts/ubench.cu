// ubench.cu -- issue-rate microbenchmarks that decide kernel design questions on B200 (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
// Every test: SMs*8 CTAs x 256 threads, 8 independent chains per thread, result in G thread-instructions/s of the
// instruction named first (mixes also print the total).  Not part of the product; results are quoted in DESIGN.md.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8

__device__ __forceinline__ float frcp(float a) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

template <int WHICH>
__global__ void __launch_bounds__(256) k(double *out, int iters, int salt) {
    const int tid = threadIdx.x;
    if constexpr (WHICH == 0) {  // FFMA, 3 register sources
        float a[CHAINS], m = 0.999f + 1e-6f * tid, c = 1e-7f * (tid + salt);
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) a[j] = 1.f + j + tid;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) a[j] = fmaf(a[j], m, c);
        float t = 0;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j];
        if (t == 12345.f) out[0] = t;
    } else if constexpr (WHICH == 1) {  // FFMA2 (fma.rn.f32x2)
        float2 a[CHAINS], m = make_float2(0.999f + 1e-6f * tid, 0.998f + 1e-6f * tid), c = make_float2(1e-7f * (tid + salt), 2e-7f * tid);
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) a[j] = make_float2(1.f + j + tid, 2.f + j);
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) a[j] = __ffma2_rn(a[j], m, c);
        float t = 0;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j].x + a[j].y;
        if (t == 12345.f) out[0] = t;
    } else if constexpr (WHICH == 2) {  // FADD2
        float2 a[CHAINS], c = make_float2(1e-7f * (tid + salt), 2e-7f * tid);
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) a[j] = make_float2(1.f + j + tid, 2.f + j);
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) a[j] = __fadd2_rn(a[j], c);
        float t = 0;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j].x + a[j].y;
        if (t == 12345.f) out[0] = t;
    } else if constexpr (WHICH == 3) {  // FFMA2 : MUFU.RCP = 4 : 1 (per 8 packed FMAs, 2 rcp)
        float2 a[CHAINS], m = make_float2(0.999f + 1e-6f * tid, 0.998f + 1e-6f * tid), c = make_float2(1e-7f * (tid + salt), 2e-7f * tid);
        float r0 = 1.5f + tid, r1 = 2.5f + tid;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) a[j] = make_float2(1.f + j + tid, 2.f + j);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) a[j] = __ffma2_rn(a[j], m, c);
            r0 = frcp(r0) + 1.f;
            r1 = frcp(r1) + 1.f;
        }
        float t = r0 + r1;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j].x + a[j].y;
        if (t == 12345.f) out[0] = t;
    } else if constexpr (WHICH == 4) {  // DFMA alone, 2 reg sources + 1 from the chain
        double a[CHAINS], m = 0.999 + 1e-9 * tid, c = 1e-7 * (tid + salt);
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) a[j] = 1.0 + j + tid;
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) a[j] = fma(a[j], m, c);
        double t = 0;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j];
        if (t == 12345.0) out[0] = t;
    } else if constexpr (WHICH == 5) {  // I2F.F64.S32 alone
        int v[CHAINS];
        double a[CHAINS];
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) { v[j] = tid + j + salt; a[j] = 0; }
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) { a[j] = __int2double_rn(v[j]); v[j] = __double2loint(a[j]) ^ i; }
        double t = 0;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j];
        if (t == 12345.0) out[0] = t;
    } else if constexpr (WHICH == 6) {  // DFMA : I2F = 8 : 1
        double a[CHAINS], m = 0.999 + 1e-9 * tid, c = 1e-7 * (tid + salt);
        int v = tid + salt;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) a[j] = 1.0 + j + tid;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) a[j] = fma(a[j], m, c);
            double d = __int2double_rn(v);
            v = __double2loint(d) ^ i;
        }
        double t = v;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j];
        if (t == 12345.0) out[0] = t;
    } else if constexpr (WHICH == 7 || WHICH == 8 || WHICH == 12) {  // DFMA : integer ALU = 1 : 1 (7), 1 : 2 (8), 2 : 3 (12)
        double a[CHAINS], m = 0.999 + 1e-9 * tid, c = 1e-7 * (tid + salt);
        unsigned v[CHAINS];
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) { a[j] = 1.0 + j + tid; v[j] = tid * 7 + j + salt; }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) {
                a[j] = fma(a[j], m, c);
                v[j] = (v[j] >> 3) ^ (v[j] + 0x9e3779b9u);             // SHF + IADD/LOP3: ~2 ALU ops
                if constexpr (WHICH == 8) v[j] = (v[j] << 5) ^ (v[j] + 0x7f4a7c15u);
                if constexpr (WHICH == 12) { if (j & 1) v[j] = (v[j] << 5) ^ (v[j] + 0x7f4a7c15u); }
            }
        }
        double t = 0;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j] + v[j];
        if (t == 12345.0) out[0] = t;
    } else if constexpr (WHICH == 9) {  // F2I.S32.F64 alone
        double a[CHAINS];
        int v[CHAINS];
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) { a[j] = 1.5 + j + tid + salt; v[j] = 0; }
        for (int i = 0; i < iters; ++i)
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) { v[j] = __double2int_rn(a[j]); a[j] = __hiloint2double(__double2hiint(a[j]), v[j] ^ i); }
        double t = 0;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j] + v[j];
        if (t == 12345.0) out[0] = t;
    } else if constexpr (WHICH == 10) {  // DFMA three fresh register sources
        double a[CHAINS], b[CHAINS], c[CHAINS];
        const double rnd = 1e-12 * (double)((clock64() + tid) & 1023);
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) { a[j] = 1.0 + j; b[j] = 0.999999 + rnd * (j + 1); c[j] = 1.000001 - rnd * (2 * j + 1); }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) a[j] = fma(b[j], c[(j + 1) & 7], a[j]);
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) b[j] = fma(a[j], c[j], b[(j + 3) & 7]);
        }
        double t = 0;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j] + b[j];
        if (t == 12345.0) out[0] = t;
    } else if constexpr (WHICH == 11) {  // DFMA acc[j] = fma(x[j], r, acc[j]) with ONE shared r per group of 8 (operand reuse candidates)
        double a[CHAINS], x[CHAINS];
        const double rnd = 1e-12 * (double)((clock64() + tid) & 1023);
        double r = 0.999 + rnd;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) { a[j] = 1.0 + j; x[j] = 0.5 + rnd * (j + 1); }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) a[j] = fma(x[j], r, a[j]);
            r = r + rnd;  // 1 DADD per 8 DFMA
        }
        double t = 0;
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) t += a[j];
        if (t == 12345.0) out[0] = t;
    }
}

// LDS.128 / LDS.64 table look-ups with lane-dependent indices (general-class tables)
template <int BYTES>
__global__ void __launch_bounds__(256) k_lds(double *out, int iters, int stride, int entries_mask) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < (entries_mask + 1) * BYTES / 8; i += 256) ((double *)sm)[i] = i * 1e-3;
    __syncthreads();
    unsigned idx = (threadIdx.x & 31) * stride + (threadIdx.x >> 5) * 37;
    double acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned e = (idx + j * 131 + i * 7) & entries_mask;
            if constexpr (BYTES == 16) {
                const double2 v = *reinterpret_cast<const double2 *>(sm + e * 16);
                acc += v.x * v.y;
            } else {
                acc += *reinterpret_cast<const double *>(sm + e * 8);
            }
        }
    }
    if (acc == 12345.0) out[0] = acc;
}

template <int WHICH>
static double run(int sms, int per_iter_first, const char *name, double ms_target = 30.0) {
    double *d;
    cudaMalloc(&d, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    int iters = 2000;
    double best = 0;
    const int blocks = sms * 8;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k<WHICH><<<blocks, 256>>>(d, iters, rep);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        double rate = (double)blocks * 256 * iters * per_iter_first / (ms * 1e-3) / 1e9;
        if (rep > 0 && rate > best) best = rate;
        if (ms < ms_target) iters = (int)(iters * ms_target / (ms > 0.01f ? ms : 0.01f));
    }
    printf("%-58s %10.1f G inst/s\n", name, best);
    cudaFree(d);
    return best;
}

template <int BYTES>
static void run_lds(int sms, int stride, int entries, const char *name) {
    double *d;
    cudaMalloc(&d, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    int iters = 4000;
    const int blocks = sms * 4;
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        k_lds<BYTES><<<blocks, 256, entries * BYTES>>>(d, iters, stride, entries - 1);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        double rate = (double)blocks * 256 * iters * 8 / (ms * 1e-3) / 1e9;
        if (rate > best) best = rate;
    }
    printf("%-58s %10.1f G thread-loads/s  (%.2f warp-loads/clk/SM at 1.965 GHz)\n", name, best, best * 1e9 / 32 / sms / 1.965e9);
    cudaFree(d);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    run<0>(sms, 8, "FFMA (3 reg sources)");
    run<1>(sms, 8, "FFMA2 packed (instructions; x2 = flops/2)");
    run<2>(sms, 8, "FADD2 packed (instructions)");
    run<3>(sms, 8, "FFMA2 in a 8 FFMA2 : 2 MUFU.RCP : 2 FADD mix (FFMA2 rate)");
    run<4>(sms, 8, "DFMA (chain + 2 reg)");
    run<5>(sms, 8, "I2F.F64.S32 (+ 1 LOP3 each)");
    run<9>(sms, 8, "F2I.S32.F64 (+ PRMT/LOP3 each)");
    run<6>(sms, 8, "DFMA in a 8 DFMA : 1 I2F mix (DFMA rate)");
    run<7>(sms, 8, "DFMA with ~2 integer ALU ops per DFMA (DFMA rate)");
    run<12>(sms, 8, "DFMA with ~3 integer ALU ops per DFMA (DFMA rate)");
    run<8>(sms, 8, "DFMA with ~4 integer ALU ops per DFMA (DFMA rate)");
    run<10>(sms, 16, "DFMA three fresh register-pair sources");
    run<11>(sms, 8, "DFMA acc = fma(x_j, r, acc_j), r shared by 8 (reuse)");
    run_lds<16>(sms, 1, 1024, "LDS.128 1024-entry table, lanes consecutive");
    run_lds<16>(sms, 3, 1024, "LDS.128 1024-entry table, lane stride 3");
    run_lds<16>(sms, 21, 1024, "LDS.128 1024-entry table, lane stride 21");
    run_lds<16>(sms, 0, 1024, "LDS.128 1024-entry table, all lanes same (broadcast)");
    run_lds<16>(sms, 1, 128, "LDS.128 128-entry table, lanes consecutive");
    run_lds<8>(sms, 1, 1024, "LDS.64 1024-entry table, lanes consecutive");
    run_lds<8>(sms, 3, 1024, "LDS.64 1024-entry table, lane stride 3");
    run_lds<8>(sms, 21, 1024, "LDS.64 1024-entry table, lane stride 21");
    return 0;
}
