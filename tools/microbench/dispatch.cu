// Issue/dispatch microbenchmark for the FP64 pipe of sm_100a: does a non-FP64 instruction issued
// between DFMAs cost FP64 throughput?  Each variant runs 8 independent DFMA chains per thread plus
// a configurable number of other instructions per 8 DFMAs.  Prints cycles per DFMA per SMSP.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dispatch dispatch.cu && ./dispatch
#include <cstdio>
#include <cuda_runtime.h>

template <int IADD, int FFMA, int MUFU, int LDS>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b, int ia, float fa) {
    __shared__ double2 sm[256];
    sm[threadIdx.x] = make_double2(a, b);
    __syncthreads();
    double x[8];
    int n[8];
    float f[8];
    double m[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x * 1e-3 + j; n[j] = threadIdx.x + j; f[j] = threadIdx.x * 1e-3f + j; }
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = 1.0 + threadIdx.x + j;
    double2 acc = make_double2(0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                x[j] = fma(x[j], a, b);
                if (j < IADD) n[j] = n[j] * 3 + ia;          // IMAD/IADD-class
                if (j < FFMA) f[j] = fmaf(f[j], fa, 1.0f);
                if (j < MUFU) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(m[j & 3]));
                if (j < LDS) { double2 v = sm[(threadIdx.x + j + n[0]) & 255]; acc.x += 0; asm volatile("" :: "d"(v.x), "d"(v.y)); }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + n[j] + f[j];
    for (int j = 0; j < 4; ++j) s += m[j];
    if (s == 123.456) out[0] = s + acc.x;
}

template <int IADD, int FFMA, int MUFU, int LDS>
void run(const char *name, int sms) {
    double *out;
    cudaMalloc(&out, 64);
    const int iters = 2048, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<IADD, FFMA, MUFU, LDS><<<blocks, 256>>>(out, iters, 0.999999, 1e-7, 7, 0.9999f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // warp-level DFMAs per SMSP
    const double dfma_per_smsp = (double)iters * 32 * 8 /*warps per block*/ * 8 /*blocks per SM*/ / 4.0;
    const double cycles = best * 1e-3 * clk * 1e3;
    printf("%-34s %8.3f ms   %.3f cycles per warp-DFMA per SMSP   (%.1f TFLOP/s)\n", name, best, cycles / dfma_per_smsp,
           2.0 * iters * 32 * 256.0 * blocks / (best * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int s = p.multiProcessorCount;
    run<0, 0, 0, 0>("8 DFMA", s);
    run<2, 0, 0, 0>("8 DFMA + 2 IMAD", s);
    run<4, 0, 0, 0>("8 DFMA + 4 IMAD", s);
    run<8, 0, 0, 0>("8 DFMA + 8 IMAD", s);
    run<0, 4, 0, 0>("8 DFMA + 4 FFMA", s);
    run<0, 8, 0, 0>("8 DFMA + 8 FFMA", s);
    run<0, 0, 1, 0>("8 DFMA + 1 MUFU.RCP64H", s);
    run<0, 0, 2, 0>("8 DFMA + 2 MUFU.RCP64H", s);
    run<0, 0, 0, 2>("8 DFMA + 2 LDS.128", s);
    run<0, 0, 0, 4>("8 DFMA + 4 LDS.128", s);
    run<4, 0, 1, 2>("8 DFMA + 4 IMAD + 1 MUFU + 2 LDS", s);
    return 0;
}
