// Do FP64 tensor-core MMAs (mma.sync m8n8k4 f64, SASS DMMA) run beside the vector FP64 pipe on sm_100a, or on it?
// Three kernels with the same structure: DFMA chains only, DMMA chains only, both interleaved.  If the mixed
// kernel takes max(t_dfma, t_dmma) the two are separate pipes and the linear forms of the likelihood term
// (num, D1, D2 = [x y r2 1] * walker constants, K = 4) could move to the tensor pipe; if it takes the sum they
// share the FP64 datapath.  Prints cycles per warp instruction per SM sub-partition.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NFMA, int NMMA>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b) {
    double x[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x * 1e-3 + j; c[j] = j; }
    const double fa = a + threadIdx.x * 1e-9, fb = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < NFMA) x[j] = fma(x[j], a, b);
                if (j < NMMA) dmma(c[2 * (j & 3)], c[2 * (j & 3) + 1], fa, fb);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + c[j];
    if (s == 123.456) out[0] = s;
}

template <int NFMA, int NMMA>
float run(const char *name, int sms, int clk_khz) {
    double *out;
    cudaMalloc(&out, 64);
    const int iters = 2048, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<NFMA, NMMA><<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    // warp instructions of each kind per SM sub-partition: 8 CTAs x 8 warps / 4 sub-partitions = 16 warps
    const double cycles = best * 1e-3 * clk_khz * 1e3;
    const double per_smsp = 16.0 * iters * 4.0;
    printf("%-34s %8.3f ms | %6.2f cycles per DFMA", name, best, NFMA ? cycles / (per_smsp * NFMA) : 0.0);
    printf(" | %6.2f cycles per DMMA\n", NMMA ? cycles / (per_smsp * NMMA) : 0.0);
    cudaFree(out);
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, clk);
    const float a = run<8, 0>("8 DFMA", p.multiProcessorCount, clk);
    const float b4 = run<0, 4>("4 DMMA m8n8k4", p.multiProcessorCount, clk);
    const float b8 = run<0, 8>("8 DMMA m8n8k4", p.multiProcessorCount, clk);
    const float m84 = run<8, 4>("8 DFMA + 4 DMMA interleaved", p.multiProcessorCount, clk);
    const float m88 = run<8, 8>("8 DFMA + 8 DMMA interleaved", p.multiProcessorCount, clk);
    const float m82 = run<8, 2>("8 DFMA + 2 DMMA interleaved", p.multiProcessorCount, clk);
    printf("mixed 8+4: %.3f ms vs sum %.3f, max %.3f\n", m84, a + b4, a > b4 ? a : b4);
    printf("mixed 8+8: %.3f ms vs sum %.3f, max %.3f\n", m88, a + b8, a > b8 ? a : b8);
    (void)m82;
    return 0;
}
