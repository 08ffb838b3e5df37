// FP64 pipe throughput versus register-operand pattern on sm_100a.
//   A  x = fma(x, a, b)          a, b shared by every chain (operand reuse cache hits)
//   B  x_j = fma(x_j, y_j, z_j)  three distinct register pairs per instruction, 8 chains
//   C  x_j = fma(x_j, y_j, z_k)  rotating third operand
//   D  x_j = x_j * y_j           DMUL, two distinct operands
//   E  x_j = x_j + y_j           DADD
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        x[j] = threadIdx.x * 1e-3 + j;
        y[j] = 0.999 + 1e-6 * (threadIdx.x + j) * a;
        z[j] = 1e-7 * (threadIdx.x + 3 * j) * b;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (MODE == 0) x[j] = fma(x[j], a, b);
                if (MODE == 1) x[j] = fma(x[j], y[j], z[j]);
                if (MODE == 2) x[j] = fma(x[j], y[j], z[(j + u) & 7]);
                if (MODE == 3) x[j] = x[j] * y[j];
                if (MODE == 4) x[j] = x[j] + z[j];
                if (MODE == 5) x[j] = fma(y[j], z[j], x[j]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char *name, int sms, int blocks_per_sm) {
    double *out;
    cudaMalloc(&out, 64);
    const int iters = 2048, blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double per_smsp = (double)iters * 32 * 8 * blocks_per_sm / 4.0;
    printf("%-44s %d CTA/SM  %.3f cycles per warp-instr per SMSP\n", name, blocks_per_sm, best * 1e-3 * clk * 1e3 / per_smsp);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int s = p.multiProcessorCount;
    for (int b : {2, 4, 8}) {
        run<0>("A fma(x, a, b) shared operands", s, b);
        run<1>("B fma(x_j, y_j, z_j) 3 distinct", s, b);
        run<2>("C fma(x_j, y_j, z_k) rotating", s, b);
        run<5>("F fma(y_j, z_j, x_j) accumulate", s, b);
        run<3>("D x_j * y_j", s, b);
        run<4>("E x_j + z_j", s, b);
    }
    return 0;
}
