// Roofline denominators measured on the device the caller is on (include/mcd_b200.h):
//   mcd_measure_fp64_peak      dependent-free DFMA chains, FMA = 2 flop
//   mcd_measure_read_bandwidth streaming read of a buffer larger than L2
// MEASURED_PEAKS.json (driver-written) carries HBM copy bandwidth and bf16 tensor throughput only;
// the likelihood kernel is bound by the FP64 pipe, so its denominator is measured here, in the
// same process and under the same clocks as the kernel it is compared with.
#include <cstdio>

#include "mcd_internal.h"

namespace {

constexpr int kChains = 8;

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
    double x[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < kChains; ++k) x[k] = fma(x[k], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k];
    if (s == 123.456) out[0] = s;   // keeps the chains alive, practically never taken
}

__global__ void __launch_bounds__(256) read_kernel(const double2 *__restrict__ src, long long n2, double *out) {
    double s = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 v = __ldg(src + i);
        s += v.x + v.y;
    }
    if (s == 123.456) out[0] = s;
}

}  // namespace

extern "C" int mcd_measure_fp64_peak(int32_t device, double *tflops_out, double *ms_out) {
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
    double *out = nullptr;
    if (cudaMalloc(&out, 64) != cudaSuccess) return -2;
    const int blocks = prop.multiProcessorCount * 8;
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return -2; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    const double flops = 2.0 * kChains * 8.0 * iters * 256.0 * blocks;
    if (tflops_out) *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}

extern "C" int mcd_measure_read_bandwidth(int32_t device, int64_t bytes, double *gbs_out) {
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
    if (bytes < (1 << 20)) bytes = 1 << 20;
    bytes &= ~(int64_t)15;
    double2 *buf = nullptr;
    double *out = nullptr;
    if (cudaMalloc(&buf, (size_t)bytes) != cudaSuccess) return -2;
    if (cudaMalloc(&out, 64) != cudaSuccess) { cudaFree(buf); return -2; }
    cudaMemset(buf, 0, (size_t)bytes);
    const int blocks = prop.multiProcessorCount * 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        read_kernel<<<blocks, 256>>>(buf, bytes / 16, out);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); cudaFree(out); return -2; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(out);
    if (gbs_out) *gbs_out = (double)bytes / (best * 1e-3) / 1e9;
    return 0;
}
