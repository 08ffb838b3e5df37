// Background-population log-likelihood columns (the step before the hot path: evaluated once per
// model object, analysis/runner.py:102, model.py:562-563).
//
//   mcd_single_stars_lnlike : background/single_stars.py:42-77, log-mean-exp of M Gaussian kernels
//                             per star WITHOUT the M x N intermediate the reference materialises
//                             (single_stars.py:73); thread = star, background velocities broadcast
//                             from shared memory, two passes (max, then sum) like the reference.
//   mcd_gaussian_lnlike     : background/gaussian.py:23-28.
#include <algorithm>
#include <cmath>

#include "mcd_internal.h"
#include "mcd_math.cuh"

namespace {

constexpr int kBgTile = 1024;   // background velocities per shared-memory tile
constexpr int kBgBlock = 128;

__global__ void __launch_bounds__(kBgBlock) single_stars_kernel(const double *__restrict__ v_bg, long long m,
                                                                 const double *__restrict__ v,
                                                                 const double *__restrict__ verr, long long n,
                                                                 double sigma_int, double *__restrict__ out) {
    __shared__ double tile[kBgTile];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const double vi = live ? v[i] : 0.0;
    const double ei = live ? verr[i] : 1.0;
    const double norm = sigma_int * sigma_int + ei * ei;          // single_stars.py:72
    const double two_norm = 2. * norm;

    // pass 1: exp_coeff_max = max_j -(v_j - v_i)^2 / (2 norm)        (single_stars.py:73-74)
    double best = -INFINITY;
    for (long long base = 0; base < m; base += kBgTile) {
        const int cnt = (int)min((long long)kBgTile, m - base);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) tile[j] = v_bg[base + j];
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const double d = tile[j] - vi;
            best = fmax(best, -(d * d) / two_norm);
        }
    }
    // pass 2: sum_j exp(exp_coeff - max) / sqrt(2 pi norm)            (single_stars.py:75-76)
    const double root = sqrt(2. * M_PI * norm);
    double sum = 0.0;
    for (long long base = 0; base < m; base += kBgTile) {
        const int cnt = (int)min((long long)kBgTile, m - base);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) tile[j] = v_bg[base + j];
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const double d = tile[j] - vi;
            sum += exp(-(d * d) / two_norm - best) / root;
        }
    }
    if (live) out[i] = best + log(sum) - log((double)m);
}

__global__ void gaussian_kernel(const double *__restrict__ v, const double *__restrict__ verr, long long n, double mean,
                                double sigma, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double norm = verr[i] * verr[i] + sigma * sigma;            // gaussian.py:25
    const double d = v[i] - mean;
    const double exponent = -0.5 * (d * d) / norm;                    // gaussian.py:26
    out[i] = -0.5 * log(2. * M_PI * norm) + exponent;                 // gaussian.py:28
}

struct DeviceBuffers {
    double *p[4] = {nullptr, nullptr, nullptr, nullptr};
    ~DeviceBuffers() {
        for (auto q : p) cudaFree(q);
    }
};

}  // namespace

extern "C" int mcd_single_stars_lnlike(int32_t device, const double *v_bg, int64_t m, const double *v,
                                       const double *verr, int64_t n, double sigma_int, double *out_host) {
    if (m <= 0 || n < 0 || !v_bg || (n > 0 && (!v || !verr || !out_host))) return -1;
    if (n == 0) return 0;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    DeviceBuffers b;
    if (cudaMalloc(&b.p[0], sizeof(double) * m) != cudaSuccess) return -2;
    for (int k = 1; k < 4; ++k)
        if (cudaMalloc(&b.p[k], sizeof(double) * n) != cudaSuccess) return -2;
    if (cudaMemcpy(b.p[0], v_bg, sizeof(double) * m, cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    if (cudaMemcpy(b.p[1], v, sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    if (cudaMemcpy(b.p[2], verr, sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    const long long grid = (n + kBgBlock - 1) / kBgBlock;
    single_stars_kernel<<<(unsigned)grid, kBgBlock>>>(b.p[0], m, b.p[1], b.p[2], n, sigma_int, b.p[3]);
    if (cudaGetLastError() != cudaSuccess) return -2;
    if (cudaMemcpy(out_host, b.p[3], sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    return 0;
}

extern "C" int mcd_gaussian_lnlike(int32_t device, const double *v, const double *verr, int64_t n, double mean,
                                   double sigma, double *out_host) {
    if (n < 0 || (n > 0 && (!v || !verr || !out_host))) return -1;
    if (n == 0) return 0;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    DeviceBuffers b;
    for (int k = 0; k < 3; ++k)
        if (cudaMalloc(&b.p[k], sizeof(double) * n) != cudaSuccess) return -2;
    if (cudaMemcpy(b.p[0], v, sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    if (cudaMemcpy(b.p[1], verr, sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    const long long grid = (n + 255) / 256;
    gaussian_kernel<<<(unsigned)grid, 256>>>(b.p[0], b.p[1], n, mean, sigma, b.p[2]);
    if (cudaGetLastError() != cudaSuccess) return -2;
    if (cudaMemcpy(out_host, b.p[2], sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    return 0;
}
