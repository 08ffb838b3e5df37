// Accuracy of the MUFU.RCP64H / MUFU.RSQ64H seeds (rcp.approx.ftz.f64, rsqrt.approx.ftz.f64) and of
// the refined values, over mantissas in [1, 4) and a few exponents.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#define MCD_NEWTON NEWTON
#include "../../mcmc_dynamics_b200/csrc/mcd_math.cuh"

__global__ void k(double *out, int n) {
    double e_rcp = 0, e_rsq = 0, f_rcp = 0, f_rsq = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        for (int ex = -2; ex <= 2; ++ex) {
            const double x = ldexp(1.0 + 3.0 * (i + 0.37) / n, 37 * ex);
            const double r = 1.0 / x, q = 1.0 / sqrt(x);
            e_rcp = fmax(e_rcp, fabs(mcd::rcp_seed(x) - r) / r);
            e_rsq = fmax(e_rsq, fabs(mcd::rsqrt_seed(x) - q) / q);
            f_rcp = fmax(f_rcp, fabs(mcd::fast_rcp(x) - r) / r);
            f_rsq = fmax(f_rsq, fabs(mcd::fast_rsqrt(x) - q) / q);
        }
    }
    // crude max over threads
    atomicMax((unsigned long long *)&out[0], __double_as_longlong(e_rcp));
    atomicMax((unsigned long long *)&out[1], __double_as_longlong(e_rsq));
    atomicMax((unsigned long long *)&out[2], __double_as_longlong(f_rcp));
    atomicMax((unsigned long long *)&out[3], __double_as_longlong(f_rsq));
}

int main() {
    double *d, h[4];
    cudaMalloc(&d, 32);
    cudaMemset(d, 0, 32);
    k<<<592, 256>>>(d, 1 << 26);
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("NEWTON=%d  seed rcp %.3e (2^%.1f)  seed rsqrt %.3e (2^%.1f)  refined rcp %.3e  refined rsqrt %.3e\n", NEWTON, h[0],
           log2(h[0]), h[1], log2(h[1]), h[2], h[3]);
    return 0;
}
