// Background-population log-likelihood columns (the step before the hot path: evaluated once per
// model object, analysis/runner.py:102, model.py:562-563).
//
//   mcd_single_stars_lnlike : background/single_stars.py:42-77, log-mean-exp of M Gaussian kernels
//                             per star WITHOUT the M x N intermediate the reference materialises
//                             (single_stars.py:73).
//   mcd_gaussian_lnlike     : background/gaussian.py:23-28.
//
// single_stars_kernel, one pass over the M x N pairs:
//   * the reference shifts every exponent of a star by its maximum, max_j -(v_j - v_i)^2 / (2 norm_i)
//     (single_stars.py:74).  That maximum belongs to the background velocity NEAREST to v_i, so with the
//     background velocities sorted once on the host it is a binary search per star (exact, not an
//     estimate) instead of a first pass over all M;
//   * 1 / (2 norm_i) is taken once per star; exp(-w/2) comes from the table-driven exp of the mixture
//     kernels (mcd_math.cuh: 2^n T[j] (1 + q(r)), 10 FP64 instructions, 4e-17) instead of a libm call
//     per pair; 1 / sqrt(2 pi norm_i) leaves the sum (it is the same for all j);
//   * `lanes` threads share a star (a power of two <= 32, chosen by the host so that small catalogues
//     still fill the GPU) and add their partial sums with warp shuffles in a fixed order;
//   * the background velocities are staged in shared memory tile by tile and read as broadcasts.
// Per pair: 3 FP64 instructions for w, 10 for the exponential, 1 FMA to accumulate.
#include <algorithm>
#include <cmath>
#include <vector>

#include "mcd_internal.h"
#include "mcd_math.cuh"

namespace {

constexpr int kBgTile = 2048;   // background velocities per shared-memory tile (16 KB)
constexpr int kBgBlock = 256;

// index of the element of the ascending array `a[0..m)` nearest to x
__device__ __forceinline__ long long nearest_index(const double *__restrict__ a, long long m, double x) {
    long long lo = 0, hi = m;              // first element >= x
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(&a[mid]) < x) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return 0;
    if (lo == m) return m - 1;
    return (x - __ldg(&a[lo - 1]) <= __ldg(&a[lo]) - x) ? lo - 1 : lo;
}

__global__ void __launch_bounds__(kBgBlock) single_stars_kernel(const double *__restrict__ v_bg_sorted, long long m,
                                                                 const double *__restrict__ v,
                                                                 const double *__restrict__ verr, long long n,
                                                                 double sigma_int, int lanes, double *__restrict__ out) {
    __shared__ double tile[kBgTile];
    __shared__ double s_exp2[64];
    if (threadIdx.x < 64) s_exp2[threadIdx.x] = mcd::kExp2Table[threadIdx.x];
    const int stars_per_block = kBgBlock / lanes;
    const int lane = threadIdx.x % lanes;
    const long long i = (long long)blockIdx.x * stars_per_block + threadIdx.x / lanes;
    const bool live = i < n;
    const double vi = live ? v[i] : 0.0;
    const double ei = live ? verr[i] : 1.0;
    const double norm = sigma_int * sigma_int + ei * ei;          // single_stars.py:72
    // w_j = (d_j^2 - d_min^2) / norm >= 0 and the shifted exponent of single_stars.py:75 is -w_j / 2
    const double inv_norm = 1.0 / norm;
    double d_min = 0.0;
    if (live) d_min = __ldg(&v_bg_sorted[nearest_index(v_bg_sorted, m, vi)]) - vi;
    const double shift = -(d_min * d_min) * inv_norm;             // = 2 * exp_coeff_max (single_stars.py:74)

    double sum = 0.0;
    for (long long base = 0; base < m; base += kBgTile) {
        const int cnt = (int)min((long long)kBgTile, m - base);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += kBgBlock) tile[j] = v_bg_sorted[base + j];
        __syncthreads();
        for (int j = lane; j < cnt; j += lanes) {
            const double d = tile[j] - vi;
            const double w = fma(d * d, inv_norm, shift);
            double mant;
            int expo;
            mcd::exp_neg_half_table(w, s_exp2, mant, expo);
            // w >= 0 up to rounding, so expo <= 0 (+1 for the table's mantissa range); terms below 2^-1022
            // relative to the largest one (which is exactly 1) are dropped, as exp() underflowing does in
            // the reference
            sum = fma(mant, mcd::pow2_flush(min(expo, 0)), sum);
        }
    }
    // lanes of a star are adjacent threads of one warp: butterfly in a fixed order
    for (int off = lanes >> 1; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    // max + log(sum / sqrt(2 pi norm)) - log(M)                    (single_stars.py:75-76)
    if (live && lane == 0) out[i] = 0.5 * shift + log(sum) - 0.5 * log(mcd::kTwoPi * norm) - log((double)m);
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

// Runner._calculate_lnlike (analysis/runner.py:240-286): the log-likelihood of the catalogue for model curves
// v_los[N], sigma_los[N] that the CALLER computed (a user-defined model on top of Runner): Gaussian sum, or the
// max-shifted two-component mixture when the model carries a fixed background column.  Library log / exp /
// division as in the reference; per-thread grid-stride sums, a fixed-order block reduction, per-block partials
// added by one block -- the result does not depend on scheduling.
constexpr int kCurveBlock = 256;

__device__ __forceinline__ double block_sum(double x, double *red) {
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    double total = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kCurveBlock / 32; ++w) total += red[w];
    return total;            // valid in thread 0
}

__global__ void __launch_bounds__(kCurveBlock) curve_lnlike_kernel(
    const double *__restrict__ v, const double *__restrict__ verr, const double *__restrict__ pmember,
    const double *__restrict__ lbg, const double *__restrict__ v_los, const double *__restrict__ sigma_los, long long n,
    double *__restrict__ partial) {
    __shared__ double red[kCurveBlock / 32];
    double sum = 0.0;
    for (long long i = (long long)blockIdx.x * kCurveBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kCurveBlock) {
        const double norm = verr[i] * verr[i] + sigma_los[i] * sigma_los[i];      // runner.py:261
        const double d = v[i] - v_los[i];
        const double exponent = -0.5 * (d * d) / norm;                            // runner.py:262
        const double lm = -0.5 * log(mcd::kTwoPi * norm) + exponent;              // runner.py:269-270,280
        if (!lbg) {
            sum += lm;
        } else {
            const double lb = lbg[i], m = pmember[i];
            const double mx = fmax(lm, lb);                                       // runner.py:282
            sum += mx + log(m * exp(lm - mx) + (1.0 - m) * exp(lb - mx));         // runner.py:283-284
        }
    }
    const double total = block_sum(sum, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kCurveBlock) curve_sum_kernel(const double *__restrict__ partial, int count,
                                                                double *__restrict__ out) {
    __shared__ double red[kCurveBlock / 32];
    double sum = 0.0;
    for (int i = threadIdx.x; i < count; i += kCurveBlock) sum += partial[i];
    const double total = block_sum(sum, red);
    if (threadIdx.x == 0) out[0] = total;
}

// Stream-ordered scratch of one call: everything is allocated, used and freed on one stream, so a call
// costs no device-wide synchronisation and no synchronous cudaMalloc/cudaFree.
struct CallScratch {
    cudaStream_t stream = nullptr;
    void *p[4] = {nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    cudaError_t open() { return cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking); }
    cudaError_t alloc(double **out, size_t count) {
        const cudaError_t err = cudaMallocAsync(&p[n], sizeof(double) * std::max<size_t>(count, 1), stream);
        *out = static_cast<double *>(p[n]);
        if (err == cudaSuccess) ++n;
        return err;
    }
    ~CallScratch() {
        if (!stream) return;
        for (int k = 0; k < n; ++k) cudaFreeAsync(p[k], stream);
        cudaStreamSynchronize(stream);
        cudaStreamDestroy(stream);
    }
};

#define BG_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t err__ = (call);                                                                       \
        if (err__ != cudaSuccess) return mcd::set_error(-2, "%s failed: %s", #call, cudaGetErrorString(err__)); \
    } while (0)

// threads per star: enough that the grid fills the GPU even for a radial bin of a few hundred stars
int lanes_for(long long n, int sm_count) {
    const long long want = (long long)sm_count * 4 * kBgBlock;
    int lanes = 1;
    while (lanes < 32 && n * lanes < want) lanes <<= 1;
    return lanes;
}

}  // namespace

cudaError_t mcd::launch_single_stars(const double *v_bg_sorted_dev, long long m, const double *v_dev, const double *verr_dev,
                                     long long n, double sigma_int, double *out_dev, int sm_count, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int lanes = lanes_for(n, sm_count);
    const int per_block = kBgBlock / lanes;
    const long long grid = (n + per_block - 1) / per_block;
    single_stars_kernel<<<(unsigned)grid, kBgBlock, 0, stream>>>(v_bg_sorted_dev, m, v_dev, verr_dev, n, sigma_int, lanes,
                                                                 out_dev);
    return cudaGetLastError();
}

// `scratch_dev` holds curve_lnlike_blocks(n, sm_count) doubles; `out_dev` one
int mcd::curve_lnlike_blocks(long long n, int sm_count) {
    const long long want = (n + kCurveBlock - 1) / kCurveBlock;
    return (int)std::max<long long>(1, std::min<long long>(want, (long long)std::max(1, sm_count) * 8));
}

cudaError_t mcd::launch_curve_lnlike(const double *v_dev, const double *verr_dev, const double *pmember_dev,
                                     const double *lbg_dev, const double *v_los_dev, const double *sigma_los_dev, long long n,
                                     double *scratch_dev, double *out_dev, int sm_count, cudaStream_t stream) {
    const int blocks = curve_lnlike_blocks(n, sm_count);
    curve_lnlike_kernel<<<blocks, kCurveBlock, 0, stream>>>(v_dev, verr_dev, pmember_dev, lbg_dev, v_los_dev, sigma_los_dev,
                                                            n, scratch_dev);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    curve_sum_kernel<<<1, kCurveBlock, 0, stream>>>(scratch_dev, blocks, out_dev);
    return cudaGetLastError();
}

static int device_sm_count(int device, int *count) {
    BG_CUDA(cudaDeviceGetAttribute(count, cudaDevAttrMultiProcessorCount, device));
    return 0;
}

extern "C" int mcd_single_stars_lnlike(int32_t device, const double *v_bg, int64_t m, const double *v,
                                       const double *verr, int64_t n, double sigma_int, double *out_host) {
    if (m <= 0 || n < 0 || !v_bg || (n > 0 && (!v || !verr || !out_host)))
        return mcd::set_error(-1, "mcd_single_stars_lnlike: bad argument (m = %lld, n = %lld)", (long long)m, (long long)n);
    if (n == 0) return 0;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return mcd::set_error(-3, "no CUDA device is visible: the B200 path has no CPU fallback");
    BG_CUDA(cudaSetDevice(device));
    int sms = 0;
    if (int rc = device_sm_count(device, &sms)) return rc;
    // ascending background velocities: the nearest one to a star is then a binary search away
    std::vector<double> sorted(v_bg, v_bg + m);
    std::sort(sorted.begin(), sorted.end());
    CallScratch s;
    BG_CUDA(s.open());
    double *bg_dev, *v_dev, *e_dev, *out_dev;
    BG_CUDA(s.alloc(&bg_dev, (size_t)m));
    BG_CUDA(s.alloc(&v_dev, (size_t)n));
    BG_CUDA(s.alloc(&e_dev, (size_t)n));
    BG_CUDA(s.alloc(&out_dev, (size_t)n));
    BG_CUDA(cudaMemcpyAsync(bg_dev, sorted.data(), sizeof(double) * m, cudaMemcpyHostToDevice, s.stream));
    BG_CUDA(cudaMemcpyAsync(v_dev, v, sizeof(double) * n, cudaMemcpyHostToDevice, s.stream));
    BG_CUDA(cudaMemcpyAsync(e_dev, verr, sizeof(double) * n, cudaMemcpyHostToDevice, s.stream));
    BG_CUDA(mcd::launch_single_stars(bg_dev, m, v_dev, e_dev, n, sigma_int, out_dev, sms, s.stream));
    BG_CUDA(cudaMemcpyAsync(out_host, out_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, s.stream));
    BG_CUDA(cudaStreamSynchronize(s.stream));
    return 0;
}

extern "C" int mcd_single_stars_lnlike_device(int32_t device, const double *v_bg_sorted_dev, int64_t m, const double *v_dev,
                                              const double *verr_dev, int64_t n, double sigma_int, double *out_dev,
                                              void *stream) {
    if (m <= 0 || n < 0 || !v_bg_sorted_dev || (n > 0 && (!v_dev || !verr_dev || !out_dev)))
        return mcd::set_error(-1, "mcd_single_stars_lnlike_device: bad argument");
    if (n == 0) return 0;
    BG_CUDA(cudaSetDevice(device));
    int sms = 0;
    if (int rc = device_sm_count(device, &sms)) return rc;
    BG_CUDA(mcd::launch_single_stars(v_bg_sorted_dev, m, v_dev, verr_dev, n, sigma_int, out_dev, sms,
                                     static_cast<cudaStream_t>(stream)));
    return 0;
}

extern "C" int mcd_gaussian_lnlike(int32_t device, const double *v, const double *verr, int64_t n, double mean,
                                   double sigma, double *out_host) {
    if (n < 0 || (n > 0 && (!v || !verr || !out_host)))
        return mcd::set_error(-1, "mcd_gaussian_lnlike: bad argument (n = %lld)", (long long)n);
    if (n == 0) return 0;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return mcd::set_error(-3, "no CUDA device is visible: the B200 path has no CPU fallback");
    BG_CUDA(cudaSetDevice(device));
    CallScratch s;
    BG_CUDA(s.open());
    double *v_dev, *e_dev, *out_dev;
    BG_CUDA(s.alloc(&v_dev, (size_t)n));
    BG_CUDA(s.alloc(&e_dev, (size_t)n));
    BG_CUDA(s.alloc(&out_dev, (size_t)n));
    BG_CUDA(cudaMemcpyAsync(v_dev, v, sizeof(double) * n, cudaMemcpyHostToDevice, s.stream));
    BG_CUDA(cudaMemcpyAsync(e_dev, verr, sizeof(double) * n, cudaMemcpyHostToDevice, s.stream));
    const long long grid = (n + 255) / 256;
    gaussian_kernel<<<(unsigned)grid, 256, 0, s.stream>>>(v_dev, e_dev, n, mean, sigma, out_dev);
    BG_CUDA(cudaGetLastError());
    BG_CUDA(cudaMemcpyAsync(out_host, out_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, s.stream));
    BG_CUDA(cudaStreamSynchronize(s.stream));
    return 0;
}
