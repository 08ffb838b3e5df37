// Counter-based random numbers for the device-resident ensemble sampler: Philox4x32-10
// (Salmon et al. 2011).  counter = (step, half, walker, purpose), key = seed.  Shared by the
// sampler's own kernels (mcd_sampler.cu) and the likelihood kernel's fused proposal/acceptance.
#pragma once
#include <stdint.h>

namespace mcd {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// two uniforms in [0, 1) with 53 random bits each
__device__ __forceinline__ void uniforms(unsigned long long seed, uint32_t step, uint32_t half, uint32_t walker,
                                         uint32_t purpose, double &u0, double &u1) {
    const uint4 r = philox4x32_10(make_uint4(step, half, walker, purpose), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const unsigned long long a = ((unsigned long long)r.x << 32) | r.y, b = ((unsigned long long)r.z << 32) | r.w;
    u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
    u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace mcd
