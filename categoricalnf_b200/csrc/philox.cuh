// Counter-based RNG (Philox4x32-10) and the logistic prior sampler.
// Replaces torch.distributions.Uniform.sample + LogisticDistribution.shift_x
// (reference layers/flows/distributions.py:117-127, 139-145) inside the encode kernel.
#pragma once
#include <stdint.h>

namespace cnf {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

// Four uniforms in [0, 1) (24-bit mantissas) for counter `ctr` under key `seed`.
__device__ __forceinline__ void philox_uniform4(unsigned long long seed, unsigned long long ctr, float (&u)[4]) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x1BD11BDAu, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = (float)(c[j] >> 8) * (1.0f / 16777216.0f);
}

// z0 / sigma for one uniform draw: squeeze into (eps/2, 1 - eps/2) in float32 exactly like the
// reference (two rounded float32 ops), then the logit in float64 (distributions.py:121-123).
__device__ __forceinline__ float logistic_from_uniform(float u, float eps) {
    const float v = __fadd_rn(__fmul_rn(u, 1.0f - eps), 0.5f * eps);
    return (float)(-log(1.0 / (double)v - 1.0));
}

}  // namespace cnf
