// Device random stream: Philox4x32-10 (Salmon et al., SC'11), counter based.
//
// The reference draws from one process-global std::mt19937 (src/random.hpp:677-680), which is
// inherently sequential.  The engine's free-running stream is keyed instead:
//   key     = 64-bit seed
//   counter = (slot, unit, generation lo, generation hi)
// so a draw depends only on WHAT it is for (generation, candidate, coordinate pair / tag), never
// on which GPU or thread produced it -- candidate sharding across GPUs does not change the draws.
// Parity with the reference goes through the draw-replay hooks instead (bbgpu_cma_sample h_z,
// bbgpu_de_generation draws), not through this stream.
#pragma once

#include "common.cuh"

namespace bb {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t) (((uint64_t) a * (uint64_t) b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
        uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4 { c0, c1, c2, c3 };
}

// (double) w, exactly, without the conversion unit: the integer is dropped into the mantissa of 2^52 and 2^52 is
// subtracted (one DADD).  I2F.F64 issues on the quarter-rate XU pipe, which ncu showed saturated (134 %) in the
// DE build and the fused Sep sampling kernel when every coordinate converted its uniform draw there.
__host__ __device__ __forceinline__ double u32_to_double(uint32_t w) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(0x43300000, (int) w) - 4503599627370496.0;
#else
    return (double) w;
#endif
}

// 53-bit uniform in [0,1): (a >> 5) * 2^26 + (b >> 6), an exact 53-bit integer, times 2^-53
__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return fma(u32_to_double(a >> 5), 67108864.0, u32_to_double(b >> 6)) * (1.0 / 9007199254740992.0);
#else
    return (double) ((((uint64_t) (a >> 5)) << 26) | (uint64_t) (b >> 6)) * (1.0 / 9007199254740992.0);
#endif
}

#ifdef __CUDACC__
// two independent N(0,1) variates from one Philox block (Box-Muller, fp64)
__device__ __forceinline__ void normal_pair(const Philox4 &r, double &z0, double &z1) {
    const double u1 = 1.0 - u53(r.x, r.y); // (0,1]
    const double u2 = u53(r.z, r.w);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}
#endif

} // namespace bb
