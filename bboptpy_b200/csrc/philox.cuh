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
#include <cmath>
#include <cstring>

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

// ---- Box-Muller kernels.  libdevice's log / sqrt / sincospi cost ~125 instructions per pair with their special-case
// paths; the draws below cost ~75 and keep fp64 accuracy (|error| <= ~2 ulp of the result, checked against long double
// in tools/normal_accuracy.cpp).  The same code runs on the host (tests) and on the device (every sampling kernel).
#include "normal_tab.inc"

#ifdef __CUDACC__
static __device__ const double BB_LOGTAB[256] = BB_LOGTAB_INIT;
#endif
static const double BB_LOGTAB_H[256] = BB_LOGTAB_INIT;
struct bb_double2 { double x, y; };
#ifdef __CUDA_ARCH__
#define BB_LOGTAB_AT(i) (*reinterpret_cast<const double2*>(&BB_LOGTAB[2 * (i)]))
#else
#define BB_LOGTAB_AT(i) (bb::bb_double2 { bb::BB_LOGTAB_H[2 * (i)], bb::BB_LOGTAB_H[2 * (i) + 1] })
#endif

__host__ __device__ __forceinline__ double bb_hilo2double(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double((int) hi, (int) lo);
#else
    const uint64_t b = ((uint64_t) hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}

__host__ __device__ __forceinline__ uint32_t bb_double2hi(double d) {
#ifdef __CUDA_ARCH__
    return (uint32_t) __double2hiint(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return (uint32_t) (b >> 32);
#endif
}

__host__ __device__ __forceinline__ uint32_t bb_double2lo(double d) {
#ifdef __CUDA_ARCH__
    return (uint32_t) __double2loint(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return (uint32_t) b;
#endif
}

// -2 ln(u) for u in [2^-53, 1]: u = 2^e m, m in [1, 2); a 128-entry table on the top mantissa bits gives rc ~ 1/m and
// lc = ln(1/rc) (entries with m >= 1.414 are folded to m/2 so that the sum does not cancel near u = 1);
// r = m rc - 1 exactly (one fma), |r| <= 2^-8, log1p(r) by its series through r^7.
__host__ __device__ __forceinline__ double neg2log_unit(double u) {
    const uint32_t hi = bb_double2hi(u), lo = bb_double2lo(u);
    const int idx = (int) ((hi >> 13) & 127u);
    const int e = (int) (hi >> 20) - 1023 + (idx >= 53 ? 1 : 0);
    const double m = bb_hilo2double((hi & 0x000fffffu) | 0x3ff00000u, lo);
    const auto tb = BB_LOGTAB_AT(idx);
    const double r = fma(m, tb.x, -1.0);
    double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    p = fma(p, r, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    const double l1p = fma(p * r, r, r);
    const double ed = u32_to_double((uint32_t) (e + 64)) - 64.0; // exact; avoids the conversion unit on the device
    const double t = -2.0 * (fma(ed, BB_LN2, tb.y) + l1p);
    return t > 0.0 ? t : 0.0;
}

// sqrt(t) for t >= 0 through one reciprocal-square-root seed and a third-order correction (no division, no special path)
__host__ __device__ __forceinline__ double sqrt_nonneg(double t) {
#ifdef __CUDA_ARCH__
    const double tt = fmax(t, 1e-300);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(tt));
    const double e = fma(-tt * y, y, 1.0);
    y = fma(y * e, fma(e, 0.375, 0.5), y);   // ~1 ulp of 1/sqrt
    const double s = t * y;                  // sqrt, ~1.5 ulp; one Newton step on the square root itself:
    return fma(fma(-s, s, t), 0.5 * y, s);
#else
    return sqrt(t);
#endif
}

// (sin, cos)(2 pi u) for u in [0, 1): y = 4u quarter turns = q + r, |r| <= 1/2; Taylor series of sin / cos (pi r / 2)
// (error < 5e-17), then the quadrant's swap and signs.
__host__ __device__ __forceinline__ void sincos_2pi_unit(double u, double &sn, double &cs) {
    const double SC[8] = BB_SINC_INIT;
    const double CC[9] = BB_COSC_INIT;
    const double y = 4.0 * u;
    const double yk = y + 6755399441055744.0; // 1.5 * 2^52: rint(y) sits in the low mantissa bits
    const uint32_t q = bb_double2lo(yk) & 3u;
    const double r = y - (yk - 6755399441055744.0);
    const double r2 = r * r;
    double ps = SC[7], pc = CC[8];
#pragma unroll
    for (int k = 6; k >= 0; k--) ps = fma(ps, r2, SC[k]);
#pragma unroll
    for (int k = 7; k >= 0; k--) pc = fma(pc, r2, CC[k]);
    ps *= r;
    const bool swap = (q & 1u) != 0u;
    double s0 = swap ? pc : ps, c0 = swap ? ps : pc;
    if (q & 2u) s0 = -s0;            // quadrants 2, 3: sin < 0 part
    if ((q + 1u) & 2u) c0 = -c0;     // quadrants 1, 2
    sn = s0;
    cs = c0;
}

// two independent N(0,1) variates from one Philox block (Box-Muller, fp64)
__host__ __device__ __forceinline__ void normal_pair(const Philox4 &r, double &z0, double &z1) {
    const double u1 = 1.0 - u53(r.x, r.y); // (0,1]
    const double u2 = u53(r.z, r.w);
    const double rad = sqrt_nonneg(neg2log_unit(u1));
    double s, c;
    sincos_2pi_unit(u2, s, c);
    z0 = rad * c;
    z1 = rad * s;
}

} // namespace bb
