// Device objective registry: Sphere, Ellipsoid, Rosenbrock, Rastrigin, Ackley.
//
// Replaces the per-candidate callback `_f._f(&x[0])` of the reference
// (cma/base_cmaes.cpp:214-217, de/jade.cpp:159, de/shade.cpp:159) by batched evaluation:
// one warp per candidate, coalesced 256-byte row segments, warp-shuffle reduction.
//
// Fitness drives every ranking and every greedy DE selection, so the arithmetic follows a
// fixed SPEC (DESIGN.md "Objective spec") that a CPU can reproduce bit for bit:
//   * 32 interleaved accumulators: lane l adds term(i) for i = l, l+32, ... in ascending i;
//   * xor butterfly 16,8,4,2,1 to combine the lanes;
//   * every multiply/add is an explicitly rounded __dmul_rn/__dadd_rn (never contracted to FMA);
//   * cos(2 pi x) and exp(x) are fixed polynomials over + - * rint (no libdevice calls).
#pragma once

#include "common.cuh"

namespace bb {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

// rint(x) (round to nearest, ties to even) on the fp64 pipe: adding and subtracting 2^52 with the sign of x rounds at
// the unit place.  Same value as rint for every finite x (|x| >= 2^52 is already integral); the FRND instruction it
// replaces issues on the quarter-rate XU pipe.
__device__ __forceinline__ double rint_rn(double x) {
    const double t = copysign(4503599627370496.0, x);
    return fabs(x) < 4503599627370496.0 ? sub(add(x, t), t) : x;
}

// Taylor coefficients of the cosine branch [0] (t^18 .. t^2) and of the sine branch [1] (t^19 .. t^3, divided by t)
static __constant__ double COS2PI_COEF[2][9] = {
    { -1.5619206968586225e-16, 4.779477332387385e-14, -1.1470745597729725e-11, 2.08767569878681e-09, -2.755731922398589e-07,
      2.48015873015873e-05, -0.001388888888888889, 0.041666666666666664, -0.5 },
    { -8.22063524662433e-18, 2.8114572543455206e-15, -7.647163731819816e-13, 1.6059043836821613e-10, -2.505210838544172e-08,
      2.7557319223985893e-06, -0.0001984126984126984, 0.008333333333333333, -0.16666666666666666 } };

// cos(2 pi x): reduce to r = x - rint(x) in [-1/2, 1/2], fold to [0, 1/8] with the quarter- and
// eighth-period symmetries, then a degree-18 Taylor polynomial in t = 2 pi * folded (t <= pi/4).
__device__ __forceinline__ double cos2pi(double x) {
    const double r = sub(x, rint_rn(x));
    double a = fabs(r);
    bool neg = false;
    if (a > 0.25) { a = sub(0.5, a); neg = true; }
    // the sine branch (a > 1/8) and the cosine branch share one Horner recurrence: only the
    // coefficients and the final multiply differ, so lanes never diverge (same bits as two branches)
    const bool big = a > 0.125;
    const double t = mul(big ? sub(0.25, a) : a, 6.283185307179586);
    const double u = mul(t, t);
#ifdef BB_COS2PI_TABLE
    // coefficient pairs from constant memory, indexed by the branch: one load per Horner step instead of materialising
    // and selecting two 64-bit immediates.  Same bits.  Pays where the term count dominates (the DE build: the objective
    // terms were 41 % of its instructions, +6 % evaluations/s); measured slower in the register-starved batched CMA kernel.
    const double *c = COS2PI_COEF[big ? 1 : 0];
    double p = c[0];
#pragma unroll
    for (int k = 1; k < 9; k++) p = add(mul(p, u), c[k]);
    p = add(mul(p, u), 1.0);
#else
    double p = big ? -8.22063524662433e-18 : -1.5619206968586225e-16;
    p = add(mul(p, u), big ? 2.8114572543455206e-15 : 4.779477332387385e-14);
    p = add(mul(p, u), big ? -7.647163731819816e-13 : -1.1470745597729725e-11);
    p = add(mul(p, u), big ? 1.6059043836821613e-10 : 2.08767569878681e-09);
    p = add(mul(p, u), big ? -2.505210838544172e-08 : -2.755731922398589e-07);
    p = add(mul(p, u), big ? 2.7557319223985893e-06 : 2.48015873015873e-05);
    p = add(mul(p, u), big ? -0.0001984126984126984 : -0.001388888888888889);
    p = add(mul(p, u), big ? 0.008333333333333333 : 0.041666666666666664);
    p = add(mul(p, u), big ? -0.16666666666666666 : -0.5);
    p = add(mul(p, u), 1.0);
#endif
    const double v = big ? mul(p, t) : p;
    return neg ? -v : v;
}

// exp(x) = 2^k * P13(r), k = rint(x log2 e), r = x - k ln2 (two-part), |r| <= 0.35
__device__ __forceinline__ double exp_poly(double x) {
    if (x < -700.0) return 0.0;
    if (x > 700.0) x = 700.0;
    const double k = rint(mul(x, 1.4426950408889634));
    double r = sub(x, mul(k, 0.6931471803691238));
    r = sub(r, mul(k, 1.9082149292705877e-10));
    double p = 1.6059043836821613e-10;
    p = add(mul(p, r), 2.08767569878681e-09);
    p = add(mul(p, r), 2.505210838544172e-08);
    p = add(mul(p, r), 2.755731922398589e-07);
    p = add(mul(p, r), 2.7557319223985893e-06);
    p = add(mul(p, r), 2.48015873015873e-05);
    p = add(mul(p, r), 0.0001984126984126984);
    p = add(mul(p, r), 0.001388888888888889);
    p = add(mul(p, r), 0.008333333333333333);
    p = add(mul(p, r), 0.041666666666666664);
    p = add(mul(p, r), 0.16666666666666666);
    p = add(mul(p, r), 0.5);
    p = add(mul(p, r), 1.0);
    p = add(mul(p, r), 1.0);
    const long long ki = (long long) k;
    const double two_k = __longlong_as_double((ki + 1023LL) << 52);
    return mul(p, two_k);
}

// term(i) of the objective sum (spec: DESIGN.md "Objective spec").  `xn` = x[i+1] (only Rosenbrock reads
// it).  Objectives with two sums (Ackley) return the second term through t2.  Sums run over i in
// [0, obj_terms(id, n)).
__device__ __forceinline__ int obj_term_count(int id, int n) { return id == BBGPU_OBJ_ROSENBROCK ? n - 1 : n; }

__device__ __forceinline__ double obj_term(int id, double v, double xn, double auxi, double &t2) {
    switch (id) {
    case BBGPU_OBJ_SPHERE:
        return mul(v, v);
    case BBGPU_OBJ_ELLIPSOID:
        return mul(auxi, mul(v, v));
    case BBGPU_OBJ_ROSENBROCK: {
        const double t1 = sub(xn, mul(v, v));
        const double tt = sub(1.0, v);
        return add(mul(100.0, mul(t1, t1)), mul(tt, tt));
    }
    case BBGPU_OBJ_RASTRIGIN:
        return sub(mul(v, v), mul(10.0, cos2pi(v)));
    case BBGPU_OBJ_ACKLEY:
        t2 = cos2pi(v);
        return mul(v, v);
    default:
        return nan("");
    }
}

// objective value from the reduced sum(s)
__device__ __forceinline__ double obj_finish(int id, int n, double s1, double s2) {
    switch (id) {
    case BBGPU_OBJ_RASTRIGIN:
        return add(mul(10.0, (double) n), s1);
    case BBGPU_OBJ_ACKLEY: {
        const double t1 = mul(20.0, exp_poly(mul(-0.2, sqrt(__ddiv_rn(s1, (double) n)))));
        const double t2 = exp_poly(__ddiv_rn(s2, (double) n));
        return add(sub(20.0, t1), sub(2.718281828459045, t2));
    }
    default:
        return s1;
    }
}

// One warp evaluates one candidate row (global or shared memory): lane l accumulates term(i) for
// i = l, l+32, ... in ascending order, then the xor butterfly.  All 32 lanes return the value.
// `aux` = ellipsoid coefficient table (device, length n) or nullptr.
template<typename Ptr>
__device__ __forceinline__ double warp_objective(int id, int n, Ptr x, const double *__restrict__ aux) {
    const int lane = threadIdx.x & 31;
    const int nt = obj_term_count(id, n);
    double acc = 0.0, acc2 = 0.0;
    for (int i = lane; i < nt; i += 32) {
        const double v = x[i];
        const double xn = id == BBGPU_OBJ_ROSENBROCK ? x[i + 1] : 0.0;
        const double a = id == BBGPU_OBJ_ELLIPSOID ? aux[i] : 0.0;
        double t2 = 0.0;
        acc = add(acc, obj_term(id, v, xn, a, t2));
        if (id == BBGPU_OBJ_ACKLEY) acc2 = add(acc2, t2);
    }
    const double s1 = warp_sum(acc);
    const double s2 = id == BBGPU_OBJ_ACKLEY ? warp_sum(acc2) : 0.0;
    return obj_finish(id, n, s1, s2);
}

// The same value with the objective known at compile time: the term loop unrolls, four rows' worth of loads and term chains
// are in flight, and the accumulation keeps its order (the bits are those of warp_objective).  For the standalone evaluation
// kernel, where a generic loop left every iteration waiting for its own loads (41 us for 64 rows of 10^4 coordinates).
template<int ID, typename Ptr>
__device__ __forceinline__ double warp_objective_id(int n, Ptr x, const double *__restrict__ aux) {
    const int lane = threadIdx.x & 31;
    const int nt = obj_term_count(ID, n);
    double acc = 0.0, acc2 = 0.0;
    int i = lane;
    for (; i + 96 < nt; i += 128) {
        double v[4], xn[4], a[4], t1[4], t2[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            v[u] = x[i + 32 * u];
            xn[u] = ID == BBGPU_OBJ_ROSENBROCK ? x[i + 32 * u + 1] : 0.0;
            a[u] = ID == BBGPU_OBJ_ELLIPSOID ? aux[i + 32 * u] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            t2[u] = 0.0;
            t1[u] = obj_term(ID, v[u], xn[u], a[u], t2[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            acc = add(acc, t1[u]);
            if (ID == BBGPU_OBJ_ACKLEY) acc2 = add(acc2, t2[u]);
        }
    }
    for (; i < nt; i += 32) {
        const double v = x[i];
        const double xn = ID == BBGPU_OBJ_ROSENBROCK ? x[i + 1] : 0.0;
        const double a = ID == BBGPU_OBJ_ELLIPSOID ? aux[i] : 0.0;
        double t2 = 0.0;
        acc = add(acc, obj_term(ID, v, xn, a, t2));
        if (ID == BBGPU_OBJ_ACKLEY) acc2 = add(acc2, t2);
    }
    const double s1 = warp_sum(acc);
    const double s2 = ID == BBGPU_OBJ_ACKLEY ? warp_sum(acc2) : 0.0;
    return obj_finish(ID, n, s1, s2);
}

// The same value from per-term buffers in shared memory (terms computed by a whole CTA in parallel):
// called by ONE warp; t1/t2 hold term(i) for i < obj_term_count.  Same accumulation order as above.
__device__ __forceinline__ double warp_reduce_terms(int id, int n, const double *t1, const double *t2) {
    const int lane = threadIdx.x & 31;
    const int nt = obj_term_count(id, n);
    double acc = 0.0, acc2 = 0.0;
    for (int i = lane; i < nt; i += 32) {
        acc = add(acc, t1[i]);
        if (id == BBGPU_OBJ_ACKLEY) acc2 = add(acc2, t2[i]);
    }
    const double s1 = warp_sum(acc);
    const double s2 = id == BBGPU_OBJ_ACKLEY ? warp_sum(acc2) : 0.0;
    return obj_finish(id, n, s1, s2);
}

} // namespace bb

// Objective handle behind the C ABI.
struct bbgpu_objective {
    bbgpu_ctx *ctx = nullptr;
    int mode = 0; // 0 builtin, 1 user device function, 2 batched host callback
    int id = 0;
    int n = 0;
    double *d_aux = nullptr; // ellipsoid coefficients
    bbgpu_device_objective_fn dev_fn = nullptr;
    bbgpu_host_objective_fn host_fn = nullptr;
    void *user = nullptr;
    int64_t evals = 0;
};

namespace bb {
// enqueue evaluation of `count` rows (leading dimension ld) on the ctx stream
int evaluate(bbgpu_ctx *ctx, bbgpu_objective *obj, const double *d_X, int64_t ld, int count, double *d_f);
// the same for the rows with mask[i] != 0 when the objective is a user function (a registry objective evaluates every row);
// work: count * (round_up(n, 8) + 1) doubles + count + 8 ints of device scratch; synchronises the stream for a user objective
int evaluate_masked(bbgpu_ctx *ctx, bbgpu_objective *obj, const double *d_X, int64_t ld, int count, double *d_f, const int *d_mask,
        void *work);
inline size_t evaluate_masked_work_bytes(int n, int count) {
    return (size_t) count * ((size_t) round_up(n, 8) + 1) * sizeof(double) + ((size_t) count + 8) * sizeof(int);
}
}
