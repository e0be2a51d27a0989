/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C statement of the synthetic objective SPEC (DESIGN.md "Objective spec").
 * The reference ships no objectives (its only one is the Python lambda in
 * /root/reference/examples/multivariate.py:6-10, the README Rosenbrock form), so the
 * registry objectives are ours to define.  To make fitness -- and therefore every
 * ranking and every greedy DE selection -- bit-identical between this CPU oracle and
 * the CUDA registry (bboptpy_b200/csrc/objectives.cuh, written independently from the
 * same spec), the spec fixes the exact IEEE-754 operation order:
 *
 *   reduce(term, nt): 32 interleaved accumulators, acc[i % 32] += term(i) for i ascending,
 *                     then the xor butterfly 16,8,4,2,1:  acc[l] = acc[l] + acc[l ^ off].
 *   no fused multiply-add anywhere (compile with -ffp-contract=off);
 *   cos(2 pi x) and exp(x) are the polynomial forms below, built only from + - * rint.
 *
 * Compile with:  gcc/g++ -O2 -ffp-contract=off   (no -ffast-math).
 */
#ifndef BBO_OBJECTIVES_REF_H_
#define BBO_OBJECTIVES_REF_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

enum { BBO_SPHERE = 0, BBO_ELLIPSOID = 1, BBO_ROSENBROCK = 2, BBO_RASTRIGIN = 3, BBO_ACKLEY = 4 };

static inline double bbo_cos2pi(double x) {
    static const double C[10] = { 1.0, -0.5, 0.041666666666666664, -0.001388888888888889,
        2.48015873015873e-05, -2.755731922398589e-07, 2.08767569878681e-09,
        -1.1470745597729725e-11, 4.779477332387385e-14, -1.5619206968586225e-16 };
    static const double S[10] = { 1.0, -0.16666666666666666, 0.008333333333333333,
        -0.0001984126984126984, 2.7557319223985893e-06, -2.505210838544172e-08,
        1.6059043836821613e-10, -7.647163731819816e-13, 2.8114572543455206e-15,
        -8.22063524662433e-18 };
    const double TWO_PI = 6.283185307179586;
    double r = x - rint(x);
    double a = fabs(r);
    int neg = 0;
    if (a > 0.25) { a = 0.5 - a; neg = 1; }
    double v;
    if (a > 0.125) {
        double t = (0.25 - a) * TWO_PI;
        double u = t * t;
        double s = S[9];
        for (int k = 8; k >= 0; k--) { s = s * u; s = s + S[k]; }
        v = s * t;
    } else {
        double t = a * TWO_PI;
        double u = t * t;
        double c = C[9];
        for (int k = 8; k >= 0; k--) { c = c * u; c = c + C[k]; }
        v = c;
    }
    return neg ? -v : v;
}

static inline double bbo_exp(double x) {
    static const double E[14] = { 1.0, 1.0, 0.5, 0.16666666666666666, 0.041666666666666664,
        0.008333333333333333, 0.001388888888888889, 0.0001984126984126984,
        2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07,
        2.505210838544172e-08, 2.08767569878681e-09, 1.6059043836821613e-10 };
    const double LOG2E = 1.4426950408889634;
    const double LN2_HI = 0.6931471803691238, LN2_LO = 1.9082149292705877e-10;
    if (x < -700.0) return 0.0;
    if (x > 700.0) x = 700.0;
    double k = rint(x * LOG2E);
    double r = x - k * LN2_HI;
    r = r - k * LN2_LO;
    double p = E[13];
    for (int j = 12; j >= 0; j--) { p = p * r; p = p + E[j]; }
    int64_t ki = (int64_t) k;
    uint64_t bits = (uint64_t) (ki + 1023) << 52;
    double two_k;
    memcpy(&two_k, &bits, 8);
    return p * two_k;
}

/* ellipsoid coefficient table: 10^(6 i/(n-1)); 1 for n==1 */
static inline void bbo_ellipsoid_coef(int n, double *coef) {
    for (int i = 0; i < n; i++) {
        coef[i] = (n > 1) ? pow(10.0, 6.0 * (double) i / ((double) n - 1.0)) : 1.0;
    }
}

static inline double bbo_butterfly(double *acc) {
    for (int off = 16; off >= 1; off >>= 1) {
        double nw[32];
        for (int l = 0; l < 32; l++) nw[l] = acc[l] + acc[l ^ off];
        for (int l = 0; l < 32; l++) acc[l] = nw[l];
    }
    return acc[0];
}

/* aux = ellipsoid coefficient table (only read for BBO_ELLIPSOID) */
static inline double bbo_objective(int id, int n, const double *x, const double *aux) {
    double acc[32], acc2[32];
    for (int l = 0; l < 32; l++) { acc[l] = 0.0; acc2[l] = 0.0; }
    switch (id) {
    case BBO_SPHERE:
        for (int i = 0; i < n; i++) acc[i & 31] = acc[i & 31] + x[i] * x[i];
        return bbo_butterfly(acc);
    case BBO_ELLIPSOID:
        for (int i = 0; i < n; i++) {
            double sq = x[i] * x[i];
            acc[i & 31] = acc[i & 31] + aux[i] * sq;
        }
        return bbo_butterfly(acc);
    case BBO_ROSENBROCK:
        for (int i = 0; i + 1 < n; i++) {
            double t1 = x[i + 1] - x[i] * x[i];
            double t2 = 1.0 - x[i];
            double a = 100.0 * (t1 * t1);
            double b = t2 * t2;
            acc[i & 31] = acc[i & 31] + (a + b);
        }
        return bbo_butterfly(acc);
    case BBO_RASTRIGIN: {
        for (int i = 0; i < n; i++) {
            double sq = x[i] * x[i];
            double c = 10.0 * bbo_cos2pi(x[i]);
            acc[i & 31] = acc[i & 31] + (sq - c);
        }
        double s = bbo_butterfly(acc);
        return 10.0 * (double) n + s;
    }
    case BBO_ACKLEY: {
        for (int i = 0; i < n; i++) {
            acc[i & 31] = acc[i & 31] + x[i] * x[i];
            acc2[i & 31] = acc2[i & 31] + bbo_cos2pi(x[i]);
        }
        double s1 = bbo_butterfly(acc);
        double s2 = bbo_butterfly(acc2);
        double t1 = 20.0 * bbo_exp(-0.2 * sqrt(s1 / (double) n));
        double t2 = bbo_exp(s2 / (double) n);
        return (20.0 - t1) + (2.718281828459045 - t2);
    }
    default:
        return NAN;
    }
}

#endif
