/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * CPU restatement of the reference's population hot path, written from the algorithm the
 * reference implements (flat row-major arrays, one struct per family), exporting the same
 * C API (oracle/bbo_api.h) as the compiled reference (oracle/_ref/libbbo_ref.so).
 * Every function cites the reference file:line it follows.  The floating-point operation
 * ORDER is kept identical to the reference on purpose: tests/test_oracle_pinning.py asserts
 * this port is bit-identical to the compiled reference (state after every phase, every
 * generation) and tests/golden/ holds vectors generated from the reference itself.
 *
 * It additionally records the DE family's *tagged draws* (the reference consumes a
 * data-dependent number of draws per candidate: jade.cpp:105-183), which is what the GPU
 * engine replays.
 *
 * Build: oracle/Makefile (g++ -O2 -ffp-contract=off).
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../bbo_api.h"
#include "../objectives_ref.h"
#include "bbo_rng.h"

namespace {

BboMt19937 g_eng;          /* the one process-global engine (random.hpp:677-680) */
BboNormal g_test_normal;   /* only for bbo_rng_draw kind 0 */
bool g_seeded = false;

inline BboMt19937 &eng() {
    if (!g_seeded) { g_eng.seed(5489u); g_seeded = true; }
    return g_eng;
}

/* ---- src/blas.cpp:154-181 (scaled 2-norm, order dependent) */
double port_nrm2(int n, const double *x) {
    if (n < 1) return 0.0;
    if (n == 1) return std::fabs(x[0]);
    double scale = 0.0, ssq = 1.0;
    for (int i = 0; i < n; i++) {
        if (x[i] != 0.0) {
            const double a = std::fabs(x[i]);
            if (scale < a) {
                ssq = 1.0 + ssq * (scale / a) * (scale / a);
                scale = a;
            } else {
                ssq = ssq + (a / scale) * (a / scale);
            }
        }
    }
    return scale * std::sqrt(ssq);
}

/* ---- src/math_utils.tpp:27-29,39-51 */
inline double port_sign(double x, double y) { return y >= 0.0 ? std::fabs(x) : -std::fabs(x); }
/* NOTE: cmaes.cpp:411,433 call `hypot(p, x)` unqualified with double arguments; overload
 * resolution picks libm's ::hypot (a non-template exact match) over the template in
 * math_utils.tpp:39-51, so the reference's QL iteration uses glibc hypot.  Pinned by
 * tests/test_oracle_pinning.py (using the template form instead breaks bit-equality). */
inline double port_hypot(double a, double b) { return ::hypot(a, b); }

struct Objective {
    int id = 0, n = 0;
    std::vector<double> aux;
    void bind(int obj, int dim) {
        id = obj;
        n = dim;
        aux.clear();
        if (obj == BBO_ELLIPSOID) {
            aux.resize(dim);
            bbo_ellipsoid_coef(dim, aux.data());
        }
    }
    double operator()(const double *x) const {
        return bbo_objective(id, n, x, aux.empty() ? nullptr : aux.data());
    }
};

/* ---- cma/base_cmaes.cpp:32-43 */
struct Ring {
    int cap = 0, head = -1, len = 0;
    std::vector<double> v;
    void reset(int c) { cap = c; head = -1; len = 0; v.assign(c, 0.0); }
    void add(double x) {
        head = (head + 1) % cap;
        v[head] = x;
        if (len < cap) len++;
    }
    double get(int i) const { return v[(cap + head - i) % cap]; }
};

struct RankRec {
    int index;
    double value;
    static bool less(const RankRec &a, const RankRec &b) { return a.value < b.value; }
};

struct PortCma {
    /* constructor arguments */
    int kind, mfev, lambda;
    double tol, sigma0, eigenrate, alphacov;
    bool bound, adjustlr;
    /* problem */
    Objective f;
    int n = 0;
    std::vector<double> lower, upper;
    /* BaseCmaes state (base_cmaes.h:52-63) */
    int mit = 0, mu = 0, it = 0, hlen = 0, ik = 0, fev = 0;
    double mueff = 0, cc = 0, cs = 0, c1 = 0, cmu = 0, damps = 0, chi = 0, sigma = 0, fbest = 0,
            fworst = 0;
    Ring best, kth;
    int ibw[4] = { 0, 0, 0, 0 };
    double ybw[4] = { 0, 0, 0, 0 };
    std::vector<double> xmean, xold, weights, tmp, pc, ps, arx;
    std::vector<double> last_z; /* the N(0,1) draws of the last sample(), draw order (replay feed) */
    std::vector<RankRec> rank;
    /* Cmaes / ActiveCmaes state (cmaes.h:42-46, active_cmaes.h:42-44) */
    int flag = 0, eigenlastev = 0;
    double eigenfreq = 0;
    std::vector<double> d, B, C, isC;
    double cm = 1, cneg = 0, alphaold = 0.5;
    std::vector<double> ycoeff;
    /* SepCmaes state (sep_cmaes.h:39-43): C and d are length n */
    double ccov = 0;
    /* CholeskyCmaes state (cholesky_cmaes.h:34-37); stol arrives in the eigenrate argument */
    std::vector<double> A, mattmp, dmean;
    /* LmCmaes state (lm_cmaes.h:38-46); memory / rademacher / usenew arrive in eigenrate / alphacov / adjustlr */
    bool lm_new = true, lm_adaptmemory = true;
    int lm_samplemode = 1, memlen = 0, memsize = 0, nsteps = 0, lm_t = 1;
    double stolmin = 1e-16, sqrt1mc1 = 0, zstar = 0.25, lm_s = 0, lm_ccc = 0;
    std::vector<int> jarr, larr;
    std::vector<double> lm_b, lm_d, az, fp, pcmat, vmat;
    std::vector<RankRec> mixed;
    std::vector<double> last_i0; /* replay feed: the subset start of every sampled pair (selectSubset) */
    BboNormal Z;
    /* eigenbasis-replay feed (bbo_cma_eig_log) */
    struct EigEvent { int fev, n; std::vector<double> B, D; };
    bool eiglog = false;
    std::vector<EigEvent> eigevents;

    /* ---- cma/base_cmaes.cpp:54-134 */
    void base_init(const double *guess) {
        mu = lambda / 2;
        mit = mfev / lambda;
        arx.assign((size_t) lambda * n, 0.0);
        for (int i = 0; i < 4; i++) { ibw[i] = 0; ybw[i] = 0.0; }
        rank.assign(lambda, RankRec { 0, 0.0 });
        weights.assign(mu, 0.0);
        double sum = 0.0;
        for (int i = 0; i < mu; i++) {
            weights[i] = std::log(0.5 * (lambda + 1.0)) - std::log(i + 1.0);
            sum += weights[i];
        }
        const double inv = 1.0 / sum;
        for (int i = 0; i < mu; i++) weights[i] *= inv;
        double lenw = 0.0;
        for (int i = 0; i < mu; i++) lenw = lenw + weights[i] * weights[i];
        mueff = 1.0 / lenw;
        chi = std::sqrt(n) * (1.0 - 1.0 / (4.0 * n) + 1.0 / (21.0 * n * n));
        sigma = sigma0;
        cc = (4.0 + mueff / n) / (n + 4.0 + 2.0 * mueff / n);
        cs = (mueff + 2.0) / (5.0 + n + mueff);
        c1 = 2.0 / ((1.3 + n) * (1.3 + n) + mueff);
        cmu = std::min(1.0 - c1, 2.0 * (mueff - 2.0 + 1.0 / mueff) / ((2.0 + n) * (2.0 + n) + mueff));
        damps = 1.0 + cs + 2.0 * std::max(0.0, std::sqrt((mueff - 1.0) / (n + 1.0)) - 1.0);
        pc.assign(n, 0.0);
        ps.assign(n, 0.0);
        tmp.assign(n, 0.0);
        xold.assign(n, 0.0);
        xmean.assign(guess, guess + n);
        it = fev = 0;
        hlen = 10 + (int) std::ceil((30.0 * n) / lambda);
        ik = (int) std::ceil(0.1 + lambda / 4.0);
        best.reset(hlen);
        kth.reset(hlen);
        fbest = -std::numeric_limits<double>::infinity();
        fworst = std::numeric_limits<double>::infinity();
    }

    /* ---- cma/cmaes.cpp:44-63, cma/active_cmaes.cpp:42-69, cma/sep_cmaes.cpp:40-67 */
    void init(const double *guess) {
        base_init(guess);
        if (kind == 4) { /* cma/lm_cmaes.cpp:42-89 */
            if (lm_adaptmemory) memsize = (int) (2.0 * std::sqrt(1.0 * n));
            memlen = 0;
            if (lm_new) {
                nsteps = n;
                lm_t = std::max(1, (int) std::log(1.0 * n));
                cc = 0.5 / std::sqrt(1.0 * n);
            } else {
                nsteps = memsize;
                lm_t = 1;
                cc = 1.0 / memsize;
            }
            cs = 0.3;
            c1 = 0.1 / std::log(n + 1.0);
            cmu = NAN;
            lm_ccc = std::sqrt(cc * (2.0 - cc) * mueff);
            damps = 1.0;
            sqrt1mc1 = std::sqrt(1.0 - c1);
            zstar = 0.25;
            lm_s = 0.0;
            jarr.assign(memsize, 0);
            larr.assign(memsize, 0);
            lm_b.assign(memsize, 0.0);
            lm_d.assign(memsize, 0.0);
            az.assign(n, 0.0);
            fp.assign(lambda, 0.0);
            /* `_pcmat.resize` / `_vmat.resize` keep an already sized matrix (:82-83); every row that is read has been
             * written since (rows jarr[0 .. memlen-1]), so stale contents are never observed */
            pcmat.resize((size_t) memsize * n, 0.0);
            vmat.resize((size_t) memsize * n, 0.0);
            mixed.assign((size_t) 2 * lambda, RankRec { 0, 0.0 });
            flag = 0;
            return;
        }
        if (kind == 3) { /* cma/cholesky_cmaes.cpp:40-52 */
            dmean.assign(n, 0.0);
            A.assign((size_t) n * n, 0.0);
            mattmp.assign((size_t) n * n, 0.0);
            for (int i = 0; i < n; i++) A[(size_t) i * n + i] = 1.0;
            flag = 0;
            return;
        }
        if (kind == 2) {
            cc = 4.0 / (n + 4.0);
            cs = (mueff + 2.0) / (3.0 + n + mueff);
            damps = 1.0 + cs + 2.0 * std::max(0.0, std::sqrt((mueff - 1.0) / (n + 1.0)) - 1.0);
            ccov = 2.0 / ((n + std::sqrt(2.0)) * (n + std::sqrt(2.0)) * mueff);
            ccov += std::min(1.0, (2.0 * mueff - 1.0) / ((n + 2.0) * (n + 2.0) + mueff))
                    * (1.0 - 1.0 / mueff);
            if (adjustlr) ccov *= ((n + 2.0) / 3.0);
            d.assign(n, 1.0);
            C.assign(n, 1.0);
            flag = 0;
            return;
        }
        eigenfreq = eigenrate * lambda / (c1 + cmu) / n;
        eigenlastev = 0;
        d.assign(n, 1.0);
        /* cmaes.cpp:53-56: `_b.resize(...)` / `_c.resize(...)` do NOT clear an already sized
         * matrix (only `_invsqrtc` is cleared first), so on a re-init of the same object -- every
         * IPOP/BIPOP restart -- the off-diagonal entries of B and C survive and only the
         * diagonals are reset to 1 (:58-60).  Reproduced, not fixed. */
        if (B.size() != (size_t) n * n) B.assign((size_t) n * n, 0.0);
        if (C.size() != (size_t) n * n) C.assign((size_t) n * n, 0.0);
        isC.assign((size_t) n * n, 0.0);
        for (int i = 0; i < n; i++) B[(size_t) i * n + i] = C[(size_t) i * n + i] = isC[(size_t) i * n + i] = 1.0;
        flag = 0;
        if (kind == 1) {
            cm = 1.0;
            alphaold = 0.5;
            cc = (4.0 + 0.0 * mueff / n) / (n + 4.0 + 0.0 * 2.0 * mueff / n);
            cs = (mueff + 2.0) / (3.0 + n + mueff);
            c1 = alphacov * std::min(1.0, lambda / 6.0) / ((n + 1.3) * (n + 1.3) + mueff);
            cmu = 1.0 - c1;
            cmu = std::min(cmu, alphacov * (mueff - 2.0 + 1.0 / mueff)
                    / ((2.0 + n) * (2.0 + n) + alphacov * mueff / 2.0));
            cneg = (1.0 - cmu) * (alphacov / 8.0) * mueff / (std::pow(n + 2.0, 1.5) + 2.0 * mueff);
            damps = 1.0 + cs + 2.0 * std::max(0.0, std::sqrt((mueff - 1.0) / (n + 1.0)) - 1.0);
            eigenfreq = eigenrate * (1.0 / (c1 + cmu + cneg)) / n;
            eigenlastev = 0;
            ycoeff.assign(mu, 0.0);
        }
    }

    /* ---- cma/base_cmaes.cpp:136-148 */
    void set_params(int np, double sg, int mf) {
        lambda = np;
        sigma0 = sg;
        mfev = mf;
        bound = false;
    }

    /* ---- cma/cmaes.cpp:65-80 and cma/sep_cmaes.cpp:69-79 (draw order: candidate-major) */
    /* ---- cma/lm_cmaes.cpp:308-322 (algorithm 6): draws one normal only when m > 1 */
    int lm_select_subset(int m, int k) {
        if (m <= 1) return 0;
        int msigma = 4;
        if (k == 0) msigma *= 10;
        int mstar = (int) (msigma * std::fabs(Z.draw(eng())));
        mstar = std::min(mstar, m);
        return m - mstar;
    }

    /* ---- cma/lm_cmaes.cpp:91-127: mirrored sampling, one base vector per PAIR of candidates */
    void sample_lm() {
        const int pairs = (lambda + 1) / 2;
        last_z.assign((size_t) pairs * n, 0.0);
        last_i0.assign(pairs, 0.0);
        int sign = 1;
        for (int k = 0; k < lambda; k++) {
            if (sign == 1) {
                double *zrow = &last_z[(size_t) (k / 2) * n];
                if (lm_samplemode == 0) {
                    for (int i = 0; i < n; i++) tmp[i] = az[i] = zrow[i] = Z.draw(eng());
                } else {
                    for (int i = 0; i < n; i++) tmp[i] = az[i] = zrow[i] = eng().uniform_real(0.0, 1.0) < 0.5 ? 1.0 : -1.0;
                }
                const int i0 = lm_new ? lm_select_subset(memlen, k) : 0;
                last_i0[k / 2] = i0;
                for (int i = i0; i < memlen; i++) {
                    const int j = jarr[i];
                    const double *v = &vmat[(size_t) j * n], *p = &pcmat[(size_t) j * n];
                    double dot = 0.0;
                    for (int q = 0; q < n; q++) dot = dot + v[q] * tmp[q]; /* the ORIGINAL draw, not az (:117-119) */
                    dot = lm_b[j] * dot;
                    for (int q = 0; q < n; q++) az[q] = sqrt1mc1 * az[q];
                    for (int q = 0; q < n; q++) az[q] = az[q] + dot * p[q];
                }
            }
            double *x = &arx[(size_t) k * n];
            const double da = sign * sigma;
            for (int q = 0; q < n; q++) x[q] = xmean[q] + da * az[q];
            sign = -sign;
        }
    }

    void sample() {
        if (kind == 4) { sample_lm(); return; }
        last_z.resize((size_t) lambda * n);
        for (int k = 0; k < lambda; k++) {
            double *x = &arx[(size_t) k * n];
            double *zrow = &last_z[(size_t) k * n];
            if (kind == 3) { /* cma/cholesky_cmaes.cpp:120-135: x = m + sigma A z, full rows of A */
                for (int i = 0; i < n; i++) zrow[i] = Z.draw(eng());
                for (int i = 0; i < n; i++) {
                    const double *row = &A[(size_t) i * n];
                    double s = 0.0;
                    for (int j = 0; j < n; j++) s = s + row[j] * zrow[j];
                    x[i] = xmean[i] + sigma * s;
                    if (bound) x[i] = std::max(lower[i], std::min(x[i], upper[i]));
                }
            } else if (kind == 2) {
                for (int i = 0; i < n; i++) {
                    zrow[i] = Z.draw(eng());
                    x[i] = xmean[i] + sigma * d[i] * zrow[i];
                    if (bound) x[i] = std::max(lower[i], std::min(x[i], upper[i]));
                }
            } else {
                for (int i = 0; i < n; i++) {
                    zrow[i] = Z.draw(eng());
                    tmp[i] = d[i] * zrow[i];
                }
                for (int i = 0; i < n; i++) {
                    const double *row = &B[(size_t) i * n];
                    double s = 0.0;
                    for (int j = 0; j < n; j++) s = s + row[j] * tmp[j];
                    x[i] = xmean[i] + sigma * s;
                    if (bound) x[i] = std::max(lower[i], std::min(x[i], upper[i]));
                }
            }
        }
    }

    /* ---- cma/base_cmaes.cpp:211-230 */
    void evalsort() {
        if (kind == 4 && it > 0) /* cma/lm_cmaes.cpp:228-238 */
            for (int k = 0; k < lambda; k++) fp[k] = rank[k].value;
        for (int k = 0; k < lambda; k++) {
            rank[k].index = k;
            rank[k].value = f(&arx[(size_t) k * n]);
        }
        fev += lambda;
        std::sort(rank.begin(), rank.end(), RankRec::less);
        ibw[0] = rank[0].index;
        ibw[1] = rank[1].index;
        ibw[2] = rank[lambda - 2].index;
        ibw[3] = rank[lambda - 1].index;
        ybw[0] = rank[0].value;
        ybw[1] = rank[1].value;
        ybw[2] = rank[lambda - 2].value;
        ybw[3] = rank[lambda - 1].value;
    }

    /* ---- cma/base_cmaes.cpp:176-189 */
    void update_sigma() {
        const double pslen = port_nrm2(n, ps.data());
        sigma *= std::exp(std::min(1.0, (cs / damps) * (pslen / chi - 1.0)));
        if (rank[0].value == rank[ik].value) sigma *= std::exp(0.2 + cs / damps);
        if (it >= hlen && fworst - fbest == 0.0) sigma *= std::exp(0.2 + cs / damps);
    }

    /* weighted recombination, shared by the three variants
     * (cmaes.cpp:85-96, active_cmaes.cpp:74-85, sep_cmaes.cpp:84-95) */
    void recombine() {
        std::copy(xmean.begin(), xmean.end(), xold.begin());
        for (int i = 0; i < n; i++) {
            double sum = 0.0;
            for (int k = 0; k < mu; k++) sum += weights[k] * arx[(size_t) rank[k].index * n + i];
            xmean[i] = (kind == 1) ? xold[i] * (1.0 - cm) + sum * cm : sum;
            if (bound) xmean[i] = std::max(lower[i], std::min(xmean[i], upper[i]));
        }
    }

    int hsig_flag() const {
        const double pslen = port_nrm2(n, ps.data());
        const double denom = 1.0 - std::pow(1.0 - cs, 2.0 * fev / lambda);
        return (pslen / std::sqrt(denom) / chi < 1.4 + 2.0 / (n + 1.0)) ? 1 : 0;
    }

    /* ---- cma/cmaes.cpp:82-149 */
    void update_vanilla() {
        recombine();
        const double csc = std::sqrt(cs * (2.0 - cs) * mueff);
        for (int i = 0; i < n; i++) {
            ps[i] *= (1.0 - cs);
            for (int j = 0; j < n; j++) {
                ps[i] += csc * isC[(size_t) i * n + j] * (xmean[j] - xold[j]) / sigma;
            }
        }
        const int hsig = hsig_flag();
        const double ccc = std::sqrt(cc * (2.0 - cc) * mueff);
        for (int i = 0; i < n; i++) {
            pc[i] = (1.0 - cc) * pc[i] + hsig * ccc * (xmean[i] - xold[i]) / sigma;
        }
        const double c2 = (1.0 - hsig) * cc * (2.0 - cc);
        for (int i = 0; i < n; i++) {
            for (int j = 0; j <= i; j++) {
                double &cij = C[(size_t) i * n + j];
                double sum = (1.0 - c1 - cmu) * cij + c1 * (pc[i] * pc[j] + c2 * cij);
                for (int k = 0; k < mu; k++) {
                    const double *x = &arx[(size_t) rank[k].index * n];
                    const double di = (x[i] - xold[i]) / sigma;
                    const double dj = (x[j] - xold[j]) / sigma;
                    sum += cmu * weights[k] * di * dj;
                }
                cij = sum;
            }
        }
        update_sigma();
        eigen();
    }

    /* ---- cma/active_cmaes.cpp:71-168 */
    void update_active() {
        recombine();
        const double csc = std::sqrt(cs * (2.0 - cs) * mueff);
        for (int i = 0; i < n; i++) {
            ps[i] *= (1.0 - cs);
            for (int j = 0; j < n; j++) {
                ps[i] += csc * isC[(size_t) i * n + j] * (xmean[j] - xold[j]) / (cm * sigma);
            }
        }
        const int hsig = hsig_flag();
        const double ccc = std::sqrt(cc * (2.0 - cc) * mueff);
        for (int i = 0; i < n; i++) {
            pc[i] = (1.0 - cc) * pc[i] + hsig * ccc * (xmean[i] - xold[i]) / (cm * sigma);
        }
        /* Mahalanobis-norm ratios for the negative update (:115-132) */
        for (int q = 0; q < mu; q++) {
            const double *xt = &arx[(size_t) rank[lambda - mu + 1 + q - 1].index * n];
            const double *xb = &arx[(size_t) rank[lambda - q - 1].index * n];
            double ssqtop = 0.0, ssqbot = 0.0;
            for (int j = 0; j < n; j++) {
                double tt = 0.0, tb = 0.0;
                for (int l = 0; l < n; l++) {
                    tt += isC[(size_t) j * n + l] * (xt[l] - xold[l]);
                    tb += isC[(size_t) j * n + l] * (xb[l] - xold[l]);
                }
                ssqtop += tt * tt;
                ssqbot += tb * tb;
            }
            ssqbot = std::max(ssqbot, 1e-8);
            ycoeff[q] = ssqtop / ssqbot;
        }
        const double c2 = (1.0 - hsig) * cc * (2.0 - cc);
        const double cmu1 = cmu + cneg * (1.0 - alphaold);
        for (int i = 0; i < n; i++) {
            for (int j = 0; j <= i; j++) {
                double &cij = C[(size_t) i * n + j];
                double sum = (1.0 - c1 - cmu + cneg * alphaold) * cij
                        + c1 * (pc[i] * pc[j] + c2 * cij);
                for (int k = 0; k < mu; k++) {
                    const double *x = &arx[(size_t) rank[k].index * n];
                    const double di = (x[i] - xold[i]) / sigma;
                    const double dj = (x[j] - xold[j]) / sigma;
                    sum += cmu1 * weights[k] * di * dj;
                }
                for (int k = 0; k < mu; k++) {
                    const double *x = &arx[(size_t) rank[lambda - k - 1].index * n];
                    const double di = (x[i] - xold[i]) / sigma;
                    const double dj = (x[j] - xold[j]) / sigma;
                    sum -= cneg * weights[k] * ycoeff[k] * di * dj;
                }
                cij = sum;
            }
        }
        update_sigma();
        eigen();
    }

    /* ---- cma/sep_cmaes.cpp:81-135 (note :101 multiplies by the VARIANCE c_i: kept) */
    void update_sep() {
        recombine();
        const double csc = std::sqrt(cs * (2.0 - cs) * mueff);
        for (int i = 0; i < n; i++) {
            ps[i] *= (1.0 - cs);
            ps[i] += csc * C[i] * (xmean[i] - xold[i]) / sigma;
        }
        const int hsig = hsig_flag();
        const double ccc = std::sqrt(cc * (2.0 - cc) * mueff);
        for (int i = 0; i < n; i++) {
            pc[i] = (1.0 - cc) * pc[i] + hsig * ccc * (xmean[i] - xold[i]) / sigma;
        }
        for (int i = 0; i < n; i++) {
            double sum = (1.0 - ccov) * C[i] + (ccov / mueff) * pc[i] * pc[i];
            for (int k = 0; k < mu; k++) {
                const double di = (arx[(size_t) rank[k].index * n + i] - xold[i]) / sigma;
                sum += ccov * (1.0 - 1.0 / mueff) * weights[k] * di * di;
            }
            C[i] = sum;
            d[i] = std::sqrt(C[i]);
        }
        update_sigma();
    }

    /* ---- cma/cholesky_cmaes.cpp:163-177: rank-one update of the factor in mattmp by beta * tmp tmp^T */
    void chol_rank1(double beta) {
        double b = 1.0;
        for (int j = 0; j < n; j++) {
            const double ajj = mattmp[(size_t) j * n + j];
            const double alfaj = tmp[j];
            const double gam = ajj * ajj * b + beta * alfaj * alfaj;
            const double a1jj = mattmp[(size_t) j * n + j] = std::sqrt(gam / b);
            for (int k = j + 1; k < n; k++) {
                tmp[k] -= (alfaj / ajj) * mattmp[(size_t) k * n + j];
                mattmp[(size_t) k * n + j] = (a1jj / ajj) * mattmp[(size_t) k * n + j] + (a1jj * beta * alfaj) / gam * tmp[k];
            }
            b += beta * (alfaj / ajj) * (alfaj / ajj);
        }
    }

    /* ---- cma/cholesky_cmaes.cpp:54-118 */
    void update_chol() {
        std::copy(xmean.begin(), xmean.end(), xold.begin());
        for (int i = 0; i < n; i++) {
            double sum = 0.0;
            for (int k = 0; k < mu; k++) sum += weights[k] * arx[(size_t) rank[k].index * n + i];
            xmean[i] = sum;
            if (bound) xmean[i] = std::max(lower[i], std::min(xmean[i], upper[i]));
            dmean[i] = (xmean[i] - xold[i]) / sigma;
        }
        const double ccc = std::sqrt(cc * (2.0 - cc) * mueff);
        for (int i = 0; i < n; i++) pc[i] = (1.0 - cc) * pc[i] + ccc * dmean[i];
        const double acoeff = std::sqrt(1.0 - c1 - cmu);
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++) mattmp[(size_t) i * n + j] = acoeff * A[(size_t) i * n + j];
        std::copy(pc.begin(), pc.end(), tmp.begin());
        chol_rank1(c1);
        /* the reference walks the FIRST mu candidates by index (not by rank) against the NEW mean (:85-90): kept */
        for (int i = 0; i < mu; i++) {
            for (int j = 0; j < n; j++) tmp[j] = (arx[(size_t) i * n + j] - xmean[j]) / sigma;
            chol_rank1(cmu * weights[i]);
        }
        /* A(t)^-1 (m(t+1) - m(t)) / sigma with the OLD factor (:93-98) */
        std::copy(dmean.begin(), dmean.end(), tmp.begin());
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int j = 0; j < i; j++) s = s + A[(size_t) i * n + j] * tmp[j];
            tmp[i] -= s;
            tmp[i] /= A[(size_t) i * n + i];
        }
        const double csc = std::sqrt(cs * (2.0 - cs) * mueff);
        for (int i = 0; i < n; i++) ps[i] = (1.0 - cs) * ps[i] + csc * tmp[i];
        A = mattmp;
        update_sigma();
    }

    /* ---- cma/lm_cmaes.cpp:260-272 (algorithm 4) on tmp */
    void lm_ainvz(int jlen) {
        const double c = 1.0 / sqrt1mc1;
        for (int i = 0; i < jlen; i++) {
            const int idx = jarr[i];
            const double *v = &vmat[(size_t) idx * n];
            double dot = 0.0;
            for (int q = 0; q < n; q++) dot = dot + v[q] * tmp[q];
            dot = lm_d[idx] * dot;
            for (int q = 0; q < n; q++) tmp[q] = c * tmp[q];
            for (int q = 0; q < n; q++) tmp[q] = tmp[q] + (-dot) * v[q];
        }
    }

    /* ---- cma/lm_cmaes.cpp:274-306 (algorithm 5) */
    int lm_update_set() {
        const int itt = it / lm_t;
        int imin = 1;
        if (itt < memsize) {
            jarr[itt] = itt;
        } else if (memsize > 1) {
            int iminval = larr[jarr[1]] - larr[jarr[0]];
            for (int i = 1; i < memsize - 1; i++) {
                const int val = larr[jarr[i + 1]] - larr[jarr[i]];
                if (val < iminval) {
                    iminval = val;
                    imin = i + 1;
                }
            }
            if (iminval >= nsteps) imin = 0;
            const int jtmp = jarr[imin];
            for (int i = imin; i < memsize - 1; i++) jarr[i] = jarr[i + 1];
            jarr[memsize - 1] = jtmp;
        }
        const int jcur = jarr[std::min(itt, memsize - 1)];
        larr[jcur] = it;
        return imin == 1 ? 0 : imin;
    }

    /* ---- cma/lm_cmaes.cpp:240-258 */
    void lm_update_sigma() {
        if (it == 0) return;
        for (int k = 0; k < lambda; k++) {
            mixed[k].index = k;
            mixed[k].value = fp[k];
            mixed[k + lambda].index = k + lambda;
            mixed[k + lambda].value = rank[k].value;
        }
        std::sort(mixed.begin(), mixed.end(), RankRec::less);
        double zpsr = 0.0;
        for (int k = 0; k < (lambda << 1); k++) {
            const double f = (1.0 * k) / lambda;
            if (mixed[k].index < lambda) zpsr += f;
            else zpsr -= f;
        }
        zpsr /= (1.0 * lambda);
        zpsr -= zstar;
        lm_s = (1.0 - cs) * lm_s + cs * zpsr;
        sigma *= std::exp(lm_s / damps);
    }

    /* ---- cma/lm_cmaes.cpp:129-181 */
    void update_lm() {
        std::copy(xmean.begin(), xmean.end(), xold.begin());
        std::fill(xmean.begin(), xmean.end(), 0.0);
        for (int k = 0; k < mu; k++) {
            const double *x = &arx[(size_t) rank[k].index * n];
            for (int q = 0; q < n; q++) xmean[q] = xmean[q] + weights[k] * x[q];
        }
        if (bound)
            for (int i = 0; i < n; i++) xmean[i] = std::max(lower[i], std::min(xmean[i], upper[i]));
        for (int i = 0; i < n; i++) pc[i] = (1.0 - cc) * pc[i] + lm_ccc * (xmean[i] - xold[i]) / sigma;
        if (it % lm_t == 0) {
            const int imin = lm_update_set();
            if (memlen < memsize) memlen++;
            int jcur = jarr[memlen - 1];
            std::copy(pc.begin(), pc.end(), pcmat.begin() + (size_t) jcur * n);
            for (int i = imin; i < memlen; i++) {
                jcur = jarr[i];
                std::copy(pcmat.begin() + (size_t) jcur * n, pcmat.begin() + (size_t) (jcur + 1) * n, tmp.begin());
                lm_ainvz(i);
                std::copy(tmp.begin(), tmp.end(), vmat.begin() + (size_t) jcur * n);
                double vnrm2 = 0.0;
                for (int q = 0; q < n; q++) vnrm2 = vnrm2 + tmp[q] * tmp[q];
                const double sqrtc1 = std::sqrt(1.0 + (c1 / (1.0 - c1)) * vnrm2);
                lm_b[jcur] = (sqrt1mc1 / vnrm2) * (sqrtc1 - 1.0);
                lm_d[jcur] = (1.0 / (sqrt1mc1 * vnrm2)) * (1.0 - 1.0 / sqrtc1);
            }
        }
        lm_update_sigma();
    }

    void update() {
        if (kind == 4) { update_lm(); return; }
        if (kind == 0) update_vanilla();
        else if (kind == 1) update_active();
        else if (kind == 3) update_chol();
        else update_sep();
    }

    /* ---- cma/base_cmaes.cpp:191-209 */
    void history() {
        if (it >= mit) return;
        best.add(rank[0].value);
        kth.add(rank[ik].value);
        if (best.len == best.cap) {
            fbest = std::numeric_limits<double>::infinity();
            fworst = -std::numeric_limits<double>::infinity();
            for (double fx : best.v) {
                fbest = std::min(fx, fbest);
                fworst = std::max(fx, fworst);
            }
        }
    }

    void iterate() {
        sample();
        evalsort();
        update();
        history();
        it++;
    }

    /* ---- Householder tridiagonalisation, cma/cmaes.cpp:285-381 (EISPACK tred2).
     * V = B (row-major), d = eigenvalue work vector, e = tmp (off-diagonal). */
    void tridiagonalize() {
        double *V = B.data();
        double *e = tmp.data();
        auto at = [&](int r, int c) -> double& { return V[(size_t) r * n + c]; };
        for (int j = 0; j < n; j++) d[j] = at(n - 1, j);
        for (int i = n - 1; i > 0; i--) {
            double scale = 0.0, h = 0.0;
            for (int k = 0; k < i; k++) scale += std::fabs(d[k]);
            if (scale == 0.0) {
                e[i] = d[i - 1];
                for (int j = 0; j < i; j++) {
                    d[j] = at(i - 1, j);
                    at(i, j) = 0.0;
                    at(j, i) = 0.0;
                }
            } else {
                for (int k = 0; k < i; k++) {
                    d[k] /= scale;
                    h += d[k] * d[k];
                }
                double fq = d[i - 1];
                double g = std::sqrt(h);
                if (fq > 0) g = -g;
                e[i] = scale * g;
                h = h - fq * g;
                d[i - 1] = fq - g;
                for (int j = 0; j < i; j++) e[j] = 0.0;
                for (int j = 0; j < i; j++) {
                    fq = d[j];
                    at(j, i) = fq;
                    g = e[j] + at(j, j) * fq;
                    for (int k = j + 1; k <= i - 1; k++) {
                        g += at(k, j) * d[k];
                        e[k] += at(k, j) * fq;
                    }
                    e[j] = g;
                }
                const double invh = 1.0 / h;
                for (int k = 0; k < i; k++) e[k] *= invh;
                fq = 0.0;
                for (int k = 0; k < i; k++) fq = fq + e[k] * d[k];
                const double hh = fq / (h + h);
                if (-hh != 0.0) {
                    for (int k = 0; k < i; k++) e[k] += (-hh) * d[k];
                }
                for (int j = 0; j < i; j++) {
                    fq = d[j];
                    g = e[j];
                    for (int k = j; k <= i - 1; k++) at(k, j) -= (fq * e[k] + g * d[k]);
                    d[j] = at(i - 1, j);
                    at(i, j) = 0.0;
                }
            }
            d[i] = h;
        }
        for (int i = 0; i < n - 1; i++) {
            at(n - 1, i) = at(i, i);
            at(i, i) = 1.0;
            const double h = d[i + 1];
            if (h != 0.0) {
                for (int k = 0; k <= i; k++) d[k] = at(k, i + 1) / h;
                for (int j = 0; j <= i; j++) {
                    double g = 0.0;
                    for (int k = 0; k <= i; k++) g += at(k, i + 1) * at(k, j);
                    for (int k = 0; k <= i; k++) at(k, j) -= g * d[k];
                }
            }
            for (int k = 0; k <= i; k++) at(k, i + 1) = 0.0;
        }
        for (int j = 0; j < n; j++) {
            d[j] = at(n - 1, j);
            at(n - 1, j) = 0.0;
        }
        at(n - 1, n - 1) = 1.0;
        e[0] = 0.0;
    }

    /* ---- implicit QL with eigenvector accumulation + ascending sort,
     * cma/cmaes.cpp:383-478 (EISPACK tql2) */
    void ql_implicit() {
        double *V = B.data();
        double *e = tmp.data();
        auto at = [&](int r, int c) -> double& { return V[(size_t) r * n + c]; };
        for (int i = 1; i < n; i++) e[i - 1] = e[i];
        e[n - 1] = 0.0;
        double shift = 0.0, tst1 = 0.0;
        const double eps = std::pow(2.0, -52.0);
        for (int l = 0; l < n; l++) {
            tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
            int m = l;
            for (m = l; m < n; m++) {
                if (std::fabs(e[m]) <= eps * tst1) break;
            }
            if (m >= n) break;
            if (m > l) {
                do {
                    double g = d[l];
                    double p = (d[l + 1] - g) / (2.0 * e[l]);
                    double r = port_hypot(p, 1.0);
                    r = port_sign(r, p);
                    d[l] = e[l] / (p + r);
                    d[l + 1] = e[l] * (p + r);
                    const double dl1 = d[l + 1];
                    double h = g - d[l];
                    for (int i = l + 2; i < n; i++) d[i] -= h;
                    shift += h;
                    p = d[m];
                    double c = 1.0, c2 = c, c3 = c;
                    const double el1 = e[l + 1];
                    double s = 0.0, s2 = 0.0;
                    for (int i = m - 1; i >= l; i--) {
                        c3 = c2;
                        c2 = c;
                        s2 = s;
                        g = c * e[i];
                        h = c * p;
                        r = port_hypot(p, e[i]);
                        e[i + 1] = s * r;
                        s = e[i] / r;
                        c = p / r;
                        p = c * d[i] - s * g;
                        d[i + 1] = h + s * (c * g + s * d[i]);
                        for (int k = 0; k < n; k++) {
                            h = at(k, i + 1);
                            at(k, i + 1) = s * at(k, i) + c * h;
                            at(k, i) = c * at(k, i) - s * h;
                        }
                    }
                    p = -s * s2 * c3 * el1 * e[l] / dl1;
                    e[l] = s * p;
                    d[l] = c * p;
                } while (std::fabs(e[l]) > eps * tst1);
            }
            d[l] += shift;
            e[l] = 0.0;
        }
        for (int i = 0; i < n - 1; i++) {
            int k = i;
            double p = d[i];
            for (int j = i + 1; j < n; j++) {
                if (d[j] < p) {
                    k = j;
                    p = d[j];
                }
            }
            if (k != i) {
                d[k] = d[i];
                d[i] = p;
                for (int j = 0; j < n; j++) std::swap(at(j, i), at(j, k));
            }
        }
    }

    /* ---- cma/cmaes.cpp:229-283 */
    void eigen() {
        if (fev - eigenlastev <= eigenfreq) return;
        for (int i = 0; i < n; i++) {
            for (int j = 0; j <= i; j++) {
                B[(size_t) i * n + j] = B[(size_t) j * n + i] = C[(size_t) i * n + j];
            }
        }
        eigenlastev = fev;
        tridiagonalize();
        ql_implicit();
        if (d[0] <= 0.0) {
            for (int i = 0; i < n; i++) d[i] = std::max(d[i], 0.0);
            const double sh = d[n - 1] / 1e14;
            for (int i = 0; i < n; i++) {
                C[(size_t) i * n + i] += sh;
                d[i] += sh;
            }
        }
        if (d[n - 1] > 1e14 * d[0]) {
            const double sh = d[n - 1] / 1e14 - d[0];
            for (int i = 0; i < n; i++) {
                C[(size_t) i * n + i] += sh;
                d[i] += sh;
            }
        }
        for (int i = 0; i < n; i++) d[i] = std::sqrt(d[i]);
        for (int i = 0; i < n; i++) {
            for (int j = 0; j <= i; j++) {
                double sum = 0.0;
                for (int k = 0; k < n; k++) sum += B[(size_t) i * n + k] / d[k] * B[(size_t) j * n + k];
                isC[(size_t) i * n + j] = isC[(size_t) j * n + i] = sum;
            }
        }
        if (eiglog) eigevents.push_back(EigEvent { fev, n, B, d });
    }

    /* ---- cma/cmaes.cpp:151-227 and cma/sep_cmaes.cpp:137-207 */
    bool converged() {
        if (kind == 4) { /* cma/lm_cmaes.cpp:183-226 */
            if (it >= mit) return true;
            if (sigma < stolmin) return true;
            if (it >= hlen && fworst - fbest < tol) return true;
            if (best.len >= n && kth.len >= n) {
                int eq = 0;
                for (int i = 0; i < n; i++) {
                    if (best.get(i) == kth.get(i)) {
                        eq++;
                        if (3 * eq >= n) return true;
                    }
                }
            }
            return false;
        }
        if (kind == 3) { /* cma/cholesky_cmaes.cpp:137-161 */
            if (std::fabs(ybw[0] - ybw[3]) > tol) return false;
            int count = 0;
            double mean = 0.0, m2 = 0.0;
            for (int k = 0; k < lambda; k++) {
                const double x = port_nrm2(n, &arx[(size_t) k * n]);
                count++;
                const double delta = x - mean;
                mean += delta / count;
                const double delta2 = x - mean;
                m2 += delta * delta2;
            }
            return m2 <= (lambda - 1) * eigenrate * eigenrate;
        }
        if (it >= mit) { flag = 1; return true; }
        if (it >= hlen && fworst - fbest < tol) { flag = 2; return true; }
        if (best.len >= n && kth.len >= n) {
            int eq = 0;
            for (int i = 0; i < n; i++) {
                if (best.get(i) == kth.get(i)) {
                    eq++;
                    if (3 * eq >= n) { flag = 3; return true; }
                }
            }
        }
        bool all = true;
        for (int i = 0; i < n; i++) {
            const double spread = (kind == 2) ? d[i] : std::sqrt(C[(size_t) i * n + i]);
            if (std::max(pc[i], spread) * sigma / sigma0 >= tol) { all = false; break; }
        }
        if (all) { flag = 4; return true; }
        if (sigma / sigma0 > 1.0e20 * d[n - 1]) { flag = 5; return true; }
        if (d[n - 1] > 1.0e7 * d[0]) { flag = 7; return true; }
        const int iaxis = n - 1 - ((it - 1) % n);
        if (kind == 2) {
            if (xmean[iaxis] == xmean[iaxis] + 0.1 * sigma * d[iaxis]) { flag = 8; return true; }
            for (int i = 0; i < n; i++) {
                if (xmean[i] == xmean[i] + 0.2 * sigma * d[i]) { flag = 9; return true; }
            }
            return false;
        }
        all = true;
        for (int i = 0; i < n; i++) {
            /* reads ROW iaxis of B, as the reference does (cmaes.cpp:205-217) */
            if (xmean[i] != xmean[i] + 0.1 * sigma * d[iaxis] * B[(size_t) iaxis * n + i]) {
                all = false;
                break;
            }
        }
        if (all) { flag = 8; return true; }
        for (int i = 0; i < n; i++) {
            if (xmean[i] == xmean[i] + 0.2 * sigma * std::sqrt(C[(size_t) i * n + i])) {
                flag = 9;
                return true;
            }
        }
        return false;
    }

    /* ---- cma/base_cmaes.cpp:162-174,232-238 */
    bool optimize(const double *guess, std::vector<double> &xout) {
        init(guess);
        bool conv = false;
        while (fev < mfev) {
            iterate();
            if (converged()) { conv = true; break; }
        }
        best_solution(xout);
        return conv;
    }

    void best_solution(std::vector<double> &xout) const {
        if (it <= 0) xout = xmean;
        else xout.assign(arx.begin() + (size_t) ibw[0] * n, arx.begin() + (size_t) (ibw[0] + 1) * n);
    }
};

/* ---- nes/xnes.cpp.  Its tred2 / tql2 (:255-449) are the same EISPACK / JAMA routines as Cmaes's (cmaes.cpp:285-478, equal
 * line for line), so the eigensolver of PortCma is reused as a workspace. */
struct XnesPoint {
    double f;
    std::vector<double> x, z;
    static bool less(const XnesPoint &a, const XnesPoint &b) { return a.f < b.f; }
};

struct PortXnes {
    int mfev = 0, n = 0, np = 0, fev = 0;
    double tol = 0, a0 = 1, etamu = 1, etasigma = 0, etab = 0, sigma = 0, Gsigma = 0;
    Objective f;
    std::vector<double> lower, upper, u, mu, Gdelta, B, G, c;
    std::vector<XnesPoint> points;
    BboNormal Z;
    PortCma eig; /* eig.B = matrix in / eigenvectors out, eig.d = eigenvalues, eig.tmp = off-diagonal workspace */

    void init() { /* :50-108; the guess is not used (:82: `_mu` starts at the origin) */
        np = 4 + (int) (3.0 * std::log(n));
        etasigma = 3.0 * (3.0 + std::log(n)) / (5.0 * n * std::sqrt(n));
        etab = etasigma;
        u.assign(np, 0.0);
        double sumu = 0.0;
        for (int i = 1; i <= np; i++) {
            const double ui = std::max(0.0, std::log(1 + 0.5 * np) - std::log(i));
            u[i - 1] = ui;
            sumu += ui;
        }
        for (int i = 0; i < np; i++) u[i] = u[i] / sumu - (1.0 / np);
        mu.assign(n, 0.0);
        sigma = a0;
        B.assign((size_t) n * n, 0.0);
        for (int i = 0; i < n; i++) B[(size_t) i * n + i] = a0 / sigma;
        points.assign(np, XnesPoint { 0.0, std::vector<double>(n), std::vector<double>(n) });
        fev = 0;
        Gdelta.assign(n, 0.0);
        G.assign((size_t) n * n, 0.0);
        c.assign((size_t) n * n, 0.0);
        eig.n = n;
        eig.B.assign((size_t) n * n, 0.0);
        eig.d.assign(n, 0.0);
        eig.tmp.assign(n, 0.0);
    }

    void exponential() { /* :227-253 */
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++) eig.B[(size_t) i * n + j] = eig.B[(size_t) j * n + i] = c[(size_t) i * n + j];
        eig.tridiagonalize();
        eig.ql_implicit();
        for (int i = 0; i < n; i++) eig.d[i] = std::exp(eig.d[i]);
        for (int i = 0; i < n; i++) {
            for (int j = 0; j <= i; j++) {
                double sum = 0.0;
                for (int k = 0; k < n; k++) sum += eig.B[(size_t) i * n + k] * eig.d[k] * eig.B[(size_t) j * n + k];
                c[(size_t) i * n + j] = c[(size_t) j * n + i] = sum;
            }
        }
    }

    void iterate() { /* :110-191 */
        for (int i = 0; i < np; i++) {
            XnesPoint &p = points[i];
            for (int d = 0; d < n; d++) p.z[d] = Z.draw(eng());
            for (int d = 0; d < n; d++) {
                p.x[d] = mu[d];
                for (int j = 0; j < n; j++) p.x[d] += sigma * B[(size_t) d * n + j] * p.z[j];
            }
            p.f = f(p.x.data());
        }
        fev += np;
        std::sort(points.begin(), points.end(), XnesPoint::less);
        for (int d = 0; d < n; d++) {
            Gdelta[d] = 0.0;
            for (int i = 0; i < np; i++) Gdelta[d] += u[i] * points[i].z[d];
        }
        for (int d1 = 0; d1 < n; d1++) {
            for (int d2 = 0; d2 < n; d2++) {
                double g = 0.0;
                for (int i = 0; i < np; i++) {
                    const double z1 = points[i].z[d1], z2 = points[i].z[d2];
                    g += u[i] * (z1 * z2 - ((d1 == d2) ? 1.0 : 0.0));
                }
                G[(size_t) d1 * n + d2] = g;
            }
        }
        Gsigma = 0.0;
        for (int d = 0; d < n; d++) Gsigma += G[(size_t) d * n + d] / n;
        for (int d = 0; d < n; d++) G[(size_t) d * n + d] -= Gsigma * 1.0;
        for (int d = 0; d < n; d++)
            for (int j = 0; j < n; j++) mu[d] += etamu * sigma * B[(size_t) d * n + j] * Gdelta[j];
        sigma *= std::exp(0.5 * etasigma * Gsigma);
        for (size_t q = 0; q < (size_t) n * n; q++) c[q] = 0.5 * etab * G[q];
        exponential();
        for (int d1 = 0; d1 < n; d1++) {
            for (int d2 = 0; d2 < n; d2++) {
                double g = 0.0;
                for (int j = 0; j < n; j++) g += B[(size_t) d1 * n + j] * c[(size_t) j * n + d2];
                G[(size_t) d1 * n + d2] = g;
            }
        }
        B = G;
    }

    bool converged() const { return std::fabs(points[0].f - points[np - 1].f) < tol; } /* :211-219 */

    bool optimize(std::vector<double> &xout) { /* :197-209 */
        init();
        bool conv = false;
        while (fev < mfev) {
            iterate();
            if (converged()) { conv = true; break; }
        }
        xout = points[0].x;
        return conv;
    }
};

/* ---- de/sansde.cpp (Yang, Tang, Yao 2008: self-adaptive DE with neighbourhood search) */
struct PortSansde {
    int mfev = 0, np = 0, ncrref = 5, npup = 50, ncrup = 25;
    double tol = 0;
    bool repaircr = true;
    Objective f;
    int n = 0, it = 0, fev = 0;
    double p = 0.5, fp = 0.5, crrec = 0, crdeltaf = 0, crm = 0.5;
    int pns[2] = { 0, 0 }, pnf[2] = { 0, 0 };
    double fpns[2] = { 0, 0 }, fpnf[2] = { 0, 0 };
    std::vector<double> lower, upper, work, sw_x, sw_f, sw_cr;
    BboNormal Z;
    /* tagged draws + outcome of the last generation (the GPU engine replays the draws) */
    std::vector<double> t_zcr, t_ifstrat, t_F, t_r1, t_r2, t_r3, t_istrat, t_jrand, t_u, t_succ, t_ftrial, t_cr1;

    void init() { /* :54-92 */
        fev = 0;
        it = 0;
        p = 0.5;
        fp = 0.5;
        crrec = crdeltaf = 0.0;
        crm = 0.5;
        pns[0] = pns[1] = pnf[0] = pnf[1] = 0;
        fpns[0] = fpns[1] = fpnf[0] = fpnf[1] = 0.0;
        work.assign(n, 0.0);
        sw_x.assign((size_t) np * n, 0.0);
        sw_f.assign(np, 0.0);
        sw_cr.assign(np, 0.5);
        for (int i = 0; i < np; i++) {
            double *x = &sw_x[(size_t) i * n];
            for (int j = 0; j < n; j++) x[j] = eng().uniform_real(lower[j], upper[j]);
            sw_f[i] = f(x);
            fev++;
        }
    }

    void sort_swarm() { /* std::sort(_swarm, compare_fitness): whole points move, their CR with them */
        struct Rec { double f; int pos; };
        std::vector<Rec> rec(np);
        for (int i = 0; i < np; i++) rec[i] = Rec { sw_f[i], i };
        std::sort(rec.begin(), rec.end(), [](const Rec &a, const Rec &b) { return a.f < b.f; });
        std::vector<double> nx((size_t) np * n), nf(np), nc(np);
        for (int i = 0; i < np; i++) {
            nf[i] = rec[i].f;
            nc[i] = sw_cr[rec[i].pos];
            std::copy(sw_x.begin() + (size_t) rec[i].pos * n, sw_x.begin() + (size_t) (rec[i].pos + 1) * n, nx.begin() + (size_t) i * n);
        }
        sw_x.swap(nx);
        sw_f.swap(nf);
        sw_cr.swap(nc);
    }

    void iterate() { /* :94-212 */
        sort_swarm();
        for (auto *v : { &t_zcr, &t_ifstrat, &t_F, &t_r1, &t_r2, &t_r3, &t_istrat, &t_jrand, &t_succ, &t_ftrial, &t_cr1 }) v->assign(np, 0.0);
        t_u.assign((size_t) np * n, 0.0);
        for (int i = 0; i < np; i++) {
            t_zcr[i] = NAN;
            if (it % ncrref == 0) {
                const double z = Z.draw(eng());
                t_zcr[i] = z;
                sw_cr[i] = z * 0.1 + crm;
                sw_cr[i] = std::max(0.0, std::min(sw_cr[i], 1.0));
            }
            const int ifstrat = eng().uniform_real(0.0, 1.0) < fp ? 0 : 1;
            double F = -1.0;
            while (F < 0.0) {
                if (ifstrat == 0) F = Z.draw(eng()) * 0.3 + 0.5;
                else F = std::tan(M_PI * (eng().uniform_real(0.0, 1.0) - 0.5)); /* sampleCauchy, :288-290 */
                F = std::min(F, 1.0);
            }
            int r1 = i;
            while (r1 == i) r1 = eng().uniform_int(0, np - 1);
            int r2 = i;
            while (r2 == i || r2 == r1) r2 = eng().uniform_int(0, np - 1);
            int r3 = i;
            while (r3 == i || r3 == r1 || r3 == r2) r3 = eng().uniform_int(0, np - 1);
            const int istrat = eng().uniform_real(0.0, 1.0) < p ? 0 : 1;
            /* mutate(), :250-286 */
            const double *x = &sw_x[(size_t) i * n], *best = &sw_x[0], *x1 = &sw_x[(size_t) r1 * n], *x2 = &sw_x[(size_t) r2 * n],
                    *x3 = &sw_x[(size_t) r3 * n];
            if (istrat == 0) {
                for (int q = 0; q < n; q++) work[q] = x1[q] + F * (x2[q] - x3[q]);
            } else {
                for (int q = 0; q < n; q++) work[q] = x[q] + F * (best[q] - x[q]) + F * (x1[q] - x2[q]);
            }
            const int irand = eng().uniform_int(0, n - 1);
            double cr1 = 0.0;
            const double CR = sw_cr[i];
            for (int q = 0; q < n; q++) {
                const double u = eng().uniform_real(0.0, 1.0);
                t_u[(size_t) i * n + q] = u;
                if (u <= CR || q == irand) cr1 += 1.0;
                else work[q] = x[q];
            }
            cr1 = repaircr ? cr1 / n : CR;
            for (int q = 0; q < n; q++) {
                if (work[q] < lower[q]) work[q] = (lower[q] + x[q]) / 2.0;
                else if (work[q] > upper[q]) work[q] = (upper[q] + x[q]) / 2.0;
            }
            const double fwork = f(work.data());
            fev++;
            t_ifstrat[i] = ifstrat; t_F[i] = F; t_r1[i] = r1; t_r2[i] = r2; t_r3[i] = r3; t_istrat[i] = istrat; t_jrand[i] = irand;
            t_ftrial[i] = fwork; t_cr1[i] = cr1;
            if (fwork < sw_f[i]) {
                t_succ[i] = 1.0;
                pns[istrat]++;
                const double deltaf = sw_f[i] - fwork;
                crrec += cr1 * deltaf;
                crdeltaf += deltaf;
                fpns[ifstrat] += F;
                sw_f[i] = fwork;
                std::copy(work.begin(), work.end(), sw_x.begin() + (size_t) i * n);
            } else {
                pnf[istrat]++;
                fpnf[ifstrat] += F;
            }
        }
        it++;
        if (it > 0 && it % npup == 0) {
            const int pnsf0 = pns[0] + pnf[0], pnsf1 = pns[1] + pnf[1];
            p = (1.0 * pns[0] * pnsf1) / (pns[1] * pnsf0 + pns[0] * pnsf1);
            pns[0] = pns[1] = pnf[0] = pnf[1] = 0;
        }
        if (it > 0 && it % ncrup == 0) {
            if (crdeltaf > 0) crm = crrec / crdeltaf;
            crrec = crdeltaf = 0.0;
        }
        if (it > 0 && it % ncrup == 0) {
            const double fpnsf0 = fpns[0] + fpnf[0], fpnsf1 = fpns[1] + fpnf[1];
            fp = (1.0 * fpns[0] * fpnsf1) / (fpns[1] * fpnsf0 + fpns[0] * fpnsf1);
            fpns[0] = fpns[1] = fpnf[0] = fpnf[1] = 0.0;
        }
    }

    bool converged() { /* through solution(), which sorts first (:214-217,235-253) */
        sort_swarm();
        int count = 0;
        double mean = 0.0, m2 = 0.0;
        for (int i = 0; i < np; i++) {
            const double r = port_nrm2(n, &sw_x[(size_t) i * n]);
            count++;
            const double dl = r - mean;
            mean += dl / count;
            m2 += dl * (r - mean);
        }
        return m2 <= (np - 1) * tol * tol;
    }

    bool optimize(std::vector<double> &xout) { /* :219-233 */
        init();
        bool conv = false;
        while (fev < mfev) {
            iterate();
            /* the reference's private converged() does not sort; only the Welford order differs, which the statistic
             * does not depend on beyond rounding -- pinned against the reference's optimize() below */
            int count = 0;
            double mean = 0.0, m2 = 0.0;
            for (int i = 0; i < np; i++) {
                const double r = port_nrm2(n, &sw_x[(size_t) i * n]);
                count++;
                const double dl = r - mean;
                mean += dl / count;
                m2 += dl * (r - mean);
            }
            if (m2 <= (np - 1) * tol * tol) { conv = true; break; }
        }
        sort_swarm();
        xout.assign(sw_x.begin(), sw_x.begin() + n);
        return conv;
    }
};

/* ---- de/ssde.cpp (Zhao et al. 2022: spherical search with DE) */
struct PortSsde {
    int mfev = 0, npinit = 0, patience = 1000, npmin = 4, H = 100;
    double tol = 0, ptop = 0.11;
    bool usede = true, repaircr = true;
    Objective f;
    int n = 0, np = 0, fev = 0, convcount = 0, k = 1, kCR = 1;
    std::vector<int> iwork, partner; /* partner[d]: the index paired with d by randA(), -1 when unpaired */
    std::vector<double> lower, upper, b, z, y, work, L1, L2, LCR, SR, SC, SCR, w, wCR, sw_x, sw_f;
    double sinv = 0, cosv = 0;
    BboNormal Z;
    std::vector<double> t_iwork, t_iL, t_zrank, t_p, t_q, t_r, t_pbest, t_cauchy, t_ub, t_succ, t_f, t_zcr, t_j0, t_ucr, t_redraw,
            t_x0; /* the uniform points init() drew, before the opposition step and the sort */

    double cauchy() { return std::tan(M_PI * (eng().uniform_real(0.0, 1.0) - 0.5)); } /* :455-457 */

    void sort_swarm() { /* std::sort(_swarm, point::compare_fitness) */
        struct Rec { double f; int pos; };
        const int cnt = (int) sw_f.size();
        std::vector<Rec> rec(cnt);
        for (int i = 0; i < cnt; i++) rec[i] = Rec { sw_f[i], i };
        std::sort(rec.begin(), rec.end(), [](const Rec &a, const Rec &c) { return a.f < c.f; });
        std::vector<double> nx((size_t) cnt * n), nf(cnt);
        for (int i = 0; i < cnt; i++) {
            nf[i] = rec[i].f;
            std::copy(sw_x.begin() + (size_t) rec[i].pos * n, sw_x.begin() + (size_t) (rec[i].pos + 1) * n, nx.begin() + (size_t) i * n);
        }
        sw_x.swap(nx);
        sw_f.swap(nf);
    }

    void init() { /* :58-131 */
        fev = 0;
        convcount = 0;
        b.assign(n, 0.0); z.assign(n, 0.0); y.assign(n, 0.0); work.assign(n, 0.0);
        iwork.resize(n);
        for (int i = 0; i < n; i++) iwork[i] = i;
        partner.assign(n, -1);
        np = npinit;
        sw_x.assign((size_t) np * n, 0.0);
        sw_f.assign(np, 0.0);
        for (int i = 0; i < np; i++) {
            double *x = &sw_x[(size_t) i * n];
            for (int j = 0; j < n; j++) x[j] = eng().uniform_real(lower[j], upper[j]);
            sw_f[i] = f(x);
            ++fev;
        }
        t_x0 = sw_x;
        if (usede) { /* opposition-based points */
            sw_x.resize((size_t) 2 * np * n);
            sw_f.resize((size_t) 2 * np);
            for (int i = 0; i < np; i++) {
                double *x = &sw_x[(size_t) (np + i) * n];
                for (int j = 0; j < n; j++) x[j] = lower[j] + upper[j] - sw_x[(size_t) i * n + j];
                sw_f[np + i] = f(x);
                ++fev;
            }
        }
        sort_swarm();
        sw_x.resize((size_t) npinit * n);
        sw_f.resize(npinit);
        L1.assign(H, 0.5); L2.assign(H, 0.5); LCR.assign(H, 0.5);
        SR.clear(); SC.clear(); SCR.clear(); w.clear(); wCR.clear();
        k = 1;
        kCR = 1;
    }

    /* std::shuffle of libstdc++ 13 (bits/stl_algo.h:3742-3805) on the global engine: two swap positions per draw */
    void shuffle_iwork() {
        const uint64_t urange = (uint64_t) n;
        if (n == 0) return;
        if (0xffffffffull / urange >= urange) {
            int i = 1;
            if ((urange % 2) == 0) {
                const int pos = eng().uniform_int(0, 1);
                std::swap(iwork[i++], iwork[pos]);
            }
            while (i != n) {
                const uint64_t swap_range = (uint64_t) i + 1;
                const uint64_t b1 = swap_range + 1;
                const uint64_t x = (uint64_t) (uint32_t) eng().uniform_int(0, (int) (swap_range * b1 - 1));
                std::swap(iwork[i++], iwork[x / b1]);
                std::swap(iwork[i++], iwork[x % b1]);
            }
            return;
        }
        for (int i = 1; i < n; i++) std::swap(iwork[i], iwork[eng().uniform_int(0, i)]);
    }

    void rand_a() { /* :394-433: identity, then every consecutive pair of the shuffled indices becomes a rotation block */
        shuffle_iwork();
        partner.assign(n, -1);
        sinv = std::sin(1e-12);
        cosv = std::cos(1e-12);
        for (int ii = 1; ii <= n / 2; ii++) {
            const int i = 2 * (ii - 1);
            partner[iwork[i]] = iwork[i + 1];
            partner[iwork[i + 1]] = iwork[i];
        }
    }
    /* A[d][k] for the two structurally non-zero columns of row d */
    double a_entry(int d, int kk) const {
        if (partner[d] < 0) return d == kk ? 1.0 : 0.0;
        if (d == kk) return sinv;
        /* A[first][second] = cos, A[second][first] = -cos, first = the one earlier in iwork */
        return first_of_pair[d] ? cosv : -cosv;
    }
    std::vector<char> first_of_pair;

    void trial_point(int i, double ci) { /* :435-457, the dense loops restricted to the non-zero entries in column order */
        for (int d = 0; d < n; d++) {
            double s = 0.0;
            const int q = partner[d];
            if (q < 0) s = s + 1.0 * z[d];
            else {
                const int k1 = std::min(d, q), k2 = std::max(d, q);
                s = s + a_entry(d, k1) * z[k1];
                s = s + a_entry(d, k2) * z[k2];
            }
            work[d] = s * b[d];
        }
        const double *x = &sw_x[(size_t) i * n];
        for (int d = 0; d < n; d++) {
            double v = x[d];
            const int q = partner[d];
            if (q < 0) v = v + ci * 1.0 * work[d];
            else {
                const int k1 = std::min(d, q), k2 = std::max(d, q);
                v = v + ci * a_entry(k1, d) * work[k1];
                v = v + ci * a_entry(k2, d) * work[k2];
            }
            y[d] = std::max(lower[d], std::min(upper[d], v));
        }
    }

    void iterate() { /* :134-339 */
        const double oldbestf = sw_f[0];
        rand_a();
        first_of_pair.assign(n, 0);
        for (int ii = 1; ii <= n / 2; ii++) first_of_pair[iwork[2 * (ii - 1)]] = 1;
        t_iwork.assign(iwork.begin(), iwork.end());
        SR.clear(); SC.clear(); SCR.clear(); w.clear(); wCR.clear();
        for (auto *v : { &t_iL, &t_zrank, &t_p, &t_q, &t_r, &t_pbest, &t_cauchy, &t_succ, &t_f, &t_zcr, &t_j0 }) v->assign(np, 0.0);
        t_ub.assign((size_t) np * n, 0.0);
        t_ucr.assign(usede ? (size_t) np * n : 0, NAN);
        t_redraw.assign(usede ? (size_t) np * n : 0, NAN);
        for (int i = 0; i < np; i++) {
            const int iL = eng().uniform_int(0, H - 1);
            const double zr = Z.draw(eng());
            double prank = zr * 0.1 + L1[iL];
            prank = std::max(0.0, std::min(1.0, prank));
            for (int r = 0; r < n; r++) {
                const double u = eng().uniform_real(0.0, 1.0);
                t_ub[(size_t) i * n + r] = u;
                b[r] = u < prank ? 1.0 : 0.0;
            }
            int pi = i, qi = i, ri = i;
            while (pi == i) pi = eng().uniform_int(0, np - 1);
            while (qi == i || qi == pi) qi = eng().uniform_int(0, np - 1);
            while (ri == i || ri == pi || ri == qi) ri = eng().uniform_int(0, np - 1);
            const int itop = std::max(1, (int) (ptop * np));
            const int pbest = eng().uniform_int(0, itop);
            const double R = (1.0 * fev) / mfev;
            const double *xp = &sw_x[(size_t) pi * n], *xq = &sw_x[(size_t) qi * n], *xr = &sw_x[(size_t) ri * n],
                    *xi = &sw_x[(size_t) i * n], *xb = &sw_x[(size_t) pbest * n], *x0 = &sw_x[0];
            if (usede) {
                const double *lead = R < 0.333 ? xp : (R < 0.666 ? xb : x0);
                for (int d = 0; d < n; d++) z[d] = lead[d] + xq[d] - xr[d] - xi[d] + R * (xb[d] - xq[d]);
            } else {
                const double *lead = i < 0.5 * np ? xp : xb;
                for (int d = 0; d < n; d++) z[d] = lead[d] + xq[d] - xr[d] - xi[d];
            }
            double cv = cauchy();
            double ci = cv * 0.1 + L2[iL];
            while (ci <= 0.0) {
                cv = cauchy();
                ci = cv * 0.1 + L2[iL];
            }
            ci = std::min(1.0, ci);
            trial_point(i, ci);
            const double fy = f(y.data());
            ++fev;
            t_iL[i] = iL; t_zrank[i] = zr; t_p[i] = pi; t_q[i] = qi; t_r[i] = ri; t_pbest[i] = pbest; t_cauchy[i] = cv; t_f[i] = fy;
            if (fy <= sw_f[i]) {
                t_succ[i] = 1.0;
                if (fy < sw_f[i]) {
                    SR.push_back(prank);
                    SC.push_back(ci);
                    w.push_back(sw_f[i] - fy);
                }
                std::copy(y.begin(), y.end(), sw_x.begin() + (size_t) i * n);
                sw_f[i] = fy;
            } else if (usede) {
                const double zc = Z.draw(eng());
                const double CRi = std::max(0.0, std::min(1.0, zc * 0.1 + LCR[iL]));
                double cr1 = 0.0;
                const int j0 = eng().uniform_int(0, n - 1);
                t_zcr[i] = zc; t_j0[i] = j0;
                for (int d = 0; d < n; d++) {
                    bool take = d == j0;
                    if (!take) {
                        const double u = eng().uniform_real(0.0, 1.0);
                        t_ucr[(size_t) i * n + d] = u;
                        take = u <= CRi;
                    }
                    if (take) {
                        work[d] = xp[d] + R * (x0[d] - xq[d]) + R * (x0[d] - xr[d]);
                        if (work[d] < lower[d] || work[d] > upper[d]) {
                            work[d] = eng().uniform_real(lower[d], upper[d]);
                            t_redraw[(size_t) i * n + d] = work[d];
                        }
                        cr1 += 1.0;
                    } else {
                        work[d] = xi[d];
                    }
                }
                cr1 = repaircr ? cr1 / n : CRi;
                const double fu = f(work.data());
                ++fev;
                t_f[i] = fu;
                if (fu <= sw_f[i]) {
                    t_succ[i] = 2.0;
                    if (fu < sw_f[i]) {
                        SCR.push_back(cr1);
                        wCR.push_back(sw_f[i] - fu);
                    }
                    std::copy(work.begin(), work.end(), sw_x.begin() + (size_t) i * n);
                    sw_f[i] = fu;
                }
            }
        }
        const int lenS = (int) SR.size();
        if (lenS > 0) {
            double rn = 0.0, rd = 0.0, cn = 0.0, cd = 0.0;
            for (int i = 0; i < lenS; i++) {
                rn += w[i] * SR[i] * SR[i];
                rd += w[i] * SR[i];
                cn += w[i] * SC[i] * SC[i];
                cd += w[i] * SC[i];
            }
            L1[k - 1] = rn / rd;
            L2[k - 1] = cn / cd;
            ++k;
            if (k > H) k = 1;
        }
        const int lenSCR = (int) SCR.size();
        if (usede && lenSCR > 0) {
            double cn = 0.0, cd = 0.0;
            for (int i = 0; i < lenSCR; i++) {
                cn += wCR[i] * SCR[i] * SCR[i];
                cd += wCR[i] * SCR[i];
            }
            LCR[kCR - 1] = cn / cd;
            ++kCR;
            if (kCR > H) kCR = 1;
        }
        sort_swarm();
        if (sw_f[0] < oldbestf) convcount = 0;
        else ++convcount;
        const int npnew = (int) (npinit + (npmin - npinit) * (1.0 * fev) / mfev);
        while (np > npnew) --np;
        sw_x.resize((size_t) np * n);
        sw_f.resize(np);
    }

    bool converged() const { /* :370-392 */
        int count = 0;
        double mean = 0.0, m2 = 0.0;
        for (int i = 0; i < np; i++) {
            const double r = port_nrm2(n, &sw_x[(size_t) i * n]);
            count++;
            const double dl = r - mean;
            mean += dl / count;
            m2 += dl * (r - mean);
        }
        if (m2 <= (np - 1) * tol * tol) return true;
        return convcount > patience;
    }

    bool optimize(std::vector<double> &xout) { /* :346-360 */
        init();
        bool conv = false;
        while (fev < mfev) {
            iterate();
            if (converged()) { conv = true; break; }
        }
        sort_swarm();
        xout.assign(sw_x.begin(), sw_x.begin() + n);
        return conv;
    }
};

void cma_bind(PortCma *p, int obj, int n, const double *lower, const double *upper) {
    p->n = n;
    p->f.bind(obj, n);
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
}

int put(const std::vector<double> &v, double *out, int cap) {
    for (int i = 0; i < (int) v.size() && i < cap; i++) out[i] = v[i];
    return (int) v.size();
}

/* ======================================================================================
 * DE family: JADE (de/jade.cpp) and L-SHADE (de/shade.cpp), with the tagged-draw recorder.
 * ==================================================================================== */
struct PortDe {
    int kind, mfev, npinit, h, npmin;
    double tol, pelite, cdamp, sigma_cr;
    bool archive, repaircr;
    Objective f;
    int n = 0, np = 0, fev = 0, kslot = 1;
    double mucr = 0.5, muf = 0.5;
    std::vector<double> lower, upper, work, MCR, MF;
    std::vector<double> sw_x, sw_f;           /* np x n, np */
    std::vector<std::vector<double>> arch;
    BboNormal Z;
    /* tagged draws of the last generation (what the GPU engine replays) */
    std::vector<double> t_zcr, t_cauchy, t_ibest, t_r1, t_r2, t_jrand, t_slot, t_succ, t_u,
            t_ri, t_pi, t_trim, t_F, t_CR, t_ftrial;

    double cauchy() { return std::tan(M_PI * (eng().uniform_real(0.0, 1.0) - 0.5)); } /* jade.cpp:292-294 */

    /* std::sort(_swarm, point::compare_fitness): the permutation produced by introsort
     * depends only on comparison outcomes, so sorting (f, position) records is equivalent */
    void sort_swarm() {
        struct Rec { double f; int pos; };
        std::vector<Rec> rec(np);
        for (int i = 0; i < np; i++) rec[i] = Rec { sw_f[i], i };
        std::sort(rec.begin(), rec.end(), [](const Rec &a, const Rec &b) { return a.f < b.f; });
        std::vector<double> nx((size_t) np * n), nf(np);
        for (int i = 0; i < np; i++) {
            nf[i] = rec[i].f;
            std::copy(sw_x.begin() + (size_t) rec[i].pos * n, sw_x.begin() + (size_t) (rec[i].pos + 1) * n,
                    nx.begin() + (size_t) i * n);
        }
        sw_x.swap(nx);
        sw_f.swap(nf);
    }

    /* ---- de/jade.cpp:64-96, de/shade.cpp:56-94 */
    void init() {
        fev = 0;
        mucr = muf = 0.5;
        MCR.assign(h, 0.5);
        MF.assign(h, 0.5);
        work.assign(n, 0.0);
        arch.clear();
        kslot = 1;
        np = npinit;
        sw_x.assign((size_t) np * n, 0.0);
        sw_f.assign(np, 0.0);
        for (int i = 0; i < np; i++) {
            double *x = &sw_x[(size_t) i * n];
            for (int j = 0; j < n; j++) x[j] = eng().uniform_real(lower[j], upper[j]);
            sw_f[i] = f(x);
            fev++;
        }
        if (kind == 1) sort_swarm();
    }

    void clear_trace(int cnt) {
        for (auto *v : { &t_zcr, &t_cauchy, &t_ibest, &t_r1, &t_r2, &t_jrand, &t_slot, &t_succ,
                &t_ri, &t_pi, &t_F, &t_CR, &t_ftrial }) v->assign(cnt, 0.0);
        t_u.assign((size_t) cnt * n, 0.0);
        t_trim.clear();
    }

    /* ---- de/jade.cpp:247-264 == de/shade.cpp:277-294 */
    double mutate(int i, const double *x, const double *bestx, const double *xr1,
            const double *xr2, double F, double CR) {
        const int jr = eng().uniform_int(0, n - 1);
        t_jrand[i] = jr;
        double cr1 = 0.0;
        for (int j = 0; j < n; j++) {
            bool cross = (j == jr);
            if (!cross) {
                const double u = eng().uniform_real(0.0, 1.0);
                t_u[(size_t) i * n + j] = u;
                cross = u < CR;
            }
            if (cross) {
                work[j] = x[j] + F * (bestx[j] - x[j]) + F * (xr1[j] - xr2[j]);
                cr1 += 1.0;
            } else {
                work[j] = x[j];
            }
        }
        return repaircr ? cr1 / n : CR;
    }

    /* ---- de/jade.cpp:98-206 and de/shade.cpp:96-236 */
    void iterate() {
        if (kind == 0) sort_swarm();
        clear_trace(np);
        std::vector<double> scr, sf, sw;
        for (int i = 0; i < np; i++) {
            double CR, F;
            int nelite;
            if (kind == 0) {
                const double z = Z.draw(eng());
                t_zcr[i] = z;
                CR = z * 0.1 + mucr;
                CR = std::max(0.0, std::min(CR, 1.0));
                F = -1.0;
                double c = 0.0;
                while (F < 0.0) {
                    c = cauchy();
                    F = c * 0.1 + muf;
                    F = std::min(F, 1.0);
                }
                t_cauchy[i] = c;
                nelite = std::max(1, (int) (pelite * np));
            } else {
                const int ri = eng().uniform_int(0, h - 1);
                t_ri[i] = ri;
                const double z = Z.draw(eng());
                t_zcr[i] = z;
                CR = std::max(0.0, std::min(z * 0.1 + MCR[ri], 1.0));
                double c = cauchy();
                F = std::min(1.0, c * 0.1 + MF[ri]);
                while (F <= 0) {
                    c = cauchy();
                    F = std::min(1.0, c * 0.1 + MF[ri]);
                }
                t_cauchy[i] = c;
                const double pi = eng().uniform_real(std::min(2.0 / n, 0.2), 0.2);
                t_pi[i] = pi;
                nelite = std::max(1, (int) (pi * np));
            }
            t_F[i] = F;
            t_CR[i] = CR;
            const int ibest = eng().uniform_int(0, nelite - 1);
            int r1 = i;
            while (r1 == i) r1 = eng().uniform_int(0, np - 1);
            int r2 = i;
            const int larch = (int) arch.size();
            while (r2 == i || r2 == r1) r2 = eng().uniform_int(0, np + larch - 1);
            t_ibest[i] = ibest;
            t_r1[i] = r1;
            t_r2[i] = r2;
            t_slot[i] = -1;
            double *xi = &sw_x[(size_t) i * n];
            const double *x2 = (r2 >= np) ? arch[r2 - np].data() : &sw_x[(size_t) r2 * n];
            const double cr1 = mutate(i, xi, &sw_x[(size_t) ibest * n], &sw_x[(size_t) r1 * n], x2, F, CR);
            for (int j = 0; j < n; j++) {
                if (work[j] < lower[j]) work[j] = (lower[j] + xi[j]) / 2.0;
                else if (work[j] > upper[j]) work[j] = (upper[j] + xi[j]) / 2.0;
            }
            const double fw = f(work.data());
            t_ftrial[i] = fw;
            fev++;
            if (fw <= sw_f[i]) {
                t_succ[i] = 1;
                const bool strict = fw < sw_f[i];
                if (archive && (kind == 0 || strict)) {
                    if (larch >= np) {
                        const int slot = eng().uniform_int(0, np - 1);
                        t_slot[i] = slot;
                        std::copy(xi, xi + n, arch[slot].begin());
                    } else {
                        arch.emplace_back(xi, xi + n);
                    }
                }
                if (kind == 0) {
                    scr.push_back(cr1);
                    sf.push_back(F);
                } else if (strict) {
                    t_succ[i] = 2;
                    scr.push_back(cr1);
                    sf.push_back(F);
                    sw.push_back(sw_f[i] - fw);
                }
                sw_f[i] = fw;
                std::copy(work.begin(), work.end(), xi);
            }
        }
        const int ns = (int) scr.size();
        if (kind == 0) {
            /* jade.cpp:186-205 with meanPow :266-277 and std :279-290 (running means) */
            double meancr = 0.0, meanf = 0.0;
            if (ns > 0) {
                double mean = 0.0, m2 = 0.0;
                for (int i = 0; i < ns; i++) {
                    const double dl = scr[i] - mean;
                    mean += dl / (i + 1.0);
                    m2 += dl * (scr[i] - mean);
                }
                const double sd = std::sqrt(m2 / ns);
                double p1 = 0.0, p2 = 0.0;
                for (int i = 0; i < ns; i++) {
                    p1 += (scr[i] - p1) / (i + 1.0);
                    p2 += (scr[i] * scr[i] - p2) / (i + 1.0);
                }
                meancr = (sd > sigma_cr) ? std::sqrt(p2) : p1;
                double f1 = 0.0, f2 = 0.0;
                for (int i = 0; i < ns; i++) {
                    f1 += (sf[i] - f1) / (i + 1.0);
                    f2 += (sf[i] * sf[i] - f2) / (i + 1.0);
                }
                meanf = f2 / f1;
            }
            mucr = (1.0 - cdamp) * mucr + cdamp * meancr;
            muf = (1.0 - cdamp) * muf + cdamp * meanf;
            return;
        }
        /* shade.cpp:186-235 */
        if (ns > 0) {
            double crn = 0.0, crd = 0.0, fn = 0.0, fd = 0.0;
            for (int i = 0; i < ns; i++) {
                crn += sw[i] * scr[i];
                crd += sw[i];
                fn += sw[i] * sf[i] * sf[i];
                fd += sw[i] * sf[i];
            }
            MCR[kslot - 1] = crn / crd;
            MF[kslot - 1] = fn / fd;
            kslot++;
            if (kslot > h) kslot = 1;
        }
        sort_swarm();
        const int npnew = (int) std::round((npmin - npinit) * ((1.0 * fev) / mfev) + npinit);
        if (npnew < np) {
            sw_x.resize((size_t) npnew * n);
            sw_f.resize(npnew);
            np = npnew;
        }
        if (archive) {
            int larch = (int) arch.size();
            while (larch > npnew) {
                const int r = eng().uniform_int(0, larch - 1);
                t_trim.push_back(r);
                arch.erase(arch.begin() + r);
                --larch;
            }
        }
    }

    /* ---- de/jade.cpp:228-245 == de/shade.cpp:258-275 */
    bool converged() const {
        int count = 0;
        double mean = 0.0, m2 = 0.0;
        for (int i = 0; i < np; i++) {
            const double r = port_nrm2(n, &sw_x[(size_t) i * n]);
            count++;
            const double dl = r - mean;
            mean += dl / count;
            m2 += dl * (r - mean);
        }
        return m2 <= (np - 1) * tol * tol;
    }

    bool optimize(std::vector<double> &xout) {
        init();
        bool conv = false;
        while (fev < mfev) {
            iterate();
            if (converged()) { conv = true; break; }
        }
        sort_swarm();
        xout.assign(sw_x.begin(), sw_x.begin() + n);
        return conv;
    }
};

} // namespace

extern "C" {

const char *bbo_kind(void) { return "port"; }

void bbo_seed(uint32_t s) {
    g_eng.seed(s);
    g_seeded = true;
    g_test_normal.reset();
}

void bbo_rng_draw(int kind, int count, double a, double b, double *out) {
    for (int i = 0; i < count; i++) {
        switch (kind) {
        case 0: out[i] = g_test_normal.draw(eng()); break;
        case 1: out[i] = eng().uniform_real(a, b); break;
        case 2: out[i] = (double) eng().uniform_int((int) a, (int) b); break;
        default: out[i] = (double) eng().next(); break;
        }
    }
}

double bbo_eval(int obj, int n, const double *x) {
    Objective f;
    f.bind(obj, n);
    return f(x);
}

void *bbo_cma_new(int kind, int mfev, double tol, int np, double sigma0, int bound,
        double eigenrate, double alphacov, int adjustlr) {
    PortCma *p = new PortCma();
    p->kind = kind;
    p->mfev = mfev;
    p->tol = tol;
    p->lambda = np;
    p->sigma0 = sigma0;
    p->bound = bound != 0;
    p->eigenrate = eigenrate;
    p->alphacov = alphacov;
    p->adjustlr = adjustlr != 0;
    if (kind == 4) { /* LmCmaes(mfev, tol, np, memory, sigma0, bound, rademacher, usenew), lm_cmaes.h:62-70 */
        p->memsize = (int) eigenrate;
        p->lm_adaptmemory = p->memsize <= 0;
        p->lm_samplemode = alphacov != 0.0 ? 1 : 0;
        p->lm_new = adjustlr != 0;
    }
    return p;
}

void bbo_cma_free(void *h) { delete static_cast<PortCma*>(h); }

void bbo_cma_init(void *h, int obj, int n, const double *lower, const double *upper,
        const double *guess) {
    PortCma *p = static_cast<PortCma*>(h);
    cma_bind(p, obj, n, lower, upper);
    p->init(guess);
}

void bbo_cma_set_params(void *h, int np, double sigma, int mfev) {
    static_cast<PortCma*>(h)->set_params(np, sigma, mfev);
}
void bbo_cma_sample(void *h) { static_cast<PortCma*>(h)->sample(); }
void bbo_cma_evalsort(void *h) { static_cast<PortCma*>(h)->evalsort(); }
void bbo_cma_update(void *h) { static_cast<PortCma*>(h)->update(); }
void bbo_cma_history(void *h) {
    PortCma *p = static_cast<PortCma*>(h);
    p->history();
    p->it++;
}
void bbo_cma_iterate(void *h) { static_cast<PortCma*>(h)->iterate(); }
int bbo_cma_converged(void *h) { return static_cast<PortCma*>(h)->converged() ? 1 : 0; }

int bbo_cma_optimize(void *h, int obj, int n, const double *lower, const double *upper,
        const double *guess, double *xout, int *fev) {
    PortCma *p = static_cast<PortCma*>(h);
    cma_bind(p, obj, n, lower, upper);
    std::vector<double> x;
    const bool conv = p->optimize(guess, x);
    for (int i = 0; i < n; i++) xout[i] = x[i];
    *fev = p->fev;
    return conv ? 1 : 0;
}

int bbo_cma_geti(void *h, const char *name) {
    PortCma *p = static_cast<PortCma*>(h);
    const std::string k(name);
    if (k == "n") return p->n;
    if (k == "lambda") return p->lambda;
    if (k == "mu") return p->mu;
    if (k == "it") return p->it;
    if (k == "fev") return p->fev;
    if (k == "mfev") return p->mfev;
    if (k == "mit") return p->mit;
    if (k == "hlen") return p->hlen;
    if (k == "ik") return p->ik;
    if (k == "bound") return p->bound ? 1 : 0;
    if (k == "best_len") return p->best.len;
    if (k == "flag") return p->flag;
    if (k == "eigenlastev") return p->eigenlastev;
    if (k == "memlen") return p->memlen;
    if (k == "memsize") return p->memsize;
    if (k == "nsteps") return p->nsteps;
    if (k == "t") return p->lm_t;
    return -1;
}

double bbo_cma_getd(void *h, const char *name) {
    PortCma *p = static_cast<PortCma*>(h);
    const std::string k(name);
    if (k == "sigma") return p->sigma;
    if (k == "sigma0") return p->sigma0;
    if (k == "tol") return p->tol;
    if (k == "mueff") return p->mueff;
    if (k == "cc") return p->cc;
    if (k == "cs") return p->cs;
    if (k == "c1") return p->c1;
    if (k == "cmu") return p->cmu;
    if (k == "damps") return p->damps;
    if (k == "chi") return p->chi;
    if (k == "fbest") return p->fbest;
    if (k == "fworst") return p->fworst;
    if (k == "eigenfreq") return p->eigenfreq;
    if (k == "cneg") return p->cneg;
    if (k == "cm") return p->cm;
    if (k == "alphaold") return p->alphaold;
    if (k == "ccov") return p->ccov;
    if (k == "stol") return p->eigenrate;
    if (k == "s") return p->lm_s;
    if (k == "ccc") return p->lm_ccc;
    if (k == "sqrt1mc1") return p->sqrt1mc1;
    return NAN;
}

int bbo_cma_getv(void *h, const char *name, double *out, int cap) {
    PortCma *p = static_cast<PortCma*>(h);
    const std::string k(name);
    if (k == "xmean") return put(p->xmean, out, cap);
    if (k == "xold") return put(p->xold, out, cap);
    if (k == "pc") return put(p->pc, out, cap);
    if (k == "ps") return put(p->ps, out, cap);
    if (k == "weights") return put(p->weights, out, cap);
    if (k == "arx") return put(p->arx, out, cap);
    if (k == "z") return put(p->last_z, out, cap);
    if (k == "hist_best") return put(p->best.v, out, cap);
    if (k == "hist_kth") return put(p->kth.v, out, cap);
    if (k == "D") return put(p->d, out, cap);
    if (k == "C") return put(p->C, out, cap);
    if (k == "B" && p->kind != 2) return put(p->B, out, cap);
    if (k == "invsqrtC" && p->kind != 2) return put(p->isC, out, cap);
    if (k == "ycoeff" && p->kind == 1) return put(p->ycoeff, out, cap);
    if (p->kind == 4) {
        if (k == "pcmat") return put(p->pcmat, out, cap);
        if (k == "vmat") return put(p->vmat, out, cap);
        if (k == "b") return put(p->lm_b, out, cap);
        if (k == "d") return put(p->lm_d, out, cap);
        if (k == "fp") return put(p->fp, out, cap);
        if (k == "i0") return put(p->last_i0, out, cap);
        if (k == "jarr" || k == "larr") {
            const std::vector<int> &v = k == "jarr" ? p->jarr : p->larr;
            for (int i = 0; i < (int) v.size() && i < cap; i++) out[i] = v[i];
            return (int) v.size();
        }
    }
    if (k == "A" && p->kind == 3) return put(p->A, out, cap);
    if (k == "dmean" && p->kind == 3) return put(p->dmean, out, cap);
    if (k == "ybw") {
        for (int i = 0; i < 4 && i < cap; i++) out[i] = p->ybw[i];
        return 4;
    }
    if (k == "ibw") {
        for (int i = 0; i < 4 && i < cap; i++) out[i] = p->ibw[i];
        return 4;
    }
    if (k == "fit_val") {
        for (int i = 0; i < p->lambda && i < cap; i++) out[i] = p->rank[i].value;
        return p->lambda;
    }
    if (k == "fit_idx") {
        for (int i = 0; i < p->lambda && i < cap; i++) out[i] = p->rank[i].index;
        return p->lambda;
    }
    return -1;
}

void bbo_cma_reset_z(void *h) { static_cast<PortCma*>(h)->Z.reset(); }

void bbo_cma_eig_log(void *h, int enable) {
    PortCma *p = static_cast<PortCma*>(h);
    p->eigevents.clear();
    p->eiglog = enable != 0;
}

int bbo_cma_eig_log_count(void *h) { return (int) static_cast<PortCma*>(h)->eigevents.size(); }

int bbo_cma_eig_log_get(void *h, int k, int *fev, double *B, double *D) {
    PortCma *p = static_cast<PortCma*>(h);
    if (k < 0 || k >= (int) p->eigevents.size()) return 0;
    const PortCma::EigEvent &e = p->eigevents[k];
    if (fev) *fev = e.fev;
    if (B) std::copy(e.B.begin(), e.B.end(), B);
    if (D) std::copy(e.D.begin(), e.D.begin() + e.n, D);
    return e.n;
}

double bbo_cma_time_eig(void *h) {
    PortCma *p = static_cast<PortCma*>(h);
    if (p->kind == 2) return 0.0;
    p->eigenlastev = p->fev - static_cast<int>(p->eigenfreq) - 1;
    const auto t0 = std::chrono::steady_clock::now();
    p->eigen();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* ---- restart drivers: cma/ipop_cmaes.cpp:65-189 and cma/bipop_cmaes.cpp:61-267 */
int bbo_restart_optimize(int kind, void *base, int mfev, double sigma0, int flag,
        double ksigmadec, int boundlambda, int maxlargeruns, double kbudget, int obj, int n,
        const double *lower, const double *upper, const double *guess, double *xout, int *fevout,
        double *trace, int cap) {
    PortCma *b = static_cast<PortCma*>(base);
    cma_bind(b, obj, n, lower, upper);
    int rows = 0;
    auto log_row = [&](int lam, double sg, double fx, int fevt, int regime) {
        if (trace && rows < cap) {
            double *r = trace + (size_t) rows * 5;
            r[0] = lam; r[1] = sg; r[2] = fx; r[3] = fevt; r[4] = regime;
        }
        rows++;
    };
    int fev = 0;
    auto max_evals = [&](int lam) {
        const int maxit = (int) (100.0 + 50.0 * (n + 3) * (n + 3) / std::sqrt(1.0 * lam));
        return std::min(maxit * lam, mfev - fev);
    };
    std::vector<double> x, xbest, x0(n);
    double fx, fbest;
    const int lambdadef = 4 + (int) (3.0 * std::log(1.0 * n));
    auto run = [&](int lam, double sg, int budget, const double *start) {
        b->set_params(lam, sg, budget);
        b->optimize(start, x);
        fx = b->f(x.data());
        fev += b->fev + 1;
        return b->fev;
    };
    if (kind == 0) {
        /* IPOP / NIPOP */
        int lambda = lambdadef;
        const int lambdamax = 10 * n * n;
        double sigma = sigma0;
        run(lambdadef, sigma0, max_evals(lambdadef), guess);
        fbest = fx;
        xbest = x;
        log_row(lambdadef, sigma0, fx, fev, 0);
        while (fev < mfev) {
            for (int i = 0; i < n; i++) x0[i] = eng().uniform_real(lower[i], upper[i]);
            if (boundlambda) {
                lambda <<= 1;
                if (lambda > lambdamax) {
                    if (lambda - lambdamax < lambdamax - (lambda >> 1)) lambda = lambdamax;
                    else lambda = lambdadef;
                }
            } else {
                lambda <<= 1;
            }
            if (flag) {
                sigma /= ksigmadec;
                sigma = std::max(sigma, 0.01 * sigma0);
            }
            run(lambda, sigma, max_evals(lambda), x0.data());
            if (fx < fbest) {
                fbest = fx;
                xbest = x;
            }
            log_row(lambda, sigma, fx, fev, 0);
        }
    } else {
        /* BIPOP / NBIPOP */
        run(lambdadef, sigma0, max_evals(lambdadef), guess);
        int largebudget = 0, smallbudget = 0, largerestarts = 0, bestregime = 1;
        int largelambda = 0;
        fbest = fx;
        xbest = x;
        log_row(lambdadef, sigma0, fx, fev, 0);
        while (true) {
            for (int i = 0; i < n; i++) x0[i] = eng().uniform_real(lower[i], upper[i]);
            int regime;
            if (flag) {
                if (bestregime == 1) regime = (largebudget <= smallbudget * kbudget) ? 1 : 2;
                else regime = (smallbudget <= kbudget * largebudget) ? 2 : 1;
            } else {
                regime = (largebudget <= smallbudget) ? 1 : 2;
            }
            if (regime == 1) {
                largelambda = (int) (lambdadef * std::pow(2, largerestarts + 1));
                double largesigma = sigma0;
                if (flag) {
                    largesigma = sigma0 * std::pow(1.0 / ksigmadec, largerestarts + 1);
                    largesigma = std::max(largesigma, 0.01 * sigma0);
                }
                const int used = run(largelambda, largesigma, max_evals(largelambda), x0.data());
                if (fx < fbest) { fbest = fx; xbest = x; bestregime = 1; }
                largebudget += used;
                largerestarts++;
                log_row(largelambda, largesigma, fx, fev, 1);
            } else {
                const double u = eng().uniform_real(0.0, 1.0);
                const int smalllambda = (int) (lambdadef * std::pow((0.5 * largelambda) / lambdadef, u * u));
                const double smallsigma = sigma0 * std::pow(10.0, -2.0 * eng().uniform_real(0.0, 1.0));
                int budget = max_evals(smalllambda);
                budget = std::min(budget, largebudget >> 1);
                const int used = run(smalllambda, smallsigma, budget, x0.data());
                if (fx < fbest) { fbest = fx; xbest = x; bestregime = 2; }
                smallbudget += used;
                log_row(smalllambda, smallsigma, fx, fev, 2);
            }
            if (largerestarts >= maxlargeruns) break;
            if (fev >= mfev) break;
        }
    }
    for (int i = 0; i < n; i++) xout[i] = xbest[i];
    *fevout = fev;
    return rows;
}

void *bbo_xnes_new(int mfev, double tol, double a0, double etamu) {
    PortXnes *p = new PortXnes();
    p->mfev = mfev;
    p->tol = tol;
    p->a0 = a0;
    p->etamu = etamu;
    return p;
}
void bbo_xnes_free(void *h) { delete static_cast<PortXnes*>(h); }

static void xnes_bind(PortXnes *p, int obj, int n, const double *lower, const double *upper) {
    p->n = n;
    p->f.bind(obj, n);
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
}

void bbo_xnes_init(void *h, int obj, int n, const double *lower, const double *upper, const double *guess) {
    (void) guess;
    PortXnes *p = static_cast<PortXnes*>(h);
    xnes_bind(p, obj, n, lower, upper);
    p->init();
}
void bbo_xnes_iterate(void *h) { static_cast<PortXnes*>(h)->iterate(); }
int bbo_xnes_converged(void *h) { return static_cast<PortXnes*>(h)->converged() ? 1 : 0; }

int bbo_xnes_optimize(void *h, int obj, int n, const double *lower, const double *upper, const double *guess, double *xout,
        int *fev) {
    (void) guess;
    PortXnes *p = static_cast<PortXnes*>(h);
    xnes_bind(p, obj, n, lower, upper);
    std::vector<double> x;
    const bool conv = p->optimize(x);
    for (int i = 0; i < n; i++) xout[i] = x[i];
    *fev = p->fev;
    return conv ? 1 : 0;
}

int bbo_xnes_geti(void *h, const char *name) {
    PortXnes *p = static_cast<PortXnes*>(h);
    const std::string k(name);
    if (k == "n") return p->n;
    if (k == "np") return p->np;
    if (k == "fev") return p->fev;
    return -1;
}

double bbo_xnes_getd(void *h, const char *name) {
    PortXnes *p = static_cast<PortXnes*>(h);
    const std::string k(name);
    if (k == "sigma") return p->sigma;
    if (k == "etasigma") return p->etasigma;
    if (k == "etab") return p->etab;
    if (k == "etamu") return p->etamu;
    if (k == "Gsigma") return p->Gsigma;
    return NAN;
}

int bbo_xnes_getv(void *h, const char *name, double *out, int cap) {
    PortXnes *p = static_cast<PortXnes*>(h);
    const std::string k(name);
    if (k == "u") return put(p->u, out, cap);
    if (k == "mu") return put(p->mu, out, cap);
    if (k == "B") return put(p->B, out, cap);
    if (k == "Gdelta") return put(p->Gdelta, out, cap);
    if (k == "x" || k == "z") {
        int len = 0;
        for (const XnesPoint &pt : p->points)
            for (double v : (k == "x" ? pt.x : pt.z)) {
                if (len < cap) out[len] = v;
                len++;
            }
        return len;
    }
    if (k == "f") {
        for (int i = 0; i < (int) p->points.size() && i < cap; i++) out[i] = p->points[i].f;
        return (int) p->points.size();
    }
    return -1;
}

void *bbo_sansde_new(int mfev, int np, double tol, int repaircr, int crref, int pupdate, int crupdate) {
    PortSansde *p = new PortSansde();
    p->mfev = mfev; p->np = np; p->tol = tol; p->repaircr = repaircr != 0;
    p->ncrref = crref; p->npup = pupdate; p->ncrup = crupdate;
    return p;
}
void bbo_sansde_free(void *h) { delete static_cast<PortSansde*>(h); }

static void sansde_bind(PortSansde *p, int obj, int n, const double *lower, const double *upper) {
    p->n = n;
    p->f.bind(obj, n);
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
}

void bbo_sansde_init(void *h, int obj, int n, const double *lower, const double *upper, const double *guess) {
    (void) guess;
    PortSansde *p = static_cast<PortSansde*>(h);
    sansde_bind(p, obj, n, lower, upper);
    p->init();
}
void bbo_sansde_iterate(void *h) { static_cast<PortSansde*>(h)->iterate(); }
int bbo_sansde_converged(void *h) { return static_cast<PortSansde*>(h)->converged() ? 1 : 0; }

int bbo_sansde_optimize(void *h, int obj, int n, const double *lower, const double *upper, const double *guess, double *xout,
        int *fev) {
    (void) guess;
    PortSansde *p = static_cast<PortSansde*>(h);
    sansde_bind(p, obj, n, lower, upper);
    std::vector<double> x;
    const bool conv = p->optimize(x);
    for (int i = 0; i < n; i++) xout[i] = x[i];
    *fev = p->fev;
    return conv ? 1 : 0;
}

int bbo_sansde_geti(void *h, const char *name) {
    PortSansde *p = static_cast<PortSansde*>(h);
    const std::string k(name);
    if (k == "n") return p->n;
    if (k == "np") return p->np;
    if (k == "fev") return p->fev;
    if (k == "it") return p->it;
    if (k == "pns0") return p->pns[0];
    if (k == "pns1") return p->pns[1];
    if (k == "pnf0") return p->pnf[0];
    if (k == "pnf1") return p->pnf[1];
    return -1;
}

double bbo_sansde_getd(void *h, const char *name) {
    PortSansde *p = static_cast<PortSansde*>(h);
    const std::string k(name);
    if (k == "p") return p->p;
    if (k == "fp") return p->fp;
    if (k == "crm") return p->crm;
    if (k == "crrec") return p->crrec;
    if (k == "crdeltaf") return p->crdeltaf;
    if (k == "fpns0") return p->fpns[0];
    if (k == "fpns1") return p->fpns[1];
    if (k == "fpnf0") return p->fpnf[0];
    if (k == "fpnf1") return p->fpnf[1];
    return NAN;
}

int bbo_sansde_getv(void *h, const char *name, double *out, int cap) {
    PortSansde *p = static_cast<PortSansde*>(h);
    const std::string k(name);
    if (k == "swarm_x") return put(p->sw_x, out, cap);
    if (k == "swarm_f") return put(p->sw_f, out, cap);
    if (k == "swarm_cr") return put(p->sw_cr, out, cap);
    const std::pair<const char*, const std::vector<double>*> tags[] = { { "t_zcr", &p->t_zcr }, { "t_ifstrat", &p->t_ifstrat },
            { "t_F", &p->t_F }, { "t_r1", &p->t_r1 }, { "t_r2", &p->t_r2 }, { "t_r3", &p->t_r3 }, { "t_istrat", &p->t_istrat },
            { "t_jrand", &p->t_jrand }, { "t_u", &p->t_u }, { "t_succ", &p->t_succ }, { "t_ftrial", &p->t_ftrial },
            { "t_cr1", &p->t_cr1 } };
    for (const auto &t : tags)
        if (k == t.first) return put(*t.second, out, cap);
    return -1;
}

void *bbo_ssde_new(int mfev, int npinit, double tol, int patience, int npmin, double ptop, int h, int usede, int repaircr) {
    PortSsde *p = new PortSsde();
    p->mfev = mfev; p->npinit = npinit; p->tol = tol; p->patience = patience; p->npmin = npmin; p->ptop = ptop; p->H = h;
    p->usede = usede != 0; p->repaircr = repaircr != 0;
    return p;
}
void bbo_ssde_free(void *h) { delete static_cast<PortSsde*>(h); }

static void ssde_bind(PortSsde *p, int obj, int n, const double *lower, const double *upper) {
    p->n = n;
    p->f.bind(obj, n);
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
}

void bbo_ssde_init(void *h, int obj, int n, const double *lower, const double *upper, const double *guess) {
    (void) guess;
    PortSsde *p = static_cast<PortSsde*>(h);
    ssde_bind(p, obj, n, lower, upper);
    p->init();
}
void bbo_ssde_iterate(void *h) { static_cast<PortSsde*>(h)->iterate(); }
int bbo_ssde_converged(void *h) { return static_cast<PortSsde*>(h)->converged() ? 1 : 0; }

int bbo_ssde_optimize(void *h, int obj, int n, const double *lower, const double *upper, const double *guess, double *xout,
        int *fev) {
    (void) guess;
    PortSsde *p = static_cast<PortSsde*>(h);
    ssde_bind(p, obj, n, lower, upper);
    std::vector<double> x;
    const bool conv = p->optimize(x);
    for (int i = 0; i < n; i++) xout[i] = x[i];
    *fev = p->fev;
    return conv ? 1 : 0;
}

int bbo_ssde_geti(void *h, const char *name) {
    PortSsde *p = static_cast<PortSsde*>(h);
    const std::string k(name);
    if (k == "n") return p->n;
    if (k == "np") return p->np;
    if (k == "fev") return p->fev;
    if (k == "convcount") return p->convcount;
    if (k == "k") return p->k;
    if (k == "kcr") return p->kCR;
    return -1;
}

double bbo_ssde_getd(void *, const char *) { return NAN; }

int bbo_ssde_getv(void *h, const char *name, double *out, int cap) {
    PortSsde *p = static_cast<PortSsde*>(h);
    const std::string k(name);
    if (k == "swarm_x") return put(p->sw_x, out, cap);
    if (k == "swarm_f") return put(p->sw_f, out, cap);
    if (k == "L1") return put(p->L1, out, cap);
    if (k == "L2") return put(p->L2, out, cap);
    if (k == "LCR") return put(p->LCR, out, cap);
    if (k == "iwork") {
        for (int i = 0; i < (int) p->iwork.size() && i < cap; i++) out[i] = p->iwork[i];
        return (int) p->iwork.size();
    }
    const std::pair<const char*, const std::vector<double>*> tags[] = { { "t_iwork", &p->t_iwork }, { "t_iL", &p->t_iL },
            { "t_zrank", &p->t_zrank }, { "t_p", &p->t_p }, { "t_q", &p->t_q }, { "t_r", &p->t_r }, { "t_pbest", &p->t_pbest },
            { "t_cauchy", &p->t_cauchy }, { "t_ub", &p->t_ub }, { "t_succ", &p->t_succ }, { "t_f", &p->t_f }, { "t_zcr", &p->t_zcr },
            { "t_j0", &p->t_j0 }, { "t_ucr", &p->t_ucr }, { "t_redraw", &p->t_redraw }, { "t_x0", &p->t_x0 } };
    for (const auto &t : tags)
        if (k == t.first) return put(*t.second, out, cap);
    return -1;
}

void *bbo_de_new(int kind, int mfev, int np, double tol, int archive, int repaircr,
        double pelite, double cdamp, double sigma, int h, int npmin) {
    PortDe *p = new PortDe();
    p->kind = kind;
    p->mfev = mfev;
    p->npinit = np;
    p->tol = tol;
    p->archive = archive != 0;
    p->repaircr = repaircr != 0;
    p->pelite = pelite;
    p->cdamp = cdamp;
    p->sigma_cr = sigma;
    p->h = (kind == 1) ? h : 1;
    p->npmin = npmin;
    return p;
}

void bbo_de_free(void *h) { delete static_cast<PortDe*>(h); }

static void de_bind(PortDe *p, int obj, int n, const double *lower, const double *upper) {
    p->n = n;
    p->f.bind(obj, n);
    p->lower.assign(lower, lower + n);
    p->upper.assign(upper, upper + n);
}

void bbo_de_init(void *h, int obj, int n, const double *lower, const double *upper,
        const double *guess) {
    (void) guess; /* unused by the reference too (jade.cpp:64-96) */
    PortDe *p = static_cast<PortDe*>(h);
    de_bind(p, obj, n, lower, upper);
    p->init();
}

void bbo_de_iterate(void *h) { static_cast<PortDe*>(h)->iterate(); }
int bbo_de_converged(void *h) { return static_cast<PortDe*>(h)->converged() ? 1 : 0; }

int bbo_de_optimize(void *h, int obj, int n, const double *lower, const double *upper,
        const double *guess, double *xout, int *fev) {
    (void) guess;
    PortDe *p = static_cast<PortDe*>(h);
    de_bind(p, obj, n, lower, upper);
    std::vector<double> x;
    const bool conv = p->optimize(x);
    for (int i = 0; i < n; i++) xout[i] = x[i];
    *fev = p->fev;
    return conv ? 1 : 0;
}

int bbo_de_geti(void *h, const char *name) {
    PortDe *p = static_cast<PortDe*>(h);
    const std::string k(name);
    if (k == "n") return p->n;
    if (k == "np") return p->np;
    if (k == "fev") return p->fev;
    if (k == "larch") return (int) p->arch.size();
    if (k == "k") return p->kslot;
    return -1;
}

double bbo_de_getd(void *h, const char *name) {
    PortDe *p = static_cast<PortDe*>(h);
    const std::string k(name);
    if (k == "mucr") return p->mucr;
    if (k == "muf") return p->muf;
    return NAN;
}

int bbo_de_getv(void *h, const char *name, double *out, int cap) {
    PortDe *p = static_cast<PortDe*>(h);
    const std::string k(name);
    if (k == "swarm_x") return put(p->sw_x, out, cap);
    if (k == "swarm_f") return put(p->sw_f, out, cap);
    if (k == "arch") {
        int len = 0;
        for (const auto &row : p->arch)
            for (double v : row) {
                if (len < cap) out[len] = v;
                len++;
            }
        return len;
    }
    if (k == "MCR" && p->kind == 1) return put(p->MCR, out, cap);
    if (k == "MF" && p->kind == 1) return put(p->MF, out, cap);
    /* tagged draws of the generation just run (port only) */
    if (k == "t_zcr") return put(p->t_zcr, out, cap);
    if (k == "t_cauchy") return put(p->t_cauchy, out, cap);
    if (k == "t_ibest") return put(p->t_ibest, out, cap);
    if (k == "t_r1") return put(p->t_r1, out, cap);
    if (k == "t_r2") return put(p->t_r2, out, cap);
    if (k == "t_jrand") return put(p->t_jrand, out, cap);
    if (k == "t_slot") return put(p->t_slot, out, cap);
    if (k == "t_succ") return put(p->t_succ, out, cap);
    if (k == "t_u") return put(p->t_u, out, cap);
    if (k == "t_ri") return put(p->t_ri, out, cap);
    if (k == "t_pi") return put(p->t_pi, out, cap);
    if (k == "t_trim") return put(p->t_trim, out, cap);
    if (k == "t_F") return put(p->t_F, out, cap);
    if (k == "t_CR") return put(p->t_CR, out, cap);
    if (k == "t_ftrial") return put(p->t_ftrial, out, cap);
    return -1;
}

} // extern "C"
