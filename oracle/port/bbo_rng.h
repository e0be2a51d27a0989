/*
 * ORACLE (test infrastructure).  Restatement of the random-number stack the reference uses.
 *
 * The reference draws everything from ONE process-global std::mt19937 through
 * effolkronium::random_static (/root/reference/src/random.hpp:677-680; get(Dist&) :639-642,
 * get(int,int) :311-318, get(double,double) :329-336).  The VALUES are defined by a
 * third-party dependency that is not under /root/reference: libstdc++ <random> of the system
 * GCC (13.3.0 here; unpinned by the reference's setup.py).  Published algorithms restated:
 *   mt19937                    Matsumoto & Nishimura 1998 (32-bit, n=624, m=397)
 *   generate_canonical<53>     two 32-bit draws, low word first:  (a + b*2^32) / 2^64,
 *                              clamped below 1            (bits/random.tcc, generate_canonical)
 *   uniform_real(a,b)          canonical*(b-a)+a
 *   normal_distribution        Marsaglia polar method with ONE cached value; returns y*mult
 *                              first and keeps x*mult       (bits/random.tcc, normal_distribution)
 *   uniform_int(a,b)           Lemire's nearly-divisionless method on 64-bit products; consumes
 *                              an engine draw even for a 1-element range (bits/uniform_int_dist.h)
 * Pinned bit-for-bit against the real library by tests/test_oracle_pinning.py (through
 * oracle/_ref/libbbo_ref.so) and by tests/golden/rng_*.json.
 */
#ifndef BBO_RNG_H_
#define BBO_RNG_H_

#include <cmath>
#include <cstdint>

struct BboMt19937 {
    uint32_t mt[624];
    int idx;

    void seed(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; i++) {
            mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t) i;
        }
        idx = 624;
    }

    void refill() {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
            uint32_t v = mt[(i + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            mt[i] = v;
        }
        idx = 0;
    }

    uint32_t next() {
        if (idx >= 624) refill();
        uint32_t y = mt[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }

    /* generate_canonical<double,53> */
    double canonical() {
        double lo = (double) next();
        double hi = (double) next();
        double sum = lo + hi * 4294967296.0;
        double ret = sum / 18446744073709551616.0;
        if (ret >= 1.0) ret = std::nextafter(1.0, 0.0);
        return ret;
    }

    double uniform_real(double a, double b) {
        return canonical() * (b - a) + a;
    }

    /* closed range [a, b] */
    int uniform_int(int a, int b) {
        uint32_t urange = (uint32_t) b - (uint32_t) a;
        if (urange == 0xffffffffu) return (int) (next() + (uint32_t) a);
        uint32_t erange = urange + 1u;
        uint64_t product = (uint64_t) next() * (uint64_t) erange;
        uint32_t low = (uint32_t) product;
        if (low < erange) {
            uint32_t threshold = (0u - erange) % erange;
            while (low < threshold) {
                product = (uint64_t) next() * (uint64_t) erange;
                low = (uint32_t) product;
            }
        }
        return (int) ((uint32_t) (product >> 32) + (uint32_t) a);
    }
};

/* std::normal_distribution<double>(0,1): stateful (one cached variate) */
struct BboNormal {
    bool has_saved = false;
    double saved = 0.0;

    void reset() { has_saved = false; }

    double draw(BboMt19937 &g) {
        double ret;
        if (has_saved) {
            has_saved = false;
            ret = saved;
        } else {
            double x, y, r2;
            do {
                x = 2.0 * g.canonical() - 1.0;
                y = 2.0 * g.canonical() - 1.0;
                r2 = x * x + y * y;
            } while (r2 > 1.0 || r2 == 0.0);
            const double mult = std::sqrt(-2.0 * std::log(r2) / r2);
            saved = x * mult;
            has_saved = true;
            ret = y * mult;
        }
        return ret * 1.0 + 0.0;
    }
};

#endif
