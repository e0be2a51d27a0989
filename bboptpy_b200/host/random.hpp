// Host random layer of the accelerated path.
//
// The reference draws everything from one process-global std::mt19937 behind
// effolkronium::random_static (/root/reference/src/random.hpp:677-680) and exposes no seed.
// Here the per-candidate draws move to a device Philox stream (csrc/philox.cuh), and the host
// keeps
//   * Random::seed(s)         one global seed: keys the Philox streams and the host engine;
//   * Random::engine()        a host std::mt19937 for the restart drivers' few scalar draws
//                             (ipop_cmaes.cpp:115-117, bipop_cmaes.cpp:112-114,241,244);
//   * the DRAW-REPLAY hook    ReplayStream: a host mt19937 + libstdc++ distributions consumed in
//                             the reference's own draw order (cmaes.cpp:66-69: candidate-major),
//                             so an optimizer in replay mode sees bit-for-bit the draws the
//                             reference sees after effolkronium::random_static::seed(s).
#ifndef BBOPT_HOST_RANDOM_HPP_
#define BBOPT_HOST_RANDOM_HPP_

#include <cstdint>
#include <random>
#include <vector>

namespace bbopt {

class Random {
public:
    static std::mt19937 &engine() {
        static std::mt19937 e(initial_seed());
        return e;
    }
    static uint64_t &philox_seed() {
        static uint64_t s = ((uint64_t) initial_seed() << 32) ^ 0x9E3779B97F4A7C15ull;
        return s;
    }
    // seeds BOTH the host engine and the device streams (and puts the process in a
    // reproducible state, which the reference cannot be from Python)
    static void seed(uint32_t s) {
        engine().seed(s);
        philox_seed() = ((uint64_t) s << 32) ^ 0x9E3779B97F4A7C15ull;
        counter() = 0;
    }
    // every optimizer instance takes a distinct Philox key derived from the global seed
    static uint64_t next_stream_key() { return philox_seed() + 0xD1B54A32D192ED03ull * (++counter()); }
    static double get(double a, double b) { return std::uniform_real_distribution<double>(a, b)(engine()); }
    static int get(int a, int b) { return std::uniform_int_distribution<int>(a, b)(engine()); }

private:
    static uint32_t initial_seed() {
        std::random_device rd;
        return rd();
    }
    static uint64_t &counter() {
        static uint64_t c = 0;
        return c;
    }
};

// Draw-replay source: reproduces the reference's N(0,1) stream for the CMA-ES family.
// `normal` is stateful like the reference's public `_Z` member (cmaes.h:48-49): its cached
// Box-Muller/polar value survives init() and IPOP/BIPOP restarts.
struct ReplayStream {
    bool enabled = false;
    bool use_global = false; // share Random::engine() (needed when restart drivers also draw)
    std::mt19937 own;
    std::normal_distribution<double> normal { 0., 1. };

    void enable(uint32_t seed) {
        enabled = true;
        use_global = false;
        own.seed(seed);
        normal.reset();
    }
    void enable_global() {
        enabled = true;
        use_global = true;
    }
    std::mt19937 &eng() { return use_global ? Random::engine() : own; }
    void fill_normals(std::vector<double> &z) {
        std::mt19937 &e = eng();
        for (double &v : z) v = normal(e);
    }
};

} // namespace bbopt

#endif
