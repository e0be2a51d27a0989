// Accuracy of philox.cuh's Box-Muller kernels against long double (host path always; device path when a GPU is present).
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/normal_accuracy.cu -o tools/bin/normal_accuracy
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../bboptpy_b200/csrc/philox.cuh"

static double ulps(double got, long double want) {
    if (want == 0.0L) return std::fabs(got) / 4.9e-324;
    const double w = (double) want;
    const double ulp = std::nextafter(std::fabs(w), INFINITY) - std::fabs(w);
    return (double) (fabsl((long double) got - want) / ulp);
}

__global__ void pairs_kernel(const uint32_t *c, int count, double *z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const bb::Philox4 r { c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3] };
    bb::normal_pair(r, z[2 * i], z[2 * i + 1]);
}

int main() {
    std::mt19937_64 g(1);
    const long double PI = 3.14159265358979323846264338327950288L;
    double worst_log = 0, worst_log_abs = 0, worst_s = 0, worst_c = 0, worst_abs_sc = 0;
    const int N = 4000000;
    for (int k = 0; k < N; k++) {
        uint32_t a = (uint32_t) g(), b = (uint32_t) g();
        if (k < 64) { a = 0xffffffffu; b = 0xffffffffu - (uint32_t) k * 64u; }      // u1 near 2^-53
        else if (k < 128) { a = 0; b = (uint32_t) (k - 64) * 64u; }                  // u1 near 1
        else if (k < 4096) { a = (uint32_t) ((k * 2654435761u) >> (k % 27)); }       // small u53 -> u1 close to 1 at many scales
        const double u1 = 1.0 - bb::u53(a, b), u2 = bb::u53(b, a);
        const double t = bb::neg2log_unit(u1);
        const long double tw = -2.0L * logl((long double) u1);
        if (u1 <= 0.999) worst_log = std::fmax(worst_log, ulps(t, tw));
        else worst_log_abs = std::fmax(worst_log_abs, std::fabs((double) (t - tw)));
        if (u1 == 1.0 && t != 0.0) { printf("t(1) = %g\n", t); return 1; }
        double s, c;
        bb::sincos_2pi_unit(u2, s, c);
        const long double sw = sinl(2.0L * PI * (long double) u2), cw = cosl(2.0L * PI * (long double) u2);
        worst_abs_sc = std::fmax(worst_abs_sc, std::fmax(std::fabs((double) (s - sw)), std::fabs((double) (c - cw))));
        if (fabsl(sw) > 1e-3L) worst_s = std::fmax(worst_s, ulps(s, sw));
        if (fabsl(cw) > 1e-3L) worst_c = std::fmax(worst_c, ulps(c, cw));
    }
    printf("host: -2 ln u max %.2f ulp for u <= 0.999, max abs error %.3g for u in (0.999, 1]; sin %.2f ulp, cos %.2f ulp (|value| > 1e-3), max abs error %.3g\n", worst_log, worst_log_abs, worst_s, worst_c, worst_abs_sc);
    int dev = 0;
    if (cudaGetDeviceCount(&dev) != cudaSuccess || dev == 0) { printf("no GPU: device path not checked\n"); return worst_log < 4 && worst_log_abs < 1e-18 && worst_abs_sc < 4e-16 ? 0 : 1; }
    const int M = 1 << 20;
    std::vector<uint32_t> c(4 * M);
    for (auto &v : c) v = (uint32_t) g();
    uint32_t *dc; double *dz;
    cudaMalloc(&dc, c.size() * 4); cudaMalloc(&dz, 2 * M * 8);
    cudaMemcpy(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice);
    pairs_kernel<<<(M + 255) / 256, 256>>>(dc, M, dz);
    std::vector<double> z(2 * M);
    if (cudaMemcpy(z.data(), dz, z.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("device run failed\n"); return 1; }
    double worst_rel = 0, worst_abs = 0, sum = 0, sum2 = 0;
    for (int i = 0; i < M; i++) {
        const long double u1 = 1.0L - (long double) bb::u53(c[4 * i], c[4 * i + 1]), u2 = bb::u53(c[4 * i + 2], c[4 * i + 3]);
        const long double rad = sqrtl(-2.0L * logl(u1));
        const long double w0 = rad * cosl(2.0L * PI * u2), w1 = rad * sinl(2.0L * PI * u2);
        const double e0 = std::fabs((double) (z[2 * i] - w0)), e1 = std::fabs((double) (z[2 * i + 1] - w1));
        worst_abs = std::fmax(worst_abs, std::fmax(e0, e1));
        if (fabsl(w0) > 1e-3L) worst_rel = std::fmax(worst_rel, e0 / (double) fabsl(w0));
        if (fabsl(w1) > 1e-3L) worst_rel = std::fmax(worst_rel, e1 / (double) fabsl(w1));
        sum += z[2 * i] + z[2 * i + 1];
        sum2 += z[2 * i] * z[2 * i] + z[2 * i + 1] * z[2 * i + 1];
    }
    printf("device: %d pairs, max abs error %.3g, max rel error %.3g (|z| > 1e-3), mean %.5f, var %.5f\n", M, worst_abs, worst_rel,
            sum / (2.0 * M), sum2 / (2.0 * M));
    return worst_abs < 5e-15 && worst_rel < 2e-15 ? 0 : 1;
}
