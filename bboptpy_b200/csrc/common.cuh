// Shared plumbing for libbbgpu.so: error reporting, the context object, small device helpers.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bbgpu.h"

namespace bb {

// thread-local last-error string behind bbgpu_last_error()
inline std::string &last_error() {
    static thread_local std::string e;
    return e;
}

inline int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

#define BB_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return bb::fail(BBGPU_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,   \
                            cudaGetErrorString(e__));                                      \
    } while (0)

#define BB_TRY(call)                                                                       \
    do {                                                                                   \
        int s__ = (call);                                                                  \
        if (s__ != BBGPU_OK) return s__;                                                   \
    } while (0)

#define BB_ARG(cond)                                                                       \
    do {                                                                                   \
        if (!(cond)) return bb::fail(BBGPU_ERR_ARG, "%s:%d invalid argument: %s", __FILE__, \
                                     __LINE__, #cond);                                     \
    } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int ceil_div(int64_t a, int64_t b) { return (int) ((a + b - 1) / b); }

} // namespace bb

struct cusolverDnContext;

// One device + one stream.  Everything the library launches goes through launches++ so the
// bench can report how many of OUR kernels ran in the timed region.
struct bbgpu_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    cusolverDnContext *solver = nullptr; // created lazily by the eig path
    // scratch: pinned host staging + device staging, grown on demand
    void *h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
    void *d_scratch = nullptr;
    size_t d_scratch_bytes = 0;
    // measurement
    bool profiling = false;
    struct ProfRec { int tag; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof;
    double prof_ms[16] = { 0 };
    int64_t prof_n[16] = { 0 };
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    void *d_flush = nullptr;
    void *d_ranks = nullptr;        // persistent zeroed counting-rank buffer (rank.cuh)
    cudaStream_t aux = nullptr;     // side stream: next generation's draws overlap the covariance update
    void *dc_tables = nullptr;      // node tables of the tridiagonal divide & conquer (stedc.cu): a function of n alone
    size_t dc_tables_bytes = 0;
    int dc_tables_n = 0;

    int pinned(size_t bytes, void **out);
    int scratch(size_t bytes, void **out);
};

namespace bb {

// after every kernel launch: count it and surface launch-configuration errors early
inline int launched(bbgpu_ctx *ctx, const char *what) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(BBGPU_ERR_CUDA, "launch %s -> %s", what, cudaGetErrorString(e));
    return BBGPU_OK;
}

// NVTX range names of the phases (bbgpu_profile_tag order): visible in nsys / ncu --nvtx timelines
inline const char *prof_tag_name(int tag) {
    static const char *names[] = { "bbgpu:sample_gemm", "bbgpu:cov_syrk", "bbgpu:norm_gemm", "bbgpu:evaluate", "bbgpu:rank",
            "bbgpu:philox", "bbgpu:eig", "bbgpu:update_misc", "bbgpu:sep_sample", "bbgpu:de_round", "bbgpu:de_misc", "bbgpu:comm",
            "bbgpu:cmab_step" };
    return tag >= 0 && tag < (int) (sizeof(names) / sizeof(names[0])) ? names[tag] : "bbgpu:other";
}

// one phase of the path: an NVTX range around its launches (always; a no-op without an attached tool) and, in
// profiling mode, device-time accounting with CUDA events
struct ProfScope {
    bbgpu_ctx *ctx;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int tag;
    ProfScope(bbgpu_ctx *c, int t) : ctx(c), tag(t) {
        nvtxRangePushA(prof_tag_name(t));
        if (!ctx->profiling) return;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, ctx->stream);
    }
    ~ProfScope() {
        nvtxRangePop();
        if (!e0) return;
        cudaEventRecord(e1, ctx->stream);
        ctx->prof.push_back({ tag, e0, e1 });
    }
};

template<typename T> int dev_alloc(T **p, size_t count) {
    *p = nullptr;
    if (count == 0) return BBGPU_OK;
    cudaError_t e = cudaMalloc((void**) p, count * sizeof(T));
    if (e != cudaSuccess) return fail(BBGPU_ERR_CUDA, "cudaMalloc(%zu bytes) -> %s", count * sizeof(T),
                                      cudaGetErrorString(e));
    return BBGPU_OK;
}

// ---- device helpers ------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
    // xor butterfly 16,8,4,2,1: every lane ends with the same bits (a+b == b+a in IEEE)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// a / b correctly rounded from r = RN(1 / b) in three instructions (Markstein 1990: q = RN(a r), then
// q' = RN(q + RN(a - b q) r) is the correctly rounded quotient for every b whose significand is not all ones; that one
// significand falls back to the division).  The kernels that divide a whole row by sigma were issue-bound on the ~25
// instructions of a generic fp64 division; the result bits are the reference's `(x - xold) / sigma` either way.
__device__ __forceinline__ double div_by(double a, double b, double r) {
    const double q = __dmul_rn(a, r);
    return __fma_rn(__fma_rn(-b, q, a), r, q);
}
__device__ __forceinline__ bool div_by_ok(double b) {
    return ((unsigned long long) __double_as_longlong(b) & 0x000fffffffffffffull) != 0x000fffffffffffffull && isfinite(b) && b != 0.0;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// deterministic block sum (blockDim.x multiple of 32, <= 1024); result valid in all threads
__device__ __forceinline__ double block_sum(double v, double *smem32) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem32[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = (lane < nw) ? smem32[lane] : 0.0;
    t = warp_sum(t);
    return t;
}

__device__ __forceinline__ double block_max(double v, double *smem32) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) smem32[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = (lane < nw) ? smem32[lane] : -INFINITY;
    t = warp_max(t);
    return t;
}
#endif

} // namespace bb
