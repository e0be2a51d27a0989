// Fitness ranking: replaces std::sort(_fitness, cmaes_index::compare_fitness)
// (cma/base_cmaes.cpp:221) and std::sort(_swarm, point::compare_fitness) (de/jade.cpp:101).
//
// Keys are the fp64 fitness mapped to order-preserving u64, paired with the candidate index and
// compared as a (key, index) tuple: a TOTAL order, so the permutation is deterministic.  The
// reference's std::sort is unstable (tie order is implementation defined); parity cases are
// tie-free and tests assert that.  NaN sorts last.
//
// lambda <= 16384: one CTA, keys + indices live in shared memory (12 B x 16384 = 192 KB).
// larger: the same bitonic network over global memory, one launch per (k, j) stage.
#pragma once

#include "common.cuh"

namespace bb {

__device__ __forceinline__ unsigned long long sortable_key(double f) {
    const long long b = __double_as_longlong(f);
    return (unsigned long long) (b ^ ((b >> 63) | (long long) 0x8000000000000000ull));
}

__device__ __forceinline__ bool tuple_greater(unsigned long long ka, unsigned ia, unsigned long long kb, unsigned ib) {
    return ka > kb || (ka == kb && ia > ib);
}

constexpr int RANK_SMEM_MAX = 16384;

static __global__ void __launch_bounds__(1024) rank_smem_kernel(const double *__restrict__ f, int count, int npad,
        int *__restrict__ idx_out, double *__restrict__ val_out) {
    extern __shared__ __align__(16) unsigned char raw[];
    unsigned long long *keys = (unsigned long long*) raw;
    unsigned *ids = (unsigned*) (keys + npad);
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        keys[i] = i < count ? sortable_key(f[i]) : ~0ull;
        ids[i] = i < count ? (unsigned) i : 0xffffffffu;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = threadIdx.x; p < (npad >> 1); p += blockDim.x) {
                const int lo = 2 * j * (p / j) + (p % j);
                const int hi = lo + j;
                const bool asc = (lo & k) == 0;
                const unsigned long long ka = keys[lo], kb = keys[hi];
                const unsigned ia = ids[lo], ib = ids[hi];
                if (tuple_greater(ka, ia, kb, ib) == asc) {
                    keys[lo] = kb; keys[hi] = ka;
                    ids[lo] = ib; ids[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const unsigned id = ids[i];
        idx_out[i] = (int) id;
        val_out[i] = f[id];
    }
}

static __global__ void rank_global_init(const double *__restrict__ f, int count, int npad, unsigned long long *keys,
        unsigned *ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    keys[i] = i < count ? sortable_key(f[i]) : ~0ull;
    ids[i] = i < count ? (unsigned) i : 0xffffffffu;
}

static __global__ void rank_global_step(unsigned long long *keys, unsigned *ids, int npad, int k, int j) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (npad >> 1)) return;
    const int lo = 2 * j * (p / j) + (p % j);
    const int hi = lo + j;
    const bool asc = (lo & k) == 0;
    const unsigned long long ka = keys[lo], kb = keys[hi];
    const unsigned ia = ids[lo], ib = ids[hi];
    if (tuple_greater(ka, ia, kb, ib) == asc) {
        keys[lo] = kb; keys[hi] = ka;
        ids[lo] = ib; ids[hi] = ia;
    }
}

static __global__ void rank_global_finish(const double *__restrict__ f, const unsigned *__restrict__ ids, int count,
        int *__restrict__ idx_out, double *__restrict__ val_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned id = ids[i];
    idx_out[i] = (int) id;
    val_out[i] = f[id];
}

// d_f[count] -> d_idx[count] (permutation, ascending fitness), d_val[count] (sorted fitness)
inline int rank_fitness(bbgpu_ctx *ctx, const double *d_f, int count, int *d_idx, double *d_val) {
    if (count <= 0) return BBGPU_OK;
    int npad = 2;
    while (npad < count) npad <<= 1;
    if (npad <= RANK_SMEM_MAX) {
        const size_t bytes = (size_t) npad * 12;
        BB_CUDA(cudaFuncSetAttribute(rank_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (RANK_SMEM_MAX * 12)));
        const int threads = npad / 2 < 1024 ? (npad / 2 < 32 ? 32 : npad / 2) : 1024;
        rank_smem_kernel<<<1, threads, bytes, ctx->stream>>>(d_f, count, npad, d_idx, d_val);
        return launched(ctx, "rank_smem_kernel");
    }
    void *sp = nullptr;
    BB_TRY(ctx->scratch((size_t) npad * 12, &sp));
    unsigned long long *keys = (unsigned long long*) sp;
    unsigned *ids = (unsigned*) (keys + npad);
    rank_global_init<<<ceil_div(npad, 256), 256, 0, ctx->stream>>>(d_f, count, npad, keys, ids);
    BB_TRY(launched(ctx, "rank_global_init"));
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            rank_global_step<<<ceil_div(npad / 2, 256), 256, 0, ctx->stream>>>(keys, ids, npad, k, j);
            BB_TRY(launched(ctx, "rank_global_step"));
        }
    rank_global_finish<<<ceil_div(count, 256), 256, 0, ctx->stream>>>(d_f, ids, count, d_idx, d_val);
    return launched(ctx, "rank_global_finish");
}

} // namespace bb
