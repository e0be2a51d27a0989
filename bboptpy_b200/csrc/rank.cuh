// Fitness ranking: replaces std::sort(_fitness, cmaes_index::compare_fitness)
// (cma/base_cmaes.cpp:221) and std::sort(_swarm, point::compare_fitness) (de/jade.cpp:101).
//
// Keys are the fp64 fitness mapped to order-preserving u64, paired with the candidate index and
// compared as a (key, index) tuple: a TOTAL order, so the permutation is deterministic.  The
// reference's std::sort is unstable (tie order is implementation defined); parity cases are
// tie-free and tests assert that.  NaN sorts last.
//
// lambda < 8192: counting rank over the whole chip (rank_count_kernel).
// lambda <= 65536: bitonic sort of 1024-candidate chunks in shared memory (one CTA per chunk), then every candidate's
//   rank = its position in its chunk + its lower bounds in the other sorted chunks (binary searches, one thread per
//   (candidate, chunk) pair): O(lambda log^2 1024 + lambda (lambda/1024) log 1024) instead of lambda^2 -- 110 us of
//   counting at lambda = 16384 (20 % of a candidate-sharded generation on 8 GPUs) become a few microseconds.
// larger: a bitonic network over global memory, one launch per (k, j) stage.
#pragma once

#include "common.cuh"

namespace bb {

// NaN of either sign (x86's default NaN 0xFFF8... has the sign bit set) maps to the largest key, so it sorts last and
// can never become the best candidate; -0.0 and +0.0 share a key (f + 0.0 turns -0.0 into +0.0).
__device__ __forceinline__ unsigned long long sortable_key(double f) {
    if (f != f) return ~0ull;
    const long long b = __double_as_longlong(f + 0.0);
    return (unsigned long long) (b ^ ((b >> 63) | (long long) 0x8000000000000000ull));
}

__device__ __forceinline__ bool tuple_greater(unsigned long long ka, unsigned ia, unsigned long long kb, unsigned ib) {
    return ka > kb || (ka == kb && ia > ib);
}

constexpr int RANK_COUNT_MAX = 65536;

// Counting rank: rank_i = #{ j : (key_j, j) < (key_i, i) }.  O(lambda^2) compares, but spread over the
// whole chip (grid = lambda/256 x lambda/256 CTAs, keys of the j-slice staged in shared memory) it is a
// few microseconds at lambda = 4096 where a single-CTA bitonic network took 70.  Integer atomics only,
// so the result is deterministic; (key, index) is a total order, so the ranks are a permutation.
static __global__ void __launch_bounds__(256) rank_count_kernel(const double *__restrict__ f, int count,
        int *__restrict__ ranks) {
    __shared__ unsigned long long sk[256];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int j0 = blockIdx.y * 256;
    const int j = j0 + threadIdx.x;
    sk[threadIdx.x] = j < count ? sortable_key(f[j]) : ~0ull;
    __syncthreads();
    if (i >= count) return;
    const unsigned long long ki = sortable_key(f[i]);
    const int lim = min(256, count - j0);
    int cnt = 0;
#pragma unroll 8
    for (int q = 0; q < lim; q++) {
        const unsigned long long kj = sk[q];
        cnt += (kj < ki) || (kj == ki && (j0 + q) < i);
    }
    if (cnt) atomicAdd(ranks + i, cnt);
}

// ---- chunk sort + cross ranks ----------------------------------------------------------------------------------------
constexpr int RANK_CHUNK = 1024;
constexpr int RANK_CHUNK_MIN = 8191; // counts above this take the chunk path (measured: 22 us against 14 for counting at 4096)

// one CTA sorts candidates [c * 1024, (c + 1) * 1024) by (key, index) in shared memory: bitonic network, 512 threads, one
// compare-exchange per thread and stage; pads (index beyond count) carry the largest tuple and end up last
static __global__ void __launch_bounds__(512) rank_chunk_sort_kernel(const double *__restrict__ f, int count,
        unsigned long long *__restrict__ skeys, unsigned *__restrict__ sids) {
    __shared__ unsigned long long k[RANK_CHUNK];
    __shared__ unsigned id[RANK_CHUNK];
    const int base = blockIdx.x * RANK_CHUNK;
    for (int q = threadIdx.x; q < RANK_CHUNK; q += 512) {
        const int i = base + q;
        k[q] = i < count ? sortable_key(f[i]) : ~0ull;
        id[q] = i < count ? (unsigned) i : 0xffffffffu;
    }
    __syncthreads();
    for (int kk = 2; kk <= RANK_CHUNK; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            const int p = threadIdx.x;
            const int lo = 2 * j * (p / j) + (p % j), hi = lo + j;
            const bool asc = (lo & kk) == 0;
            const unsigned long long ka = k[lo], kb = k[hi];
            const unsigned ia = id[lo], ib = id[hi];
            if (tuple_greater(ka, ia, kb, ib) == asc) {
                k[lo] = kb; k[hi] = ka;
                id[lo] = ib; id[hi] = ia;
            }
            __syncthreads();
        }
    for (int q = threadIdx.x; q < RANK_CHUNK; q += 512) {
        skeys[base + q] = k[q];
        sids[base + q] = id[q];
    }
}

// thread (e, c): candidate at sorted position e (of chunk e / 1024) against sorted chunk c: the number of tuples of c
// below it (its own position when c is its chunk), added to the candidate's rank.  Integer atomics: deterministic.
static __global__ void __launch_bounds__(256) rank_cross_kernel(const unsigned long long *__restrict__ skeys,
        const unsigned *__restrict__ sids, int count, int nchunks, int *__restrict__ ranks) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int c = blockIdx.y;
    if (e >= nchunks * RANK_CHUNK) return;
    const unsigned me = sids[e];
    if (me >= (unsigned) count) return; // a pad
    int cnt;
    if (e / RANK_CHUNK == c) {
        cnt = e % RANK_CHUNK;
    } else {
        const unsigned long long ke = skeys[e];
        const unsigned long long *ck = skeys + (size_t) c * RANK_CHUNK;
        const unsigned *ci = sids + (size_t) c * RANK_CHUNK;
        int lo = 0, hi = RANK_CHUNK; // first position whose tuple is not below (ke, me)
#pragma unroll
        for (int step = 0; step < 11; step++) { // 1024 -> 512 -> ... -> 1 -> 0
            const int mid = min((lo + hi) >> 1, RANK_CHUNK - 1);
            const unsigned long long km = ck[mid];
            const bool below = lo < hi && (km < ke || (km == ke && ci[mid] < me));
            hi = (below || lo >= hi) ? hi : mid;
            lo = below ? mid + 1 : lo;
        }
        cnt = lo;
    }
    if (cnt) atomicAdd(ranks + me, cnt);
}

// optional per-call report of the ends of the ranking (fused into the scatter kernel)
struct RankEnds {
    int *ibw;      // [4] indices of best, 2nd best, 2nd worst, worst   (nullptr -> no report)
    double *ybw;   // [4] their fitness
    double *f_kth; // fitness at sorted position ik
    int ik;
};

// scatter by rank; re-zeroes the rank buffer for the next call (so no memset launch is needed)
static __global__ void __launch_bounds__(256) rank_scatter_kernel(const double *__restrict__ f, int count,
        int *__restrict__ ranks, int *__restrict__ idx_out, double *__restrict__ val_out, RankEnds ends) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    const int r = ranks[i];
    ranks[i] = 0;
    const double v = f[i];
    idx_out[r] = i;
    val_out[r] = v;
    if (ends.ibw != nullptr) {
        const int last = count - 1, second_last = count > 1 ? count - 2 : 0, second = count > 1 ? 1 : 0;
        if (r == 0) { ends.ibw[0] = i; ends.ybw[0] = v; }
        if (r == second) { ends.ibw[1] = i; ends.ybw[1] = v; }
        if (r == second_last) { ends.ibw[2] = i; ends.ybw[2] = v; }
        if (r == last) { ends.ibw[3] = i; ends.ybw[3] = v; }
        if (r == (ends.ik < count ? ends.ik : last)) *ends.f_kth = v;
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
inline int rank_fitness(bbgpu_ctx *ctx, const double *d_f, int count, int *d_idx, double *d_val,
        const RankEnds *ends = nullptr) {
    if (count <= 0) return BBGPU_OK;
    if (count <= RANK_COUNT_MAX) {
        if (ctx->d_ranks == nullptr) { // persistent, zero on entry to every call (the scatter kernel re-zeroes it)
            BB_CUDA(cudaMalloc(&ctx->d_ranks, (size_t) RANK_COUNT_MAX * sizeof(int)));
            BB_CUDA(cudaMemsetAsync(ctx->d_ranks, 0, (size_t) RANK_COUNT_MAX * sizeof(int), ctx->stream));
        }
        int *ranks = (int*) ctx->d_ranks;
        const int blocks = ceil_div(count, 256);
        static const bool force_count = [] { const char *e = getenv("BBGPU_RANK"); return e && e[0] == 'c'; }();
        if (count > RANK_CHUNK_MIN && !force_count) {
            const int nchunks = ceil_div(count, RANK_CHUNK);
            void *sp = nullptr;
            BB_TRY(ctx->scratch((size_t) nchunks * RANK_CHUNK * 12, &sp));
            unsigned long long *skeys = (unsigned long long*) sp;
            unsigned *sids = (unsigned*) (skeys + (size_t) nchunks * RANK_CHUNK);
            rank_chunk_sort_kernel<<<nchunks, 512, 0, ctx->stream>>>(d_f, count, skeys, sids);
            BB_TRY(launched(ctx, "rank_chunk_sort_kernel"));
            rank_cross_kernel<<<dim3(nchunks * RANK_CHUNK / 256, nchunks), 256, 0, ctx->stream>>>(skeys, sids, count, nchunks, ranks);
            BB_TRY(launched(ctx, "rank_cross_kernel"));
        } else {
            rank_count_kernel<<<dim3(blocks, blocks), 256, 0, ctx->stream>>>(d_f, count, ranks);
            BB_TRY(launched(ctx, "rank_count_kernel"));
        }
        RankEnds e = ends ? *ends : RankEnds { nullptr, nullptr, nullptr, 0 };
        rank_scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(d_f, count, ranks, d_idx, d_val, e);
        return launched(ctx, "rank_scatter_kernel");
    }
    int npad = 2;
    while (npad < count) npad <<= 1;
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
