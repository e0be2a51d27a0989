// LM-CMA-ES on the device (reference: cma/lm_cmaes.cpp, Loshchilov 2014 / 2015).
//
// The covariance is implicit in `memory` direction vectors: rows of PC (the stored evolution paths, `_pcmat`) and V
// (`_vmat`) with the scalars b, d; `jarr[0 .. memlen-1]` names the live rows in application order.  The bookkeeping of
// jarr / larr (algorithm 5) is integer control logic and stays with the host class; the O(memory n) vector work per
// candidate and the O(memory^2 n) rebuild of V are here.
#pragma once

#include "common.cuh"
#include "philox.cuh"

namespace bb {

constexpr int LM_MAXMEM = 1024; // direction vectors a handle can hold (dots staged in shared memory)
constexpr int LM_GROUP = 8;     // inner products reduced together

// samplePopulation (lm_cmaes.cpp:91-127), one CTA per PAIR of candidates (mirrored sampling):
//   z_p        the base vector: replayed, or Rademacher / Gaussian from Philox block (coordinate pair, p, generation)
//   i0         selectSubset (algorithm 6, :308-322): replayed, or memlen - min(memlen, int(msigma |N(0,1)|))
//   dot_i      b_j (v_j . z_p), j = jarr[i], for i = i0 .. memlen-1 -- with the ORIGINAL z_p (:117-119), so all of them
//              are independent: LM_GROUP at a time through one block reduction
//   az         z_p, then az <- sqrt(1-c1) az + dot_i pc_j in the order of i (algorithm 3, :120-121), element-local
//   x_{2p}, x_{2p+1} = m +- sigma az (:124)
__global__ void __launch_bounds__(256) lm_sample_kernel(double *__restrict__ X, double *__restrict__ Zp, int64_t ld, int n, int lambda,
        const double *__restrict__ xmean, double sigma, const double *__restrict__ PC, const double *__restrict__ V,
        const double *__restrict__ b, const int *__restrict__ jarr, int memlen, const int *__restrict__ i0_in, bool replay,
        int rademacher, int usenew, double sqrt1mc1, uint64_t seed, uint64_t generation) {
    __shared__ double dots[LM_MAXMEM];
    __shared__ double red[LM_GROUP][8];
    __shared__ int s_i0;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *z = Zp + (int64_t) p * ld;
    if (!replay) {
        for (int jp = tid; 2 * jp < n; jp += blockDim.x) {
            const Philox4 r = philox4x32_10((uint32_t) jp, (uint32_t) p, (uint32_t) generation, (uint32_t) (generation >> 32) ^ 0x4c4d0000u,
                    (uint32_t) seed, (uint32_t) (seed >> 32));
            double z0, z1;
            if (rademacher) {
                z0 = (r.x & 1u) ? 1.0 : -1.0;
                z1 = (r.z & 1u) ? 1.0 : -1.0;
            } else {
                normal_pair(r, z0, z1);
            }
            z[2 * jp] = z0;
            if (2 * jp + 1 < n) z[2 * jp + 1] = z1;
        }
    }
    if (tid == 0) {
        int i0 = 0;
        if (replay) i0 = i0_in[p];
        else if (usenew && memlen > 1) {
            const Philox4 r = philox4x32_10(0xffffffffu, (uint32_t) p, (uint32_t) generation, (uint32_t) (generation >> 32) ^ 0x4c4d0000u,
                    (uint32_t) seed, (uint32_t) (seed >> 32));
            double g0, g1;
            normal_pair(r, g0, g1);
            const int msigma = p == 0 ? 40 : 4; // k == 0: the first candidate of the generation (:313-316)
            const double m = (double) msigma * fabs(g0);
            const int mstar = m >= (double) memlen ? memlen : (int) m;
            i0 = memlen - mstar;
        }
        s_i0 = i0;
    }
    __syncthreads();
    const int i0 = s_i0;
    for (int g0 = i0; g0 < memlen; g0 += LM_GROUP) {
        double acc[LM_GROUP];
        const double *vr[LM_GROUP];
#pragma unroll
        for (int u = 0; u < LM_GROUP; u++) {
            acc[u] = 0.0;
            vr[u] = V + (int64_t) jarr[min(g0 + u, memlen - 1)] * ld;
        }
        for (int e = tid; e < n; e += blockDim.x) {
            const double ze = z[e];
#pragma unroll
            for (int u = 0; u < LM_GROUP; u++) acc[u] = fma(vr[u][e], ze, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < LM_GROUP; u++) {
            const double s = warp_sum(acc[u]);
            if (lane == 0) red[u][warp] = s;
        }
        __syncthreads();
        if (tid < LM_GROUP && g0 + tid < memlen) {
            double s = 0.0;
            for (int w = 0; w < 8; w++) s += red[tid][w];
            dots[g0 + tid] = __dmul_rn(b[jarr[g0 + tid]], s);
        }
        __syncthreads();
    }
    const bool second = 2 * p + 1 < lambda;
    double *x0 = X + (int64_t) (2 * p) * ld, *x1 = X + (int64_t) (2 * p + 1) * ld;
    for (int e = tid; e < n; e += blockDim.x) {
        double a = z[e];
        for (int i = i0; i < memlen; i++)
            a = __dadd_rn(__dmul_rn(sqrt1mc1, a), __dmul_rn(dots[i], PC[(int64_t) jarr[i] * ld + e]));
        const double me = xmean[e];
        x0[e] = __dadd_rn(me, __dmul_rn(sigma, a));
        if (second) x1[e] = __dadd_rn(me, __dmul_rn(-sigma, a));
    }
}

// ---- the same sampling, split over column slabs -------------------------------------------------------------------------
// lm_sample_kernel gives a pair to ONE CTA: lambda = 64 puts 32 CTAs on 148 SMs, each streaming all of V and PC (2 x 16 MB at
// n = 10^4, memory 200) -- 250 us of a 330 us generation.  Split: grid (pair, slab).  lm_dots_kernel draws the slab's part of
// z_p and writes the partial inner products of its columns, lm_apply_kernel adds the partials in slab order (deterministic),
// scales them by b_j and runs the element-local recurrence on its slab.  Same arithmetic per element; the inner products are
// summed in a different (still fixed) order.
constexpr int LM_MAXSLABS = 16;

// i0 of pair p (selectSubset, algorithm 6, :308-322) from the pair's own Philox block
__device__ __forceinline__ int lm_draw_i0(int p, int memlen, int usenew, uint64_t seed, uint64_t generation) {
    if (!(usenew && memlen > 1)) return 0;
    const Philox4 r = philox4x32_10(0xffffffffu, (uint32_t) p, (uint32_t) generation, (uint32_t) (generation >> 32) ^ 0x4c4d0000u,
            (uint32_t) seed, (uint32_t) (seed >> 32));
    double g0, g1;
    normal_pair(r, g0, g1);
    const int msigma = p == 0 ? 40 : 4; // k == 0: the first candidate of the generation (:313-316)
    const double m = (double) msigma * fabs(g0);
    const int mstar = m >= (double) memlen ? memlen : (int) m;
    return memlen - mstar;
}

__global__ void __launch_bounds__(256) lm_dots_kernel(double *__restrict__ Zp, int64_t ld, int n, int slab_len, const double *__restrict__ V,
        const int *__restrict__ jarr, int memlen, int *__restrict__ i0_io, bool replay, int rademacher, int usenew, uint64_t seed,
        uint64_t generation, double *__restrict__ part, int part_ld) {
    const int p = blockIdx.x, sl = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e0 = sl * slab_len, e1 = min(n, e0 + slab_len); // slab_len is even: Philox pairs never straddle slabs
    double *z = Zp + (int64_t) p * ld;
    if (!replay) {
        for (int jp = e0 / 2 + tid; 2 * jp < e1; jp += blockDim.x) {
            const Philox4 r = philox4x32_10((uint32_t) jp, (uint32_t) p, (uint32_t) generation, (uint32_t) (generation >> 32) ^ 0x4c4d0000u,
                    (uint32_t) seed, (uint32_t) (seed >> 32));
            double z0, z1;
            if (rademacher) {
                z0 = (r.x & 1u) ? 1.0 : -1.0;
                z1 = (r.z & 1u) ? 1.0 : -1.0;
            } else {
                normal_pair(r, z0, z1);
            }
            z[2 * jp] = z0;
            if (2 * jp + 1 < n) z[2 * jp + 1] = z1;
        }
    }
    const int i0 = replay ? i0_io[p] : lm_draw_i0(p, memlen, usenew, seed, generation);
    if (!replay && sl == 0 && tid == 0) i0_io[p] = i0;
    __syncthreads();
    // A warp per direction vector, four vectors in flight: no block-wide reduction, no barrier.  (The first version reduced
    // eight inner products at a time over the whole CTA: 25 rounds of load latency + two barriers each at memory 200.)
    double *out = part + ((int64_t) p * gridDim.y + sl) * part_ld;
    constexpr int NW = 8, VPW = 4;
    for (int g0 = i0 + warp * VPW; g0 < memlen; g0 += NW * VPW) {
        const double *vr[VPW];
        double acc[VPW];
#pragma unroll
        for (int u = 0; u < VPW; u++) {
            acc[u] = 0.0;
            vr[u] = V + (int64_t) jarr[min(g0 + u, memlen - 1)] * ld;
        }
        for (int e = e0 + lane; e < e1; e += 32) {
            const double ze = z[e];
#pragma unroll
            for (int u = 0; u < VPW; u++) acc[u] = fma(vr[u][e], ze, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < VPW; u++) {
            const double sum = warp_sum(acc[u]);
            if (lane == 0 && g0 + u < memlen) out[g0 + u] = sum;
        }
    }
}

__global__ void __launch_bounds__(256) lm_apply_kernel(double *__restrict__ X, const double *__restrict__ Zp, int64_t ld, int n, int lambda,
        int slab_len, const double *__restrict__ xmean, double sigma, const double *__restrict__ PC, const double *__restrict__ b,
        const int *__restrict__ jarr, int memlen, const int *__restrict__ i0_arr, double sqrt1mc1, const double *__restrict__ part,
        int part_ld) {
    __shared__ double dots[LM_MAXMEM];
    __shared__ int srow[LM_MAXMEM];
    const int p = blockIdx.x, sl = blockIdx.y, slabs = gridDim.y, tid = threadIdx.x;
    const int e0 = sl * slab_len, e1 = min(n, e0 + slab_len);
    const int i0 = i0_arr[p];
    for (int i = i0 + tid; i < memlen; i += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < slabs; t++) s = __dadd_rn(s, part[((int64_t) p * slabs + t) * part_ld + i]);
        const int j = jarr[i];
        dots[i] = __dmul_rn(b[j], s);
        srow[i] = j;
    }
    __syncthreads();
    const double *z = Zp + (int64_t) p * ld;
    const bool second = 2 * p + 1 < lambda;
    double *x0 = X + (int64_t) (2 * p) * ld, *x1 = X + (int64_t) (2 * p + 1) * ld;
    // the recurrence is a dependent chain per element: a thread runs EPT elements side by side so that their loads overlap
    constexpr int EPT = 4;
    for (int base = e0; base < e1; base += EPT * 256) {
        double a[EPT];
        int e[EPT];
#pragma unroll
        for (int k = 0; k < EPT; k++) {
            e[k] = base + tid + k * 256;
            a[k] = e[k] < e1 ? z[e[k]] : 0.0;
            e[k] = min(e[k], e1 - 1); // lanes past the end redo the last element (same loads), they do not store
        }
#pragma unroll 2
        for (int i = i0; i < memlen; i++) {
            const double di = dots[i];
            const double *row = PC + (int64_t) srow[i] * ld;
#pragma unroll
            for (int k = 0; k < EPT; k++) a[k] = __dadd_rn(__dmul_rn(sqrt1mc1, a[k]), __dmul_rn(di, row[e[k]]));
        }
#pragma unroll
        for (int k = 0; k < EPT; k++) {
            const int ek = base + tid + k * 256;
            if (ek < e1) {
                const double me = xmean[ek];
                x0[ek] = __dadd_rn(me, __dmul_rn(sigma, a[k]));
                if (second) x1[ek] = __dadd_rn(me, __dmul_rn(-sigma, a[k]));
            }
        }
    }
}

// p_c (lm_cmaes.cpp:146-149: (1 - cc) pc + (ccc (m - m_old)) / sigma, no hsig) and, when the memory is due (:151-177):
// PC[jarr[memlen-1]] <- p_c; for i = imin .. memlen-1: t <- PC[jarr[i]], pushed through the inverse transforms of rows
// jarr[0 .. i-1] (algorithm 4, `ainvz`: every inner product uses the UPDATED t, a genuine chain), V[jarr[i]] <- t, then
// b, d from |t|^2.  One CTA: the chain is up to memory^2 / 2 dependent reductions of length n.
__global__ void __launch_bounds__(1024) lm_memory_kernel(int n, int64_t ld, double *__restrict__ pc, const double *__restrict__ xmean,
        const double *__restrict__ xold, double sigma, double cc, double ccc, double c1, double sqrt1mc1, int do_memory, int imin,
        int memlen, const int *__restrict__ jarr, double *__restrict__ PC, double *__restrict__ V, double *__restrict__ b,
        double *__restrict__ d, double *__restrict__ t) {
    __shared__ double red[32];
    const int tid = threadIdx.x;
    for (int e = tid; e < n; e += blockDim.x)
        pc[e] = __dadd_rn(__dmul_rn(1.0 - cc, pc[e]), __ddiv_rn(__dmul_rn(ccc, __dsub_rn(xmean[e], xold[e])), sigma));
    if (!do_memory) return;
    __syncthreads();
    {
        double *dst = PC + (int64_t) jarr[memlen - 1] * ld;
        for (int e = tid; e < n; e += blockDim.x) dst[e] = pc[e];
    }
    if (do_memory == 2) return; // the rebuild runs in lm_rebuild_kernel
    __syncthreads();
    const double c = 1.0 / sqrt1mc1;
    for (int i = imin; i < memlen; i++) {
        const int jcur = jarr[i];
        const double *src = PC + (int64_t) jcur * ld;
        for (int e = tid; e < n; e += blockDim.x) t[e] = src[e];
        for (int k = 0; k < i; k++) {
            const int idx = jarr[k];
            const double *v = V + (int64_t) idx * ld;
            double s = 0.0;
            for (int e = tid; e < n; e += blockDim.x) s = fma(v[e], t[e], s);
            s = block_sum(s, red);
            const double dot = __dmul_rn(d[idx], s);
            for (int e = tid; e < n; e += blockDim.x) t[e] = __dadd_rn(__dmul_rn(c, t[e]), __dmul_rn(-dot, v[e]));
        }
        double *vdst = V + (int64_t) jcur * ld;
        double s = 0.0;
        for (int e = tid; e < n; e += blockDim.x) {
            const double te = t[e];
            vdst[e] = te;
            s = fma(te, te, s);
        }
        s = block_sum(s, red);
        if (tid == 0) {
            const double sqrtc1 = sqrt(1.0 + (c1 / (1.0 - c1)) * s);
            b[jcur] = (sqrt1mc1 / s) * (sqrtc1 - 1.0);
            d[jcur] = (1.0 / (sqrt1mc1 * s)) * (1.0 - 1.0 / sqrtc1);
        }
        __syncthreads(); // b / d / V of this row are read by the next rows' chains
    }
}

// ---- the rebuild as a wavefront over the SMs ---------------------------------------------------------------------------------
// lm_memory_kernel walks the rebuilt rows one after the other: row i is a chain of i dependent n-reductions (≈215 cycles each,
// the floor of a block-wide sum), 20 000 of them for a full rebuild at memory 200: 2.2 ms, every few generations.  But row i
// needs the NEW row k only at its step k, and the rows below imin never change: with one row per CTA, every row runs the steps
// k < imin at once and then follows the row in front of it one step behind -- the critical path is one step + one finalisation
// per rebuilt row instead of i steps.  Rows imin + me, imin + me + G, ... belong to CTA me; a finished row is published by a
// flag (release after the CTA's stores; the consumer's thread 0 acquires it, the barrier carries it to the others; the row
// itself is read past L1).  Dependencies only point to smaller rows and every CTA takes its rows in increasing order, so the
// cooperative launch (all CTAs resident) cannot deadlock.  Arithmetic per row as in lm_memory_kernel: same element-to-thread
// assignment, same block sum.
constexpr int LM_TP = 12; // elements of t a thread keeps in registers: n <= 12288, beyond that t lives in global scratch

__global__ void __launch_bounds__(1024, 1) lm_rebuild_kernel(int n, int64_t ld, double c1, double sqrt1mc1, int imin, int memlen,
        const int *__restrict__ jarr, const double *__restrict__ PC, double *V, double *b, double *d, int *flags, double *tglob) {
    __shared__ double red[32];
    const int tid = threadIdx.x, G = gridDim.x, me = blockIdx.x;
    const double c = 1.0 / sqrt1mc1;
    const bool inreg = n <= LM_TP * 1024;
    double *tg = tglob ? tglob + (int64_t) me * ld : nullptr;
    for (int i = imin + me; i < memlen; i += G) {
        const int jcur = jarr[i];
        const double *src = PC + (int64_t) jcur * ld;
        double t[LM_TP];
        if (inreg) {
#pragma unroll
            for (int q = 0; q < LM_TP; q++) {
                const int e = tid + q * 1024;
                t[q] = e < n ? src[e] : 0.0;
            }
        } else {
            for (int e = tid; e < n; e += 1024) tg[e] = src[e];
        }
        for (int k = 0; k < i; k++) {
            const int idx = jarr[k];
            if (k >= imin) { // rebuilt by this launch: wait for its owner
                if (tid == 0) {
                    int seen = 0, spins = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flags + k) : "memory");
                    } while (seen == 0 && ++spins < (1 << 24));
                    if (seen == 0) flags[memlen] = 1; // gave up (a lost CTA must not hang the device); the word is there for a debugger
                }
                __syncthreads();
            }
            const double *v = V + (int64_t) idx * ld;
            double s = 0.0;
            if (inreg) {
                double vk[LM_TP];
#pragma unroll
                for (int q = 0; q < LM_TP; q++) {
                    const int e = tid + q * 1024;
                    vk[q] = e < n ? __ldcg(v + e) : 0.0;
                }
#pragma unroll
                for (int q = 0; q < LM_TP; q++)
                    if (tid + q * 1024 < n) s = fma(vk[q], t[q], s);
                s = block_sum(s, red);
                const double dot = __dmul_rn(__ldcg(d + idx), s);
#pragma unroll
                for (int q = 0; q < LM_TP; q++) t[q] = __dadd_rn(__dmul_rn(c, t[q]), __dmul_rn(-dot, vk[q]));
            } else {
                for (int e = tid; e < n; e += 1024) s = fma(__ldcg(v + e), tg[e], s);
                s = block_sum(s, red);
                const double dot = __dmul_rn(__ldcg(d + idx), s);
                for (int e = tid; e < n; e += 1024) tg[e] = __dadd_rn(__dmul_rn(c, tg[e]), __dmul_rn(-dot, __ldcg(v + e)));
            }
        }
        double *vdst = V + (int64_t) jcur * ld;
        double s = 0.0;
        if (inreg) {
#pragma unroll
            for (int q = 0; q < LM_TP; q++) {
                const int e = tid + q * 1024;
                if (e < n) {
                    vdst[e] = t[q];
                    s = fma(t[q], t[q], s);
                }
            }
        } else {
            for (int e = tid; e < n; e += 1024) {
                const double te = tg[e];
                vdst[e] = te;
                s = fma(te, te, s);
            }
        }
        s = block_sum(s, red);
        if (tid == 0) {
            const double sqrtc1 = sqrt(1.0 + (c1 / (1.0 - c1)) * s);
            b[jcur] = (sqrt1mc1 / s) * (sqrtc1 - 1.0);
            d[jcur] = (1.0 / (sqrt1mc1 * s)) * (1.0 - 1.0 / sqrtc1);
        }
        __threadfence(); // this thread's part of the row (and b / d) before the flag
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(flags + i), "r"(1) : "memory");
    }
}

} // namespace bb
