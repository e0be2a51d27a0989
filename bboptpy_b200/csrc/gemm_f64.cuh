// fp64 tensor-core GEMM / SYRK for the CMA-ES hot loops (sm_100a).
//
// tcgen05.mma has no .kind::f64 (ptxas rejects it), so the fp64 tensor path on B200 is the
// warp-level `mma.sync.aligned.m8n8k4.f64` (SASS DMMA.8x8x4).  One CTA = 8 warps = a
// 128 x (16*NA) output tile; each warp owns 32 x (8*NA) = 4 x NA DMMA atoms (2*4*NA fp64
// accumulators per lane).  Operands stream through a 4-stage cp.async ring in shared memory
// with row pitches chosen so every fragment load is bank-conflict free (pitch = 4 mod 16
// doubles).  The kernel is DMMA-issue bound by construction: per k4-step a warp issues
// 4+NA 8-byte shared loads for 4*NA DMMAs.
//
// Two operand forms cover every dense contraction of the hot path:
//   FORM 1  C[M,N] = A[M,K] * B[K,N]   A row-major, K contiguous
//           - sampling      X = m + sigma * Z * (B diag(D))^T     (cma/cmaes.cpp:65-80)
//           - Active norms  ||C^-1/2 (x_k - m_old)||^2            (cma/active_cmaes.cpp:115-132)
//   FORM 2  C[M,N] = sum_k s[k] * At[k,M]^T * B[k,N]   both operands "k-major rows"
//           - rank-mu / active covariance update                  (cma/cmaes.cpp:125-142,
//                                                                  cma/active_cmaes.cpp:135-161)
//           - C^-1/2 = B D^-1 B^T                                 (cma/cmaes.cpp:274-282)
// Epilogues are fused: mean shift + sigma scale + box clip (sampling), row sum of squares
// (Mahalanobis norms, deterministic per-N-tile partials), decay + rank-one + lower-triangle
// store (covariance).
#pragma once

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace bb {

constexpr int G_THREADS = 256;

enum { EPI_STORE = 0, EPI_SAMPLE = 1, EPI_ROWSQ = 2, EPI_SYRK = 3 };

struct GemmArgs {
    const double *A;
    int64_t lda;
    const double *B;
    int64_t ldb;
    int M, N, K;
    const double *kscale; // FORM2: per-k coefficient applied to the A operand (nullptr -> 1)
    bool kscale_padded;   // kscale is readable (and zero) up to the next multiple of 32 past K: TMA path allowed
    // epilogue
    double *C;
    int64_t ldc;
    const double *vec;    // SAMPLE: mean[N]; SYRK: pc[n]
    const double *lo, *hi;// SAMPLE: bounds (nullptr -> unbounded)
    double s0, s1;        // SAMPLE: s0 = sigma; SYRK: s0 = decay, s1 = c1
    const double *dscal;  // SYRK: device scalar c2 = (1-hsig) cc (2-cc)
    double *partial;      // ROWSQ: partial[tileN * M + row]
    const int *Kdev;      // optional: the reduction length on the DEVICE (<= K); rows [*Kdev, K) are not visited.  Lets a
                          // data-dependent row count (compacted selections) reach the kernel without a host round trip
};

// CTA = 8 warps as 4 (M) x 2 (N); a warp owns MA x NA DMMA atoms -> CTA tile (32 MA) x (16 NA)
// BK = k chunk per pipeline stage, STAGES = depth of the cp.async ring
template<int MA, int NA, int BK, int STAGES> struct GemmSmem {
    static constexpr int BM = MA * 32;
    static constexpr int BN = NA * 16;
    static constexpr int LDA1 = BK + 4; // FORM1 A tile pitch [BM][BK+4]   (pitch = 4 mod 16 doubles:
    static constexpr int LDA2 = BM + 4; // FORM2 A tile pitch [BK][BM+4]    conflict-free fragment loads)
    static constexpr int LDB = BN + 4;
    static constexpr int ASZ = (BM * LDA1 > BK * LDA2) ? BM * LDA1 : BK * LDA2;
    static constexpr int BSZ = BK * LDB;
    static constexpr int STAGE = ASZ + BSZ + BK; // doubles per stage (A, B, kscale)
    static constexpr size_t BYTES = (size_t) STAGE * STAGES * sizeof(double);
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template<int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" :: "n"(N));
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
            : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// rows x cols (doubles) tile copy global -> shared in 16-byte chunks with zero fill outside
// [0,rlim) x [0,climit).  gcol0/cols are even; pitches are even; base pointers 16-byte aligned.
template<int ROWS, int COLS, int PITCH>
__device__ __forceinline__ void load_tile(double *smem, const double *g, int64_t ld, int grow0, int gcol0,
        int rlim, int clim) {
    constexpr int CPR = COLS / 2; // chunks per row
    constexpr int TOTAL = ROWS * CPR;
#pragma unroll
    for (int c = threadIdx.x; c < TOTAL; c += G_THREADS) {
        const int r = c / CPR, cc = (c % CPR) * 2;
        const int gr = grow0 + r, gc = gcol0 + cc;
        int valid = (gr < rlim) ? (clim - gc) : 0;
        valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid);
        const double *src = valid > 0 ? g + (int64_t) gr * ld + gc : g;
        cp_async16(smem + r * PITCH + cc, src, valid * 8);
    }
}

template<int FORM, int MA, int NA, int BK, int STAGES, int EPI>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_f64_kernel(const GemmArgs p) {
    using S = GemmSmem<MA, NA, BK, STAGES>;
    constexpr int BN = S::BN;
    constexpr int G_BM = S::BM;
    constexpr int G_BK = BK;
    constexpr int G_STAGES = STAGES;
    constexpr int G_LDA1 = S::LDA1;
    constexpr int G_LDA2 = S::LDA2;
    constexpr int G_ASZ = S::ASZ;
    extern __shared__ __align__(16) double smem[];

    const int tiles_n = (p.N + BN - 1) / BN;
    const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
    const int m0 = tm * G_BM, n0 = tn * BN;
    if (EPI == EPI_SYRK && n0 > m0 + G_BM - 1) return; // tile strictly above the diagonal

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 1, wn = warp & 1; // 4 x 2 warps
    const int g = lane >> 2, t = lane & 3;

    double acc[MA][NA][2];
#pragma unroll
    for (int i = 0; i < MA; i++)
#pragma unroll
        for (int j = 0; j < NA; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = ((p.Kdev ? min(*p.Kdev, p.K) : p.K) + G_BK - 1) / G_BK;

    auto issue = [&](int kc) {
        if (kc < nk) {
            double *st = smem + (size_t) (kc % G_STAGES) * S::STAGE;
            const int k0 = kc * G_BK;
            if (FORM == 1) load_tile<G_BM, G_BK, G_LDA1>(st, p.A, p.lda, m0, k0, p.M, p.K);
            else load_tile<G_BK, G_BM, G_LDA2>(st, p.A, p.lda, k0, m0, p.K, p.M);
            load_tile<G_BK, BN, S::LDB>(st + G_ASZ, p.B, p.ldb, k0, n0, p.K, p.N);
            if (FORM == 2 && threadIdx.x < G_BK) {
                const int k = k0 + threadIdx.x;
                const bool ok = (p.kscale != nullptr) && k < p.K;
                cp_async8(st + G_ASZ + S::BSZ + threadIdx.x, ok ? p.kscale + k : p.B, ok ? 8 : 0);
            }
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < G_STAGES - 1; s++) issue(s);

    for (int kc = 0; kc < nk; kc++) {
        cp_async_wait<G_STAGES - 2>();
        __syncthreads();
        issue(kc + G_STAGES - 1);
        const double *As = smem + (size_t) (kc % G_STAGES) * S::STAGE;
        const double *Bs = As + G_ASZ;
        const double *Ks = Bs + S::BSZ;
#pragma unroll
        for (int ks = 0; ks < G_BK / 4; ks++) {
            double a[MA], b[NA];
            const int kk = ks * 4 + t;
            if (FORM == 1) {
#pragma unroll
                for (int i = 0; i < MA; i++) a[i] = As[(wm * (MA * 8) + i * 8 + g) * G_LDA1 + kk];
            } else {
                const double sc = (p.kscale != nullptr) ? Ks[kk] : 1.0;
#pragma unroll
                for (int i = 0; i < MA; i++) a[i] = As[kk * G_LDA2 + wm * (MA * 8) + i * 8 + g] * sc;
            }
#pragma unroll
            for (int j = 0; j < NA; j++) b[j] = Bs[kk * S::LDB + wn * (NA * 8) + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MA; i++)
#pragma unroll
                for (int j = 0; j < NA; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue.  Lane (g,t) holds C[row g][cols 2t, 2t+1] of every atom.
    if (EPI == EPI_ROWSQ) {
        __syncthreads(); // all warps are done reading the operand ring; reuse it
        double *rs = smem; // [2][BM]
#pragma unroll
        for (int i = 0; i < MA; i++) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NA; j++) {
                s = __dadd_rn(s, __dmul_rn(acc[i][j][0], acc[i][j][0]));
                s = __dadd_rn(s, __dmul_rn(acc[i][j][1], acc[i][j][1]));
            }
            s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
            s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
            if (t == 0) rs[wn * G_BM + wm * (MA * 8) + i * 8 + g] = s;
        }
        __syncthreads();
        if (threadIdx.x < G_BM) {
            const int row = m0 + threadIdx.x;
            if (row < p.M) p.partial[(int64_t) tn * p.M + row] = __dadd_rn(rs[threadIdx.x], rs[G_BM + threadIdx.x]);
        }
        return;
    }

    double c2 = 0.0;
    if (EPI == EPI_SYRK) c2 = *p.dscal;
#pragma unroll
    for (int i = 0; i < MA; i++) {
        const int row = m0 + wm * (MA * 8) + i * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NA; j++) {
            const int col = n0 + wn * (NA * 8) + j * 8 + t * 2;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int cc = col + e;
                if (cc >= p.N) continue;
                double v = acc[i][j][e];
                double *dst = p.C + (int64_t) row * p.ldc + cc;
                if (EPI == EPI_SAMPLE) {
                    v = p.vec[cc] + p.s0 * v;
                    if (p.lo != nullptr) v = fmax(p.lo[cc], fmin(v, p.hi[cc]));
                    *dst = v;
                } else if (EPI == EPI_SYRK) {
                    if (cc > row) continue; // lower triangle only (cma/cmaes.cpp:127)
                    const double old = *dst;
                    const double base = p.s0 * old + p.s1 * (p.vec[row] * p.vec[cc] + c2 * old);
                    *dst = base + v;
                } else {
                    *dst = v;
                }
            }
        }
    }
}

template<int FORM, int MA, int NA, int BK, int STAGES, int EPI>
inline int launch_gemm_tile(bbgpu_ctx *ctx, const GemmArgs &a) {
    using S = GemmSmem<MA, NA, BK, STAGES>;
    BB_CUDA(cudaFuncSetAttribute(gemm_f64_kernel<FORM, MA, NA, BK, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) S::BYTES));
    const int tiles = ceil_div(a.M, S::BM) * ceil_div(a.N, S::BN);
    gemm_f64_kernel<FORM, MA, NA, BK, STAGES, EPI><<<tiles, G_THREADS, S::BYTES, ctx->stream>>>(a);
    return launched(ctx, "gemm_f64_kernel");
}

// number of tiles the kernel actually works on (SYRK skips tiles strictly above the diagonal)
inline int64_t live_tiles(int M, int N, int BM, int BN, bool lower_only) {
    const int64_t tm = ceil_div(M, BM), tn = ceil_div(N, BN);
    if (!lower_only) return tm * tn;
    int64_t cnt = 0;
    for (int64_t i = 0; i < tm; i++) cnt += std::min<int64_t>(tn, (i * BM + BM - 1) / BN + 1);
    return cnt;
}

// pick the CTA tile (128x112, 128x128 or 64x64) that needs the least tile-waves x tile-area on this
// device.  The small tile matters for the covariance SYRK: n = 1000 has only 44 lower 128x112 tiles
// for 148 SMs, but 136 lower 64x64 tiles (one full wave).
inline int pick_tile(bbgpu_ctx *ctx, const GemmArgs &a, bool lower);

template<int FORM, int EPI>
inline int launch_gemm_cpasync(bbgpu_ctx *ctx, const GemmArgs &a, int *tiles_n_out = nullptr) {
    if (a.M <= 0 || a.N <= 0) return BBGPU_OK;
    if ((a.lda & 1) || (a.ldb & 1) || (((uintptr_t) a.A) & 15) || (((uintptr_t) a.B) & 15))
        return fail(BBGPU_ERR_ARG, "gemm operands must be 16-byte aligned with even leading dimensions");
    const int pick = pick_tile(ctx, a, EPI == EPI_SYRK);
    switch (pick) {
    case 2:
        if (tiles_n_out) *tiles_n_out = ceil_div(a.N, 64);
        return launch_gemm_tile<FORM, 2, 4, 32, 4, EPI>(ctx, a);
    case 3:
        if (tiles_n_out) *tiles_n_out = ceil_div(a.N, 112);
        return launch_gemm_tile<FORM, 4, 7, 32, 3, EPI>(ctx, a);
    case 4:
        if (tiles_n_out) *tiles_n_out = ceil_div(a.N, 128);
        return launch_gemm_tile<FORM, 4, 8, 32, 3, EPI>(ctx, a);
    case 1:
        if (tiles_n_out) *tiles_n_out = ceil_div(a.N, 128);
        return launch_gemm_tile<FORM, 4, 8, 16, 4, EPI>(ctx, a);
    default:
        if (tiles_n_out) *tiles_n_out = ceil_div(a.N, 112);
        return launch_gemm_tile<FORM, 4, 7, 16, 4, EPI>(ctx, a);
    }
}

// CTA tile choice: 0/3 = 128x112 (BK 16/32), 1/4 = 128x128 (BK 16/32), 2 = 64x64 (BK 32).  Least
// tile-waves x tile-area on this device; measured on B200 at 4096x1000x1000 (cp.async ring): 128x112
// BK=32/3 stages 28.2 TF, BK=16/4 stages 27.2 TF, 128x128 24.7-25.2 TF.  The small tile serves the
// covariance SYRK: n = 1000 has only 44 lower 128x112 tiles for 148 SMs but 136 lower 64x64 tiles.
inline int pick_tile(bbgpu_ctx *ctx, const GemmArgs &a, bool lower) {
    const int sms = ctx->sm_count;
    auto cost = [&](int BM, int BN) {
        const int64_t t = live_tiles(a.M, a.N, BM, BN, lower);
        return (double) ceil_div(t, (int64_t) sms) * BM * BN;
    };
    const double c7 = cost(128, 112), c8 = cost(128, 128), c4 = cost(64, 64) * 1.15;
    int pick = (c4 < c7 && c4 < c8) ? 2 : (c7 < c8 ? 3 : 4);
    // experiment hook: BBGPU_GEMM_VARIANT=0..4 forces a tile configuration (tools/gemm_variant.py)
    static const int forced = [] { const char *e = std::getenv("BBGPU_GEMM_VARIANT"); return e ? std::atoi(e) : -1; }();
    if (forced >= 0) pick = forced;
    return pick;
}

} // namespace bb
