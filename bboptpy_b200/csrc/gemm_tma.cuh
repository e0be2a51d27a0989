// fp64 DMMA GEMM / SYRK fed by TMA (sm_100a): the warp-specialised successor of gemm_f64.cuh.
//
// Same math, same CTA tile (8 DMMA warps as 4 x 2, MA x NA atoms each) and the same fused epilogues
// as gemm_f64.cuh, but the operand ring is filled by ONE producer thread issuing
// `cp.async.bulk.tensor.2d` (SASS UTMALDG) against two tensor maps, and stages are handed over with
// mbarriers (full: TMA transaction bytes; empty: one arrive per consumer warp).  There is no
// CTA-wide barrier in the main loop and the DMMA warps issue no copy instructions at all.
//
// Shared-memory layout without padding AND without bank conflicts: a fragment load of m8n8k4 reads,
// per warp, an 8 x 4 (A, k contiguous) or 4 x 8 (B, n contiguous) patch.  TMA cannot pad rows, so each
// operand is fetched as narrow boxes whose rows are exactly one fragment wide:
//   k-contiguous operand (FORM 1 A):  BK/4 boxes of [rows][4 doubles]  -> smem [k/4][row][4]
//   n-contiguous operand (B, FORM 2 A): width/8 boxes of [BK][8 doubles] -> smem [col/8][k][8]
// so the 32 lanes of a fragment load touch 32 consecutive doubles (2 wavefronts, the minimum).
// Out-of-range rows / columns are zero-filled by the TMA unit (tensor extents = logical sizes).
#pragma once

#include <cuda.h>

#include "gemm_f64.cuh"

namespace bb {

constexpr int T_CONSUMERS = 256;          // 8 DMMA warps
constexpr int T_THREADS = T_CONSUMERS + 32; // + 1 producer warp

template<int MA, int NA, int BK, int STAGES> struct TmaSmem {
    static constexpr int BM = MA * 32;
    static constexpr int BN = NA * 16;
    static constexpr int ASZ = BM * BK;
    static constexpr int BSZ = BK * BN;
    static constexpr int STAGE = ASZ + BSZ + BK; // A tile, B tile, per-k coefficients (FORM 2)
    static constexpr size_t BYTES = (size_t) STAGE * STAGES * sizeof(double) + 2 * STAGES * sizeof(unsigned long long) + 128;
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"((unsigned) __cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"((unsigned) __cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"((unsigned) __cvta_generic_to_shared(bar)) : "memory");
}
// bounded wait: a broken pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, int parity) {
    const unsigned addr = (unsigned) __cvta_generic_to_shared(bar);
#pragma unroll 1
    for (unsigned spin = 0; spin < (1u << 28); spin++) {
        unsigned done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
// 1-D bulk copy global -> shared, completing on the same transaction barrier (SASS UBLKCP)
__device__ __forceinline__ void bulk_load_1d(void *smem, const void *gmem, int bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
            :: "r"((unsigned) __cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes),
               "r"((unsigned) __cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem, const CUtensorMap *map, int c0, int c1, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
            :: "r"((unsigned) __cvta_generic_to_shared(smem)), "l"(map), "r"(c0), "r"(c1),
               "r"((unsigned) __cvta_generic_to_shared(bar)) : "memory");
}

// mapA: FORM 1 -> box [BM rows][4];  FORM 2 -> box [BK rows][8].   mapB: box [BK rows][8].
template<int FORM, int MA, int NA, int BK, int STAGES, int EPI>
__global__ void __launch_bounds__(T_THREADS, 1) gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA,
        const __grid_constant__ CUtensorMap mapB, const GemmArgs p) {
    using S = TmaSmem<MA, NA, BK, STAGES>;
    constexpr int BM = S::BM, BN = S::BN;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double*>(smem_raw);
    unsigned long long *full = reinterpret_cast<unsigned long long*>(smem + (size_t) S::STAGE * STAGES);
    unsigned long long *empty = full + STAGES;

    const int tiles_n = (p.N + BN - 1) / BN;
    const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
    const int m0 = tm * BM, n0 = tn * BN;
    if (EPI == EPI_SYRK && n0 > m0 + BM - 1) return;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, T_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int nk = ((p.Kdev ? min(*p.Kdev, p.K) : p.K) + BK - 1) / BK;

    if (warp == T_CONSUMERS / 32) {
        // ===== TMA producer (one elected lane) =====
        if (lane == 0) {
            const bool scaled = FORM == 2 && p.kscale != nullptr;
            const int TX = (S::ASZ + S::BSZ + (scaled ? BK : 0)) * (int) sizeof(double);
            for (int kc = 0; kc < nk; kc++) {
                const int s = kc % STAGES, ph = (kc / STAGES) & 1;
                mbar_wait(empty + s, ph ^ 1);
                double *As = smem + (size_t) s * S::STAGE;
                double *Bs = As + S::ASZ;
                mbar_expect_tx(full + s, TX);
                const int k0 = kc * BK;
                if (FORM == 1) {
#pragma unroll
                    for (int q = 0; q < BK / 4; q++) tma_load_2d(As + q * (BM * 4), &mapA, k0 + q * 4, m0, full + s);
                } else {
#pragma unroll
                    for (int q = 0; q < BM / 8; q++) tma_load_2d(As + q * (BK * 8), &mapA, m0 + q * 8, k0, full + s);
                }
#pragma unroll
                for (int q = 0; q < BN / 8; q++) tma_load_2d(Bs + q * (BK * 8), &mapB, n0 + q * 8, k0, full + s);
                // per-k coefficients of this chunk (the caller pads kscale to a multiple of BK doubles)
                if (scaled) bulk_load_1d(Bs + S::BSZ, p.kscale + k0, BK * (int) sizeof(double), full + s);
            }
        }
        return;
    }

    // ===== DMMA consumers =====
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, t = lane & 3;
    double acc[MA][NA][2];
#pragma unroll
    for (int i = 0; i < MA; i++)
#pragma unroll
        for (int j = 0; j < NA; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kc = 0; kc < nk; kc++) {
        const int s = kc % STAGES, ph = (kc / STAGES) & 1;
        mbar_wait(full + s, ph);
        const double *As = smem + (size_t) s * S::STAGE;
        const double *Bs = As + S::ASZ;
        const double *Ks = Bs + S::BSZ;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ks++) {
            double a[MA], b[NA];
            if (FORM == 1) {
#pragma unroll
                for (int i = 0; i < MA; i++) a[i] = As[ks * (BM * 4) + (wm * (MA * 8) + i * 8 + g) * 4 + t];
            } else {
                // rows k >= K of the operands are zero-filled by TMA, so a stale coefficient there is harmless
                // as long as it is finite: the pad of kscale is zero
                const double sc = (p.kscale != nullptr) ? Ks[ks * 4 + t] : 1.0;
#pragma unroll
                for (int i = 0; i < MA; i++) a[i] = As[(wm * MA + i) * (BK * 8) + (ks * 4 + t) * 8 + g] * sc;
            }
#pragma unroll
            for (int j = 0; j < NA; j++) b[j] = Bs[(wn * NA + j) * (BK * 8) + (ks * 4 + t) * 8 + g];
#pragma unroll
            for (int i = 0; i < MA; i++)
#pragma unroll
                for (int j = 0; j < NA; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }

    // ---- epilogues (identical to gemm_f64.cuh; only the 256 consumer threads take part)
    if (EPI == EPI_ROWSQ) {
        asm volatile("bar.sync 1, %0;\n" :: "n"(T_CONSUMERS));
        double *rs = smem;
#pragma unroll
        for (int i = 0; i < MA; i++) {
            double sacc = 0.0;
#pragma unroll
            for (int j = 0; j < NA; j++) {
                sacc = __dadd_rn(sacc, __dmul_rn(acc[i][j][0], acc[i][j][0]));
                sacc = __dadd_rn(sacc, __dmul_rn(acc[i][j][1], acc[i][j][1]));
            }
            sacc = __dadd_rn(sacc, __shfl_xor_sync(0xffffffffu, sacc, 1));
            sacc = __dadd_rn(sacc, __shfl_xor_sync(0xffffffffu, sacc, 2));
            if (t == 0) rs[wn * BM + wm * (MA * 8) + i * 8 + g] = sacc;
        }
        asm volatile("bar.sync 1, %0;\n" :: "n"(T_CONSUMERS));
        if (threadIdx.x < BM) {
            const int row = m0 + threadIdx.x;
            if (row < p.M) p.partial[(int64_t) tn * p.M + row] = __dadd_rn(rs[threadIdx.x], rs[BM + threadIdx.x]);
        }
        return;
    }

    double c2 = 0.0;
    if (EPI == EPI_SYRK) c2 = *p.dscal;
#pragma unroll
    for (int j = 0; j < NA; j++) {
        const int col = n0 + wn * (NA * 8) + j * 8 + t * 2;
        double v0 = 0.0, v1 = 0.0, l0 = 0.0, l1 = 0.0, h0 = 0.0, h1 = 0.0;
        const bool in0 = col < p.N, in1 = col + 1 < p.N;
        if (EPI == EPI_SAMPLE || EPI == EPI_SYRK) {
            if (in0) v0 = p.vec[col];
            if (in1) v1 = p.vec[col + 1];
        }
        const bool clip = EPI == EPI_SAMPLE && p.lo != nullptr;
        if (clip) {
            if (in0) { l0 = p.lo[col]; h0 = p.hi[col]; }
            if (in1) { l1 = p.lo[col + 1]; h1 = p.hi[col + 1]; }
        }
#pragma unroll
        for (int i = 0; i < MA; i++) {
            const int row = m0 + wm * (MA * 8) + i * 8 + g;
            if (row >= p.M || !in0) continue;
            double *dst = p.C + (int64_t) row * p.ldc + col;
            double o0 = acc[i][j][0], o1 = acc[i][j][1];
            if (EPI == EPI_SAMPLE) {
                o0 = v0 + p.s0 * o0;
                o1 = v1 + p.s0 * o1;
                if (clip) {
                    o0 = fmax(l0, fmin(o0, h0));
                    o1 = fmax(l1, fmin(o1, h1));
                }
            } else if (EPI == EPI_SYRK) {
                if (col > row) continue;
                const double pr = p.vec[row];
                const double old0 = dst[0];
                o0 = (p.s0 * old0 + p.s1 * (pr * v0 + c2 * old0)) + o0;
                if (in1 && col + 1 <= row) {
                    const double old1 = dst[1];
                    o1 = (p.s0 * old1 + p.s1 * (pr * v1 + c2 * old1)) + o1;
                    dst[0] = o0;
                    dst[1] = o1;
                } else {
                    dst[0] = o0;
                }
                continue;
            }
            if (in1) *reinterpret_cast<double2*>(dst) = make_double2(o0, o1); // col even, ldc even: 16-byte aligned
            else dst[0] = o0;
        }
    }
}

// ---- host side: tensor maps through the driver entry point (no -lcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
        CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encoder() {
    static EncodeTiledFn fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
                q != cudaDriverEntryPointSuccess) ptr = nullptr;
        return (EncodeTiledFn) ptr;
    }();
    return fn;
}

// 2-D fp64 tensor [rows][cols] (cols contiguous, pitch ld doubles), box [box_rows][box_cols]
inline int make_map(CUtensorMap *map, const double *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
    EncodeTiledFn enc = tma_encoder();
    if (!enc) return fail(BBGPU_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = { (cuuint64_t) cols, (cuuint64_t) rows };
    const cuuint64_t strides[1] = { (cuuint64_t) ld * 8 };
    const cuuint32_t box[2] = { (cuuint32_t) box_cols, (cuuint32_t) box_rows };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(BBGPU_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int) r);
    return BBGPU_OK;
}

template<int FORM, int MA, int NA, int BK, int STAGES, int EPI>
inline int launch_gemm_tma_tile(bbgpu_ctx *ctx, const GemmArgs &a) {
    using S = TmaSmem<MA, NA, BK, STAGES>;
    CUtensorMap mapA, mapB;
    if (FORM == 1) BB_TRY(make_map(&mapA, a.A, a.M, a.K, a.lda, S::BM, 4));
    else BB_TRY(make_map(&mapA, a.A, a.K, a.M, a.lda, BK, 8));
    BB_TRY(make_map(&mapB, a.B, a.K, a.N, a.ldb, BK, 8));
    BB_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<FORM, MA, NA, BK, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) S::BYTES));
    const int tiles = ceil_div(a.M, S::BM) * ceil_div(a.N, S::BN);
    gemm_tma_kernel<FORM, MA, NA, BK, STAGES, EPI><<<tiles, T_THREADS, S::BYTES, ctx->stream>>>(mapA, mapB, a);
    return launched(ctx, "gemm_tma_kernel");
}

// TMA needs 16-byte aligned bases and pitches that are multiples of 16 bytes (all engine buffers are:
// 256-byte cudaMalloc alignment, pitch = multiple of 8 doubles)
inline bool tma_ok(const GemmArgs &a) {
    return tma_encoder() != nullptr && (a.lda % 2 == 0) && (a.ldb % 2 == 0) && (((uintptr_t) a.A) % 16 == 0) &&
            (((uintptr_t) a.B) % 16 == 0) && (a.kscale == nullptr || (a.kscale_padded && ((uintptr_t) a.kscale) % 16 == 0));
}

// The one entry point the engine uses.  TMA path by default; BBGPU_GEMM_TMA=0 selects the cp.async
// ring of gemm_f64.cuh (kept as the cross-check in tests/test_gpu_kernels.py).
template<int FORM, int EPI>
inline int launch_gemm(bbgpu_ctx *ctx, const GemmArgs &a, int *tiles_n_out = nullptr) {
    if (a.M <= 0 || a.N <= 0) return BBGPU_OK;
    static const bool use_tma = [] { const char *e = std::getenv("BBGPU_GEMM_TMA"); return e ? std::atoi(e) != 0 : true; }();
    if (!use_tma || !tma_ok(a)) return launch_gemm_cpasync<FORM, EPI>(ctx, a, tiles_n_out);
    switch (pick_tile(ctx, a, EPI == EPI_SYRK)) {
    case 2:
        if (tiles_n_out) *tiles_n_out = ceil_div(a.N, 64);
        return launch_gemm_tma_tile<FORM, 2, 4, 32, 4, EPI>(ctx, a);
    case 1:
    case 4:
        if (tiles_n_out) *tiles_n_out = ceil_div(a.N, 128);
        return launch_gemm_tma_tile<FORM, 4, 8, 32, 3, EPI>(ctx, a);
    default:
        if (tiles_n_out) *tiles_n_out = ceil_div(a.N, 112);
        return launch_gemm_tma_tile<FORM, 4, 7, 32, 3, EPI>(ctx, a);
    }
}

} // namespace bb
