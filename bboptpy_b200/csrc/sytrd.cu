// Householder tridiagonalisation of one dense symmetric matrix as ONE persistent cooperative kernel.
//
// Why: at the headline configuration (n = 1000) the reference decomposes C after every generation, and 9.6 ms of
// cuSOLVER's 12.0 ms go into `Dsytrd` -- 28 panel kernels, ~10 us per column, bound by launch / barrier latency and not by
// bandwidth (DESIGN.md "The eigendecomposition at n = 1000").  This kernel is the first half of the replacement (the
// second half, a tridiagonal divide & conquer, is next round's work; cuSOLVER exports no device-side tridiagonal solver,
// so this kernel is measured on its own for now: bbgpu_sytrd_host, tests/test_gpu_kernels.py, tools/sytrd_probe.py).
//
// Layout: row r of the matrix lives in the shared memory of CTA r % gridDim (rows are dealt cyclically so the shrinking
// trailing block stays balanced), BOTH triangles are kept up to date, so A v is row-local: no cross-CTA reduction.
// One grid barrier per column.  Before barrier i every CTA has written y = A22 v_i for its rows, and the owner of row
// i + 1 has published that row (its state before update i); after the barrier every CTA redundantly forms w_i (needs all
// of y), applies update i to the published row -- that is column i + 1 of the updated matrix -- to get the next
// Householder vector v_{i+1}, then makes ONE pass over its own rows that applies A22 -= v w^T + w v^T and accumulates
// y = A22 v_{i+1} for the next barrier.
// Output in LAPACK's dsytrd('L') convention: d, e, tau, and the reflectors below the subdiagonal of a column-major V
// (what cusolverDnDormtr consumes).
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace {

namespace cg = cooperative_groups;
constexpr int SY_THREADS = 256;

// Grid barrier on a monotonically increasing counter (one atomic per CTA, one polling thread per CTA): the kernel is
// launched cooperatively, so all CTAs are resident.  ~1 us at 148 CTAs against ~2.5 us for cg::grid_group::sync().
__device__ __forceinline__ void sy_grid_barrier(unsigned *counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory"); // release: the CTA's stores first
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

constexpr int SY_KMAX = 8;   // columns per thread: n <= 2048
constexpr int SY_RMAX = 16;  // rows per CTA (their scalars and partial sums live in registers): n <= 16 * gridDim

__global__ void __launch_bounds__(SY_THREADS, 1) sytrd_coop_kernel(const double *__restrict__ A, int n, int ldA, double *__restrict__ d,
        double *__restrict__ e, double *__restrict__ tau, double *__restrict__ Vcm, double *ybuf, double *rowbuf, unsigned *barrier,
        long long *prof) {
    extern __shared__ __align__(16) double sm[];
    const int G = gridDim.x, me = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = SY_THREADS / 32;
    const int np = (n + 63) & ~63;                        // row pitch: whole 64-column chunks, pad columns stay zero
    const int nrows = me < n ? (n - me + G - 1) / G : 0;  // rows me, me + G, ...
    const int cap = (n + G - 1) / G;
    double *rows = sm;                          // cap x np
    double *v = rows + (size_t) cap * np;       // current reflector v_i by global row (0 for c <= i and in the padding)
    double *w = v + np;
    double *vn = w + np;                        // next reflector
    double *red = vn + np;                      // 32
    double *part = red + 32;                    // NW x cap partial row sums of the pass
    for (int q = 0; q < cap; q++) {
        const int r = me + q * G;
        for (int c = tid; c < np; c += SY_THREADS)
            rows[(size_t) q * np + c] = (r < n && c < n) ? A[(size_t) max(r, c) * ldA + min(r, c)] : 0.0;
    }
    // ---- reflector 0 from column 0 of the input.  Thread t owns columns t, t + 256, ... of every vector in registers.
    double vreg[SY_KMAX];
    double tau_i;
    {
        double ss = 0.0;
#pragma unroll
        for (int k = 0; k < SY_KMAX; k++) {
            const int c = tid + k * SY_THREADS;
            vreg[k] = (c >= 1 && c < n) ? A[(size_t) c * ldA] : 0.0;
            if (c >= 2) ss += vreg[k] * vreg[k];
        }
        ss = bb::block_sum(ss, red);
        const double alpha = A[(size_t) ldA];
        double beta = alpha, scal = 0.0;
        tau_i = 0.0;
        if (ss != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + ss), alpha);
            tau_i = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
#pragma unroll
        for (int k = 0; k < SY_KMAX; k++) {
            const int c = tid + k * SY_THREADS;
            vreg[k] = c == 1 ? 1.0 : vreg[k] * scal;
            if (c < np) { v[c] = vreg[k]; w[c] = 0.0; vn[c] = 0.0; }
        }
        if (me == 0 && tid == 0) { d[0] = A[0]; e[0] = beta; tau[0] = tau_i; }
    }
    __syncthreads();
    // y_0 = A v_0 for the own rows r > 0; the owner of row 1 publishes it
    for (int q = wid; q < nrows; q += NW) {
        const int r = me + q * G;
        if (r < 1) continue;
        const double *row = rows + (size_t) q * np;
        double s = 0.0;
        for (int c = 1 + lane; c < n; c += 32) s = fma(row[c], v[c], s);
        s = bb::warp_sum(s);
        if (lane == 0) ybuf[r] = s;
    }
    if (1 % G == me) {
        const double *row = rows + (size_t) (1 / G) * np;
        for (int c = tid; c < n; c += SY_THREADS) rowbuf[c] = row[c];
    }
    long long tacc[6] = { 0, 0, 0, 0, 0, 0 }, t0 = clock64(), t1;
#define SY_TICK(k) do { t1 = clock64(); tacc[k] += t1 - t0; t0 = t1; } while (0)
    for (int i = 0; i < n - 1; i++) {
        const int parity = i & 1;
        SY_TICK(5);
        sy_grid_barrier(barrier, (unsigned) (i + 1) * (unsigned) G);
        SY_TICK(0);
        // y_i and row i+1 of A^(i) were written by other SMs: read them past L1 (a line of the same buffer from two steps
        // ago may still sit there).  Everything a thread needs comes in ONE round trip: its own columns plus the three
        // scalars every thread would otherwise wait a broadcast for.
        const double *y = ybuf + (size_t) parity * n, *old = rowbuf + (size_t) parity * n;
        double yk[SY_KMAX], ok[SY_KMAX];
#pragma unroll
        for (int k = 0; k < SY_KMAX; k++) {
            const int c = tid + k * SY_THREADS;
            const bool in = c > i && c < n;
            yk[k] = in ? __ldcg(y + c) : 0.0;
            ok[k] = in ? __ldcg(old + c) : 0.0;
        }
        const double y1 = __ldcg(y + i + 1), o1 = __ldcg(old + i + 1);
        const bool more = i + 2 < n;
        const double y2 = more ? __ldcg(y + i + 2) : 0.0, o2 = more ? __ldcg(old + i + 2) : 0.0;
        double dot = 0.0;
#pragma unroll
        for (int k = 0; k < SY_KMAX; k++) dot = fma(yk[k], vreg[k], dot);
        dot = bb::block_sum(dot, red);
        SY_TICK(1);
        // w_i = tau y - (tau^2 / 2)(y . v) v, every CTA redundantly (same order, same bits)
        const double half = 0.5 * tau_i * tau_i * dot;
        const double wr1 = tau_i * y1 - half;  // w_i[i+1]; v_i[i+1] = 1
        const double diag = o1 - 2.0 * wr1;    // A^(i+1)[i+1][i+1]
        // reflector i into V (column-major; rows i+2.. hold v_i[i+2..], the leading 1 at row i+1 is implicit)
        if (me == i % G) {
#pragma unroll
            for (int k = 0; k < SY_KMAX; k++) {
                const int c = tid + k * SY_THREADS;
                if (c >= i + 2 && c < n) Vcm[(size_t) i * n + c] = vreg[k];
            }
        }
        if (!more) { // last column: only d[n-1] is left
            if (me == 0 && tid == 0) d[n - 1] = diag;
            break;
        }
        // column i+1 of A^(i+1) (= row i+1 by symmetry) from the published row, then reflector i+1
        double xk[SY_KMAX], ss = 0.0;
#pragma unroll
        for (int k = 0; k < SY_KMAX; k++) {
            const int c = tid + k * SY_THREADS;
            const double wc = (c > i && c < n) ? tau_i * yk[k] - half * vreg[k] : 0.0;
            if (c < np) w[c] = wc;
            xk[k] = (c > i + 1 && c < n) ? ok[k] - wc - wr1 * vreg[k] : 0.0;
            if (c > i + 2) ss = fma(xk[k], xk[k], ss);
        }
        SY_TICK(2);
        ss = bb::block_sum(ss, red);
        const double v2 = v[i + 2];
        const double alpha = o2 - (tau_i * y2 - half * v2) - wr1 * v2;
        double beta_n = alpha, tau_n = 0.0, scal = 0.0;
        if (ss != 0.0) {
            beta_n = -copysign(sqrt(alpha * alpha + ss), alpha);
            tau_n = (beta_n - alpha) / beta_n;
            scal = 1.0 / (alpha - beta_n);
        }
#pragma unroll
        for (int k = 0; k < SY_KMAX; k++) {
            const int c = tid + k * SY_THREADS;
            vreg[k] = c == i + 2 ? 1.0 : xk[k] * scal; // 0 for c <= i+1 and beyond n
            if (c < np) vn[c] = vreg[k];
        }
        if (me == (i + 1) % G && tid == 0) { d[i + 1] = diag; e[i + 1] = beta_n; tau[i + 1] = tau_n; }
        __syncthreads();
        SY_TICK(3);
        // ---- one pass over the own rows: A22 -= v w^T + w v^T and y_{i+1} = A22 v_{i+1}.  A warp takes 64-column chunks,
        // loads w, v, v_{i+1} of the chunk once and walks the CTA's rows with them (shared-memory traffic is the bound of
        // this phase: the vectors are read once per chunk, not once per row).  Columns <= i need no mask: w = v = 0 there.
        {
            double acc[SY_RMAX], vr[SY_RMAX], wr[SY_RMAX];
            unsigned live = 0u;
#pragma unroll
            for (int q = 0; q < SY_RMAX; q++) {
                acc[q] = 0.0;
                const int r = me + q * G;
                const bool on = q < nrows && r > i;
                if (on) live |= 1u << q;
                vr[q] = on ? v[r] : 0.0;
                wr[q] = on ? w[r] : 0.0;
            }
            for (int c0 = ((i + 1) & ~63) + 64 * wid; c0 < np; c0 += 64 * NW) {
                const int c = c0 + 2 * lane;
                const double2 w2 = *reinterpret_cast<const double2*>(w + c), v2c = *reinterpret_cast<const double2*>(v + c),
                        n2 = *reinterpret_cast<const double2*>(vn + c);
#pragma unroll
                for (int q = 0; q < SY_RMAX; q++) {
                    if (live & (1u << q)) {
                        double2 *p = reinterpret_cast<double2*>(rows + (size_t) q * np + c);
                        double2 a2 = *p;
                        a2.x = a2.x - vr[q] * w2.x - wr[q] * v2c.x;
                        a2.y = a2.y - vr[q] * w2.y - wr[q] * v2c.y;
                        *p = a2;
                        acc[q] = fma(a2.x, n2.x, fma(a2.y, n2.y, acc[q]));
                    }
                }
            }
            // the 16 row sums of a warp in one butterfly: each level halves the values a lane carries (16 double shuffles
            // instead of 16 x 5); lane l ends with the total of row l >> 1
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const bool up = (lane & 16) != 0;
                const double send = up ? acc[q] : acc[q + 8], keep = up ? acc[q + 8] : acc[q];
                acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const bool up = (lane & 8) != 0;
                const double send = up ? acc[q] : acc[q + 4], keep = up ? acc[q + 4] : acc[q];
                acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const bool up = (lane & 4) != 0;
                const double send = up ? acc[q] : acc[q + 2], keep = up ? acc[q + 2] : acc[q];
                acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            {
                const bool up = (lane & 2) != 0;
                const double send = up ? acc[0] : acc[1], keep = up ? acc[1] : acc[0];
                acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
            }
            if ((lane & 1) == 0 && (lane >> 1) < nrows) part[wid * cap + (lane >> 1)] = acc[0];
            __syncthreads();
            if (tid < nrows) {
                const int r = me + tid * G;
                if (r > i + 1) {
                    double s = 0.0;
                    for (int u = 0; u < NW; u++) s += part[u * cap + tid];
                    ybuf[(size_t) (parity ^ 1) * n + r] = s;
                }
            }
        }
        if ((i + 2) % G == me) { // row i+2 of A^(i+1): the source of reflector i+2
            const double *row = rows + (size_t) ((i + 2) / G) * np;
            for (int c = tid; c < n; c += SY_THREADS) rowbuf[(size_t) (parity ^ 1) * n + c] = row[c];
        }
        SY_TICK(4);
        { double *t = v; v = vn; vn = t; } // roll: v_{i+1} becomes the current reflector
        tau_i = tau_n;
    }
    if (prof != nullptr && tid == 0 && me < 4)
        for (int k = 0; k < 6; k++) prof[me * 6 + k] = tacc[k];
#undef SY_TICK
}


// ---- register-resident variant (n <= 1024, at most SR_RQ rows per CTA: n <= 1036 on 148 SMs) ---------------------------
// Same algorithm, same exchange, same outputs as sytrd_coop_kernel; what changes is where the matrix lives.  Thread t already
// keeps columns t, t + 256, ... of every VECTOR in registers; here it also keeps those columns of the CTA's (at most 7)
// ROWS: 28 doubles.  The pass of a column -- A22 -= v w^T + w v^T, y = A22 v_next -- is then 84 register FMAs per thread
// with the reflector columns it already holds, instead of a sweep over 54 KB of shared memory (2 160 of the 7 850 cycles of
// a column, profiles/r01_ncu_eig_own.txt).  The two block-wide sums on the column's dependency chain use a flat second
// stage (eight partials read back by every thread) and the reflector scalars come from a fast reciprocal.
constexpr int SR_RQ = 7;  // rows per CTA
constexpr int SR_KQ = 4;  // columns per thread

__device__ __forceinline__ double sr_block_sum(double v, double *red8, int lane, int wid) {
    v = bb::warp_sum(v);
    __syncthreads();
    if (lane == 0) red8[wid] = v;
    __syncthreads();
    const double4 lo = *reinterpret_cast<const double4*>(red8), hi = *reinterpret_cast<const double4*>(red8 + 4);
    return ((lo.x + lo.y) + (lo.z + lo.w)) + ((hi.x + hi.y) + (hi.z + hi.w));
}

__device__ __forceinline__ double sr_rcp(double p) { // 1 / p, ~1 ulp, for a normal non-zero p
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    r = fma(r, fma(-p, r, 1.0), r);
    r = fma(r, fma(-p, r, 1.0), r);
    return r;
}

// ---- data-carrying exchange (FLAGS variant of the register-resident kernel) ---------------------------------------------------
// With the counter barrier a column pays three dependent L2 round trips between the last store of y and the first use of it:
// the release of the arrive (the stores must be performed first), the poll of the counter, the loads of y and the pivot row.
// Here an exchanged double is its own flag: the slot holds a sentinel (all bits set: a NaN no arithmetic produces) until the
// value lands, and readers poll the slots they need -- no barrier, no fence, one round trip; an aligned 8-byte access is
// single-copy atomic.  Slots are recycled through FOUR buffers: the data of step s lives in buffer s % 4, its owner resets
// the slots to the sentinel at its own step s - 3 (followed by a fence) and writes them at step s - 1 (same thread, same address).
// A CTA can be at most one step ahead of another (it cannot finish the poll of step i + 1 before every CTA has published for
// it, which each does after its own poll of step i), so at the owner's step s - 3 every reader of step s - 4 is done with the
// buffer.  The exchange area is filled with the sentinel before the launch (cudaMemset 0xFF).
// MEASURED AND NOT USED BY DEFAULT (profiles/r02_sytrd_flags_rejected.txt): 3.45 ms against 2.95 ms for the barrier at n = 1000.
// 38 000 threads polling L2 cost more than the round trip they save; a first form with {32 data bits, 32-bit tag} words,
// NCCL-LL style, doubled the polled bytes and measured 3.33 ms.
constexpr unsigned long long XQ_EMPTY = ~0ull;
__device__ __forceinline__ void xq_store(double *p, double val) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(p), "d"(val) : "memory");
}
__device__ __forceinline__ void xq_reset(double *p) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(XQ_EMPTY) : "memory");
}
__device__ __forceinline__ bool xq_load(const double *p, double &val) {
    unsigned long long b;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(b) : "l"(p) : "memory");
    if (b == XQ_EMPTY) return false;
    val = __longlong_as_double((long long) b);
    return true;
}
constexpr int XQ_SPIN_CAP = 1 << 22; // polls before a thread gives up and raises barrier[1] (a lost CTA must not hang the device)

template<bool FLAGS>
__global__ void __launch_bounds__(SY_THREADS, 1) sytrd_reg_kernel(const double *__restrict__ A, int n, int ldA, double *__restrict__ d,
        double *__restrict__ e, double *__restrict__ tau, double *__restrict__ Vcm, double *ybuf, double *rowbuf, unsigned *barrier) {
    __shared__ __align__(16) double v[SR_KQ * SY_THREADS], w[SR_KQ * SY_THREADS];
    __shared__ __align__(32) double red8[8];
    __shared__ double part[8][8];
    __shared__ double s_sc[4]; // FLAGS: y[i+1], row[i+1], y[i+2], row[i+2] from the threads that own those columns
    const int G = gridDim.x, me = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nrows = me < n ? (n - me + G - 1) / G : 0;
    double a[SR_RQ][SR_KQ];
#pragma unroll
    for (int q = 0; q < SR_RQ; q++) {
        const int r = me + q * G;
#pragma unroll
        for (int k = 0; k < SR_KQ; k++) {
            const int c = tid + k * SY_THREADS;
            a[q][k] = (r < n && c < n) ? A[(size_t) max(r, c) * ldA + min(r, c)] : 0.0;
        }
    }
    // ---- reflector 0 from column 0 of the input
    double vreg[SR_KQ];
    double tau_i;
    {
        double ss = 0.0;
#pragma unroll
        for (int k = 0; k < SR_KQ; k++) {
            const int c = tid + k * SY_THREADS;
            vreg[k] = (c >= 1 && c < n) ? A[(size_t) c * ldA] : 0.0;
            if (c >= 2) ss += vreg[k] * vreg[k];
        }
        ss = sr_block_sum(ss, red8, lane, wid);
        const double alpha = A[(size_t) ldA];
        double beta = alpha, scal = 0.0;
        tau_i = 0.0;
        if (ss != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + ss), alpha);
            tau_i = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
#pragma unroll
        for (int k = 0; k < SR_KQ; k++) {
            const int c = tid + k * SY_THREADS;
            vreg[k] = c == 1 ? 1.0 : vreg[k] * scal;
            v[c] = vreg[k];
            w[c] = 0.0;
        }
        if (me == 0 && tid == 0) { d[0] = A[0]; e[0] = beta; tau[0] = tau_i; }
    }
    __syncthreads();
    // y_0 = A v_0 for the own rows r > 0 (row sums over the thread-owned columns), and the pivot row 1
    auto row_sums = [&](const double (&vn)[SR_KQ], int i_done, int parity_out) {
        // acc[q] = sum over this thread's columns of a[q][k] * vn[k]; eight values (7 rows + a zero) reduced over the warp by
        // a butterfly that halves the values a lane carries, then over the eight warps through shared memory
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            acc[q] = 0.0;
            if (q < SR_RQ) {
#pragma unroll
                for (int k = 0; k < SR_KQ; k++) acc[q] = fma(a[q][k], vn[k], acc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const bool up = (lane & 16) != 0;
            const double send = up ? acc[q] : acc[q + 4], keep = up ? acc[q + 4] : acc[q];
            acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const bool up = (lane & 8) != 0;
            const double send = up ? acc[q] : acc[q + 2], keep = up ? acc[q + 2] : acc[q];
            acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
            const bool up = (lane & 4) != 0;
            const double send = up ? acc[0] : acc[1], keep = up ? acc[1] : acc[0];
            acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 2);
            acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
        }
        if ((lane & 3) == 0) part[wid][lane >> 2] = acc[0]; // row lane >> 2
        __syncthreads();
        if (tid < nrows) {
            const int r = me + tid * G;
            if (r > i_done) {
                double s = 0.0;
#pragma unroll
                for (int u = 0; u < 8; u++) s += part[u][tid];
                if (FLAGS) xq_store(ybuf + (size_t) (i_done & 3) * n + r, s); // consumed by step i_done
                else ybuf[(size_t) parity_out * n + r] = s;
            }
        }
    };
    auto publish_row = [&](int r, int parity_out) { // row r (owned by this CTA) into the exchange buffer
        const int qp = r / G;
#pragma unroll
        for (int q = 0; q < SR_RQ; q++) {
            if (q == qp) {
#pragma unroll
                for (int k = 0; k < SR_KQ; k++) {
                    const int c = tid + k * SY_THREADS;
                    if (c < n) {
                        if (FLAGS) xq_store(rowbuf + (size_t) ((r - 1) & 3) * n + c, a[q][k]); // row r: consumed by step r - 1
                        else rowbuf[(size_t) parity_out * n + c] = a[q][k];
                    }
                }
            }
        }
    };
    row_sums(vreg, 0, 0);
    if (1 % G == me) publish_row(1, 0);
    for (int i = 0; i < n - 1; i++) {
        const int parity = i & 1;
        const bool more = i + 2 < n;
        double yk[SR_KQ], ok[SR_KQ];
        double y1, o1, y2, o2;
        if (FLAGS) {
            const double *yp = ybuf + (size_t) (i & 3) * n, *rp = rowbuf + (size_t) (i & 3) * n;
            unsigned pend = 0u;
#pragma unroll
            for (int k = 0; k < SR_KQ; k++) {
                const int c = tid + k * SY_THREADS;
                yk[k] = 0.0;
                ok[k] = 0.0;
                if (c > i && c < n) pend |= (1u << k) | (16u << k);
            }
            int spins = 0;
            // quiet wait on ONE entry (the others arrive within the same microsecond), then sweep what is still missing
            if (pend && lane == (i & 31)) {
                const int k0 = __ffs(pend) - 1;
                bool got = false;
                while (!got) {
#pragma unroll
                    for (int k = 0; k < SR_KQ; k++)
                        if (k == k0) got = xq_load(yp + tid + k * SY_THREADS, yk[k]);
                    if (++spins > XQ_SPIN_CAP) { barrier[1] = 1u; break; }
                }
                pend &= ~(1u << k0);
            }
            __syncwarp();
            while (pend) {
#pragma unroll
                for (int k = 0; k < SR_KQ; k++) {
                    const int c = tid + k * SY_THREADS;
                    if ((pend & (1u << k)) && xq_load(yp + c, yk[k])) pend &= ~(1u << k);
                    if ((pend & (16u << k)) && xq_load(rp + c, ok[k])) pend &= ~(16u << k);
                }
                if (++spins > XQ_SPIN_CAP) { barrier[1] = 1u; break; }
            }
#pragma unroll
            for (int k = 0; k < SR_KQ; k++) {
                const int c = tid + k * SY_THREADS;
                if (c == i + 1) { s_sc[0] = yk[k]; s_sc[1] = ok[k]; }
                if (c == i + 2 && more) { s_sc[2] = yk[k]; s_sc[3] = ok[k]; }
            }
        } else {
            sy_grid_barrier(barrier, (unsigned) (i + 1) * (unsigned) G);
            const double *y = ybuf + (size_t) parity * n, *old = rowbuf + (size_t) parity * n;
#pragma unroll
            for (int k = 0; k < SR_KQ; k++) {
                const int c = tid + k * SY_THREADS;
                const bool in = c > i && c < n;
                yk[k] = in ? __ldcg(y + c) : 0.0;
                ok[k] = in ? __ldcg(old + c) : 0.0;
            }
            y1 = __ldcg(y + i + 1);
            o1 = __ldcg(old + i + 1);
            y2 = more ? __ldcg(y + i + 2) : 0.0;
            o2 = more ? __ldcg(old + i + 2) : 0.0;
        }
        double dot = 0.0;
#pragma unroll
        for (int k = 0; k < SR_KQ; k++) dot = fma(yk[k], vreg[k], dot);
        dot = sr_block_sum(dot, red8, lane, wid); // (its barriers publish s_sc)
        if (FLAGS) {
            y1 = s_sc[0];
            o1 = s_sc[1];
            y2 = more ? s_sc[2] : 0.0;
            o2 = more ? s_sc[3] : 0.0;
        }
        const double half = 0.5 * tau_i * tau_i * dot;
        const double wr1 = tau_i * y1 - half;
        const double diag = o1 - 2.0 * wr1;
        if (me == i % G) {
#pragma unroll
            for (int k = 0; k < SR_KQ; k++) {
                const int c = tid + k * SY_THREADS;
                if (c >= i + 2 && c < n) Vcm[(size_t) i * n + c] = vreg[k];
            }
        }
        if (!more) {
            if (me == 0 && tid == 0) d[n - 1] = diag;
            break;
        }
        const double v2 = v[i + 2]; // v_i[i+2], before w / v are overwritten below
        double wk[SR_KQ], xk[SR_KQ], ss = 0.0;
#pragma unroll
        for (int k = 0; k < SR_KQ; k++) {
            const int c = tid + k * SY_THREADS;
            wk[k] = (c > i && c < n) ? tau_i * yk[k] - half * vreg[k] : 0.0;
            xk[k] = (c > i + 1 && c < n) ? ok[k] - wk[k] - wr1 * vreg[k] : 0.0;
            if (c > i + 2) ss = fma(xk[k], xk[k], ss);
        }
        ss = sr_block_sum(ss, red8, lane, wid); // its barriers also order the reads of v above against the writes below
        const double alpha = o2 - (tau_i * y2 - half * v2) - wr1 * v2;
        double beta_n = alpha, tau_n = 0.0, scal = 0.0;
        if (ss != 0.0) {
            beta_n = -copysign(sqrt(alpha * alpha + ss), alpha);
            tau_n = (beta_n - alpha) * sr_rcp(beta_n);
            scal = sr_rcp(alpha - beta_n);
        }
        // the rows of this CTA need v_i and w_i at THEIR column index: through shared memory (v still holds v_i)
#pragma unroll
        for (int k = 0; k < SR_KQ; k++) w[tid + k * SY_THREADS] = wk[k];
        __syncthreads();
        double vr[SR_RQ], wr[SR_RQ];
#pragma unroll
        for (int q = 0; q < SR_RQ; q++) {
            const int r = me + q * G;
            const bool on = q < nrows && r > i;
            vr[q] = on ? v[r] : 0.0;
            wr[q] = on ? w[r] : 0.0;
        }
        double vnk[SR_KQ];
#pragma unroll
        for (int k = 0; k < SR_KQ; k++) {
            const int c = tid + k * SY_THREADS;
            vnk[k] = c == i + 2 ? 1.0 : xk[k] * scal; // 0 for c <= i+1 and beyond n
        }
        if (me == (i + 1) % G && tid == 0) { d[i + 1] = diag; e[i + 1] = beta_n; tau[i + 1] = tau_n; }
        // ---- the pass, in registers: A22 -= v w^T + w v^T on the own rows (columns <= i: w = v = 0, nothing changes)
#pragma unroll
        for (int q = 0; q < SR_RQ; q++) {
#pragma unroll
            for (int k = 0; k < SR_KQ; k++) a[q][k] = a[q][k] - vr[q] * wk[k] - wr[q] * vreg[k];
        }
        __syncthreads(); // every thread has read v_i[r] / w_i[r]
#pragma unroll
        for (int k = 0; k < SR_KQ; k++) {
            vreg[k] = vnk[k];
            v[tid + k * SY_THREADS] = vnk[k];
        }
        row_sums(vreg, i + 1, parity ^ 1);
        if ((i + 2) % G == me) publish_row(i + 2, parity ^ 1);
        if (FLAGS) {
            // recycle the slots of step i + 3 (buffer (i + 3) & 3 held step i - 1, which every CTA has consumed: see above): our y
            // slots, and the pivot row if that row (i + 4) is ours.  We fill them two steps from now.  The fence is off the
            // critical path -- the next thing these threads do is wait for the other CTAs -- and makes sure an empty slot is in L2
            // before anything this CTA publishes later can be seen.
            bool any = false;
            if (tid < nrows) {
                xq_reset(ybuf + (size_t) ((i + 3) & 3) * n + me + tid * G);
                any = true;
            }
            if ((i + 4) % G == me) {
#pragma unroll
                for (int k = 0; k < SR_KQ; k++) {
                    const int c = tid + k * SY_THREADS;
                    if (c < n) xq_reset(rowbuf + (size_t) ((i + 3) & 3) * n + c);
                }
                any = true;
            }
            if (any) __threadfence();
        }
        tau_i = tau_n;
    }
}

} // namespace

namespace bb {

// limits of the cooperative kernel on this device: shared memory for the rows of one CTA, columns per thread
bool sytrd_fits(bbgpu_ctx *ctx, int n) {
    const int G = ctx->sm_count, nrows = (n + G - 1) / G, np = (n + 63) & ~63;
    const size_t smem = ((size_t) nrows * np + 3 * (size_t) np + 32 + 8 * (size_t) nrows) * sizeof(double);
    return n >= 3 && n <= SY_KMAX * SY_THREADS && nrows <= SY_RMAX && smem <= 200 * 1024;
}

size_t sytrd_workspace_doubles(int n) { return 8 * (size_t) n + 64; } // y and pivot rows: 4 buffers x n slots each; barrier words

} // namespace bb

namespace {

// ---- Q^T from the reflectors, row by row ------------------------------------------------------------------------------
// Row r of Q^T = e_r^T H_{n-2} ... H_1 H_0 (the H_i are symmetric).  Multiplying a ROW by H_i from the right is local to the
// row: q <- q - tau_i (q . v_i) v_i^T, so a warp keeps one row in registers and walks the reflectors with no exchange
// between CTAs at all; reflectors i >= r leave e_r untouched and are skipped (the 4/3 n^3 flops of orgtr, not 2 n^3).  The
// CTA stages v_i in shared memory (double buffer, the next column is prefetched while this one is applied).  Runs on the
// side stream WHILE the tridiagonal divide & conquer runs on the main one; the back-transformation is then one GEMM.
constexpr int QR_THREADS = 256, QR_WARPS = QR_THREADS / 32;

template<int KQ>
__global__ void __launch_bounds__(QR_THREADS) q_rows_kernel(const double *__restrict__ Vcm, const double *__restrict__ tau, int n,
        double *__restrict__ R, int ldr) {
    constexpr int NP = KQ * 32;
    __shared__ __align__(16) double vbuf[2][NP];
    __shared__ double stau[NP];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r = blockIdx.x + wid * gridDim.x; // the rows of a CTA are spread over the matrix: equal work per CTA
    const int rtop = min(blockIdx.x + (QR_WARPS - 1) * (int) gridDim.x, n - 1);
    const int istart = min(rtop - 1, n - 2);    // reflectors above leave every row of this CTA untouched
    double q[KQ];
#pragma unroll
    for (int k = 0; k < KQ; k++) q[k] = (lane + 32 * k == r) ? 1.0 : 0.0;
    for (int c = tid; c < NP; c += QR_THREADS) stau[c] = c < n - 1 ? tau[c] : 0.0;
    // v_i by global row: 0 up to i, the implicit 1 at i+1, the stored part below
    auto fetch = [&](int i, int c) -> double {
        return (i < 0 || c <= i || c >= n) ? 0.0 : (c == i + 1 ? 1.0 : Vcm[(size_t) i * n + c]);
    };
    for (int c = tid; c < NP; c += QR_THREADS) vbuf[istart & 1][c] = fetch(istart, c);
    __syncthreads();
    for (int i = istart; i >= 0; i--) {
        double nxt[NP / QR_THREADS];
#pragma unroll
        for (int j = 0; j < NP / QR_THREADS; j++) nxt[j] = fetch(i - 1, tid + j * QR_THREADS);
        const double *v = vbuf[i & 1];
        if (r < n && r > i) { // warp-uniform
            const int k0 = (i + 1) >> 5;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int k = 0; k < KQ; k += 4) {
                if (k + 3 >= k0) { // v is zero up to i: whole groups of 32 below the reflector are skipped
                    s0 = fma(q[k], v[lane + 32 * k], s0);
                    s1 = fma(q[k + 1], v[lane + 32 * (k + 1)], s1);
                    s2 = fma(q[k + 2], v[lane + 32 * (k + 2)], s2);
                    s3 = fma(q[k + 3], v[lane + 32 * (k + 3)], s3);
                }
            }
            const double t = stau[i] * bb::warp_sum((s0 + s1) + (s2 + s3));
#pragma unroll
            for (int k = 0; k < KQ; k += 4) {
                if (k + 3 >= k0) {
                    q[k] = fma(-t, v[lane + 32 * k], q[k]);
                    q[k + 1] = fma(-t, v[lane + 32 * (k + 1)], q[k + 1]);
                    q[k + 2] = fma(-t, v[lane + 32 * (k + 2)], q[k + 2]);
                    q[k + 3] = fma(-t, v[lane + 32 * (k + 3)], q[k + 3]);
                }
            }
        }
        double *w = vbuf[(i + 1) & 1]; // read last in iteration i+1: every warp is past that iteration's barrier
#pragma unroll
        for (int j = 0; j < NP / QR_THREADS; j++) w[tid + j * QR_THREADS] = nxt[j];
        __syncthreads();
    }
    if (r < n) {
#pragma unroll
        for (int k = 0; k < KQ; k++) {
            const int c = lane + 32 * k;
            if (c < n) R[(size_t) r * ldr + c] = q[k];
        }
    }
}

} // namespace

namespace bb {

// R (n x n, row-major, pitch ldr) = Q^T for the reflectors of sytrd_device, enqueued on `st`
int q_rows_device(bbgpu_ctx *ctx, cudaStream_t st, int n, const double *d_Vcm, const double *d_tau, double *d_R, int ldr) {
    if (n > 2048) return fail(BBGPU_ERR_ARG, "q_rows_device: n = %d is beyond the kernel", n);
    const int grid = (n + QR_WARPS - 1) / QR_WARPS;
    if (n <= 1024) q_rows_kernel<32><<<grid, QR_THREADS, 0, st>>>(d_Vcm, d_tau, n, d_R, ldr);
    else q_rows_kernel<64><<<grid, QR_THREADS, 0, st>>>(d_Vcm, d_tau, n, d_R, ldr);
    return launched(ctx, "q_rows_kernel");
}

// device form: A row-major with pitch ldA (lower triangle read) -> d, e, tau (n each) and the reflectors in Vcm (n*n,
// column-major, LAPACK dsytrd('L') layout); ws = sytrd_workspace_doubles(n).  Enqueued on the ctx stream.
int sytrd_device(bbgpu_ctx *ctx, int n, const double *d_A, int ldA, double *d_d, double *d_e, double *d_tau, double *d_Vcm,
        double *ws) {
    if (!sytrd_fits(ctx, n)) return fail(BBGPU_ERR_ARG, "sytrd_device: n = %d is beyond the cooperative kernel", n);
    cudaStream_t st = ctx->stream;
    const int G = ctx->sm_count, nrows = (n + G - 1) / G, np = (n + 63) & ~63;
    // the register-resident kernel where the rows of a CTA fit its registers (BBGPU_SYTRD=smem keeps the other one)
    static const bool force_smem = [] { const char *e = getenv("BBGPU_SYTRD"); return e && e[0] == 's'; }();
    if (!force_smem && n <= SR_KQ * SY_THREADS && nrows <= SR_RQ) {
        // exchange: counter barrier + plain doubles (default), or self-validating slots without a barrier (BBGPU_SYTRD=flags:
        // measured slower, profiles/r02_sytrd_flags_rejected.txt)
        if (reinterpret_cast<uintptr_t>(ws) & 15) return fail(BBGPU_ERR_ARG, "sytrd_device: the workspace must be 16-byte aligned");
        static const bool use_barrier = [] { const char *e = getenv("BBGPU_SYTRD"); return !(e && e[0] == 'f'); }();
        double *dy = ws, *dr = ws + 4 * (size_t) n; // 4 buffers x n slots each (the barrier variant uses two)
        unsigned *dbar = reinterpret_cast<unsigned*>(ws + 8 * (size_t) n);
        if (use_barrier) BB_CUDA(cudaMemsetAsync(dbar, 0, 64, st));
        else {
            BB_CUDA(cudaMemsetAsync(ws, 0xFF, 8 * (size_t) n * sizeof(double), st)); // every slot empty
            BB_CUDA(cudaMemsetAsync(dbar, 0, 64, st));
        }
        int nn = n, ld = ldA;
        const double *A = d_A;
        void *args[] = { (void*) &A, (void*) &nn, (void*) &ld, (void*) &d_d, (void*) &d_e, (void*) &d_tau, (void*) &d_Vcm, (void*) &dy,
                (void*) &dr, (void*) &dbar };
        const void *fn = use_barrier ? (const void*) sytrd_reg_kernel<false> : (const void*) sytrd_reg_kernel<true>;
        BB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(SY_THREADS), args, 0, st));
        return launched(ctx, "sytrd_reg_kernel");
    }
    const size_t smem = ((size_t) nrows * np + 3 * (size_t) np + 32 + 8 * (size_t) nrows) * sizeof(double);
    BB_CUDA(cudaFuncSetAttribute(sytrd_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    double *dy = ws, *dr = ws + 2 * (size_t) n;
    unsigned *dbar = reinterpret_cast<unsigned*>(ws + 4 * (size_t) n);
    long long *dprof = nullptr;
    BB_CUDA(cudaMemsetAsync(dbar, 0, 64, st));
    int nn = n, ld = ldA;
    const double *A = d_A;
    void *args[] = { (void*) &A, (void*) &nn, (void*) &ld, (void*) &d_d, (void*) &d_e, (void*) &d_tau, (void*) &d_Vcm, (void*) &dy,
            (void*) &dr, (void*) &dbar, (void*) &dprof };
    BB_CUDA(cudaLaunchCooperativeKernel((const void*) sytrd_coop_kernel, dim3(G), dim3(SY_THREADS), args, smem, st));
    return launched(ctx, "sytrd_coop_kernel");
}

} // namespace bb

extern "C" {

// Tridiagonalise one symmetric matrix (host buffers; unit tests and the timing probe).  h_A: n*n row-major, the lower
// triangle is read.  Outputs in LAPACK dsytrd('L') convention: d[n], e[n-1], tau[n-1], and h_V (n*n COLUMN-major) with
// reflector i in rows i+2.. of column i (its implicit leading 1 at row i+1).  *ms receives the kernel time (CUDA events).
int bbgpu_sytrd_host(bbgpu_ctx *ctx, int n, const double *h_A, double *h_d, double *h_e, double *h_tau, double *h_V, int reps,
                     float *ms) {
    BB_ARG(ctx && h_A && h_d && h_e && h_tau && n >= 3 && reps >= 1);
    cudaStream_t st = ctx->stream;
    const int G = ctx->sm_count;
    const int nrows = (n + G - 1) / G;
    const int np = (n + 63) & ~63;
    if (n > SY_KMAX * SY_THREADS || nrows > SY_RMAX)
        return bb::fail(BBGPU_ERR_ARG, "bbgpu_sytrd_host: n = %d is beyond this kernel (n <= %d and <= %d rows per SM)", n, SY_KMAX * SY_THREADS, SY_RMAX);
    const size_t smem = ((size_t) nrows * np + 3 * (size_t) np + 32 + 8 * (size_t) nrows) * sizeof(double);
    if (smem > 220 * 1024) return bb::fail(BBGPU_ERR_ARG, "bbgpu_sytrd_host: n = %d needs %zu bytes of shared memory per CTA", n, smem);
    BB_CUDA(cudaFuncSetAttribute(sytrd_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0;
    BB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sytrd_coop_kernel, SY_THREADS, smem));
    if (per_sm < 1) return bb::fail(BBGPU_ERR_ARG, "bbgpu_sytrd_host: the kernel does not fit an SM");
    double *dA = nullptr, *dd = nullptr, *de = nullptr, *dt = nullptr, *dV = nullptr, *dy = nullptr, *dr = nullptr;
    BB_TRY(bb::dev_alloc(&dA, (size_t) n * n));
    BB_TRY(bb::dev_alloc(&dV, (size_t) n * n));
    BB_TRY(bb::dev_alloc(&dd, (size_t) n));
    BB_TRY(bb::dev_alloc(&de, (size_t) n));
    BB_TRY(bb::dev_alloc(&dt, (size_t) n));
    BB_TRY(bb::dev_alloc(&dy, (size_t) 2 * n));
    BB_TRY(bb::dev_alloc(&dr, (size_t) 2 * n));
    BB_CUDA(cudaMemcpyAsync(dA, h_A, (size_t) n * n * 8, cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemsetAsync(dV, 0, (size_t) n * n * 8, st));
    int nn = n, ld = n;
    unsigned *dbar = nullptr;
    BB_TRY(bb::dev_alloc(&dbar, (size_t) 1));
    long long *dprof = nullptr;
    BB_TRY(bb::dev_alloc(&dprof, (size_t) 24));
    void *args[] = { (void*) &dA, (void*) &nn, (void*) &ld, (void*) &dd, (void*) &de, (void*) &dt, (void*) &dV, (void*) &dy, (void*) &dr, (void*) &dbar, (void*) &dprof };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    // the kernel the engine would pick for this n (register-resident up to n = 1024) unless the per-phase profile of the
    // shared-memory kernel is asked for
    const bool prof_run = getenv("BBGPU_SYTRD_PROF") != nullptr;
    double *ws = nullptr;
    BB_TRY(bb::dev_alloc(&ws, bb::sytrd_workspace_doubles(n)));
    for (int r = 0; r < reps; r++) {
        BB_CUDA(cudaMemsetAsync(dbar, 0, sizeof(unsigned), st));
        cudaEventRecord(e0, st);
        if (prof_run || !bb::sytrd_fits(ctx, n)) { // (this entry accepts sizes up to the hard shared-memory limit)
            BB_CUDA(cudaLaunchCooperativeKernel((const void*) sytrd_coop_kernel, dim3(G), dim3(SY_THREADS), args, smem, st));
            BB_TRY(bb::launched(ctx, "sytrd_coop_kernel"));
        } else {
            BB_TRY(bb::sytrd_device(ctx, n, dA, n, dd, de, dt, dV, ws));
        }
        cudaEventRecord(e1, st);
        BB_CUDA(cudaEventSynchronize(e1));
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        best = t < best ? t : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = best;
    BB_CUDA(cudaMemcpyAsync(h_d, dd, (size_t) n * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(h_e, de, (size_t) (n - 1) * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaMemcpyAsync(h_tau, dt, (size_t) (n - 1) * 8, cudaMemcpyDeviceToHost, st));
    if (h_V) BB_CUDA(cudaMemcpyAsync(h_V, dV, (size_t) n * n * 8, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    if (getenv("BBGPU_SYTRD_PROF")) { // cycles per phase and column (CTAs 0, 1): barrier, load + dot, w + x, reflector, pass, rest
        long long hp[24];
        cudaMemcpy(hp, dprof, sizeof(hp), cudaMemcpyDeviceToHost);
        for (int b = 0; b < 2; b++)
            fprintf(stderr, "sytrd phases (cycles per column, CTA %d): barrier %.0f load+dot %.0f w+x %.0f reflector %.0f pass %.0f other %.0f\n", b,
                    hp[b * 6 + 0] / (double) n, hp[b * 6 + 1] / (double) n, hp[b * 6 + 2] / (double) n, hp[b * 6 + 3] / (double) n,
                    hp[b * 6 + 4] / (double) n, hp[b * 6 + 5] / (double) n);
    }
    cudaFree(dA); cudaFree(dV); cudaFree(dd); cudaFree(de); cudaFree(dt); cudaFree(dy); cudaFree(dr); cudaFree(dbar); cudaFree(dprof);
    cudaFree(ws);
    return BBGPU_OK;
}

} // extern "C"
