// Objective registry entry points + the batched evaluation kernel (see objectives.cuh).
#include <cmath>

#include "objectives.cuh"

namespace bb {

// One warp per candidate row; 8 rows per 256-thread CTA.  HBM-bound: count*n*8 bytes read once.
__global__ void __launch_bounds__(256) eval_rows_kernel(int id, int n, int count, const double *__restrict__ X,
        int64_t ld, const double *__restrict__ aux, double *__restrict__ f) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= count) return;
    const double *x = X + (int64_t) row * ld;
    double v;
    switch (id) { // (warp-uniform)
    case BBGPU_OBJ_SPHERE: v = warp_objective_id<BBGPU_OBJ_SPHERE>(n, x, aux); break;
    case BBGPU_OBJ_ELLIPSOID: v = warp_objective_id<BBGPU_OBJ_ELLIPSOID>(n, x, aux); break;
    case BBGPU_OBJ_ROSENBROCK: v = warp_objective_id<BBGPU_OBJ_ROSENBROCK>(n, x, aux); break;
    case BBGPU_OBJ_RASTRIGIN: v = warp_objective_id<BBGPU_OBJ_RASTRIGIN>(n, x, aux); break;
    case BBGPU_OBJ_ACKLEY: v = warp_objective_id<BBGPU_OBJ_ACKLEY>(n, x, aux); break;
    default: v = warp_objective(id, n, x, aux); break;
    }
    if ((threadIdx.x & 31) == 0) f[row] = v;
}

// Few long rows (LM-CMA-ES: 64 candidates of 10^4 coordinates put 64 warps on the whole GPU): a CTA per row computes the terms
// into shared memory with all its threads, one warp then adds them in the objective's order (warp_reduce_terms: same bits).
template<int ID>
__global__ void __launch_bounds__(256) eval_rows_wide_kernel(int n, int count, const double *__restrict__ X, int64_t ld,
        const double *__restrict__ aux, double *__restrict__ f) {
    extern __shared__ double terms[]; // t1[nt], Ackley: t2[nt]
    const int row = blockIdx.x;
    const double *x = X + (int64_t) row * ld;
    const int nt = obj_term_count(ID, n);
    double *t1 = terms, *t2 = terms + nt;
    for (int i = threadIdx.x; i < nt; i += blockDim.x) {
        const double v = x[i];
        const double xn = ID == BBGPU_OBJ_ROSENBROCK ? x[i + 1] : 0.0;
        const double a = ID == BBGPU_OBJ_ELLIPSOID ? aux[i] : 0.0;
        double s2 = 0.0;
        t1[i] = obj_term(ID, v, xn, a, s2);
        if (ID == BBGPU_OBJ_ACKLEY) t2[i] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const double v = warp_reduce_terms(ID, n, t1, t2);
        if (threadIdx.x == 0) f[row] = v;
    }
}

template<int ID>
static int launch_wide(bbgpu_ctx *ctx, int n, int count, const double *d_X, int64_t ld, const double *aux, double *d_f, size_t smem) {
    if (smem > 48 * 1024) BB_CUDA(cudaFuncSetAttribute(eval_rows_wide_kernel<ID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    eval_rows_wide_kernel<ID><<<count, 256, smem, ctx->stream>>>(n, count, d_X, ld, aux, d_f);
    return launched(ctx, "eval_rows_wide_kernel");
}

int evaluate(bbgpu_ctx *ctx, bbgpu_objective *obj, const double *d_X, int64_t ld, int count, double *d_f) {
    if (count <= 0) return BBGPU_OK;
    obj->evals += count;
    if (obj->mode == 0) {
        const size_t smem = (size_t) obj->n * 8 * (obj->id == BBGPU_OBJ_ACKLEY ? 2 : 1);
        if (count <= 2 * ctx->sm_count && obj->n >= 2048 && smem <= 200 * 1024) {
            switch (obj->id) {
            case BBGPU_OBJ_SPHERE: return launch_wide<BBGPU_OBJ_SPHERE>(ctx, obj->n, count, d_X, ld, obj->d_aux, d_f, smem);
            case BBGPU_OBJ_ELLIPSOID: return launch_wide<BBGPU_OBJ_ELLIPSOID>(ctx, obj->n, count, d_X, ld, obj->d_aux, d_f, smem);
            case BBGPU_OBJ_ROSENBROCK: return launch_wide<BBGPU_OBJ_ROSENBROCK>(ctx, obj->n, count, d_X, ld, obj->d_aux, d_f, smem);
            case BBGPU_OBJ_RASTRIGIN: return launch_wide<BBGPU_OBJ_RASTRIGIN>(ctx, obj->n, count, d_X, ld, obj->d_aux, d_f, smem);
            case BBGPU_OBJ_ACKLEY: return launch_wide<BBGPU_OBJ_ACKLEY>(ctx, obj->n, count, d_X, ld, obj->d_aux, d_f, smem);
            default: break;
            }
        }
        eval_rows_kernel<<<ceil_div(count, 8), 256, 0, ctx->stream>>>(obj->id, obj->n, count, d_X, ld,
                obj->d_aux, d_f);
        return launched(ctx, "eval_rows_kernel");
    }
    if (obj->mode == 1) {
        const int rc = obj->dev_fn(d_X, ld, obj->n, count, d_f, (void*) ctx->stream, obj->user);
        if (rc != 0) return fail(BBGPU_ERR_OBJECTIVE, "user device objective returned %d", rc);
        return BBGPU_OK;
    }
    // batched host callback bridge: device -> pinned host -> callback -> device
    const int n = obj->n;
    void *hp = nullptr;
    BB_TRY(ctx->pinned(((size_t) count * n + count) * sizeof(double), &hp));
    double *hX = (double*) hp, *hf = hX + (size_t) count * n;
    BB_CUDA(cudaMemcpy2DAsync(hX, (size_t) n * 8, d_X, (size_t) ld * 8, (size_t) n * 8, count,
            cudaMemcpyDeviceToHost, ctx->stream));
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int rc = obj->host_fn(hX, n, n, count, hf, obj->user);
    if (rc != 0) return fail(BBGPU_ERR_OBJECTIVE, "host objective callback returned %d", rc);
    BB_CUDA(cudaMemcpyAsync(d_f, hf, (size_t) count * 8, cudaMemcpyHostToDevice, ctx->stream));
    BB_CUDA(cudaStreamSynchronize(ctx->stream)); // the pinned buffer is reused by later calls
    return BBGPU_OK;
}

// ---- evaluation of the rows a mask names (round engines of SaNSDE / SSDE with a user objective) --------------------------------
// A registry objective is a cheap kernel: all rows are evaluated and the mask is ignored.  A user objective -- a CUDA function or
// a Python callable behind the host bridge, usually the expensive part of the whole optimisation -- is called for the marked rows
// only: their indices are listed, the rows gathered into a dense staging block, evaluated, and the values scattered back.
__global__ void __launch_bounds__(1024) mask_list_kernel(const int *__restrict__ mask, int count, int *__restrict__ list, int *__restrict__ nlist) {
    __shared__ int wsum[32], carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int k0 = 0; k0 < count; k0 += 1024) {
        const int i = k0 + tid;
        const int flag = (i < count && mask[i] != 0) ? 1 : 0;
        int x = flag;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int sacc = wsum[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, sacc, off);
                if (lane >= off) sacc += y;
            }
            wsum[lane] = sacc;
        }
        __syncthreads();
        if (flag) list[carry + (wid ? wsum[wid - 1] : 0) + x - 1] = i;
        __syncthreads();
        if (tid == 0) carry += wsum[31];
        __syncthreads();
    }
    if (tid == 0) *nlist = carry;
}

__global__ void __launch_bounds__(256) rows_gather_kernel(const double *__restrict__ X, int64_t ld, int n, const int *__restrict__ list, int rows,
        double *__restrict__ dst, int64_t ldd) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const double *src = X + (int64_t) list[r] * ld;
    double *t = dst + (int64_t) r * ldd;
    for (int c = threadIdx.x & 31; c < n; c += 32) t[c] = src[c];
}

__global__ void __launch_bounds__(256) values_scatter_kernel(const double *__restrict__ v, const int *__restrict__ list, int rows, double *__restrict__ f) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) f[list[r]] = v[r];
}

// work: device scratch of the caller, at least count ints + 8 ints + count * (round_up(n, 8) + 1) doubles (layout below)
int evaluate_masked(bbgpu_ctx *ctx, bbgpu_objective *obj, const double *d_X, int64_t ld, int count, double *d_f, const int *d_mask,
        void *work) {
    if (count <= 0) return BBGPU_OK;
    if (obj->mode == 0 || d_mask == nullptr || work == nullptr) return evaluate(ctx, obj, d_X, ld, count, d_f);
    cudaStream_t st = ctx->stream;
    const int64_t ldd = round_up(obj->n, 8);
    double *stage = (double*) work, *vals = stage + (size_t) count * ldd;
    int *list = (int*) (vals + count), *nlist = list + count;
    mask_list_kernel<<<1, 1024, 0, st>>>(d_mask, count, list, nlist);
    BB_TRY(launched(ctx, "mask_list_kernel"));
    int rows = 0;
    BB_CUDA(cudaMemcpyAsync(&rows, nlist, sizeof(int), cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    if (rows == 0) return BBGPU_OK;
    rows_gather_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(d_X, ld, obj->n, list, rows, stage, ldd);
    BB_TRY(launched(ctx, "rows_gather_kernel"));
    BB_TRY(evaluate(ctx, obj, stage, ldd, rows, vals));
    values_scatter_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(vals, list, rows, d_f);
    return launched(ctx, "values_scatter_kernel");
}

} // namespace bb

extern "C" {

int bbgpu_objective_builtin(bbgpu_ctx *ctx, int id, int n, bbgpu_objective **out) {
    BB_ARG(ctx && out && n >= 1);
    BB_ARG(id >= BBGPU_OBJ_SPHERE && id <= BBGPU_OBJ_ACKLEY);
    bbgpu_objective *o = new bbgpu_objective();
    o->ctx = ctx;
    o->mode = 0;
    o->id = id;
    o->n = n;
    if (id == BBGPU_OBJ_ELLIPSOID) {
        // coefficient table 10^(6 i/(n-1)), computed once on the host with libm pow -- the same
        // expression the oracle uses, so both sides read identical coefficients
        std::vector<double> coef(n);
        for (int i = 0; i < n; i++) coef[i] = (n > 1) ? std::pow(10.0, 6.0 * (double) i / ((double) n - 1.0)) : 1.0;
        int s = bb::dev_alloc(&o->d_aux, (size_t) bb::round_up(n, 8)); // padded like the rows: 32-byte vector reads
        if (s != BBGPU_OK) { delete o; return s; }
        cudaError_t e = cudaMemcpy(o->d_aux, coef.data(), (size_t) n * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(o->d_aux); delete o; return bb::fail(BBGPU_ERR_CUDA, "memcpy coef: %s", cudaGetErrorString(e)); }
    }
    *out = o;
    return BBGPU_OK;
}

int bbgpu_objective_device(bbgpu_ctx *ctx, bbgpu_device_objective_fn fn, void *user, int n,
        bbgpu_objective **out) {
    BB_ARG(ctx && out && fn && n >= 1);
    bbgpu_objective *o = new bbgpu_objective();
    o->ctx = ctx;
    o->mode = 1;
    o->n = n;
    o->dev_fn = fn;
    o->user = user;
    *out = o;
    return BBGPU_OK;
}

int bbgpu_objective_host(bbgpu_ctx *ctx, bbgpu_host_objective_fn fn, void *user, int n,
        bbgpu_objective **out) {
    BB_ARG(ctx && out && fn && n >= 1);
    bbgpu_objective *o = new bbgpu_objective();
    o->ctx = ctx;
    o->mode = 2;
    o->n = n;
    o->host_fn = fn;
    o->user = user;
    *out = o;
    return BBGPU_OK;
}

int bbgpu_objective_destroy(bbgpu_objective *obj) {
    if (!obj) return BBGPU_OK;
    if (obj->d_aux) cudaFree(obj->d_aux);
    delete obj;
    return BBGPU_OK;
}

int bbgpu_eval(bbgpu_ctx *ctx, bbgpu_objective *obj, const double *d_X, int64_t ld, int count,
        double *d_f) {
    BB_ARG(ctx && obj && d_X && d_f && ld >= obj->n && count >= 0);
    return bb::evaluate(ctx, obj, d_X, ld, count, d_f);
}

int bbgpu_eval_host(bbgpu_ctx *ctx, bbgpu_objective *obj, const double *h_X, int count, double *h_f) {
    BB_ARG(ctx && obj && h_X && h_f && count >= 0);
    if (count == 0) return BBGPU_OK;
    const size_t nx = (size_t) count * obj->n;
    void *dp = nullptr;
    BB_TRY(ctx->scratch((nx + count) * sizeof(double), &dp));
    double *dX = (double*) dp, *df = dX + nx;
    BB_CUDA(cudaMemcpyAsync(dX, h_X, nx * 8, cudaMemcpyHostToDevice, ctx->stream));
    BB_TRY(bb::evaluate(ctx, obj, dX, obj->n, count, df));
    BB_CUDA(cudaMemcpyAsync(h_f, df, (size_t) count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    return BBGPU_OK;
}

} // extern "C"
