// Context lifetime: device selection, stream, scratch buffers.  No CPU fallback: creation
// fails loudly when no CUDA device is usable.
#include "common.cuh"

int bbgpu_ctx::pinned(size_t bytes, void **out) {
    if (bytes > h_pinned_bytes) {
        if (h_pinned) {
            cudaStreamSynchronize(stream); // copies from / to the old buffer may still be in flight
            if (aux) cudaStreamSynchronize(aux);
            cudaFreeHost(h_pinned);
        }
        h_pinned = nullptr;
        h_pinned_bytes = 0;
        const size_t want = bb::round_up((int64_t) bytes, 1 << 20);
        cudaError_t e = cudaMallocHost(&h_pinned, want);
        if (e != cudaSuccess) return bb::fail(BBGPU_ERR_CUDA, "cudaMallocHost(%zu) -> %s", want, cudaGetErrorString(e));
        h_pinned_bytes = want;
    }
    *out = h_pinned;
    return BBGPU_OK;
}

int bbgpu_ctx::scratch(size_t bytes, void **out) {
    if (bytes > d_scratch_bytes) {
        if (d_scratch) {
            cudaStreamSynchronize(stream);
            cudaFree(d_scratch);
        }
        d_scratch = nullptr;
        d_scratch_bytes = 0;
        const size_t want = bb::round_up((int64_t) bytes, 1 << 20);
        cudaError_t e = cudaMalloc(&d_scratch, want);
        if (e != cudaSuccess) return bb::fail(BBGPU_ERR_CUDA, "cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
        d_scratch_bytes = want;
    }
    *out = d_scratch;
    return BBGPU_OK;
}

namespace bb { void destroy_solver(bbgpu_ctx *ctx); }

extern "C" {

int bbgpu_version(void) { return BBGPU_VERSION; }

const char *bbgpu_last_error(void) { return bb::last_error().c_str(); }

int bbgpu_ctx_create(int device, void *cuda_stream, bbgpu_ctx **out) {
    BB_ARG(out != nullptr);
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        return bb::fail(BBGPU_ERR_NODEVICE, "no CUDA device available (%s); libbbgpu has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    BB_ARG(device >= 0 && device < count);
    BB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    BB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        return bb::fail(BBGPU_ERR_NODEVICE, "device %d is sm_%d%d; libbbgpu is built for sm_100a only", device,
                prop.major, prop.minor);
    }
    bbgpu_ctx *ctx = new bbgpu_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cuda_stream) {
        ctx->stream = (cudaStream_t) cuda_stream;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return bb::fail(BBGPU_ERR_CUDA, "cudaStreamCreate -> %s", cudaGetErrorString(e));
        }
        ctx->own_stream = true;
    }
    *out = ctx;
    return BBGPU_OK;
}

int bbgpu_ctx_destroy(bbgpu_ctx *ctx) {
    if (!ctx) return BBGPU_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    bb::destroy_solver(ctx);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_flush) cudaFree(ctx->d_flush);
    if (ctx->d_ranks) cudaFree(ctx->d_ranks);
    if (ctx->dc_tables) cudaFree(ctx->dc_tables);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    for (auto &r : ctx->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    if (ctx->timer0) { cudaEventDestroy(ctx->timer0); cudaEventDestroy(ctx->timer1); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return BBGPU_OK;
}

int bbgpu_ctx_sync(bbgpu_ctx *ctx) {
    BB_ARG(ctx);
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    return BBGPU_OK;
}

static void prof_collect(bbgpu_ctx *ctx) {
    cudaStreamSynchronize(ctx->stream);
    for (auto &r : ctx->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && r.tag >= 0 && r.tag < 16) {
            ctx->prof_ms[r.tag] += ms;
            ctx->prof_n[r.tag] += 1;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    ctx->prof.clear();
}

int bbgpu_ctx_profile_enable(bbgpu_ctx *ctx, int on) {
    BB_ARG(ctx);
    if (!on) prof_collect(ctx);
    ctx->profiling = on != 0;
    return BBGPU_OK;
}

int bbgpu_ctx_profile_read(bbgpu_ctx *ctx, int tag, double *total_ms, int64_t *count) {
    BB_ARG(ctx && tag >= 0 && tag < BBGPU_PROF_TAGS && total_ms && count);
    prof_collect(ctx);
    *total_ms = ctx->prof_ms[tag];
    *count = ctx->prof_n[tag];
    return BBGPU_OK;
}

int bbgpu_ctx_profile_reset(bbgpu_ctx *ctx) {
    BB_ARG(ctx);
    prof_collect(ctx);
    for (int i = 0; i < 16; i++) { ctx->prof_ms[i] = 0.0; ctx->prof_n[i] = 0; }
    return BBGPU_OK;
}

int bbgpu_ctx_timer_start(bbgpu_ctx *ctx) {
    BB_ARG(ctx);
    if (!ctx->timer0) {
        BB_CUDA(cudaEventCreate(&ctx->timer0));
        BB_CUDA(cudaEventCreate(&ctx->timer1));
    }
    BB_CUDA(cudaEventRecord(ctx->timer0, ctx->stream));
    return BBGPU_OK;
}

int bbgpu_ctx_timer_stop(bbgpu_ctx *ctx, float *ms) {
    BB_ARG(ctx && ms && ctx->timer0);
    BB_CUDA(cudaEventRecord(ctx->timer1, ctx->stream));
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    BB_CUDA(cudaEventElapsedTime(ms, ctx->timer0, ctx->timer1));
    return BBGPU_OK;
}

int bbgpu_ctx_flush_l2(bbgpu_ctx *ctx) {
    BB_ARG(ctx);
    const size_t bytes = (size_t) 256 << 20;
    if (!ctx->d_flush) BB_CUDA(cudaMalloc(&ctx->d_flush, bytes));
    BB_CUDA(cudaMemsetAsync(ctx->d_flush, 0, bytes, ctx->stream));
    return BBGPU_OK;
}

int bbgpu_ctx_device(bbgpu_ctx *ctx) { return ctx ? ctx->device : -1; }

void *bbgpu_ctx_stream(bbgpu_ctx *ctx) { return ctx ? (void*) ctx->stream : nullptr; }

int64_t bbgpu_ctx_launch_count(bbgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

} // extern "C"
