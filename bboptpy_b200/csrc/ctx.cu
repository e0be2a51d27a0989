// Context lifetime: device selection, stream, scratch buffers.  No CPU fallback: creation
// fails loudly when no CUDA device is usable.
#include "common.cuh"

int bbgpu_ctx::pinned(size_t bytes, void **out) {
    if (bytes > h_pinned_bytes) {
        if (h_pinned) cudaFreeHost(h_pinned);
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
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return BBGPU_OK;
}

int bbgpu_ctx_sync(bbgpu_ctx *ctx) {
    BB_ARG(ctx);
    BB_CUDA(cudaStreamSynchronize(ctx->stream));
    return BBGPU_OK;
}

void *bbgpu_ctx_stream(bbgpu_ctx *ctx) { return ctx ? (void*) ctx->stream : nullptr; }

int64_t bbgpu_ctx_launch_count(bbgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

} // extern "C"
