// Communicator: one rank per GPU, NCCL over NVLink / NVSwitch for the two exchange steps of a
// candidate-sharded CMA-ES run (fitness all-gather, rank-mu partial all-reduce) and for the
// best-of reduce over independent runs.
//
// NCCL is bound at run time (dlopen of libnccl.so.2), not at link time: a process that already
// loaded a libnccl (e.g. torch.distributed) shares that copy, and libbbgpu.so still loads on a
// box without NCCL -- creating a communicator then fails loudly.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

struct bbgpu_comm {
    bbgpu_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double *d_pair = nullptr; // [2*world] scratch for best-of
};

namespace bb {
// a buffer of every rank mapped into every rank (cudaIpc), with the flags of the peer-memory all-reduce
struct PeerRegion {
    bbgpu_comm *c = nullptr;
    double *buf[16] = { nullptr };
    unsigned long long *flags[16] = { nullptr };
    unsigned long long *my_flags = nullptr;
    unsigned *counter = nullptr;
    unsigned long long epoch = 0;
};
}

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi &nccl() {
    static NcclApi api = [] {
        NcclApi a;
        const char *names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char *nm : names) {
            a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) {
            a.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "unknown");
            return a;
        }
#define BB_SYM(field, name) \
        a.field = (decltype(a.field)) dlsym(a.lib, name); \
        if (!a.field && a.why.empty()) a.why = std::string("libnccl lacks ") + name;
        BB_SYM(GetUniqueId, "ncclGetUniqueId")
        BB_SYM(CommInitRank, "ncclCommInitRank")
        BB_SYM(CommDestroy, "ncclCommDestroy")
        BB_SYM(AllReduce, "ncclAllReduce")
        BB_SYM(AllGather, "ncclAllGather")
        BB_SYM(Broadcast, "ncclBroadcast")
        BB_SYM(GroupStart, "ncclGroupStart")
        BB_SYM(GroupEnd, "ncclGroupEnd")
        BB_SYM(GetErrorString, "ncclGetErrorString")
#undef BB_SYM
        return a;
    }();
    return api;
}

int nccl_ready() {
    NcclApi &a = nccl();
    if (!a.why.empty()) return bb::fail(BBGPU_ERR_NODEVICE, "NCCL unavailable: %s", a.why.c_str());
    return BBGPU_OK;
}

#define BB_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            return bb::fail(BBGPU_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,       \
                            nccl().GetErrorString(r__));                                       \
    } while (0)

// ---- peer memory (NVLink / NVSwitch): an all-reduce written as ONE kernel over IPC-mapped buffers -----------------
// Two-shot, in place: rank g owns slice g of the buffer.  arrive barrier (every rank's partial is complete) -> rank g
// reads slice g from every peer, adds in rank order (the same bits on every rank, run to run), and stores the sum into
// slice g of EVERY peer -> done barrier.  Nobody but its owner ever reads a slice, so the sums can land in place.
// Flags live in each rank's own memory and are written remotely (epoch numbers, never reset).
constexpr int P2P_MAX_WORLD = 16;

struct P2PArgs {
    double *buf[P2P_MAX_WORLD];                 // the registered buffer of every rank (mapped)
    unsigned long long *flags[P2P_MAX_WORLD];   // [2 * world]: arrive[src], done[src]
    unsigned *counter;                          // local: CTAs finished with their part of the slice
    int world, rank;
    unsigned long long epoch;
    size_t count;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const P2PArgs a) {
    const int G = a.world, me = a.rank, tid = threadIdx.x;
    __shared__ int last;
    unsigned long long *mine = a.flags[me];
    if (blockIdx.x == 0 && tid < G) st_release_sys(a.flags[tid] + me, a.epoch); // this rank's partial is complete (stream order)
    if (tid < G) while (ld_acquire_sys(mine + tid) < a.epoch) { }
    __syncthreads();
    const size_t per = ((a.count + G - 1) / G + 1) & ~(size_t) 1; // even: 16-byte accesses
    const size_t lo = (size_t) me * per, hi = min(a.count, lo + per);
    for (size_t i = lo + 2 * ((size_t) blockIdx.x * blockDim.x + tid); i < hi; i += 2 * (size_t) gridDim.x * blockDim.x) {
        if (i + 1 < hi) {
            double2 s = make_double2(0.0, 0.0);
            for (int r = 0; r < G; r++) {
                const double2 v = __ldcv(reinterpret_cast<const double2*>(a.buf[r] + i)); // peers' data: never from a stale L1 line
                s.x += v.x;
                s.y += v.y;
            }
            for (int r = 0; r < G; r++) *reinterpret_cast<double2*>(a.buf[r] + i) = s;
        } else {
            double s = 0.0;
            for (int r = 0; r < G; r++) s += __ldcv(a.buf[r] + i);
            for (int r = 0; r < G; r++) a.buf[r][i] = s;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(a.counter, 1u);
        last = (t == gridDim.x - 1) ? 1 : 0;
        if (last) *a.counter = 0;
    }
    __syncthreads();
    if (last && tid < G) st_release_sys(a.flags[tid] + G + me, a.epoch); // my slice has landed everywhere
    if (blockIdx.x == 0 && tid < G) while (ld_acquire_sys(mine + G + tid) < a.epoch) { }
}

__global__ void best_pair_kernel(const double *f, double *pair, int rank) {
    pair[0] = *f;
    pair[1] = (double) rank;
}

} // namespace

namespace bb {

// used by the sharded CMA path (cma.cu); all on the ctx stream, in place
int comm_allreduce_sum(bbgpu_comm *c, double *d_buf, size_t count) {
    BB_NCCL(nccl().AllReduce(d_buf, d_buf, count, ncclFloat64, ncclSum, c->comm, c->ctx->stream));
    return BBGPU_OK;
}

// every rank contributes d_buf[rank*per .. rank*per+per)
int comm_allgather(bbgpu_comm *c, double *d_buf, size_t per) {
    BB_NCCL(nccl().AllGather(d_buf + (size_t) c->rank * per, d_buf, per, ncclFloat64, c->comm, c->ctx->stream));
    return BBGPU_OK;
}

int comm_allgather2(bbgpu_comm *c, double *d_a, double *d_b, size_t per) {
    BB_NCCL(nccl().GroupStart());
    BB_NCCL(nccl().AllGather(d_a + (size_t) c->rank * per, d_a, per, ncclFloat64, c->comm, c->ctx->stream));
    BB_NCCL(nccl().AllGather(d_b + (size_t) c->rank * per, d_b, per, ncclFloat64, c->comm, c->ctx->stream));
    BB_NCCL(nccl().GroupEnd());
    return BBGPU_OK;
}

int comm_broadcast(bbgpu_comm *c, double *d_buf, size_t count, int root) {
    BB_NCCL(nccl().Broadcast(d_buf, d_buf, count, ncclFloat64, root, c->comm, c->ctx->stream));
    return BBGPU_OK;
}

// Collective: map `d_local` (a whole cudaMalloc allocation of `bytes`) of every rank into every rank.  Returns
// BBGPU_OK with *out == nullptr when peer mapping is not possible on this box (every rank agrees): the caller keeps NCCL.
int comm_peer_region_create(bbgpu_comm *c, double *d_local, size_t bytes, PeerRegion **out) {
    *out = nullptr;
    (void) bytes;
    if (c->world > P2P_MAX_WORLD) return BBGPU_OK;
    cudaStream_t st = c->ctx->stream;
    PeerRegion *r = new PeerRegion();
    r->c = c;
    const int G = c->world, me = c->rank;
    int ok = 1;
    struct Pair { cudaIpcMemHandle_t buf, flags; };
    static_assert(sizeof(Pair) == 128, "two IPC handles");
    std::vector<Pair> all(G);
    char *d_x = nullptr;
    if (cudaMalloc((void**) &r->my_flags, sizeof(unsigned long long) * 2 * G + 64) != cudaSuccess) ok = 0;
    if (ok) {
        cudaMemset(r->my_flags, 0, sizeof(unsigned long long) * 2 * G + 64);
        r->counter = reinterpret_cast<unsigned*>(r->my_flags + 2 * G);
        if (cudaIpcGetMemHandle(&all[me].buf, d_local) != cudaSuccess) ok = 0;
        if (ok && cudaIpcGetMemHandle(&all[me].flags, r->my_flags) != cudaSuccess) ok = 0;
    }
    cudaGetLastError();
    if (cudaMalloc((void**) &d_x, sizeof(Pair) * G + 8) != cudaSuccess) { delete r; return fail(BBGPU_ERR_CUDA, "cudaMalloc failed"); }
    cudaMemcpyAsync(d_x + sizeof(Pair) * me, &all[me], sizeof(Pair), cudaMemcpyHostToDevice, st);
    ncclResult_t nr = nccl().AllGather(d_x + sizeof(Pair) * me, d_x, sizeof(Pair), ncclChar, c->comm, st);
    if (nr != ncclSuccess) { cudaFree(d_x); delete r; return fail(BBGPU_ERR_CUDA, "ncclAllGather(ipc handles) -> %s", nccl().GetErrorString(nr)); }
    cudaMemcpyAsync(all.data(), d_x, sizeof(Pair) * G, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    for (int q = 0; q < G && ok; q++) {
        if (q == me) {
            r->buf[q] = d_local;
            r->flags[q] = r->my_flags;
            continue;
        }
        void *pb = nullptr, *pf = nullptr;
        if (cudaIpcOpenMemHandle(&pb, all[q].buf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
        else if (cudaIpcOpenMemHandle(&pf, all[q].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
        r->buf[q] = (double*) pb;
        r->flags[q] = (unsigned long long*) pf;
    }
    cudaGetLastError();
    // agree: one rank without peer access means nobody uses the region
    double *d_ok = reinterpret_cast<double*>(d_x);
    const double mine = ok ? 0.0 : 1.0;
    cudaMemcpyAsync(d_ok, &mine, sizeof(double), cudaMemcpyHostToDevice, st);
    nr = nccl().AllReduce(d_ok, d_ok, 1, ncclFloat64, ncclSum, c->comm, st);
    double bad = 1.0;
    cudaMemcpyAsync(&bad, d_ok, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(d_x);
    if (nr != ncclSuccess || bad != 0.0) {
        for (int q = 0; q < G; q++) {
            if (q == me) continue;
            if (r->buf[q]) cudaIpcCloseMemHandle(r->buf[q]);
            if (r->flags[q]) cudaIpcCloseMemHandle(r->flags[q]);
        }
        if (r->my_flags) cudaFree(r->my_flags);
        delete r;
        cudaGetLastError();
        return BBGPU_OK;
    }
    *out = r;
    return BBGPU_OK;
}

int comm_peer_region_destroy(PeerRegion *r) {
    if (!r) return BBGPU_OK;
    cudaStreamSynchronize(r->c->ctx->stream);
    for (int q = 0; q < r->c->world; q++) {
        if (q == r->c->rank) continue;
        if (r->buf[q]) cudaIpcCloseMemHandle(r->buf[q]);
        if (r->flags[q]) cudaIpcCloseMemHandle(r->flags[q]);
    }
    // nobody may free an exported buffer while a peer still maps it: meet once (NCCL) after everybody closed
    nccl().AllReduce(r->c->d_pair, r->c->d_pair, 1, ncclFloat64, ncclSum, r->c->comm, r->c->ctx->stream);
    cudaStreamSynchronize(r->c->ctx->stream);
    if (r->my_flags) cudaFree(r->my_flags);
    delete r;
    return BBGPU_OK;
}

// in-place sum over the ranks of the registered buffers' first `count` doubles, one kernel over peer memory
int comm_allreduce_sum_p2p(PeerRegion *r, size_t count) {
    bbgpu_ctx *ctx = r->c->ctx;
    P2PArgs a;
    for (int q = 0; q < P2P_MAX_WORLD; q++) { a.buf[q] = r->buf[q]; a.flags[q] = r->flags[q]; }
    a.counter = r->counter;
    a.world = r->c->world;
    a.rank = r->c->rank;
    a.epoch = ++r->epoch;
    a.count = count;
    p2p_allreduce_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(a);
    return launched(ctx, "p2p_allreduce_kernel");
}

int comm_rank(const bbgpu_comm *c) { return c->rank; }
int comm_world(const bbgpu_comm *c) { return c->world; }
bbgpu_ctx *comm_ctx(const bbgpu_comm *c) { return c->ctx; }

} // namespace bb

extern "C" {

int bbgpu_comm_unique_id(void *out128) {
    BB_ARG(out128);
    BB_TRY(nccl_ready());
    static_assert(sizeof(ncclUniqueId) == BBGPU_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    BB_NCCL(nccl().GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return BBGPU_OK;
}

int bbgpu_comm_create(bbgpu_ctx *ctx, int world, int rank, const void *id128, bbgpu_comm **out) {
    BB_ARG(ctx && out && id128 && world >= 1 && rank >= 0 && rank < world);
    *out = nullptr;
    BB_TRY(nccl_ready());
    BB_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    bbgpu_comm *c = new bbgpu_comm();
    c->ctx = ctx; c->rank = rank; c->world = world;
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return bb::fail(BBGPU_ERR_CUDA, "ncclCommInitRank -> %s", nccl().GetErrorString(r));
    }
    if (cudaMalloc((void**) &c->d_pair, (size_t) 2 * world * sizeof(double)) != cudaSuccess) {
        nccl().CommDestroy(c->comm);
        delete c;
        return bb::fail(BBGPU_ERR_CUDA, "cudaMalloc failed");
    }
    *out = c;
    return BBGPU_OK;
}

int bbgpu_comm_destroy(bbgpu_comm *c) {
    if (!c) return BBGPU_OK;
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->d_pair) cudaFree(c->d_pair);
    delete c;
    return BBGPU_OK;
}

int bbgpu_comm_rank(bbgpu_comm *c) { return c ? c->rank : 0; }
int bbgpu_comm_world(bbgpu_comm *c) { return c ? c->world : 1; }

int bbgpu_comm_allreduce_sum(bbgpu_comm *c, double *d_buf, int64_t count) {
    BB_ARG(c && d_buf && count >= 0);
    return bb::comm_allreduce_sum(c, d_buf, (size_t) count);
}

// Best-of over independent runs: every rank passes the device scalar holding its best fitness; returns
// (on every rank) the winning fitness and its rank, ties to the lowest rank, and broadcasts the winner's
// n-vector d_x in place.
int bbgpu_comm_best_of(bbgpu_comm *c, const double *d_f, double *d_x, int n, double *h_fbest, int *h_owner) {
    BB_ARG(c && d_f && h_fbest && h_owner);
    cudaStream_t st = c->ctx->stream;
    best_pair_kernel<<<1, 1, 0, st>>>(d_f, c->d_pair + 2 * c->rank, c->rank);
    BB_TRY(bb::launched(c->ctx, "best_pair_kernel"));
    BB_NCCL(nccl().AllGather(c->d_pair + 2 * c->rank, c->d_pair, 2, ncclFloat64, c->comm, st));
    std::vector<double> pairs((size_t) 2 * c->world);
    BB_CUDA(cudaMemcpyAsync(pairs.data(), c->d_pair, pairs.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    int owner = 0;
    for (int r = 1; r < c->world; r++)
        if (pairs[2 * r] < pairs[2 * owner]) owner = r;
    *h_fbest = pairs[2 * owner];
    *h_owner = owner;
    if (d_x && n > 0) BB_TRY(bb::comm_broadcast(c, d_x, (size_t) n, owner));
    return BBGPU_OK;
}

// host-buffer form: f and x live on the host (the batched drivers keep their per-run bests there)
int bbgpu_comm_best_of_host(bbgpu_comm *c, double f, double *h_x, int n, double *h_fbest, int *h_owner) {
    BB_ARG(c && h_x && n > 0 && h_fbest && h_owner);
    cudaStream_t st = c->ctx->stream;
    void *sp = nullptr;
    BB_TRY(c->ctx->scratch(sizeof(double) * ((size_t) n + 1), &sp));
    double *d = (double*) sp;
    BB_CUDA(cudaMemcpyAsync(d, &f, sizeof(double), cudaMemcpyHostToDevice, st));
    BB_CUDA(cudaMemcpyAsync(d + 1, h_x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    BB_TRY(bbgpu_comm_best_of(c, d, d + 1, n, h_fbest, h_owner));
    BB_CUDA(cudaMemcpyAsync(h_x, d + 1, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    return BBGPU_OK;
}

} // extern "C"
