// dist.cu -- NCCL plumbing for the slab-partitioned solve (see dist.cuh).
#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "dist.cuh"

namespace shm3d {

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi& api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [&] {
        // SHM3D_NCCL_LIB (a path) wins; else the libnccl already mapped into the process (torch's, when the caller is a
        // torch.distributed program) or whatever the loader's search path offers -- no environment-specific paths here
        const char* env = getenv("SHM3D_NCCL_LIB");
        if (env && *env) a.h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        for (const char* n : {"libnccl.so.2", "libnccl.so"}) {
            if (!a.h) a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
        }
        for (const char* n : {"libnccl.so.2", "libnccl.so"}) {
            if (!a.h) a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!a.h) {
            a.err = std::string("cannot load libnccl.so.2 (set SHM3D_NCCL_LIB to its path): ") + dlerror();
            return;
        }
#define BIND(field, sym)                                        \
    a.field = (decltype(a.field))dlsym(a.h, sym);               \
    if (!a.field) a.err = std::string("missing NCCL symbol ") + sym;
        BIND(GetUniqueId, "ncclGetUniqueId")
        BIND(CommInitRank, "ncclCommInitRank")
        BIND(CommDestroy, "ncclCommDestroy")
        BIND(AllReduce, "ncclAllReduce")
        BIND(AllGather, "ncclAllGather")
        BIND(Send, "ncclSend")
        BIND(Recv, "ncclRecv")
        BIND(GroupStart, "ncclGroupStart")
        BIND(GroupEnd, "ncclGroupEnd")
        BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
    });
    return a;
}

#define NCCL_CHECK(expr)                                                                                  \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess)                                                                            \
            throw Error(SHM3D_ERR_NCCL, std::string(#expr) + ": " + api().GetErrorString(_r));            \
    } while (0)

}  // namespace

int Dist::unique_id(void* out128) {
    if (!out128) return SHM3D_ERR_INVALID_ARG;
    NcclApi& a = api();
    if (!a.err.empty() || !a.h) return SHM3D_ERR_NCCL;
    ncclUniqueId id;
    if (a.GetUniqueId(&id) != ncclSuccess) return SHM3D_ERR_NCCL;
    memcpy(out128, &id, sizeof(id));
    return SHM3D_OK;
}

Dist::Dist(int rank, int world, const void* nccl_id, cudaStream_t s) : rank_(rank), world_(world), stream_(s) {
    NcclApi& a = api();
    if (!a.h || !a.err.empty()) throw Error(SHM3D_ERR_NCCL, a.err.empty() ? "NCCL unavailable" : a.err);
    ncclUniqueId id;
    memcpy(&id, nccl_id, sizeof(id));
    ncclComm_t comm;
    NCCL_CHECK(a.CommInitRank(&comm, world, id, rank));
    comm_ = comm;
    SHM3D_CUDA_CHECK(cudaMalloc((void**)&d_tmp_, sizeof(unsigned int)));
}

Dist::~Dist() {
    if (comm_) api().CommDestroy((ncclComm_t)comm_);
    cudaFree(d_tmp_);
}

void Dist::exchange_halo(float* v, const LevelDims& L, cudaStream_t s) {
    NcclApi& a = api();
    ncclComm_t comm = (ncclComm_t)comm_;
    const size_t pl = L.plane();
    const size_t n = L.n();
    // slab neighbours are rank+-1 whenever that rank owns planes of this level
    NCCL_CHECK(a.GroupStart());
    if (L.k0 > 0) {  // lower neighbour exists
        NCCL_CHECK(a.Send(v, pl, ncclFloat32, rank_ - 1, comm, s));
        NCCL_CHECK(a.Recv(v - pl, pl, ncclFloat32, rank_ - 1, comm, s));
    }
    if (L.k1 < L.nz) {
        NCCL_CHECK(a.Send(v + n - pl, pl, ncclFloat32, rank_ + 1, comm, s));
        NCCL_CHECK(a.Recv(v + n, pl, ncclFloat32, rank_ + 1, comm, s));
    }
    NCCL_CHECK(a.GroupEnd());
}

void Dist::exchange_halo3(float* v, size_t cs, const LevelDims& L, cudaStream_t s) {
    NcclApi& a = api();
    ncclComm_t comm = (ncclComm_t)comm_;
    const size_t pl = L.plane();
    const size_t n = L.n();
    NCCL_CHECK(a.GroupStart());
    for (int c = 0; c < 3; c++) {
        float* p = v + (size_t)c * cs;
        if (L.k0 > 0) {
            NCCL_CHECK(a.Send(p, pl, ncclFloat32, rank_ - 1, comm, s));
            NCCL_CHECK(a.Recv(p - pl, pl, ncclFloat32, rank_ - 1, comm, s));
        }
        if (L.k1 < L.nz) {
            NCCL_CHECK(a.Send(p + n - pl, pl, ncclFloat32, rank_ + 1, comm, s));
            NCCL_CHECK(a.Recv(p + n, pl, ncclFloat32, rank_ + 1, comm, s));
        }
    }
    NCCL_CHECK(a.GroupEnd());
}

void Dist::allreduce(double* dev, int n, cudaStream_t s) {
    NCCL_CHECK(api().AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)comm_, s));
}

void Dist::allgather(float* full, size_t count_per_rank, cudaStream_t s) {
    // in place: rank r's contribution already sits at full + r * count_per_rank
    NCCL_CHECK(api().AllGather(full + (size_t)rank_ * count_per_rank, full, count_per_rank, ncclFloat32, (ncclComm_t)comm_, s));
}

void Dist::send_plane(const float* p, size_t n, int peer, cudaStream_t s) {
    NCCL_CHECK(api().Send(p, n, ncclFloat32, peer, (ncclComm_t)comm_, s));
}
void Dist::recv_plane(float* p, size_t n, int peer, cudaStream_t s) {
    NCCL_CHECK(api().Recv(p, n, ncclFloat32, peer, (ncclComm_t)comm_, s));
}

void Dist::gather_slabs(const float* local, float* full, size_t plane, int nz, int root, cudaStream_t s) {
    NcclApi& a = api();
    ncclComm_t comm = (ncclComm_t)comm_;
    int k0, k1;
    slab_range(rank_, world_, nz, k0, k1);
    if (rank_ == root)
        SHM3D_CUDA_CHECK(cudaMemcpyAsync(full + (size_t)k0 * plane, local, (size_t)(k1 - k0) * plane * sizeof(float),
                                         cudaMemcpyDeviceToDevice, s));
    NCCL_CHECK(a.GroupStart());
    if (rank_ == root) {
        for (int r = 0; r < world_; r++) {
            if (r == root) continue;
            int r0, r1;
            slab_range(r, world_, nz, r0, r1);
            if (r1 > r0) NCCL_CHECK(a.Recv(full + (size_t)r0 * plane, (size_t)(r1 - r0) * plane, ncclFloat32, r, comm, s));
        }
    } else if (k1 > k0) {
        NCCL_CHECK(a.Send(local, (size_t)(k1 - k0) * plane, ncclFloat32, root, comm, s));
    }
    NCCL_CHECK(a.GroupEnd());
}

void Dist::p2p_round(const std::vector<P2P>& sends, const std::vector<P2P>& recvs, cudaStream_t s) {
    NcclApi& a = api();
    ncclComm_t comm = (ncclComm_t)comm_;
    NCCL_CHECK(a.GroupStart());
    for (const P2P& t : sends) NCCL_CHECK(a.Send(t.ptr, t.count, ncclFloat32, t.peer, comm, s));
    for (const P2P& t : recvs) NCCL_CHECK(a.Recv(t.ptr, t.count, ncclFloat32, t.peer, comm, s));
    NCCL_CHECK(a.GroupEnd());
}

unsigned int Dist::allreduce_max_host(unsigned int v) {
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(d_tmp_, &v, sizeof(v), cudaMemcpyHostToDevice, stream_));
    NCCL_CHECK(api().AllReduce(d_tmp_, d_tmp_, 1, ncclUint32, ncclMax, (ncclComm_t)comm_, stream_));
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(&v, d_tmp_, sizeof(v), cudaMemcpyDeviceToHost, stream_));
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(stream_));
    return v;
}

void Dist::attach(Projector& P) {
    P.reduce_hook_ = [this](double* rhs, int m, cudaStream_t s) { this->allreduce(rhs, m, s); };
}

}  // namespace shm3d
