// dist.cuh -- z-slab partition of the grid across the GPUs of one node, NCCL over NVLink.
//
// The reference is single-process; this is new plumbing for the B200 build:
//  * contiguous z-slabs (k is the slowest index, src/signed_heat_grid_solver.cpp:507);
//  * one ghost plane per side exchanged with grouped ncclSend/ncclRecv before every stencil read;
//  * scalar / m-vector all-reduces for the CG dot products and the constraint gathers.
// NCCL is bound at run time (dlopen of the libnccl.so.2 that torch ships) so the library loads -- and its
// single-GPU path works -- on boxes without NCCL.
#pragma once
#include "projector.cuh"

namespace shm3d {

inline void slab_range(int rank, int world, int nz, int& k0, int& k1) {
    // even split; if nz is not divisible the first (nz % world) ranks take one extra plane
    int base = nz / world, rem = nz % world;
    k0 = rank * base + (rank < rem ? rank : rem);
    k1 = k0 + base + (rank < rem ? 1 : 0);
}

class Dist {
  public:
    Dist(int rank, int world, const void* nccl_id, cudaStream_t s);
    ~Dist();
    static int unique_id(void* out128);

    int rank() const { return rank_; }
    int world() const { return world_; }

    // fill the ghost planes of an interior pointer v (plane below <- rank-1's last plane, plane above <- rank+1's first)
    void exchange_halo(float* v, const LevelDims& L, cudaStream_t s);
    // three component-major padded vectors at once (stride between components)
    void exchange_halo3(float* v, size_t comp_stride, const LevelDims& L, cudaStream_t s);
    void allreduce(double* dev, int n, cudaStream_t s);  // in-place sum
    void allgather(float* full, size_t count_per_rank, cudaStream_t s);  // in place, rank r owns [r*count, (r+1)*count)
    void send_plane(const float* p, size_t n, int peer, cudaStream_t s);  // point-to-point (chained prefix sums)
    void recv_plane(float* p, size_t n, int peer, cudaStream_t s);
    // the slabs of a float field (rank r owns planes [k0_r, k1_r) of nz) gathered on rank `root`: `full` (root only) gets
    // nz*plane floats, every rank sends its `local` slab -- grouped ncclSend/ncclRecv over NVLink (row N3 on slab contexts)
    void gather_slabs(const float* local, float* full, size_t plane, int nz, int root, cudaStream_t s);
    // one grouped round of point-to-point transfers (Steps 1-2 computed on cyclically assigned z-chunks, then moved to
    // the slab owners); sends and receives to / from one peer must be listed in the same order on both sides
    struct P2P {
        float* ptr;
        size_t count;
        int peer;
    };
    void p2p_round(const std::vector<P2P>& sends, const std::vector<P2P>& recvs, cudaStream_t s);
    unsigned int allreduce_max_host(unsigned int v);
    void attach(Projector& P);  // hook the projector's gather to an all-reduce

  private:
    int rank_, world_;
    void* comm_ = nullptr;
    cudaStream_t stream_;
    unsigned int* d_tmp_ = nullptr;
};

}  // namespace shm3d
