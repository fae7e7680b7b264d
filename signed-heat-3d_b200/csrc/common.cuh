// common.cuh -- shared declarations for the shm3d B200 grid solver (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/shm3d_grid.h"

namespace shm3d {

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SHM3D_CUDA_CHECK(expr)                                                                          \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw ::shm3d::Error(SHM3D_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) +   \
                                                     " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

// launch counter (per process; the C ABI reports the delta per call)
extern int64_t g_kernel_launches;
#define SHM3D_LAUNCHED() (++::shm3d::g_kernel_launches)

// RAII device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n_) {
        if (n_ <= n && p) return;
        release();
        n = n_;
        if (n) SHM3D_CUDA_CHECK(cudaMalloc((void**)&p, n * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void zero(cudaStream_t s) { if (n) SHM3D_CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void upload(const T* h, size_t cnt, cudaStream_t s) {
        alloc(cnt);
        if (cnt) SHM3D_CUDA_CHECK(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
};

// Local (slab) view of the grid.  Node (i,j,k) with k in [k0,k1) is stored at i + j*nx + (k-k0)*nx*ny.
struct GridDesc {
    int nx, ny, nz;    // global node counts
    int k0, k1;        // this rank's z range
    double bmin[3];    // position of global node (0,0,0)
    double cell;
    size_t plane() const { return (size_t)nx * ny; }
    int nzl() const { return k1 - k0; }
    size_t nlocal() const { return plane() * (size_t)nzl(); }
    size_t nglobal() const { return plane() * (size_t)nz; }
};

// ------------------------------------------------------------------------------------------------
// Sources, clustered for the summation kernel (host build: sources.cpp)
// ------------------------------------------------------------------------------------------------
struct ClusteredSources {
    // all positions are relative to `origin` (the bbox centre) so fp32 keeps full precision near the surface
    double origin[3];
    std::vector<float4> pos;     // xyz = position - origin, w = unused
    std::vector<float4> wn;      // xyz = unit normal * area * wscale
    std::vector<float4> bounds;  // per cluster: centre xyz (relative), w = radius
    std::vector<int2> range;     // per cluster: first source, count (<= 32)
    double wscale;
};

// Morton-sorted clusters of <= 32 sources with lambda * radius <= rho_max.
// Throws Error(SHM3D_ERR_NONFINITE) on non-finite positions / weights.
void build_clusters(int64_t M, const double* pos, const double* nrm, const double* area, const double origin[3],
                    double lambda, double rho_max, ClusteredSources& out);

}  // namespace shm3d
