// isosurface.cuh -- row N3: the solver's float32 field stays on the device and the consumer's work happens there:
// marching cubes identical to polyscope's registerIsosurfaceAsMesh, and plane slices through the field.
#pragma once
#include "common.cuh"

namespace shm3d {

struct IsoResult {
    int64_t n_vertices = 0, n_triangles = 0;
    double ms_device = 0;   // CUDA-event time of the kernels (count, scan, vertices, triangles)
    int64_t launches = 0;
};

class IsoSurface {
public:
    IsoSurface();
    ~IsoSurface();
    IsoSurface(const IsoSurface&) = delete;
    IsoSurface& operator=(const IsoSurface&) = delete;

    // Device-resident float field of a full (unpartitioned) grid, index i + j*nx + k*nx*ny.  Host fields are staged
    // (and doubles narrowed to float32 the way the consumer stores them) into a buffer owned by this object.
    const float* stage_field(cudaStream_t s, size_t n, const void* field, int kind);
    // the staging buffer itself, sized for n floats (slab contexts gather the ranks' slabs into it)
    float* field_buffer(size_t n) {
        field32_.alloc(n);
        return field32_.p;
    }

    // bound_min / bound_max: the float bounds the volume grid was registered with; both NULL = lattice coordinates.
    IsoResult extract(cudaStream_t s, int nx, int ny, int nz, const float* d_field, float isoval, const float* bound_min,
                      const float* bound_max);
    // result of the last extract(): float[3*nV], uint32[3*nT], numbered and ordered like MC::marching_cube's output
    const float* d_vertices() const { return verts_.p; }
    const uint32_t* d_triangles() const { return tris_.p; }
    void fetch(cudaStream_t s, float* vertices_out, uint32_t* triangles_out) const;
    void clear_result() { last_ = IsoResult(); }  // (slab contexts: ranks other than the gathering one hold no mesh)

    // out[a + b*nu] = trilinear interpolant (src/signed_heat_grid_solver.cpp:405-431) at origin + a*du + b*dv; NaN outside
    int64_t slice(cudaStream_t s, int nx, int ny, int nz, const float* d_field, const double bbox_min[3], double cell,
                  const double origin[3], const double du[3], const double dv[3], int nu, int nv, float* out_host);

private:
    DevBuf<float> field32_;
    DevBuf<double> stage64_;
    DevBuf<unsigned int> col_v_, col_t_, col_x_;
    DevBuf<unsigned long long> voff_, toff_, block_sums_;
    DevBuf<float> verts_;
    DevBuf<uint32_t> vkey_, tris_;
    DevBuf<float> slice_;
    unsigned long long* h_totals_ = nullptr;  // pinned
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    IsoResult last_;
};

}  // namespace shm3d
