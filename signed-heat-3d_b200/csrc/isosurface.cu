// isosurface.cu -- row N3: the downstream consumer of phi on the device.
//
// Reference flow being replaced: computeDistance returns N doubles to the host (8.6 GB at 1024^3), polyscope narrows
// them to float32 (deps/polyscope/include/polyscope/volume_grid.ipp:103-106) and registerIsosurfaceAsMesh
// (deps/polyscope/src/volume_grid_scalar_quantity.cpp:209-228; called from src/main.cpp:116-128) runs the sequential
// MC::marching_cube (deps/polyscope/deps/MarchingCubeCpp/include/MarchingCube/MC.h:242-315) over them.
// Here shm3d_solve_device leaves phi in HBM as float32 and four launches produce the same indexed mesh -- identical
// vertex coordinates, vertex numbering and triangle order (logic and its derivation: isosurface_core.h):
//
//   k_mc_count      one thread per lattice column (fixed j,i; marching along k), lanes along i: two coalesced rows per
//                   warp and plane, the z+1 corners as sign bits by shuffle, eight planes of loads in flight; per column:
//                   vertices created, triangles emitted, first / last non-trivial cell                 [reads phi once]
//   k_mc_scan_sums / k_mc_scan_write   exclusive scan of the (nx-1)(ny-1) column counts (per-CTA partial sums, then offsets)
//   k_mc_vertices   columns that create vertices march again (first..last non-trivial cell only) and write positions +
//                   per-column search keys
//   k_mc_triangles  columns that emit triangles march again and resolve each corner to a vertex id by locating the
//                   creating cell's column and bisecting its short key list
//
// All four are HBM-bound integer / compare work: 4 B per node for the count pass (algorithmic), and the two emit passes
// only touch columns the surface crosses.  No tensor cores, no atomics, deterministic output.
#include "isosurface.cuh"

#include <math.h>

#include "isosurface_core.h"

namespace shm3d {

using mc::Lattice;

__device__ const unsigned long long d_mc_table[256] = {
#include "mc_table.inc"
};

using mc::kLanesZ;
using mc::kRowsY;
using mc::kScanBlock;
using mc::kScanBlocks;

__device__ __forceinline__ void load_table(unsigned long long* s_tab) {
    for (int i = threadIdx.y * kLanesZ + threadIdx.x; i < 256; i += kLanesZ * kRowsY) s_tab[i] = d_mc_table[i];
    __syncthreads();
}

// Count pass.  A warp marches along lattice X (the slowest memory axis) over the nodes (y, y+1; z0 .. z0+31) and owns the 31
// columns (y; z0 .. z0+30): per plane every lane loads its own two nodes -- two coalesced 128-byte rows per warp -- and gets
// the two nodes at z+1 as sign bits from the next lane by one shuffle (lane 31 only supplies them; consecutive warps
// overlap by one node).  The loads of kCountAhead planes are issued before any of them is used, so each thread keeps
// 2 * kCountAhead requests in flight: the first version (one plane at a time, four loads per step) was latency-bound at
// 0.04 of the HBM roofline; with the loads pipelined the pass became instruction-bound (ncu: 44 instructions per
// cell step, issue slots 72 %), hence the lean step below (pointer increments, no per-lane edge cases, trivial cells decided
// on the nibbles).
constexpr int kCountAhead = 8;
constexpr int kCountCols = kLanesZ - 1;  // columns per warp

__global__ void __launch_bounds__(kLanesZ* kRowsY)
    k_mc_count(const Lattice L, const float* __restrict__ field, unsigned int* __restrict__ col_v,
               unsigned int* __restrict__ col_t, unsigned int* __restrict__ col_x) {
    __shared__ unsigned long long s_tab[256];
    load_table(s_tab);
    const int lane = threadIdx.x;
    const int y = (int)(blockIdx.y * kRowsY + threadIdx.y);
    const int z = (int)(blockIdx.x * kCountCols) + lane;
    if (y >= L.SY - 1) return;  // (warp-uniform)
    const bool owns = lane < kCountCols && z < L.SZ - 1;
    const float niso = -L.isoval;
    const float* q = field + (long long)y * L.SZ + min(z, L.SZ - 1);  // (lanes past the last node re-read it: never used)
    auto nibble = [&](float v0, float v1) {
        const unsigned mine = mc::sign_bit(niso, v0) | (mc::sign_bit(niso, v1) << 1);
        return mine | (__shfl_down_sync(0xffffffffu, mine, 1) << 2);
    };
    unsigned n_prev = nibble(q[0], q[L.SZ]);
    mc::ColumnCount cc{0u, 0u, 0, 0};
    const int ncell = L.SX - 1;
    auto step = [&](int x, float v0, float v1) {
        const unsigned n_cur = nibble(v0, v1);
        // ~98 % of the cells are entirely on one side of the level (cases 0 / 255): decided on the nibbles, before the case
        // number is assembled
        if (!((n_prev | n_cur) == 0u || (n_prev & n_cur) == 15u))
            mc::count_cell(cc, s_tab, mc::case_of_nibbles(n_prev, n_cur), x, y, z);
        n_prev = n_cur;
    };
    int x0 = 0;
    for (; x0 + kCountAhead <= ncell; x0 += kCountAhead) {
        float v0[kCountAhead], v1[kCountAhead];
#pragma unroll
        for (int u = 0; u < kCountAhead; u++) {
            q += L.strideX;
            v0[u] = q[0];
            v1[u] = q[L.SZ];
        }
#pragma unroll
        for (int u = 0; u < kCountAhead; u++) step(x0 + u, v0[u], v1[u]);
    }
    for (; x0 < ncell; x0++) {
        q += L.strideX;
        step(x0, q[0], q[L.SZ]);
    }
    if (!owns) return;
    const int c = mc::column_id(L, y, z);
    col_v[c] = cc.nv;
    col_t[c] = cc.nt;
    col_x[c] = mc::pack_range(cc.x_lo, cc.x_hi);
}

// Exclusive scan of the column counts in two launches.  Scan thread t (of kScanBlocks * kScanBlock) owns a short contiguous
// chunk (mc::scan_chunk: 8 columns at 512^2); k_mc_scan_sums leaves one partial sum per CTA, k_mc_scan_write adds the
// partial sums of the CTAs before it, scans its threads' sums, and writes the chunk.  (Round 1 scanned with ONE CTA whose
// threads walked 1 KB chunks element by element: 0.98 ms of the 1.85 ms the whole isosurface took.)
__device__ __forceinline__ void block_sum2(unsigned long long& a, unsigned long long& c, unsigned long long* sa,
                                           unsigned long long* sc) {  // sums over the CTA, returned to every thread
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    __syncthreads();
    if (lane == 0) {
        sa[w] = a;
        sc[w] = c;
    }
    __syncthreads();
    a = c = 0;
    for (int i = 0; i < kScanBlock / 32; i++) {
        a += sa[i];
        c += sc[i];
    }
}

__global__ void __launch_bounds__(kScanBlock)
    k_mc_scan_sums(const unsigned int* __restrict__ col_v, const unsigned int* __restrict__ col_t, int n,
                   unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long sa[kScanBlock / 32], sc[kScanBlock / 32];
    int b, e;
    mc::scan_chunk(n, blockIdx.x * kScanBlock + threadIdx.x, b, e);
    unsigned long long a, c;
    mc::scan_chunk_sum(col_v, col_t, b, e, a, c);
    block_sum2(a, c, sa, sc);
    if (threadIdx.x == 0) {
        block_sums[2 * blockIdx.x] = a;
        block_sums[2 * blockIdx.x + 1] = c;
    }
}

__global__ void __launch_bounds__(kScanBlock)
    k_mc_scan_write(const unsigned int* __restrict__ col_v, const unsigned int* __restrict__ col_t, int n,
                    const unsigned long long* __restrict__ block_sums, unsigned long long* __restrict__ voff,
                    unsigned long long* __restrict__ toff) {
    __shared__ unsigned long long sa[kScanBlock / 32], sc[kScanBlock / 32];
    __shared__ unsigned long long sv[kScanBlock], st[kScanBlock];
    const int t = threadIdx.x;
    // everything the CTAs before this one hold
    unsigned long long pa = 0, pc = 0;
    for (int i = t; i < (int)blockIdx.x; i += kScanBlock) {
        pa += block_sums[2 * i];
        pc += block_sums[2 * i + 1];
    }
    block_sum2(pa, pc, sa, sc);
    int b, e;
    mc::scan_chunk(n, blockIdx.x * kScanBlock + t, b, e);
    unsigned long long a, c;
    mc::scan_chunk_sum(col_v, col_t, b, e, a, c);
    sv[t] = a;
    st[t] = c;
    __syncthreads();
    for (int off = 1; off < kScanBlock; off <<= 1) {  // inclusive scan of the threads' chunk sums
        unsigned long long x = t >= off ? sv[t - off] : 0ull, y = t >= off ? st[t - off] : 0ull;
        __syncthreads();
        sv[t] += x;
        st[t] += y;
        __syncthreads();
    }
    mc::scan_chunk_write(col_v, col_t, b, e, pa + sv[t] - a, pc + st[t] - c, voff, toff);
    if (blockIdx.x == gridDim.x - 1 && t == kScanBlock - 1) {
        voff[n] = pa + sv[t];
        toff[n] = pc + st[t];
    }
}

__global__ void __launch_bounds__(kLanesZ* kRowsY)
    k_mc_vertices(const Lattice L, const float* __restrict__ field, const unsigned long long* __restrict__ voff,
                  const unsigned int* __restrict__ col_x, float* __restrict__ vertices, uint32_t* __restrict__ vkey) {
    int y, z;
    if (!mc::thread_column(L, blockIdx.x, blockIdx.y, threadIdx.x, threadIdx.y, y, z)) return;
    const int c = mc::column_id(L, y, z);
    const unsigned long long v0 = voff[c];
    if (voff[c + 1] == v0) return;  // this column creates nothing: no need to read it again
    int x_lo, x_hi;
    mc::unpack_range(col_x[c], x_lo, x_hi);  // only the cells between the column's first and last non-trivial case
    mc::VertexVisitor vv{&L, y, z, v0, vertices, vkey};
    mc::march_column(L, field, y, z, vv, x_lo, x_hi);
}

__global__ void __launch_bounds__(kLanesZ* kRowsY)
    k_mc_triangles(const Lattice L, const float* __restrict__ field, const unsigned long long* __restrict__ voff,
                   const unsigned long long* __restrict__ toff, const unsigned int* __restrict__ col_x,
                   const uint32_t* __restrict__ vkey, uint32_t* __restrict__ triangles) {
    __shared__ unsigned long long s_tab[256];
    load_table(s_tab);
    int y, z;
    if (!mc::thread_column(L, blockIdx.x, blockIdx.y, threadIdx.x, threadIdx.y, y, z)) return;
    const int c = mc::column_id(L, y, z);
    const unsigned long long t0 = toff[c];
    if (toff[c + 1] == t0) return;
    int x_lo, x_hi;
    mc::unpack_range(col_x[c], x_lo, x_hi);
    mc::TriangleVisitor tv{&L, s_tab, voff, vkey, y, z, t0, triangles};
    mc::march_column(L, field, y, z, tv, x_lo, x_hi);
}

__global__ void k_narrow_f64(const double* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i];  // round to nearest, like the consumer's static_cast (volume_grid.ipp:103-106)
}

// trilinear interpolant of the node values at origin + a*du + b*dv (src/signed_heat_grid_solver.cpp:405-431), fp64
// arithmetic on the float32 field; NaN where the point's cell is not inside the grid
__global__ void k_slice(int nx, int ny, int nz, const float* __restrict__ field, double bx, double by, double bz, double cell,
                        double ox, double oy, double oz, double ux, double uy, double uz, double vx, double vy, double vz,
                        int nu, int nv, float* __restrict__ out) {
    const size_t total = (size_t)nu * nv;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const double a = (double)(q % nu), b = (double)(q / nu);
        // every product and sum rounded on its own (no FMA contraction), like the reference's expressions evaluated by a
        // compiler that does not contract and like the oracle's numpy: one ulp in q decides which cell a sample on a
        // cell face belongs to
        const double qx = __dadd_rn(__dadd_rn(ox, __dmul_rn(a, ux)), __dmul_rn(b, vx));
        const double qy = __dadd_rn(__dadd_rn(oy, __dmul_rn(a, uy)), __dmul_rn(b, vy));
        const double qz = __dadd_rn(__dadd_rn(oz, __dmul_rn(a, uz)), __dmul_rn(b, vz));
        const double fi = floor((qx - bx) / cell), fj = floor((qy - by) / cell), fk = floor((qz - bz) / cell);
        float r = nanf("");
        if (fi >= 0 && fj >= 0 && fk >= 0 && fi < nx - 1 && fj < ny - 1 && fk < nz - 1) {
            const int i = (int)fi, j = (int)fj, k = (int)fk;
            const double tx = (qx - __dadd_rn(bx, __dmul_rn((double)i, cell))) / cell,
                         ty = (qy - __dadd_rn(by, __dmul_rn((double)j, cell))) / cell,
                         tz = (qz - __dadd_rn(bz, __dmul_rn((double)k, cell))) / cell;
            const float* p = field + (size_t)i + (size_t)j * nx + (size_t)k * nx * ny;
            const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
            const double v00 = p[0] * (1. - tx) + p[1] * tx, v01 = p[sz] * (1. - tx) + p[sz + 1] * tx;
            const double v10 = p[sy] * (1. - tx) + p[sy + 1] * tx, v11 = p[sy + sz] * (1. - tx) + p[sy + sz + 1] * tx;
            const double v0 = v00 * (1. - ty) + v10 * ty, v1 = v01 * (1. - ty) + v11 * ty;
            r = (float)(v0 * (1. - tz) + v1 * tz);
        }
        out[q] = r;
    }
}

// ------------------------------------------------------------------------------------------------ host side
IsoSurface::IsoSurface() {
    SHM3D_CUDA_CHECK(cudaHostAlloc((void**)&h_totals_, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    SHM3D_CUDA_CHECK(cudaEventCreate(&ev0_));
    SHM3D_CUDA_CHECK(cudaEventCreate(&ev1_));
}

IsoSurface::~IsoSurface() {
    if (h_totals_) cudaFreeHost(h_totals_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
}

const float* IsoSurface::stage_field(cudaStream_t s, size_t n, const void* field, int kind) {
    if (kind == SHM3D_FIELD_DEVICE_F32) return (const float*)field;
    field32_.alloc(n);
    if (kind == SHM3D_FIELD_HOST_F32) {
        SHM3D_CUDA_CHECK(cudaMemcpyAsync(field32_.p, field, n * sizeof(float), cudaMemcpyHostToDevice, s));
        return field32_.p;
    }
    if (kind != SHM3D_FIELD_HOST_F64) throw Error(SHM3D_ERR_INVALID_ARG, "unknown field kind");
    const size_t chunk = (size_t)1 << 26;  // 512 MB of doubles per hop
    stage64_.alloc(n < chunk ? n : chunk);
    const double* h = (const double*)field;
    for (size_t b = 0; b < n; b += chunk) {
        const size_t cnt = n - b < chunk ? n - b : chunk;
        SHM3D_CUDA_CHECK(cudaMemcpyAsync(stage64_.p, h + b, cnt * sizeof(double), cudaMemcpyHostToDevice, s));
        k_narrow_f64<<<148 * 8, 256, 0, s>>>(stage64_.p, field32_.p + b, cnt);
        SHM3D_LAUNCHED();
        SHM3D_CUDA_CHECK(cudaGetLastError());
    }
    return field32_.p;
}

IsoResult IsoSurface::extract(cudaStream_t s, int nx, int ny, int nz, const float* d_field, float isoval,
                              const float* bound_min, const float* bound_max) {
    if (nx < 2 || ny < 2 || nz < 2) throw Error(SHM3D_ERR_INVALID_ARG, "isosurface: the grid needs at least 2 nodes per axis");
    if ((bound_min == nullptr) != (bound_max == nullptr))
        throw Error(SHM3D_ERR_INVALID_ARG, "isosurface: give both bounds or neither");
    if ((long long)(nx - 1) * (ny - 1) > 0x7fffffffLL) throw Error(SHM3D_ERR_INVALID_ARG, "isosurface: too many columns");
    const Lattice L = mc::make_lattice(nx, ny, nz, isoval, bound_min, bound_max);
    const int nc = L.ncols();
    col_v_.alloc((size_t)nc);
    col_t_.alloc((size_t)nc);
    col_x_.alloc((size_t)nc);
    voff_.alloc((size_t)nc + 1);
    toff_.alloc((size_t)nc + 1);
    unsigned gx, gy;
    mc::launch_grid(L, gx, gy);
    const dim3 block(kLanesZ, kRowsY), grid(gx, gy);
    if (grid.y > 65535u) throw Error(SHM3D_ERR_INVALID_ARG, "isosurface: ny too large");
    if (L.SX > 65535) throw Error(SHM3D_ERR_INVALID_ARG, "isosurface: nz too large");  // (the packed per-column cell range)
    IsoResult res;
    SHM3D_CUDA_CHECK(cudaEventRecord(ev0_, s));
    const dim3 grid_count((unsigned)((L.SZ - 1 + kCountCols - 1) / kCountCols), gy);  // 31 columns per warp
    k_mc_count<<<grid_count, block, 0, s>>>(L, d_field, col_v_.p, col_t_.p, col_x_.p);
    SHM3D_LAUNCHED();
    block_sums_.alloc(2 * (size_t)kScanBlocks);
    k_mc_scan_sums<<<kScanBlocks, kScanBlock, 0, s>>>(col_v_.p, col_t_.p, nc, block_sums_.p);
    SHM3D_LAUNCHED();
    k_mc_scan_write<<<kScanBlocks, kScanBlock, 0, s>>>(col_v_.p, col_t_.p, nc, block_sums_.p, voff_.p, toff_.p);
    SHM3D_LAUNCHED();
    SHM3D_CUDA_CHECK(cudaGetLastError());
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(&h_totals_[0], voff_.p + nc, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(&h_totals_[1], toff_.p + nc, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(s));
    const unsigned long long nV = h_totals_[0], nT = h_totals_[1];
    if (nV >= 0xffffffffULL) throw Error(SHM3D_ERR_INVALID_ARG, "isosurface: more than 2^32 vertices");
    res.launches = 2;
    if (nV > 0) {
        verts_.alloc((size_t)(3 * nV));
        vkey_.alloc((size_t)nV);
        k_mc_vertices<<<grid, block, 0, s>>>(L, d_field, voff_.p, col_x_.p, verts_.p, vkey_.p);
        SHM3D_LAUNCHED();
        res.launches++;
    }
    if (nT > 0) {  // a triangle only references edges that cross, i.e. vertices that exist
        tris_.alloc((size_t)(3 * nT));
        k_mc_triangles<<<grid, block, 0, s>>>(L, d_field, voff_.p, toff_.p, col_x_.p, vkey_.p, tris_.p);
        SHM3D_LAUNCHED();
        res.launches++;
    }
    SHM3D_CUDA_CHECK(cudaGetLastError());
    SHM3D_CUDA_CHECK(cudaEventRecord(ev1_, s));
    SHM3D_CUDA_CHECK(cudaEventSynchronize(ev1_));
    float ms = 0;
    SHM3D_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
    res.ms_device = ms;
    res.n_vertices = (int64_t)nV;
    res.n_triangles = (int64_t)nT;
    last_ = res;
    return res;
}

void IsoSurface::fetch(cudaStream_t s, float* vertices_out, uint32_t* triangles_out) const {
    if (vertices_out && last_.n_vertices)
        SHM3D_CUDA_CHECK(cudaMemcpyAsync(vertices_out, verts_.p, (size_t)last_.n_vertices * 3 * sizeof(float),
                                         cudaMemcpyDeviceToHost, s));
    if (triangles_out && last_.n_triangles)
        SHM3D_CUDA_CHECK(cudaMemcpyAsync(triangles_out, tris_.p, (size_t)last_.n_triangles * 3 * sizeof(uint32_t),
                                         cudaMemcpyDeviceToHost, s));
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(s));
}

int64_t IsoSurface::slice(cudaStream_t s, int nx, int ny, int nz, const float* d_field, const double bbox_min[3], double cell,
                          const double origin[3], const double du[3], const double dv[3], int nu, int nv, float* out_host) {
    if (nu < 1 || nv < 1 || !out_host || !(cell > 0)) throw Error(SHM3D_ERR_INVALID_ARG, "slice: bad arguments");
    const size_t total = (size_t)nu * nv;
    slice_.alloc(total);
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_slice<<<blocks, 256, 0, s>>>(nx, ny, nz, d_field, bbox_min[0], bbox_min[1], bbox_min[2], cell, origin[0], origin[1],
                                   origin[2], du[0], du[1], du[2], dv[0], dv[1], dv[2], nu, nv, slice_.p);
    SHM3D_LAUNCHED();
    SHM3D_CUDA_CHECK(cudaGetLastError());
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(out_host, slice_.p, total * sizeof(float), cudaMemcpyDeviceToHost, s));
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(s));
    return 1;
}

}  // namespace shm3d
