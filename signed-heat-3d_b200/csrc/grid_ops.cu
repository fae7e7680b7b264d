// grid_ops.cu -- matrix-free operators on the regular grid for Step 3 (all HBM-bandwidth-bound).
//
// Replaces the Eigen sparse machinery of the reference: gradient() / D^T Y
// (src/signed_heat_grid_solver.cpp:336-402, :70-74), laplacian() (:278-334) and the sparse LU of the KKT
// system (:101-108) -- here a constrained multigrid-preconditioned CG whose per-iteration work is the kernels
// below.  K' = -cell^2 L is the integer 7-point Neumann stencil: (K'u)_i = sum_{in-range nbr} (u_i - u_nbr).
//
// Layout: x-fastest float arrays; every vector that is read through a stencil is allocated with one ghost
// plane below and above the local z-slab, and kernels receive the pointer to the first interior plane.
//
// Kernel shape (B200): persistent grid (<= 8 CTAs of 256 threads per SM).  Grids with nx % 4 == 0 -- every grid the
// reference can produce -- use the row-oriented kernels of grid_rows.cuh (a warp owns a grid row: boundary predicates and
// reciprocal diagonals once per row, float4 quads per lane, x-neighbours by shuffle).  The element-indexed templates
// below (V = 1) remain as the general path for other sizes and for the pointwise kernels (V = 4: x/r update, dots).
// Reductions are fp64 and deterministic: one partial per CTA, the last CTA to finish (ticket) folds them in a fixed
// order.
#include <algorithm>
#include <functional>

#include <cooperative_groups.h>

#include "proj_dev.cuh"

namespace shm3d {

int64_t g_kernel_launches = 0;

namespace {

constexpr int kT = 256;
constexpr int kMaxBlocks = 148 * 8;

// ---------------------------------------------------------------- deterministic reduction helper
struct RedScratch {
    double* partials;
    unsigned int* counter;
};

// The per-CTA partials and the ticket counter belong to the CONTEXT whose call is running on this host thread
// (set_reduction_scratch at every API entry): two contexts -- other devices, or other streams of one device -- never
// share them.
thread_local RedScratch t_scratch{nullptr, nullptr};

RedScratch red_scratch() {
    if (!t_scratch.partials) throw Error(SHM3D_ERR_INVALID_ARG, "internal: reduction scratch not bound to a context");
    return t_scratch;
}

template <int K>
__device__ __forceinline__ void block_reduce_commit(double (&v)[K], RedScratch rs, double* out) {
    __shared__ double s_w[K][kT / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_w[k][w] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            double x = 0;
#pragma unroll
            for (int i = 0; i < kT / 32; i++) x += s_w[k][i];
            rs.partials[(size_t)blockIdx.x * K + k] = x;
        }
        __threadfence();
        unsigned int ticket = atomicAdd(rs.counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
#pragma unroll
        for (int k = 0; k < K; k++) {
            double x = 0;
            for (unsigned int b = threadIdx.x; b < gridDim.x; b += kT) x += rs.partials[(size_t)b * K + k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            __syncthreads();
            if (lane == 0) s_w[k][w] = x;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0;
#pragma unroll
                for (int i = 0; i < kT / 32; i++) t += s_w[k][i];
                out[k] = t;
            }
        }
        if (threadIdx.x == 0) *rs.counter = 0;
    }
}

// ---------------------------------------------------------------- vector access helpers (V = 4 or 1)
template <int V>
struct Vec;
template <>
struct Vec<4> {
    float v[4];
    __device__ __forceinline__ static Vec ld(const float* p) {
        float4 t = *reinterpret_cast<const float4*>(p);
        Vec r;
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
        return r;
    }
    __device__ __forceinline__ void st(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <>
struct Vec<1> {
    float v[1];
    __device__ __forceinline__ static Vec ld(const float* p) { Vec r; r.v[0] = *p; return r; }
    __device__ __forceinline__ void st(float* p) const { *p = v[0]; }
};
template <int V>
__device__ __forceinline__ Vec<V> vzero() {
    Vec<V> r;
#pragma unroll
    for (int t = 0; t < V; t++) r.v[t] = 0.f;
    return r;
}

// element index -> (i0, j, kl) for the group starting at element e (e % V == 0, nx % V == 0)
__device__ __forceinline__ void decode(unsigned int e, int nx, int ny, int& i, int& j, int& kl) {
    unsigned int row = e / (unsigned int)nx;
    i = (int)(e - row * (unsigned int)nx);
    kl = (int)(row / (unsigned int)ny);
    j = (int)(row - (unsigned int)kl * (unsigned int)ny);
}

// K'u for the V nodes starting at e, and the per-node diagonal (number of in-range neighbours)
template <int V>
__device__ __forceinline__ void stencil(const float* __restrict__ u, unsigned int e, const Vec<V>& c, int i0, int j, int k,
                                        const LevelDims& L, Vec<V>& Ku, Vec<V>& dg) {
    const unsigned int pl = (unsigned int)L.nx * (unsigned int)L.ny;
    const bool ym = j > 0, yp = j < L.ny - 1, zm = k > 0, zp = k < L.nz - 1;
    Vec<V> a = ym ? Vec<V>::ld(u + e - L.nx) : vzero<V>();
    Vec<V> b = yp ? Vec<V>::ld(u + e + L.nx) : vzero<V>();
    Vec<V> d = zm ? Vec<V>::ld(u + (ptrdiff_t)e - (ptrdiff_t)pl) : vzero<V>();
    Vec<V> f = zp ? Vec<V>::ld(u + e + pl) : vzero<V>();
    const float left = (i0 > 0) ? u[e - 1] : 0.f;
    const float right = (i0 + V < L.nx) ? u[e + V] : 0.f;
    const int cyz = (int)ym + (int)yp + (int)zm + (int)zp;
#pragma unroll
    for (int t = 0; t < V; t++) {
        const int i = i0 + t;
        const float xl = (t > 0) ? c.v[t - 1] : left;
        const float xr = (t < V - 1) ? c.v[t + 1] : right;
        const float cnt = (float)(cyz + (i > 0) + (i < L.nx - 1));
        dg.v[t] = cnt;
        Ku.v[t] = cnt * c.v[t] - (xl + xr + a.v[t] + b.v[t] + d.v[t] + f.v[t]);
    }
}

#define GRID_STRIDE_GROUPS(L_, V_)                                                                   \
    const unsigned int _ng = (unsigned int)((L_).n() / (V_));                                          \
    for (unsigned int _g = blockIdx.x * kT + threadIdx.x; _g < _ng; _g += gridDim.x * kT)

// ---------------------------------------------------------------- b = cell * D'^T Y
__global__ void __launch_bounds__(kT) k_div_rhs(LevelDims L, float cell, const float* __restrict__ Y, size_t cs,
                                                float* __restrict__ b, int scrub, unsigned int* nonfinite) {
    const size_t n = L.n();
    for (size_t e = (size_t)blockIdx.x * kT + threadIdx.x; e < n; e += (size_t)gridDim.x * kT) {
        size_t row = e / (size_t)L.nx;
        const int i = (int)(e - row * (size_t)L.nx);
        const int kl = (int)(row / (size_t)L.ny);
        const int j = (int)(row - (size_t)kl * L.ny);
        const int k = L.k0 + kl;
        const size_t pl = L.plane();
        const float* Yx = Y;
        const float* Yy = Y + cs;
        const float* Yz = Y + 2 * cs;
        // per axis, line index t of n nodes, g = Y_a:
        //   b_t = [t>=1] g[t-1] - [t<=n-2] g[t] - [t==n-2] g[t+1] + [t==n-1] g[t]
        float acc = 0.f;
        {
            float g = Yx[e];
            if (i >= 1) acc += Yx[e - 1];
            if (i <= L.nx - 2) acc -= g;
            if (i == L.nx - 2) acc -= Yx[e + 1];
            if (i == L.nx - 1) acc += g;
        }
        {
            float g = Yy[e];
            if (j >= 1) acc += Yy[e - L.nx];
            if (j <= L.ny - 2) acc -= g;
            if (j == L.ny - 2) acc -= Yy[e + L.nx];
            if (j == L.ny - 1) acc += g;
        }
        {
            float g = Yz[e];
            if (k >= 1) acc += Yz[e - pl];
            if (k <= L.nz - 2) acc -= g;
            if (k == L.nz - 2) acc -= Yz[e + pl];
            if (k == L.nz - 1) acc += g;
        }
        float v = cell * acc;
        if (!isfinite(v)) {
            atomicAdd(nonfinite, 1u);
            if (scrub) v = 0.f;
        }
        b[e] = v;
    }
}

// ---------------------------------------------------------------- q = K'p, out = sum p q
template <int V>
__global__ void __launch_bounds__(kT) k_stencil_dot(LevelDims L, const float* __restrict__ p, float* __restrict__ q,
                                                    RedScratch rs, double* out) {
    double acc[1] = {0.0};
    GRID_STRIDE_GROUPS(L, V) {
        const unsigned int e = _g * V;
        int i0, j, kl;
        decode(e, L.nx, L.ny, i0, j, kl);
        Vec<V> c = Vec<V>::ld(p + e), Ku, dg;
        stencil<V>(p, e, c, i0, j, L.k0 + kl, L, Ku, dg);
        Ku.st(q + e);
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < V; t++) s = fmaf(c.v[t], Ku.v[t], s);
        acc[0] += (double)s;
    }
    block_reduce_commit<1>(acc, rs, out);
}

// ---------------------------------------------------------------- x += a p, r -= a q, out = sum r
template <int V>
__global__ void __launch_bounds__(kT) k_update_xr(LevelDims L, float* __restrict__ x, float* __restrict__ r,
                                                  const float* __restrict__ p, const float* __restrict__ q,
                                                  const double* rho, const double* pq, RedScratch rs, double* out) {
    const float a = (float)(*rho / *pq);
    double acc[1] = {0.0};
    GRID_STRIDE_GROUPS(L, V) {
        const unsigned int e = _g * V;
        Vec<V> xv = Vec<V>::ld(x + e), rv = Vec<V>::ld(r + e), pv = Vec<V>::ld(p + e), qv = Vec<V>::ld(q + e);
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < V; t++) {
            xv.v[t] = fmaf(a, pv.v[t], xv.v[t]);
            rv.v[t] = fmaf(-a, qv.v[t], rv.v[t]);
            s += rv.v[t];
        }
        xv.st(x + e);
        rv.st(r + e);
        acc[0] += (double)s;
    }
    block_reduce_commit<1>(acc, rs, out);
}

template <int V>
__global__ void __launch_bounds__(kT) k_dot_rz(LevelDims L, const float* __restrict__ r, const float* __restrict__ z,
                                               RedScratch rs, double* out) {
    double acc[2] = {0.0, 0.0};
    GRID_STRIDE_GROUPS(L, V) {
        const unsigned int e = _g * V;
        Vec<V> rv = Vec<V>::ld(r + e), zv = Vec<V>::ld(z + e);
        float s = 0.f, sz = 0.f;
#pragma unroll
        for (int t = 0; t < V; t++) {
            s = fmaf(rv.v[t], zv.v[t], s);
            sz += zv.v[t];
        }
        acc[0] += (double)s;
        acc[1] += (double)sz;
    }
    block_reduce_commit<2>(acc, rs, out);
}

template <int V>
__global__ void __launch_bounds__(kT) k_update_p(LevelDims L, float* p, const float* pin, const float* __restrict__ z,
                                                 const double* sum_z, double n_global, const double* rho_new,
                                                 const double* rho_old, int first) {
    const float mean = (float)(*sum_z / n_global);
    const float beta = first ? 0.f : (float)(*rho_new / *rho_old);
    GRID_STRIDE_GROUPS(L, V) {
        const unsigned int e = _g * V;
        Vec<V> zv = Vec<V>::ld(z + e), pv = first ? vzero<V>() : Vec<V>::ld(pin + e);
#pragma unroll
        for (int t = 0; t < V; t++) pv.v[t] = fmaf(beta, pv.v[t], zv.v[t] - mean);
        pv.st(p + e);
    }
}

// p <- (z - mean) + beta p ;  q = K'p ;  out = sum p q   in one pass: the new p at the six neighbours is recomputed from
// z and the old p there (cache hits), so the iteration reads z, p and writes p, q: 4 words instead of 3 + 2.
// p is updated in place, which is safe only because neighbours are recomputed from p_old -- hence the separate output pn.
template <int V>
__global__ void __launch_bounds__(kT) k_update_p_stencil(LevelDims L, float* __restrict__ pn, const float* __restrict__ p,
                                                         const float* __restrict__ z, float* __restrict__ q,
                                                         const double* sum_z, double n_global, const double* rho_new,
                                                         const double* rho_old, int first, RedScratch rs, double* out) {
    const float mean = (float)(*sum_z / n_global);
    const float beta = first ? 0.f : (float)(*rho_new / *rho_old);
    const unsigned int pl = (unsigned int)L.nx * (unsigned int)L.ny;
    double acc[1] = {0.0};
    GRID_STRIDE_GROUPS(L, V) {
        const unsigned int e = _g * V;
        int i0, j, kl;
        decode(e, L.nx, L.ny, i0, j, kl);
        const int k = L.k0 + kl;
        const bool ym = j > 0, yp = j < L.ny - 1, zm = k > 0, zp = k < L.nz - 1;
        auto comb = [&](const Vec<V>& zv, const Vec<V>& pv) {
            Vec<V> r;
#pragma unroll
            for (int t = 0; t < V; t++) r.v[t] = fmaf(beta, pv.v[t], zv.v[t] - mean);
            return r;
        };
        auto ldp = [&](ptrdiff_t off) { return first ? vzero<V>() : Vec<V>::ld(p + off); };
        const Vec<V> c = comb(Vec<V>::ld(z + e), ldp(e));
        const Vec<V> a = ym ? comb(Vec<V>::ld(z + e - L.nx), ldp((ptrdiff_t)e - L.nx)) : vzero<V>();
        const Vec<V> bq = yp ? comb(Vec<V>::ld(z + e + L.nx), ldp((ptrdiff_t)e + L.nx)) : vzero<V>();
        const Vec<V> d = zm ? comb(Vec<V>::ld(z + (ptrdiff_t)e - (ptrdiff_t)pl), ldp((ptrdiff_t)e - (ptrdiff_t)pl)) : vzero<V>();
        const Vec<V> f = zp ? comb(Vec<V>::ld(z + e + pl), ldp((ptrdiff_t)e + pl)) : vzero<V>();
        const float left = (i0 > 0) ? fmaf(beta, first ? 0.f : p[e - 1], z[e - 1] - mean) : 0.f;
        const float right = (i0 + V < L.nx) ? fmaf(beta, first ? 0.f : p[e + V], z[e + V] - mean) : 0.f;
        const int cyz = (int)ym + (int)yp + (int)zm + (int)zp;
        Vec<V> Ku;
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < V; t++) {
            const int i = i0 + t;
            const float xl = (t > 0) ? c.v[t > 0 ? t - 1 : 0] : left;
            const float xr = (t < V - 1) ? c.v[t < V - 1 ? t + 1 : 0] : right;
            const float cnt = (float)(cyz + (i > 0) + (i < L.nx - 1));
            Ku.v[t] = cnt * c.v[t] - (xl + xr + a.v[t] + bq.v[t] + d.v[t] + f.v[t]);
            s = fmaf(c.v[t], Ku.v[t], s);
        }
        c.st(pn + e);
        Ku.st(q + e);
        acc[0] += (double)s;
    }
    block_reduce_commit<1>(acc, rs, out);
}

__global__ void __launch_bounds__(kT) k_fill(float* p, size_t n, float v) {
    for (size_t e = (size_t)blockIdx.x * kT + threadIdx.x; e < n; e += (size_t)gridDim.x * kT) p[e] = v;
}
__global__ void __launch_bounds__(kT) k_copy(float* d, const float* s, size_t n) {
    for (size_t e = (size_t)blockIdx.x * kT + threadIdx.x; e < n; e += (size_t)gridDim.x * kT) d[e] = s[e];
}
__global__ void __launch_bounds__(kT) k_vec_sum(const float* v, size_t n, RedScratch rs, double* out) {
    double a[1] = {0.0};
    for (size_t e = (size_t)blockIdx.x * kT + threadIdx.x; e < n; e += (size_t)gridDim.x * kT) a[0] += (double)v[e];
    block_reduce_commit<1>(a, rs, out);
}
__global__ void __launch_bounds__(kT) k_axpy_const(float* v, size_t n, const double* num, double den, float sign) {
    const float a = sign * (float)(*num / den);
    for (size_t e = (size_t)blockIdx.x * kT + threadIdx.x; e < n; e += (size_t)gridDim.x * kT) v[e] += a;
}

// ---------------------------------------------------------------- multigrid
template <int V>
__global__ void __launch_bounds__(kT) k_mg_smooth0(LevelDims L, float* __restrict__ x, const float* __restrict__ b,
                                                   const double* sum_b, double n_global, float omega) {
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    GRID_STRIDE_GROUPS(L, V) {
        const unsigned int e = _g * V;
        int i0, j, kl;
        decode(e, L.nx, L.ny, i0, j, kl);
        const int k = L.k0 + kl;
        const int cyz = (j > 0) + (j < L.ny - 1) + (k > 0) + (k < L.nz - 1);
        Vec<V> bv = Vec<V>::ld(b + e), xv;
#pragma unroll
        for (int t = 0; t < V; t++) {
            const int i = i0 + t;
            xv.v[t] = omega * (bv.v[t] - shift) / (float)(cyz + (i > 0) + (i < L.nx - 1));
        }
        xv.st(x + e);
    }
}

// xo = x + omega ((b - shift) - K'x) / d.  DOT: also acc = { sum b*xo, sum xo } (the r.z and sum z of the PCG when this
// is the last sweep of the fine level: saves a full read of r and z).
template <int V, bool DOT>
__global__ void __launch_bounds__(kT) k_mg_smooth(LevelDims L, float* __restrict__ xo, const float* __restrict__ x,
                                                  const float* __restrict__ b, const double* sum_b, double n_global,
                                                  float omega, RedScratch rs, double* out) {
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    double acc[2] = {0.0, 0.0};
    GRID_STRIDE_GROUPS(L, V) {
        const unsigned int e = _g * V;
        int i0, j, kl;
        decode(e, L.nx, L.ny, i0, j, kl);
        Vec<V> c = Vec<V>::ld(x + e), bv = Vec<V>::ld(b + e), Ku, dg, o;
        stencil<V>(x, e, c, i0, j, L.k0 + kl, L, Ku, dg);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int t = 0; t < V; t++) {
            o.v[t] = c.v[t] + omega * ((bv.v[t] - shift) - Ku.v[t]) / dg.v[t];
            if (DOT) {
                s0 = fmaf(bv.v[t], o.v[t], s0);
                s1 += o.v[t];
            }
        }
        o.st(xo + e);
        if (DOT) {
            acc[0] += (double)s0;
            acc[1] += (double)s1;
        }
    }
    if (DOT) block_reduce_commit<2>(acc, rs, out);
}

// First two damped-Jacobi sweeps from a zero guess in ONE pass over b:
//   x1 = omega (b - shift)/d ;  x2 = x1 + omega ((b - shift) - K'x1)/d
// x1 at the six neighbours is recomputed from b (their d from their coordinates): 2 words of traffic instead of 5.
template <int V>
__global__ void __launch_bounds__(kT) k_mg_smooth01(LevelDims L, float* __restrict__ xo, const float* __restrict__ b,
                                                    const double* sum_b, double n_global, float omega, float omega2) {
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    const unsigned int pl = (unsigned int)L.nx * (unsigned int)L.ny;
    GRID_STRIDE_GROUPS(L, V) {
        const unsigned int e = _g * V;
        int i0, j, kl;
        decode(e, L.nx, L.ny, i0, j, kl);
        const int k = L.k0 + kl;
        const bool ym = j > 0, yp = j < L.ny - 1, zm = k > 0, zp = k < L.nz - 1;
        // neighbour counts along y and z of this row and of the adjacent rows / planes
        const int cy = (int)ym + (int)yp, cz = (int)zm + (int)zp;
        const int cy_m = (int)(j - 1 > 0) + 1, cy_p = 1 + (int)(j + 1 < L.ny - 1);  // rows j-1 / j+1 (when they exist)
        const int cz_m = (int)(k - 1 > 0) + 1, cz_p = 1 + (int)(k + 1 < L.nz - 1);
        const Vec<V> bc = Vec<V>::ld(b + e);
        const Vec<V> ba = ym ? Vec<V>::ld(b + e - L.nx) : vzero<V>();
        const Vec<V> bb = yp ? Vec<V>::ld(b + e + L.nx) : vzero<V>();
        const Vec<V> bd = zm ? Vec<V>::ld(b + (ptrdiff_t)e - (ptrdiff_t)pl) : vzero<V>();
        const Vec<V> bf = zp ? Vec<V>::ld(b + e + pl) : vzero<V>();
        const bool has_l = i0 > 0, has_r = i0 + V < L.nx;
        const float bl = has_l ? b[e - 1] : 0.f, br = has_r ? b[e + V] : 0.f;
        // x1 of the two x-end neighbours
        const int cxl = (int)(i0 - 1 > 0) + 1, cxr = 1 + (int)(i0 + V < L.nx - 1);
        const float x1l = has_l ? omega * (bl - shift) / (float)(cxl + cy + cz) : 0.f;
        const float x1r = has_r ? omega * (br - shift) / (float)(cxr + cy + cz) : 0.f;
        float x1c[V];
        int cx[V];
#pragma unroll
        for (int t = 0; t < V; t++) {
            const int i = i0 + t;
            cx[t] = (int)(i > 0) + (int)(i < L.nx - 1);
            x1c[t] = omega * (bc.v[t] - shift) / (float)(cx[t] + cy + cz);
        }
        Vec<V> o;
#pragma unroll
        for (int t = 0; t < V; t++) {
            const float d = (float)(cx[t] + cy + cz);
            const float xl = (t > 0) ? x1c[t > 0 ? t - 1 : 0] : x1l;
            const float xr = (t < V - 1) ? x1c[t < V - 1 ? t + 1 : 0] : x1r;
            const float xa = ym ? omega * (ba.v[t] - shift) / (float)(cx[t] + cy_m + cz) : 0.f;
            const float xb = yp ? omega * (bb.v[t] - shift) / (float)(cx[t] + cy_p + cz) : 0.f;
            const float xd = zm ? omega * (bd.v[t] - shift) / (float)(cx[t] + cy + cz_m) : 0.f;
            const float xf = zp ? omega * (bf.v[t] - shift) / (float)(cx[t] + cy + cz_p) : 0.f;
            const float Kx = d * x1c[t] - (xl + xr + xa + xb + xd + xf);
            o.v[t] = x1c[t] + omega2 * ((bc.v[t] - shift) - Kx) / d;
        }
        o.st(xo + e);
    }
}

template <int V>
__global__ void __launch_bounds__(kT) k_mg_residual(LevelDims L, const float* __restrict__ x,
                                                    const float* __restrict__ b, const double* sum_b, double n_global,
                                                    float* __restrict__ r) {
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    GRID_STRIDE_GROUPS(L, V) {
        const unsigned int e = _g * V;
        int i0, j, kl;
        decode(e, L.nx, L.ny, i0, j, kl);
        Vec<V> c = Vec<V>::ld(x + e), bv = Vec<V>::ld(b + e), Ku, dg, o;
        stencil<V>(x, e, c, i0, j, L.k0 + kl, L, Ku, dg);
#pragma unroll
        for (int t = 0; t < V; t++) o.v[t] = (bv.v[t] - shift) - Ku.v[t];
        o.st(r + e);
    }
}

// 1D restriction weights of the transposed clamped trilinear prolongation: coarse I gathers fine 2I-1..2I+2
__device__ __forceinline__ void rweights(int I, int nc, float (&wt)[4]) {
    wt[0] = (I > 0) ? 0.25f : 0.f;
    wt[1] = (I > 0) ? 0.75f : 1.0f;
    wt[2] = (I < nc - 1) ? 0.75f : 1.0f;
    wt[3] = (I < nc - 1) ? 0.25f : 0.f;
}

__global__ void __launch_bounds__(kT) k_mg_restrict(LevelDims Lf, LevelDims Lc, const float* __restrict__ r,
                                                    float* __restrict__ bc) {
    const unsigned int nc = (unsigned int)Lc.n();
    for (unsigned int e = blockIdx.x * kT + threadIdx.x; e < nc; e += gridDim.x * kT) {
        int I, J, Kl;
        decode(e, Lc.nx, Lc.ny, I, J, Kl);
        const int K = Lc.k0 + Kl;
        float wx[4], wy[4], wz[4];
        rweights(I, Lc.nx, wx);
        rweights(J, Lc.ny, wy);
        rweights(K, Lc.nz, wz);
        const ptrdiff_t plf = (ptrdiff_t)Lf.plane();
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (wz[c] == 0.f) continue;
            const int kf = 2 * K - 1 + c - Lf.k0;  // local fine plane (may be a ghost plane: -1 or nzl)
#pragma unroll
            for (int bq = 0; bq < 4; bq++) {
                if (wy[bq] == 0.f) continue;
                const int jf = 2 * J - 1 + bq;
                const float* row = r + (ptrdiff_t)kf * plf + (ptrdiff_t)jf * Lf.nx;
                float s = 0.f;
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    if (wx[a] == 0.f) continue;
                    s = fmaf(wx[a], row[2 * I - 1 + a], s);
                }
                acc = fmaf(wy[bq] * wz[c], s, acc);
            }
        }
        bc[e] = 0.5f * acc;
    }
}

template <int V>
__global__ void __launch_bounds__(kT) k_mg_prolong_add(LevelDims Lf, LevelDims Lc, float* __restrict__ x,
                                                       const float* __restrict__ ec) {
    GRID_STRIDE_GROUPS(Lf, V) {
        const unsigned int e = _g * V;
        int i0, j, kl;
        decode(e, Lf.nx, Lf.ny, i0, j, kl);
        const int k = Lf.k0 + kl;
        const int J0 = j >> 1, K0 = k >> 1;
        const int J1 = min(max((j & 1) ? J0 + 1 : J0 - 1, 0), Lc.ny - 1);
        const int K1 = min(max((k & 1) ? K0 + 1 : K0 - 1, 0), Lc.nz - 1);
        const ptrdiff_t plc = (ptrdiff_t)Lc.plane();
        const float* r00 = ec + (ptrdiff_t)(K0 - Lc.k0) * plc + (ptrdiff_t)J0 * Lc.nx;
        const float* r01 = ec + (ptrdiff_t)(K0 - Lc.k0) * plc + (ptrdiff_t)J1 * Lc.nx;
        const float* r10 = ec + (ptrdiff_t)(K1 - Lc.k0) * plc + (ptrdiff_t)J0 * Lc.nx;  // may be a ghost plane
        const float* r11 = ec + (ptrdiff_t)(K1 - Lc.k0) * plc + (ptrdiff_t)J1 * Lc.nx;
        Vec<V> xv = Vec<V>::ld(x + e);
#pragma unroll
        for (int t = 0; t < V; t++) {
            const int i = i0 + t;
            const int I0 = i >> 1;
            const int I1 = min(max((i & 1) ? I0 + 1 : I0 - 1, 0), Lc.nx - 1);
            // y/z-interpolated coarse values at columns I0 and I1
            const float c0 = 0.75f * (0.75f * r00[I0] + 0.25f * r01[I0]) + 0.25f * (0.75f * r10[I0] + 0.25f * r11[I0]);
            const float c1 = 0.75f * (0.75f * r00[I1] + 0.25f * r01[I1]) + 0.25f * (0.75f * r10[I1] + 0.25f * r11[I1]);
            xv.v[t] += 0.75f * c0 + 0.25f * c1;
        }
        xv.st(x + e);
    }
}

// coarsest level: dense pseudo-inverse matvec, one CTA (n3 <= 512)
__global__ void __launch_bounds__(512) k_mg_coarse(int n3, const float* __restrict__ pinv, const float* __restrict__ b,
                                                   float* __restrict__ x) {
    __shared__ float sb[512];
    int t = threadIdx.x;
    if (t < n3) sb[t] = b[t];
    __syncthreads();
    if (t < n3) {
        float acc = 0.f;
        for (int c = 0; c < n3; c++) acc = fmaf(pinv[(size_t)t * n3 + c], sb[c], acc);
        x[t] = acc;
    }
}

#include "grid_rows.cuh"

inline unsigned int nblk_rows(const LevelDims& L) {
    size_t rows = (size_t)L.ny * (size_t)L.nzl();
    size_t b = (rows + kT / 32 - 1) / (kT / 32);
    return (unsigned int)std::max<size_t>(1, std::min<size_t>(b, kMaxBlocks));
}

// TMA-staged marching kernels (grid_march.cuh): bound per host thread at every API entry like the reduction scratch
thread_local bool g_march_disabled = false;
thread_local int g_march_sms = 148;
#include "grid_march.cuh"
#include "mg_tail.cuh"
// ---------------------------------------------------------------- fastIntegration (reference integrateGreedily, :224-275)
// The reference runs a FIFO breadth-first search from node (0,0,0), visiting neighbours in the order -x,+x,-y,+y,-z,+z,
// and sets phi[q] = phi[p] + normalize(Y_p + Y_q) . (q - p) for the first p that reaches q.  On the full box the BFS
// levels are the planes i+j+k = const and, by induction, each level is dequeued in descending lexicographic (i,j,k)
// order, so the first visitor of q = (i,j,k) is (i,j,k-1) if k > 0, else (i,j-1,0) if j > 0, else (i-1,0,0)
// (checked against the literal BFS in tests/test_oracle.py).  The whole integration is therefore three prefix sums:
// along x on the line j=k=0, along y on the plane k=0, along z everywhere.
__device__ __forceinline__ float fast_step(float ax, float ay, float az, float bx, float by, float bz, int axis, float cell) {
    const float sx = ax + bx, sy = ay + by, sz = az + bz;
    const float n = sqrtf(sx * sx + sy * sy + sz * sz);
    const float c = axis == 0 ? sx : (axis == 1 ? sy : sz);
    return c / n * cell;
}

// plane k = 0 of the global grid (one CTA; only the rank that owns plane 0 runs it)
__global__ void __launch_bounds__(1024) k_fast_base(LevelDims L, float cell, const float* __restrict__ Y, size_t cs,
                                                    float* __restrict__ phi) {
    const float* Yx = Y;
    const float* Yy = Y + cs;
    const float* Yz = Y + 2 * cs;
    if (threadIdx.x == 0) {
        float acc = 0.f;
        phi[0] = 0.f;
        for (int i = 1; i < L.nx; i++) {
            acc += fast_step(Yx[i - 1], Yy[i - 1], Yz[i - 1], Yx[i], Yy[i], Yz[i], 0, cell);
            phi[i] = acc;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L.nx; i += blockDim.x) {
        float acc = phi[i];
        for (int j = 1; j < L.ny; j++) {
            const size_t a = (size_t)i + (size_t)(j - 1) * L.nx, b = a + L.nx;
            acc += fast_step(Yx[a], Yy[a], Yz[a], Yx[b], Yy[b], Yz[b], 1, cell);
            phi[b] = acc;
        }
    }
}

// z prefix sums of this rank's slab: one thread per (i,j) column.  phi and Y are padded (ghost plane below holds the
// previous rank's last plane in slab-parallel runs); plane 0 of the global grid must already be filled.
__global__ void __launch_bounds__(kT) k_fast_z(LevelDims L, float cell, const float* __restrict__ Y, size_t cs,
                                               float* __restrict__ phi) {
    const size_t pl = L.plane();
    const size_t col = (size_t)blockIdx.x * kT + threadIdx.x;
    if (col >= pl) return;
    const float* Yx = Y + col;
    const float* Yy = Y + cs + col;
    const float* Yz = Y + 2 * cs + col;
    float* ph = phi + col;
    int kl = 0;
    float acc, px, py, pz;
    if (L.k0 == 0) {  // global plane 0: given
        acc = ph[0];
        px = Yx[0];
        py = Yy[0];
        pz = Yz[0];
        kl = 1;
    } else {  // continue from the ghost plane below
        acc = *(ph - (ptrdiff_t)pl);
        px = *(Yx - (ptrdiff_t)pl);
        py = *(Yy - (ptrdiff_t)pl);
        pz = *(Yz - (ptrdiff_t)pl);
    }
    const int nzl = L.nzl();
#pragma unroll 4
    for (; kl < nzl; kl++) {
        const size_t o = (size_t)kl * pl;
        const float qx = Yx[o], qy = Yy[o], qz = Yz[o];
        acc += fast_step(px, py, pz, qx, qy, qz, 2, cell);
        ph[o] = acc;
        px = qx;
        py = qy;
        pz = qz;
    }
}

inline unsigned int nblk(size_t groups) {
    size_t b = (groups + kT - 1) / kT;
    return (unsigned int)std::max<size_t>(1, std::min<size_t>(b, kMaxBlocks));
}
inline bool vec4(const LevelDims& L) { return (L.nx % 4) == 0; }

}  // namespace

void set_reduction_scratch(double* partials, unsigned int* counter) {
    t_scratch.partials = partials;
    t_scratch.counter = counter;
}
void set_march_config(bool enabled, int sm_count) {
    g_march_disabled = !enabled;
    g_march_sms = sm_count > 0 ? sm_count : 148;
}
size_t reduction_scratch_doubles() { return (size_t)kMaxBlocks * 4; }

#define POST() \
    SHM3D_LAUNCHED(); \
    SHM3D_CUDA_CHECK(cudaGetLastError())

// dispatch on the vector width
#define VDISPATCH(L_, kern, ...)                                                    \
    do {                                                                            \
        if (vec4(L_)) kern<4><<<nblk((L_).n() / 4), kT, 0, s>>>(__VA_ARGS__);       \
        else kern<1><<<nblk((L_).n()), kT, 0, s>>>(__VA_ARGS__);                    \
    } while (0)

void launch_div_rhs(const LevelDims& L, float cell, const float* Y, size_t cs, float* b, int scrub,
                    unsigned int* nonfinite_count, cudaStream_t s) {
    k_div_rhs<<<nblk(L.n()), kT, 0, s>>>(L, cell, Y, cs, b, scrub, nonfinite_count);
    POST();
}
void launch_stencil_dot(const LevelDims& L, const float* p, float* q, double* acc, cudaStream_t s) {
    if (vec4(L)) k_row_stencil_dot<<<nblk_rows(L), kT, 0, s>>>(L, p, q, red_scratch(), acc);
    else k_stencil_dot<1><<<nblk(L.n()), kT, 0, s>>>(L, p, q, red_scratch(), acc);
    POST();
}
void launch_update_xr(const LevelDims& L, float* x, float* r, const float* p, const float* q, const double* rho,
                      const double* pq, double* acc_sum_r, cudaStream_t s) {
    VDISPATCH(L, k_update_xr, L, x, r, p, q, rho, pq, red_scratch(), acc_sum_r);
    POST();
}
void launch_dot_rz(const LevelDims& L, const float* r, const float* z, double* acc, cudaStream_t s) {
    VDISPATCH(L, k_dot_rz, L, r, z, red_scratch(), acc);
    POST();
}
void launch_update_p(const LevelDims& L, float* p, const float* pin, const float* z, const double* sum_z, double n_global,
                     const double* rho_new, const double* rho_old, int first, cudaStream_t s) {
    VDISPATCH(L, k_update_p, L, p, pin, z, sum_z, n_global, rho_new, rho_old, first);
    POST();
}
void launch_fill(float* p, size_t n, float v, cudaStream_t s) {
    if (!n) return;
    k_fill<<<nblk(n), kT, 0, s>>>(p, n, v);
    POST();
}
void launch_copy(float* dst, const float* src, size_t n, cudaStream_t s) {
    if (!n) return;
    k_copy<<<nblk(n), kT, 0, s>>>(dst, src, n);
    POST();
}
void launch_vec_sum(const float* v, size_t n, double* acc, cudaStream_t s) {
    k_vec_sum<<<nblk(n), kT, 0, s>>>(v, n, red_scratch(), acc);
    POST();
}
void launch_axpy_const(float* v, size_t n, const double* num, double den, float sign, cudaStream_t s) {
    k_axpy_const<<<nblk(n), kT, 0, s>>>(v, n, num, den, sign);
    POST();
}
void launch_mg_smooth0(const LevelDims& L, float* x, const float* b, const double* sum_b, double n_global, float omega,
                       cudaStream_t s) {
    VDISPATCH(L, k_mg_smooth0, L, x, b, sum_b, n_global, omega);
    POST();
}
void launch_mg_smooth(const LevelDims& L, float* xo, const float* x, const float* b, const double* sum_b,
                      double n_global, float omega, cudaStream_t s) {
    if (march_ok(L)) march_launch<OpSmooth<false>>(L, x, nullptr, b, xo, nullptr, {sum_b, n_global, omega}, RedScratch{}, nullptr, s);
    else if (vec4(L)) k_row_smooth<false><<<nblk_rows(L), kT, 0, s>>>(L, xo, x, b, sum_b, n_global, omega, RedScratch{}, nullptr);
    else k_mg_smooth<1, false><<<nblk(L.n()), kT, 0, s>>>(L, xo, x, b, sum_b, n_global, omega, RedScratch{}, nullptr);
    POST();
}
void launch_mg_smooth_dot(const LevelDims& L, float* xo, const float* x, const float* b, const double* sum_b,
                          double n_global, float omega, double* acc, cudaStream_t s) {
    if (march_ok(L)) march_launch<OpSmooth<true>>(L, x, nullptr, b, xo, nullptr, {sum_b, n_global, omega}, red_scratch(), acc, s);
    else if (vec4(L)) k_row_smooth<true><<<nblk_rows(L), kT, 0, s>>>(L, xo, x, b, sum_b, n_global, omega, red_scratch(), acc);
    else k_mg_smooth<1, true><<<nblk(L.n()), kT, 0, s>>>(L, xo, x, b, sum_b, n_global, omega, red_scratch(), acc);
    POST();
}
void launch_mg_smooth01(const LevelDims& L, float* xo, const float* b, const double* sum_b, double n_global, float omega,
                        float omega2, cudaStream_t s) {
    if (march_ok(L)) march_launch<OpSmooth01>(L, b, nullptr, nullptr, xo, nullptr, {sum_b, n_global, omega, omega2}, RedScratch{}, nullptr, s);
    else if (vec4(L)) k_row_smooth01<<<nblk_rows(L), kT, 0, s>>>(L, xo, b, sum_b, n_global, omega, omega2);
    else k_mg_smooth01<1><<<nblk(L.n()), kT, 0, s>>>(L, xo, b, sum_b, n_global, omega, omega2);
    POST();
}
void launch_update_p_stencil(const LevelDims& L, float* p_new, const float* p_old, const float* z, float* q,
                             const double* sum_z, double n_global, const double* rho_new, const double* rho_old, int first,
                             double* acc, cudaStream_t s) {
    if (march_ok(L))
        march_launch<OpUpdateP>(L, z, p_old, nullptr, p_new, q, {sum_z, rho_new, rho_old, n_global, first}, red_scratch(), acc, s);
    else if (vec4(L))
        k_row_update_p_stencil<<<nblk_rows(L), kT, 0, s>>>(L, p_new, p_old, z, q, sum_z, n_global, rho_new, rho_old, first,
                                                           red_scratch(), acc);
    else
        k_update_p_stencil<1><<<nblk(L.n()), kT, 0, s>>>(L, p_new, p_old, z, q, sum_z, n_global, rho_new, rho_old, first,
                                                         red_scratch(), acc);
    POST();
}
void launch_mg_residual(const LevelDims& L, const float* x, const float* b, const double* sum_b, double n_global,
                        float* r, cudaStream_t s) {
    if (march_ok(L)) march_launch<OpResidual>(L, x, nullptr, b, r, nullptr, {sum_b, n_global}, RedScratch{}, nullptr, s);
    else if (vec4(L)) k_row_residual<<<nblk_rows(L), kT, 0, s>>>(L, x, b, sum_b, n_global, r);
    else k_mg_residual<1><<<nblk(L.n()), kT, 0, s>>>(L, x, b, sum_b, n_global, r);
    POST();
}
void launch_mg_restrict(const LevelDims& Lf, const LevelDims& Lc, const float* r, float* bc, cudaStream_t s) {
    if (vec4(Lc) && Lf.nx == 2 * Lc.nx) k_row_restrict<<<nblk_rows(Lc), kT, 0, s>>>(Lf, Lc, r, bc);
    else k_mg_restrict<<<nblk(Lc.n()), kT, 0, s>>>(Lf, Lc, r, bc);
    POST();
}
void launch_mg_prolong_add(const LevelDims& Lf, const LevelDims& Lc, float* x, const float* ec, cudaStream_t s) {
    if (vec4(Lf) && Lf.nx == 2 * Lc.nx) k_row_prolong_add<<<nblk_rows(Lf), kT, 0, s>>>(Lf, Lc, x, ec);
    else k_mg_prolong_add<1><<<nblk(Lf.n()), kT, 0, s>>>(Lf, Lc, x, ec);
    POST();
}
void launch_fast_integrate_base(const LevelDims& L, float cell, const float* Y, size_t cs, float* phi, cudaStream_t s) {
    k_fast_base<<<1, 1024, 0, s>>>(L, cell, Y, cs, phi);
    POST();
}
void launch_fast_integrate_z(const LevelDims& L, float cell, const float* Y, size_t cs, float* phi, cudaStream_t s) {
    k_fast_z<<<(unsigned)((L.plane() + kT - 1) / kT), kT, 0, s>>>(L, cell, Y, cs, phi);
    POST();
}
void launch_cluster_program(const TailOp* d_ops, int n_ops, int ctas, float* v, const float* w, const double* shift_num,
                            double shift_den, cudaStream_t s) {
    if (n_ops <= 0) return;
    if (ctas <= 1) {
        k_cluster_program<<<1, kTailThreads, 0, s>>>(d_ops, n_ops, v, w, shift_num, shift_den);
        POST();
        return;
    }
    const int cs = tail_cluster_size(ctas);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(kTailThreads);
    cfg.dynamicSmemBytes = cs > 1 ? kTailSmemBytes : 0;
    cfg.stream = s;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = cs;
    at.val.clusterDim.y = at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    SHM3D_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_cluster_program, d_ops, n_ops, v, w, shift_num, shift_den));
    POST();
}

void launch_mg_coarse_solve(int n3, const float* pinv, const float* b, float* x, cudaStream_t s) {
    k_mg_coarse<<<1, 512, 0, s>>>(n3, pinv, b, x);
    POST();
}

}  // namespace shm3d
