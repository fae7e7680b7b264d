// grid_ops.cu -- matrix-free operators on the regular grid for Step 3 (all HBM-bandwidth-bound).
//
// Replaces the Eigen sparse machinery of the reference: gradient() / D^T Y
// (src/signed_heat_grid_solver.cpp:336-402, :70-74), laplacian() (:278-334) and the sparse LU of the KKT
// system (:101-108) -- here a constrained multigrid-preconditioned CG whose per-iteration work is the kernels
// below.  K' = -cell^2 L is the integer 7-point Neumann stencil: (K'u)_i = sum_{in-range nbr} (u_i - u_nbr).
//
// Layout: x-fastest float arrays; every vector that is read through a stencil is allocated with one ghost
// plane below and above the local z-slab, and kernels receive the pointer to the first interior plane.
// Reductions are fp64, deterministic: per-block partials, the last block to finish folds them in fixed order.
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace shm3d {

int64_t g_kernel_launches = 0;

namespace {

constexpr int kT = 256;

// ---------------------------------------------------------------- deterministic reduction helper
// scratch layout per reduction site: partials[blocks*K], counter.
struct RedScratch {
    double* partials;
    unsigned int* counter;
};

static double* g_partials = nullptr;
static unsigned int* g_counter = nullptr;
static size_t g_partials_cap = 0;

RedScratch red_scratch(size_t blocks, int K) {
    size_t need = blocks * (size_t)K;
    if (need > g_partials_cap) {
        if (g_partials) cudaFree(g_partials);
        g_partials_cap = need * 2 + 1024;
        SHM3D_CUDA_CHECK(cudaMalloc((void**)&g_partials, g_partials_cap * sizeof(double)));
    }
    if (!g_counter) {
        SHM3D_CUDA_CHECK(cudaMalloc((void**)&g_counter, sizeof(unsigned int)));
        SHM3D_CUDA_CHECK(cudaMemset(g_counter, 0, sizeof(unsigned int)));
    }
    return RedScratch{g_partials, g_counter};
}

template <int K>
__device__ __forceinline__ void block_reduce_commit(double (&v)[K], RedScratch rs, double* out, bool accumulate) {
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
                out[k] = accumulate ? out[k] + t : t;
            }
        }
        if (threadIdx.x == 0) *rs.counter = 0;
    }
}

__device__ __forceinline__ void decode(size_t e, int nx, int ny, int& i, int& j, int& kl) {
    size_t row = e / (size_t)nx;
    i = (int)(e - row * (size_t)nx);
    kl = (int)(row / (size_t)ny);
    j = (int)(row - (size_t)kl * ny);
}

// K'u at one node, given the centre value.  u is an interior pointer (ghost planes addressable).
__device__ __forceinline__ float stencil_at(const float* __restrict__ u, size_t idx, float c, int i, int j, int k,
                                            const LevelDims& L) {
    const size_t pl = (size_t)L.nx * L.ny;
    float s = 0.f;
    int cnt = 0;
    if (i > 0) { s += u[idx - 1]; cnt++; }
    if (i < L.nx - 1) { s += u[idx + 1]; cnt++; }
    if (j > 0) { s += u[idx - L.nx]; cnt++; }
    if (j < L.ny - 1) { s += u[idx + L.nx]; cnt++; }
    if (k > 0) { s += u[idx - pl]; cnt++; }
    if (k < L.nz - 1) { s += u[idx + pl]; cnt++; }
    return (float)cnt * c - s;
}
__device__ __forceinline__ int diag_at(int i, int j, int k, const LevelDims& L) {
    return (i > 0) + (i < L.nx - 1) + (j > 0) + (j < L.ny - 1) + (k > 0) + (k < L.nz - 1);
}

// ---------------------------------------------------------------- b = cell * D'^T Y
__global__ void __launch_bounds__(kT) k_div_rhs(LevelDims L, float cell, const float* __restrict__ Y, size_t cs,
                                                float* __restrict__ b, int scrub, unsigned int* nonfinite) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e >= L.n()) return;
    int i, j, kl;
    decode(e, L.nx, L.ny, i, j, kl);
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

// ---------------------------------------------------------------- q = K'p, acc = sum p q
__global__ void __launch_bounds__(kT) k_stencil_dot(LevelDims L, const float* __restrict__ p, float* __restrict__ q,
                                                    RedScratch rs, double* out) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    double v[1] = {0.0};
    if (e < L.n()) {
        int i, j, kl;
        decode(e, L.nx, L.ny, i, j, kl);
        float c = p[e];
        float r = stencil_at(p, e, c, i, j, L.k0 + kl, L);
        q[e] = r;
        v[0] = (double)c * (double)r;
    }
    block_reduce_commit<1>(v, rs, out, false);
}

// ---------------------------------------------------------------- x += a p, r -= a q, acc = sum r
__global__ void __launch_bounds__(kT) k_update_xr(size_t n, float* __restrict__ x, float* __restrict__ r,
                                                  const float* __restrict__ p, const float* __restrict__ q,
                                                  const double* rho, const double* pq, RedScratch rs, double* out) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    const float a = (float)(*rho / *pq);
    double v[1] = {0.0};
    if (e < n) {
        x[e] = fmaf(a, p[e], x[e]);
        float rr = fmaf(-a, q[e], r[e]);
        r[e] = rr;
        v[0] = rr;
    }
    block_reduce_commit<1>(v, rs, out, false);
}

__global__ void __launch_bounds__(kT) k_dot_rz(size_t n, const float* __restrict__ r, const float* __restrict__ z,
                                               RedScratch rs, double* out) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    double v[2] = {0.0, 0.0};
    if (e < n) {
        float zz = z[e];
        v[0] = (double)r[e] * (double)zz;
        v[1] = zz;
    }
    block_reduce_commit<2>(v, rs, out, false);
}

__global__ void __launch_bounds__(kT) k_update_p(size_t n, float* __restrict__ p, const float* __restrict__ z,
                                                 const double* sum_z, double n_global, const double* rho_new,
                                                 const double* rho_old, int first) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e >= n) return;
    const float mean = (float)(*sum_z / n_global);
    float g = z[e] - mean;
    if (first) {
        p[e] = g;
    } else {
        const float beta = (float)(*rho_new / *rho_old);
        p[e] = fmaf(beta, p[e], g);
    }
}

__global__ void __launch_bounds__(kT) k_fill(float* p, size_t n, float v) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e < n) p[e] = v;
}
__global__ void __launch_bounds__(kT) k_copy(float* d, const float* s, size_t n) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e < n) d[e] = s[e];
}
__global__ void __launch_bounds__(kT) k_vec_sum(const float* v, size_t n, RedScratch rs, double* out) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    double a[1] = {e < n ? (double)v[e] : 0.0};
    block_reduce_commit<1>(a, rs, out, false);
}
__global__ void __launch_bounds__(kT) k_axpy_const(float* v, size_t n, const double* num, double den, float sign) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e < n) v[e] += sign * (float)(*num / den);
}

// ---------------------------------------------------------------- multigrid
__global__ void __launch_bounds__(kT) k_mg_smooth0(LevelDims L, float* __restrict__ x, const float* __restrict__ b,
                                                   const double* sum_b, double n_global, float omega) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e >= L.n()) return;
    int i, j, kl;
    decode(e, L.nx, L.ny, i, j, kl);
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    x[e] = omega * (b[e] - shift) / (float)diag_at(i, j, L.k0 + kl, L);
}

__global__ void __launch_bounds__(kT) k_mg_smooth(LevelDims L, float* __restrict__ xo, const float* __restrict__ x,
                                                  const float* __restrict__ b, const double* sum_b, double n_global,
                                                  float omega) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e >= L.n()) return;
    int i, j, kl;
    decode(e, L.nx, L.ny, i, j, kl);
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    const int k = L.k0 + kl;
    float c = x[e];
    float Kx = stencil_at(x, e, c, i, j, k, L);
    xo[e] = c + omega * ((b[e] - shift) - Kx) / (float)diag_at(i, j, k, L);
}

__global__ void __launch_bounds__(kT) k_mg_residual(LevelDims L, const float* __restrict__ x,
                                                    const float* __restrict__ b, const double* sum_b, double n_global,
                                                    float* __restrict__ r) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e >= L.n()) return;
    int i, j, kl;
    decode(e, L.nx, L.ny, i, j, kl);
    const float shift = sum_b ? (float)(*sum_b / n_global) : 0.f;
    float c = x[e];
    r[e] = (b[e] - shift) - stencil_at(x, e, c, i, j, L.k0 + kl, L);
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
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e >= Lc.n()) return;
    int I, J, Kl;
    decode(e, Lc.nx, Lc.ny, I, J, Kl);
    const int K = Lc.k0 + Kl;
    float wx[4], wy[4], wz[4];
    rweights(I, Lc.nx, wx);
    rweights(J, Lc.ny, wy);
    rweights(K, Lc.nz, wz);
    const size_t plf = Lf.plane();
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (wz[c] == 0.f) continue;
        const int kf = 2 * K - 1 + c - Lf.k0;  // local fine plane (may be a ghost plane: -1 or nzl)
#pragma unroll
        for (int bq = 0; bq < 4; bq++) {
            if (wy[bq] == 0.f) continue;
            const int jf = 2 * J - 1 + bq;
            const float* row = r + (ptrdiff_t)kf * (ptrdiff_t)plf + (size_t)jf * Lf.nx;
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

__global__ void __launch_bounds__(kT) k_mg_prolong_add(LevelDims Lf, LevelDims Lc, float* __restrict__ x,
                                                       const float* __restrict__ ec) {
    size_t e = (size_t)blockIdx.x * kT + threadIdx.x;
    if (e >= Lf.n()) return;
    int i, j, kl;
    decode(e, Lf.nx, Lf.ny, i, j, kl);
    const int k = Lf.k0 + kl;
    const int I0 = i >> 1, J0 = j >> 1, K0 = k >> 1;
    const int I1 = min(max((i & 1) ? I0 + 1 : I0 - 1, 0), Lc.nx - 1);
    const int J1 = min(max((j & 1) ? J0 + 1 : J0 - 1, 0), Lc.ny - 1);
    const int K1 = min(max((k & 1) ? K0 + 1 : K0 - 1, 0), Lc.nz - 1);
    const ptrdiff_t plc = (ptrdiff_t)Lc.plane();
    const float* a0 = ec + (ptrdiff_t)(K0 - Lc.k0) * plc;
    const float* a1 = ec + (ptrdiff_t)(K1 - Lc.k0) * plc;  // may be a ghost plane
    auto at = [&](const float* pl, int J, int I) { return pl[(size_t)J * Lc.nx + I]; };
    float v0 = 0.75f * (0.75f * at(a0, J0, I0) + 0.25f * at(a0, J0, I1)) +
               0.25f * (0.75f * at(a0, J1, I0) + 0.25f * at(a0, J1, I1));
    float v1 = 0.75f * (0.75f * at(a1, J0, I0) + 0.25f * at(a1, J0, I1)) +
               0.25f * (0.75f * at(a1, J1, I0) + 0.25f * at(a1, J1, I1));
    x[e] += 0.75f * v0 + 0.25f * v1;
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

inline unsigned int nblk(size_t n) { return (unsigned int)((n + kT - 1) / kT); }

}  // namespace

#define POST() \
    SHM3D_LAUNCHED(); \
    SHM3D_CUDA_CHECK(cudaGetLastError())

void launch_div_rhs(const LevelDims& L, float cell, const float* Y, size_t cs, float* b, int scrub,
                    unsigned int* nonfinite_count, cudaStream_t s) {
    k_div_rhs<<<nblk(L.n()), kT, 0, s>>>(L, cell, Y, cs, b, scrub, nonfinite_count);
    POST();
}
void launch_stencil_dot(const LevelDims& L, const float* p, float* q, double* acc, cudaStream_t s) {
    unsigned int nb = nblk(L.n());
    k_stencil_dot<<<nb, kT, 0, s>>>(L, p, q, red_scratch(nb, 1), acc);
    POST();
}
void launch_update_xr(const LevelDims& L, float* x, float* r, const float* p, const float* q, const double* rho,
                      const double* pq, double* acc_sum_r, cudaStream_t s) {
    unsigned int nb = nblk(L.n());
    k_update_xr<<<nb, kT, 0, s>>>(L.n(), x, r, p, q, rho, pq, red_scratch(nb, 1), acc_sum_r);
    POST();
}
void launch_dot_rz(const LevelDims& L, const float* r, const float* z, double* acc, cudaStream_t s) {
    unsigned int nb = nblk(L.n());
    k_dot_rz<<<nb, kT, 0, s>>>(L.n(), r, z, red_scratch(nb, 2), acc);
    POST();
}
void launch_update_p(const LevelDims& L, float* p, const float* z, const double* sum_z, double n_global,
                     const double* rho_new, const double* rho_old, int first, cudaStream_t s) {
    k_update_p<<<nblk(L.n()), kT, 0, s>>>(L.n(), p, z, sum_z, n_global, rho_new, rho_old, first);
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
    unsigned int nb = nblk(n);
    k_vec_sum<<<nb, kT, 0, s>>>(v, n, red_scratch(nb, 1), acc);
    POST();
}
void launch_axpy_const(float* v, size_t n, const double* num, double den, float sign, cudaStream_t s) {
    k_axpy_const<<<nblk(n), kT, 0, s>>>(v, n, num, den, sign);
    POST();
}
void launch_mg_smooth0(const LevelDims& L, float* x, const float* b, const double* sum_b, double n_global, float omega,
                       cudaStream_t s) {
    k_mg_smooth0<<<nblk(L.n()), kT, 0, s>>>(L, x, b, sum_b, n_global, omega);
    POST();
}
void launch_mg_smooth(const LevelDims& L, float* xo, const float* x, const float* b, const double* sum_b,
                      double n_global, float omega, cudaStream_t s) {
    k_mg_smooth<<<nblk(L.n()), kT, 0, s>>>(L, xo, x, b, sum_b, n_global, omega);
    POST();
}
void launch_mg_residual(const LevelDims& L, const float* x, const float* b, const double* sum_b, double n_global,
                        float* r, cudaStream_t s) {
    k_mg_residual<<<nblk(L.n()), kT, 0, s>>>(L, x, b, sum_b, n_global, r);
    POST();
}
void launch_mg_restrict(const LevelDims& Lf, const LevelDims& Lc, const float* r, float* bc, cudaStream_t s) {
    k_mg_restrict<<<nblk(Lc.n()), kT, 0, s>>>(Lf, Lc, r, bc);
    POST();
}
void launch_mg_prolong_add(const LevelDims& Lf, const LevelDims& Lc, float* x, const float* ec, cudaStream_t s) {
    k_mg_prolong_add<<<nblk(Lf.n()), kT, 0, s>>>(Lf, Lc, x, ec);
    POST();
}
void launch_mg_coarse_solve(int n3, const float* pinv, const float* b, float* x, cudaStream_t s) {
    k_mg_coarse<<<1, 512, 0, s>>>(n3, pinv, b, x);
    POST();
}

}  // namespace shm3d
