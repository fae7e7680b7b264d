// kernels.cuh -- launch wrappers of the hand-written sm_100a kernels.
#pragma once
#include "common.cuh"

namespace shm3d {

// ---------------------------------------------------------------- Steps 1-2 (k_sum.cu)
struct SumParams {
    int nx, ny, nz;  // global node counts
    int k0, k1;      // local z range
    float ox, oy, oz;  // position of global node (0,0,0) relative to the source origin
    float cell;
    float lam2;  // lambda * log2(e)
    float tol;   // tau / lambda (distance units); +inf = keep everything
    int n_clusters;
    // SHM3D_FLAG_FP64_UNDERFLOW: where the TRUE sum X is so small that the reference's X.norm() loses precision or is
    // zero in double (max|X| < 2^-511 / 2^-537.5), the normalisation is evaluated like the reference's (k_sum.cu,
    // normalise_node).  The kernel holds X~ = wscale * 2^(lam2*m) * X: log2 max|X| = log2 max|X~| - lam2*m - log2(wscale),
    // uf_thr = -537.5 + log2(wscale), uf_log2_unscale = -log2(wscale); uf_enable = 0 skips it.
    int uf_enable;
    float uf_thr;
    float uf_log2_unscale;
};
// Y: component-major, component a of local node idx at Y[a*ystride + idx]
void launch_heat_sum(const SumParams& P, const float4* cl_bounds, const int2* cl_range, const float4* src_pos,
                     const float4* src_wn, float* Y, size_t ystride, unsigned long long* pair_counter,
                     cudaStream_t stream);

// Steps 1-2 at arbitrary query points (xyz relative to the source origin, w unused): Y interleaved float[n_q][3]
void launch_heat_sum_points(int n_src, const float4* src_pos, const float4* src_wn, float lam2, long long n_q,
                            const float4* qpts, float* Y, cudaStream_t stream);

// ---------------------------------------------------------------- grid operators (grid_ops.cu)
// A "level" is a cell-centred box grid of nx*ny*nzl local nodes (z-slab [k0,k1) of nz planes).
struct LevelDims {
    int nx, ny, nz;  // global
    int k0, k1;      // local slab
    __host__ __device__ int nzl() const { return k1 - k0; }
    __host__ __device__ size_t plane() const { return (size_t)nx * ny; }
    __host__ __device__ size_t n() const { return plane() * (size_t)(k1 - k0); }
};

// Vectors that take part in stencils carry one ghost plane below and one above the slab
// (filled by the halo exchange in multi-GPU runs, ignored at the physical boundary).
// A "padded" pointer p addresses local node (i,j,kl) at p[plane + i + j*nx + kl*plane].

// b = cell * D'^T Y  (= cell^2 * D^T Y), reference src/signed_heat_grid_solver.cpp:336-402 and :70-74.
// Y is component-major float[3][n + 2 planes] padded per component; nonfinite_count counts scrubbed entries.
void launch_div_rhs(const LevelDims& L, float cell, const float* Y_padded, size_t comp_stride, float* b, int scrub,
                    unsigned int* nonfinite_count, cudaStream_t s);

// q = K' p (integer 7-point Neumann stencil), and acc[0] += sum p*q (fp64).  p padded, q padded.
void launch_stencil_dot(const LevelDims& L, const float* p_padded, float* q_padded, double* acc, cudaStream_t s);

// x += alpha p ; r -= alpha q ; acc[0] += sum r (after update).  alpha = rho/pq read from device scalars.
void launch_update_xr(const LevelDims& L, float* x, float* r_padded, const float* p_padded, const float* q_padded,
                      const double* rho, const double* pq, double* acc_sum_r, cudaStream_t s);

// acc[0] += sum r*z ; acc[1] += sum z
void launch_dot_rz(const LevelDims& L, const float* r_padded, const float* z_padded, double* acc, cudaStream_t s);

// p = (z - mean_z) + beta p_in, with mean_z = sums[1]/N and beta = rho_new/rho_old from device scalars (p may alias p_in)
void launch_update_p(const LevelDims& L, float* p, const float* p_in, const float* z, const double* sum_z, double n_global,
                     const double* rho_new, const double* rho_old, int first, cudaStream_t s);

// fastIntegration (reference integrateGreedily, src/signed_heat_grid_solver.cpp:224-275) as prefix sums; Y, phi padded.
// base: global plane k = 0 (rank owning it only); z: this rank's slab, continuing from the ghost plane below.
void launch_fast_integrate_base(const LevelDims& L, float cell, const float* Y_padded, size_t comp_stride, float* phi_padded,
                                cudaStream_t s);
void launch_fast_integrate_z(const LevelDims& L, float cell, const float* Y_padded, size_t comp_stride, float* phi_padded,
                             cudaStream_t s);

// deterministic-reduction scratch of the calling context (bound per host thread at every API entry)
void set_reduction_scratch(double* partials, unsigned int* counter);
size_t reduction_scratch_doubles();
// TMA-staged marching stencil kernels (grid_march.cuh): on/off and the SM count of the calling context's device
void set_march_config(bool enabled, int sm_count);

// generic helpers
void launch_fill(float* p, size_t n, float v, cudaStream_t s);
void launch_copy(float* dst, const float* src, size_t n, cudaStream_t s);
void launch_vec_sum(const float* v, size_t n, double* acc, cudaStream_t s);
void launch_axpy_const(float* v, size_t n, const double* num, double den, float sign, cudaStream_t s);  // v += sign*num/den

// ---- multigrid (cell-centred, trilinear transfers, damped Jacobi)
// x = omega * (b - shift) / diag             (first sweep from a zero guess; shift = *mean or 0)
void launch_mg_smooth0(const LevelDims& L, float* x_padded, const float* b_padded, const double* sum_b, double n_global,
                       float omega, cudaStream_t s);
// xout = x + omega * ((b - shift) - K' x) / diag
void launch_mg_smooth(const LevelDims& L, float* xout_padded, const float* x_padded, const float* b_padded,
                      const double* sum_b, double n_global, float omega, cudaStream_t s);
// same, and acc[0] = sum b*xout, acc[1] = sum xout (fused r.z / sum z of the PCG on the last fine sweep)
void launch_mg_smooth_dot(const LevelDims& L, float* xout_padded, const float* x_padded, const float* b_padded,
                          const double* sum_b, double n_global, float omega, double* acc, cudaStream_t s);
// the first two sweeps from a zero guess in one pass over b (b must have valid ghost planes in slab-parallel runs)
void launch_mg_smooth01(const LevelDims& L, float* xout_padded, const float* b_padded, const double* sum_b,
                        double n_global, float omega, float omega2, cudaStream_t s);
// p_new = (z - mean_z) + beta p_old ; q = K' p_new ; acc[0] = sum p_new q  (p_new != p_old: neighbours are recomputed)
void launch_update_p_stencil(const LevelDims& L, float* p_new_padded, const float* p_old_padded, const float* z_padded,
                             float* q_padded, const double* sum_z, double n_global, const double* rho_new,
                             const double* rho_old, int first, double* acc, cudaStream_t s);
// bc = 0.5 * P^T (b - shift - K' x)   (P = cell-centred trilinear prolongation, clamped at the boundary)
// r = (b - shift) - K' x   (r needs ghost planes before the restriction in slab-parallel runs)
void launch_mg_residual(const LevelDims& L, const float* x_padded, const float* b, const double* sum_b, double n_global,
                        float* r_padded, cudaStream_t s);
void launch_mg_restrict(const LevelDims& Lf, const LevelDims& Lc, const float* r_padded, float* bc, cudaStream_t s);
// x += P ec
void launch_mg_prolong_add(const LevelDims& Lf, const LevelDims& Lc, float* x_padded, const float* ec_padded,
                           cudaStream_t s);
// coarsest level: x = Kc^+ b via a precomputed dense pseudo-inverse (n^3 <= 512 unknowns)
void launch_mg_coarse_solve(int n3, const float* pinv, const float* b, float* x, cudaStream_t s);

// ---------------------------------------------------------------- cluster programs (mg_tail.cuh)
// A sequence of latency-bound multigrid / projector operations executed by ONE launch (one CTA, or one thread-block
// cluster), with a CTA / cluster barrier between consecutive ops.  Vector operands are interior pointers of padded level
// vectors; kTailSlotV / kTailSlotW stand for the v / w arguments of the launch.
struct ProjDev;  // proj_dev.cuh
enum TailCode : int {
    kTSmooth0 = 0,  // o = omega a / d                    (level L)
    kTSmooth,       // o = b + omega (a - K'b) / d        (level L)
    kTResidual,     // o = a - K'b                        (level L)
    kTRestrict,     // o (level Lc) = 0.5 P^T a (level L)
    kTProlong,      // o (level L) += P a (level Lc)
    kTCoarse,       // o = pinv(a) b, dense h x h
    kTCopy,         // o = a                              (level L)
    kTGather,       // proj: rhs = A (a - b - shift); h != 0: shift = the launch's *shift_num / shift_den
    kTFwd,          // proj: forward sweep of tree height h
    kTBwd,          // proj: backward sweep of tree height h
    kTScatter       // proj: o -= D^-1 A^T sol
};
struct TailOp {
    int code;
    int h;
    float omega;
    int reserved;
    LevelDims L, Lc;
    const float* a;
    const float* b;
    float* o;
    const ProjDev* proj;
};
#define kTailSlotV (reinterpret_cast<const float*>(uintptr_t(8)))
#define kTailSlotW (reinterpret_cast<const float*>(uintptr_t(16)))
// ctas = 1: one CTA, __syncthreads() between ops (what the solver uses); > 1: a thread-block cluster of up to 16 CTAs
void launch_cluster_program(const TailOp* d_ops, int n_ops, int ctas, float* v, const float* w, const double* shift_num,
                            double shift_den, cudaStream_t s);

}  // namespace shm3d
