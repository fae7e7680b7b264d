// solver.cu -- orchestration of the grid-solver hot path and the C ABI (include/shm3d_grid.h).
//
// Mirrors SignedHeatGridSolver::computeDistance (reference src/signed_heat_grid_solver.cpp:5-114 / :116-222):
//   Steps 1-2  k_sum (k_sum.cu)            <- :48-65
//   rhs        k_div_rhs (grid_ops.cu)      <- :70-74
//   Step 3     constrained multigrid-PCG    <- :80-108 (same KKT system, solved in the null space of A)
//   shift      k_source_average             <- :110-111, :466-496
// No CPU fallback exists: every entry point needs a CUDA device.
#include <array>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>

#include "dist.cuh"
#include "isosurface.cuh"
#include "projector.cuh"

using namespace shm3d;

namespace shm3d {

static std::string g_create_error;

struct PVec {  // float vector with one ghost plane on each side of the slab
    DevBuf<float> buf;
    size_t plane = 0, n = 0;
    void alloc(const LevelDims& L, cudaStream_t s) {
        plane = L.plane();
        n = L.n();
        size_t tot = n + 2 * plane;
        if (buf.n < tot) {
            buf.alloc(tot);
        }
        SHM3D_CUDA_CHECK(cudaMemsetAsync(buf.p, 0, tot * sizeof(float), s));
    }
    float* ip() const { return buf.p + plane; }
};

struct MGLevel {
    LevelDims L;
    double bmin[3];
    double cell;
    PVec x, tmp, r;  // iterate, second iterate buffer, residual scratch
    PVec b;          // right-hand side (level >= 1)
    std::unique_ptr<Projector> proj;
    ConstraintRows rows;
    bool replicated = false;  // slab-parallel runs: this (coarse) level is held in full by every rank
};

}  // namespace shm3d

struct shm3d_ctx {
    int device = 0;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    std::string err;
    std::unique_ptr<Dist> dist;
    // device scalars
    DevBuf<double> sc;  // see enum Sc
    DevBuf<unsigned long long> counters;
    DevBuf<unsigned int> nonfinite;
    // cached across calls (sizes only grow)
    DevBuf<float4> d_spos, d_swn, d_cbounds;
    DevBuf<int2> d_crange;
    DevBuf<double> d_pos, d_area, d_nrm;
    int sm_count = 148;
    PVec Y[1];  // component-major: 3 padded components stored back to back
    DevBuf<float> Ybuf;
    DevBuf<float> Ystage, Yrecv;  // slab contexts: Steps 1-2 of the round-robin z-chunks, before / after they move to their owners
    PVec vx, vr, vp, vp2, vq;
    double* h_rho = nullptr;  // pinned ring of the PCG's rho values (lagged convergence check)
    DevBuf<float> d_pinv;
    DevBuf<double> d_phi64, shift_part;
    DevBuf<long long> d_coinc;
    DevBuf<double> red_partials;       // deterministic reductions: one partial per CTA ...
    DevBuf<unsigned int> red_counter;  // ... and the ticket counter (grid_ops.cu), owned by this context
    DevBuf<float4> d_qpts;
    DevBuf<float> d_qY;
    std::vector<MGLevel> levels;
    DevBuf<TailOp> tail_prog;          // the V-cycle below ~64^3 as one cluster program (mg_tail.cuh)
    cudaGraphExec_t pcg_graph[2] = {nullptr, nullptr};  // one PCG iteration per buffer parity, re-used across solves
    double* d_rho_ring = nullptr;      // device alias of h_rho (mapped pinned memory)
    std::unique_ptr<IsoSurface> iso;  // row N3, created on first use
};

namespace shm3d {

constexpr int kRhoRing = 64;
constexpr size_t kReplicateBelow = (size_t)128 * 128 * 128;
enum Sc { kRho = 0, kPQ, kSumR, kRZ, kSumZ, kRhoNew, kRho0, kShiftNum, kShiftDen, kTmp, kIter, kNumSc = 16 };

// ------------------------------------------------------------------------------------------------
// small device kernels that belong to the orchestration
// ------------------------------------------------------------------------------------------------
// After the r.z / sum z reduction of iteration it = sc[kIter]:  rho_new = r.g = r.z - mean(z) * sum(r)  (r is in
// range(P), so r.(A^T lam) = 0); previous rho -> kTmp (beta denominator; +inf on the first iteration: beta = 0), and
// rho goes to slot it % kRhoRing of the host-mapped ring for the lagged convergence check.  The iteration counter lives
// on the device so that the same captured graph serves every iteration.
__global__ void k_scalars(double* sc, double n_global, double* rho_ring) {
    const double mean_z = sc[kSumZ] / n_global;
    const double rho_new = sc[kRZ] - mean_z * sc[kSumR];
    const long long it = (long long)sc[kIter];
    sc[kRhoNew] = rho_new;
    if (it == 0) sc[kRho0] = rho_new;
    sc[kTmp] = it == 0 ? (double)INFINITY : sc[kRho];
    sc[kRho] = rho_new;
    sc[kIter] = (double)(it + 1);
    rho_ring[it % kRhoRing] = rho_new;
    __threadfence_system();
}

// weighted source average of the trilinear interpolant (src/signed_heat_grid_solver.cpp:405-431, :466-496)
__global__ void k_source_average(LevelDims L, double bx, double by, double bz, double cell, int64_t M,
                                 const double* __restrict__ pos, const double* __restrict__ area,
                                 const float* __restrict__ phi, double* out /* [2]: sum A*phi, sum A */) {
    double a0 = 0, a1 = 0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < M; s += (int64_t)gridDim.x * blockDim.x) {
        double qx = pos[3 * s], qy = pos[3 * s + 1], qz = pos[3 * s + 2];
        int i = (int)floor((qx - bx) / cell), j = (int)floor((qy - by) / cell), k = (int)floor((qz - bz) / cell);
        // node position rounded like bboxMin + cell*i without contraction (reference :510-514 as the oracle evaluates it)
        double tx = (qx - __dadd_rn(bx, __dmul_rn((double)i, cell))) / cell, ty = (qy - __dadd_rn(by, __dmul_rn((double)j, cell))) / cell,
               tz = (qz - __dadd_rn(bz, __dmul_rn((double)k, cell))) / cell;
        double A = area[s];
        double v = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            int di = c & 1, dj = (c >> 1) & 1, dk = c >> 2;
            int kk = k + dk;
            if (kk < L.k0 || kk >= L.k1) continue;  // owned by another rank
            double w = (di ? tx : 1. - tx) * (dj ? ty : 1. - ty) * (dk ? tz : 1. - tz);
            v += w * (double)phi[(size_t)(i + di) + (size_t)(j + dj) * L.nx + (size_t)(kk - L.k0) * L.nx * L.ny];
        }
        a0 += A * v;
        a1 += A;
    }
    __shared__ double s0[256], s1[256];
    s0[threadIdx.x] = a0;
    s1[threadIdx.x] = a1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s0[threadIdx.x] += s0[threadIdx.x + o];
            s1[threadIdx.x] += s1[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[2 * blockIdx.x] = s0[0];
        out[2 * blockIdx.x + 1] = s1[0];
    }
}
__global__ void k_poison_nodes(int n, const long long* __restrict__ idx, float* Y, size_t comp_stride) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float q = __int_as_float(0x7fc00000);
    Y[idx[t]] = q;
    Y[idx[t] + comp_stride] = q;
    Y[idx[t] + 2 * comp_stride] = q;
}
__global__ void k_fold_pairs(const double* part, int nb, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < nb; i++) {
            a += part[2 * i];
            b += part[2 * i + 1];
        }
        out[0] = a;
        out[1] = b;
    }
}
__global__ void k_finish_phi(size_t n, const float* __restrict__ x, const double* sc, double* __restrict__ out64,
                             float* __restrict__ out32) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double shift = sc[kShiftNum] / sc[kShiftDen];
    double v = (double)x[e] - shift;
    if (out64) out64[e] = v;
    if (out32) out32[e] = (float)v;
}

// ------------------------------------------------------------------------------------------------
// host-side dense helpers for the coarsest multigrid level
// ------------------------------------------------------------------------------------------------
static void jacobi_eig(std::vector<double>& A, int n, std::vector<double>& V) {  // A symmetric -> diag; V columns
    V.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < i; j++) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
        if (off < 1e-26) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = A[(size_t)p * n + q];
                if (std::fabs(apq) < 1e-300) continue;
                double th = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2 * apq);
                double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
                double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; k++) {
                    double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
                    A[(size_t)k * n + p] = c * akp - s * akq;
                    A[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
                    A[(size_t)p * n + k] = c * apk - s * aqk;
                    A[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = c * vkp - s * vkq;
                    V[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
    }
}

// dense operator of the coarsest level: e = B b solves min 1/2 e^T K' e - b^T e s.t. A e = 0 (pseudo-inverse of P K' P)
static void coarse_operator(const LevelDims& L, const ConstraintRows& rows, std::vector<float>& B) {
    const int nx = L.nx, ny = L.ny, nz = L.nz, n = nx * ny * nz;
    std::vector<double> K((size_t)n * n, 0.0);
    auto id = [&](int i, int j, int k) { return i + j * nx + k * nx * ny; };
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                int c = id(i, j, k);
                const int nb[6][3] = {{i - 1, j, k}, {i + 1, j, k}, {i, j - 1, k}, {i, j + 1, k}, {i, j, k - 1}, {i, j, k + 1}};
                for (auto& q : nb) {
                    if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= nx || q[1] >= ny || q[2] >= nz) continue;
                    K[(size_t)c * n + c] += 1;
                    K[(size_t)c * n + id(q[0], q[1], q[2])] -= 1;
                }
            }
    // orthonormal basis Q of the constraint rows (modified Gram-Schmidt, dependent rows dropped)
    std::vector<std::vector<double>> Q;
    for (int r = 0; r < rows.m; r++) {
        std::vector<double> v(n, 0.0);
        for (int c = 0; c < 8; c++) v[rows.node[(size_t)r * 8 + c]] += rows.w[(size_t)r * 8 + c];
        for (int pass = 0; pass < 2; pass++)
            for (auto& q : Q) {
                double d = 0;
                for (int i = 0; i < n; i++) d += q[i] * v[i];
                for (int i = 0; i < n; i++) v[i] -= d * q[i];
            }
        double nn = 0;
        for (int i = 0; i < n; i++) nn += v[i] * v[i];
        if (nn < 1e-16) continue;
        nn = 1 / std::sqrt(nn);
        for (int i = 0; i < n; i++) v[i] *= nn;
        Q.push_back(std::move(v));
    }
    // M = P K P with P = I - Q^T Q
    auto applyP_cols = [&](std::vector<double>& X) {  // X <- P X (columns)
        for (auto& q : Q) {
            std::vector<double> d(n, 0.0);
            for (int i = 0; i < n; i++)
                for (int c = 0; c < n; c++) d[c] += q[i] * X[(size_t)i * n + c];
            for (int i = 0; i < n; i++)
                for (int c = 0; c < n; c++) X[(size_t)i * n + c] -= q[i] * d[c];
        }
    };
    applyP_cols(K);  // P K
    // (P K) P = (P (P K)^T)^T, K symmetric
    std::vector<double> Kt((size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Kt[(size_t)i * n + j] = K[(size_t)j * n + i];
    applyP_cols(Kt);
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double v = 0.5 * (Kt[(size_t)i * n + j] + Kt[(size_t)j * n + i]);
            Kt[(size_t)i * n + j] = Kt[(size_t)j * n + i] = v;
        }
    std::vector<double> V;
    jacobi_eig(Kt, n, V);
    double lmax = 0;
    for (int i = 0; i < n; i++) lmax = std::max(lmax, std::fabs(Kt[(size_t)i * n + i]));
    B.assign((size_t)n * n, 0.f);
    std::vector<double> Bd((size_t)n * n, 0.0);
    for (int e = 0; e < n; e++) {
        double lam = Kt[(size_t)e * n + e];
        if (lam < 1e-9 * lmax) continue;
        for (int i = 0; i < n; i++) {
            double vi = V[(size_t)i * n + e] / lam;
            for (int j = 0; j < n; j++) Bd[(size_t)i * n + j] += vi * V[(size_t)j * n + e];
        }
    }
    for (size_t i = 0; i < Bd.size(); i++) B[i] = (float)Bd[i];
}

// ------------------------------------------------------------------------------------------------
// the solver
// ------------------------------------------------------------------------------------------------
struct Timer {
    cudaEvent_t a, b;
    cudaStream_t s;
    explicit Timer(cudaStream_t st) : s(st) {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~Timer() {
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    void start() { cudaEventRecord(a, s); }
    void stop() { cudaEventRecord(b, s); }
    double ms() {
        cudaEventSynchronize(b);
        float t = 0;
        cudaEventElapsedTime(&t, a, b);
        return t;
    }
};

// Per-kernel device timing for the roofline report (SHM3D_FLAG_PROFILE): event pairs recorded on the solver's
// stream around selected launches, resolved after the stream has drained.
struct EventProfiler {
    struct Rec { cudaEvent_t a, b; int slot; };
    std::vector<Rec> recs;
    bool on = false;
    void begin(cudaStream_t s, int slot) {
        if (!on) return;
        Rec r;
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        r.slot = slot;
        cudaEventRecord(r.a, s);
        recs.push_back(r);
    }
    void end(cudaStream_t s) {
        if (on) cudaEventRecord(recs.back().b, s);
    }
    void resolve(double* ms /*[nslots]*/, int64_t* count) {
        for (Rec& r : recs) {
            cudaEventSynchronize(r.b);
            float t = 0;
            cudaEventElapsedTime(&t, r.a, r.b);
            ms[r.slot] += t;
            count[r.slot]++;
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        recs.clear();
    }
};
enum ProfSlot { kProfStencil = 0, kProfVcycle, kProfProjector, kProfUpdate, kNumProf };

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Solver {
    shm3d_ctx* c;
    const shm3d_params* p;
    GridDesc G;
    LevelDims L0;
    cudaStream_t s;
    shm3d_stats st;
    float omega = 0.8f;
    // damping of sweep k of a leg of `n` sweeps.  n == 2: the two-step Chebyshev weights for the high-frequency band
    // lambda(D^-1 K) in [1/3, 2] of the 3-D 7-point stencil (error factor 0.34 per leg instead of 0.54 with 0.8, 0.8);
    // the post-smoothing leg uses them in reverse order, which keeps the V-cycle symmetric.
    bool chebyshev = true;
    float cheb_a = 0.2f, cheb_b = 2.f;  // [1/3, 2] is the textbook smoothing band; 0.2 measured best inside the PCG (83 vs 87 its)
    float sweep_omega(int k, int n) const {
        if (!chebyshev || n < 2) return omega;
        // Chebyshev roots on [a, b]: omega_k = 1 / (c + h cos((2k+1) pi / (2n)))
        const float cc = 0.5f * (cheb_a + cheb_b), hh = 0.5f * (cheb_b - cheb_a);
        return 1.f / (cc + hh * cosf((2 * k + 1) * 3.14159265f / (2 * n)));
    }
    int nu = 2;
    bool use_mg = true;
    EventProfiler prof;
    int nu_coarse = 0;           // experiment: sweeps on the projected levels (0 = same as nu)
    int cmg_from = 2;            // first level whose smoothers are projected (finer ones: plain Poisson smoother)
    bool constrained_mg = true;  // project inside the multigrid smoothers (every level) vs. plain Poisson V-cycle

    Solver(shm3d_ctx* ctx, const shm3d_params* prm) : c(ctx), p(prm), s(ctx->stream) {
        memset(&st, 0, sizeof(st));
        if (!prm) throw Error(SHM3D_ERR_INVALID_ARG, "params is NULL");
        if (prm->nx < 4 || prm->ny < 4 || prm->nz < 4) throw Error(SHM3D_ERR_INVALID_ARG, "grid must be at least 4^3");
        if (!(prm->cell > 0) || !(prm->lambda > 0) || !std::isfinite(prm->cell) || !std::isfinite(prm->lambda))
            throw Error(SHM3D_ERR_INVALID_ARG, "cell and lambda must be positive and finite");
        G.nx = prm->nx;
        G.ny = prm->ny;
        G.nz = prm->nz;
        int k0, k1;
        slab_range(c->rank, c->world, prm->nz, k0, k1);
        G.k0 = k0;
        G.k1 = k1;
        for (int a = 0; a < 3; a++) G.bmin[a] = prm->bbox_min[a];
        G.cell = prm->cell;
        L0 = LevelDims{G.nx, G.ny, G.nz, G.k0, G.k1};
        if (L0.nzl() < 1) throw Error(SHM3D_ERR_INVALID_ARG, "more ranks than grid planes");
        // the grid kernels index a rank's nodes (plus two ghost planes) with 32-bit integers
        if ((double)L0.n() + 2.0 * (double)L0.plane() >= 4294967296.0)
            throw Error(SHM3D_ERR_INVALID_ARG, "this rank's slab has 2^32 or more nodes: split the grid over more GPUs");
        nu = prm->mg_smooth > 0 ? prm->mg_smooth : 2;
        use_mg = !(prm->flags & SHM3D_FLAG_NO_MG);
        prof.on = (prm->flags & SHM3D_FLAG_PROFILE) != 0;
        constrained_mg = !(prm->flags & SHM3D_FLAG_PLAIN_MG);
        set_march_config(!(prm->flags & SHM3D_FLAG_NO_TMA), c->sm_count);
        use_tail = (prm->flags & SHM3D_FLAG_TAIL_PROGRAM) != 0;
        set_projector_chained_launches(!(prm->flags & SHM3D_FLAG_NO_PDL));
#ifdef SHM3D_TUNING_KNOBS
        if (const char* e = getenv("SHM3D_TAIL_CTAS")) tail_ctas = atoi(e);
        if (const char* e = getenv("SHM3D_TAIL_MAX_N")) tail_max_nodes = (size_t)atoi(e) * atoi(e) * atoi(e);
#endif
        use_graph = !(prm->flags & SHM3D_FLAG_NO_GRAPH);
        cmg_from = prm->mg_constrained_from == 0 ? 2 : std::max(0, prm->mg_constrained_from);
#ifdef SHM3D_TUNING_KNOBS  // experiment overrides: compiled out of the product library (csrc/build.sh -DSHM3D_TUNING_KNOBS)
        if (const char* e = getenv("SHM3D_CMG_FROM")) cmg_from = atoi(e);
        if (const char* e = getenv("SHM3D_NU_COARSE")) nu_coarse = atoi(e);
        if (const char* e = getenv("SHM3D_CHEBYSHEV")) chebyshev = atoi(e) != 0;
        if (const char* e = getenv("SHM3D_CHEB_A")) cheb_a = (float)atof(e);
#endif
        if (c->sc.n < kNumSc) c->sc.alloc(kNumSc);
        if (c->counters.n < 4) c->counters.alloc(4);
        if (c->nonfinite.n < 1) c->nonfinite.alloc(1);
    }

    // ---------------------------------------------------------------- sources
    ClusteredSources cs;
    int64_t M = 0;
    std::vector<double> h_pos, h_nrm, h_area;  // host copies when inputs are device-resident
    const double *pos = nullptr, *nrm = nullptr, *area = nullptr;

    // inside_grid: the caller goes on to the shift / the constraints, which interpolate phi at the sources (Steps 1-2 alone
    // -- shm3d_step12 -- accept sources anywhere)
    void prepare_sources(int64_t M_, const double* pos_, const double* nrm_, const double* area_, bool on_device,
                         bool inside_grid = true) {
        M = M_;
        if (M <= 0 || !pos_ || !area_) throw Error(SHM3D_ERR_INVALID_ARG, "no sources");
        double t0 = now_ms();
        if (on_device) {
            h_pos.resize(3 * M);
            h_area.resize(M);
            SHM3D_CUDA_CHECK(cudaMemcpyAsync(h_pos.data(), pos_, 3 * M * sizeof(double), cudaMemcpyDeviceToHost, s));
            SHM3D_CUDA_CHECK(cudaMemcpyAsync(h_area.data(), area_, M * sizeof(double), cudaMemcpyDeviceToHost, s));
            if (nrm_) {
                h_nrm.resize(3 * M);
                SHM3D_CUDA_CHECK(cudaMemcpyAsync(h_nrm.data(), nrm_, 3 * M * sizeof(double), cudaMemcpyDeviceToHost, s));
            }
            SHM3D_CUDA_CHECK(cudaStreamSynchronize(s));
            pos = h_pos.data();
            area = h_area.data();
            nrm = nrm_ ? h_nrm.data() : nullptr;
            c->d_pos.alloc(3 * M);
            c->d_area.alloc(M);
            SHM3D_CUDA_CHECK(cudaMemcpyAsync(c->d_pos.p, pos_, 3 * M * sizeof(double), cudaMemcpyDeviceToDevice, s));
            SHM3D_CUDA_CHECK(cudaMemcpyAsync(c->d_area.p, area_, M * sizeof(double), cudaMemcpyDeviceToDevice, s));
        } else {
            pos = pos_;
            nrm = nrm_;
            area = area_;
            c->d_pos.upload(pos, 3 * M, s);
            c->d_area.upload(area, M, s);
        }
        // every source must lie inside the node lattice: the shift (k_source_average) and the constraint rows index the
        // eight corner nodes of its cell (the reference's own evaluateFunction / trilinearCoefficients, :405-464, assume it)
        for (int64_t i = 0; inside_grid && i < M; i++)
            for (int a = 0; a < 3; a++) {
                const int na = a == 0 ? G.nx : (a == 1 ? G.ny : G.nz);
                const double t = (pos[3 * i + a] - G.bmin[a]) / G.cell;
                if (!(t >= 0.0) || !(t < (double)(na - 1)))  // (also rejects NaN)
                    throw Error(SHM3D_ERR_INVALID_ARG, "source " + std::to_string(i) + " lies outside the grid (axis " +
                                                           std::to_string(a) + "): the box must contain every source");
            }
        st.ms_h2d += now_ms() - t0;
    }

    // Sources that coincide EXACTLY (in the reference's double arithmetic) with a grid node: the reference evaluates
    // exp(-lambda*0)/0 = Inf there, Y becomes NaN and the mesh overload later scrubs the affected
    // rhs entries.  The fp32 kernel cannot see an exact coincidence, so those nodes are found here and poisoned.
    std::vector<long long> coincident;
    void find_coincident_nodes() {
        coincident.clear();
        for (int64_t s_ = 0; s_ < M; s_++) {
            long long id[3];
            bool hit = true;
            for (int a = 0; a < 3 && hit; a++) {
                const double y = pos[3 * s_ + a];
                const double t = std::floor((y - G.bmin[a]) / G.cell + 0.5);
                const int n = a == 0 ? G.nx : (a == 1 ? G.ny : G.nz);
                // nodeIndicesToPosition (reference :510-514) evaluates bboxMin + cell*i; whether the compiler contracts that
                // into an FMA is build-dependent, so coincidence is accepted within a few ulp
                hit = t >= 0 && t < n && std::fabs((G.bmin[a] + G.cell * t) - y) <= 8.9e-16 * (std::fabs(y) + G.cell);
                id[a] = (long long)t;
            }
            if (hit && id[2] >= G.k0 && id[2] < G.k1)
                coincident.push_back(id[0] + id[1] * (long long)G.nx + (id[2] - G.k0) * (long long)G.nx * G.ny);
        }
    }

    void cluster_and_upload() {
        double t0 = now_ms();
        if (!nrm) throw Error(SHM3D_ERR_INVALID_ARG, "normals are required for Steps 1-2");
        find_coincident_nodes();
        double origin[3];
        origin[0] = G.bmin[0] + 0.5 * G.cell * (G.nx - 1);
        origin[1] = G.bmin[1] + 0.5 * G.cell * (G.ny - 1);
        origin[2] = G.bmin[2] + 0.5 * G.cell * (G.nz - 1);
        build_clusters(M, pos, nrm, area, origin, p->lambda, 4.0, cs);
        c->d_spos.upload(cs.pos, s);
        c->d_swn.upload(cs.wn, s);
        c->d_cbounds.upload(cs.bounds, s);
        c->d_crange.upload(cs.range, s);
        st.n_clusters = (int)cs.bounds.size();
        st.ms_h2d += now_ms() - t0;
    }

    // ---------------------------------------------------------------- Steps 1-2
    size_t ycomp() const { return L0.n() + 2 * L0.plane(); }  // stride between padded components
    float* Ycomp(int a) const { return c->Ybuf.p + (size_t)a * ycomp() + L0.plane(); }

    void run_step12() {
        c->Ybuf.alloc(3 * ycomp());
        SumParams P;
        P.nx = G.nx;
        P.ny = G.ny;
        P.nz = G.nz;
        P.k0 = G.k0;
        P.k1 = G.k1;
        P.ox = (float)(G.bmin[0] - cs.origin[0]);
        P.oy = (float)(G.bmin[1] - cs.origin[1]);
        P.oz = (float)(G.bmin[2] - cs.origin[2]);
        P.cell = (float)G.cell;
        P.lam2 = (float)(p->lambda * 1.4426950408889634);
        // default 10: measured culling error (tools/tau_probe2.py, profiles/experiments/r02_tau_probe_all_inputs.log)
        // max|dY| <= 2.2e-5 against the fp64 fixtures on every parity input (worst: the 1430-point bunny cloud, where
        // each cluster carries weight), phi indistinguishable from tau = inf.  phi alone would tolerate tau = 6 (<= 2.6e-5
        // everywhere, k_sum 365 instead of 509 ms at 512^3) but Y degrades to 1e-3 there; tau = 10 keeps the unit vectors
        // of Step 2 within 3e-5 of the reference's
        double tau = p->cull_tau > 0 ? p->cull_tau : 10.0;
        P.tol = std::isinf(tau) ? INFINITY : (float)(tau / p->lambda);
        P.n_clusters = (int)cs.bounds.size();
        P.uf_enable = (p->flags & SHM3D_FLAG_FP64_UNDERFLOW) ? 1 : 0;
        P.uf_thr = (float)(-537.5 + std::log2(cs.wscale));
        P.uf_log2_unscale = (float)(-std::log2(cs.wscale));
        SHM3D_CUDA_CHECK(cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long), s));
        SHM3D_CUDA_CHECK(cudaMemsetAsync(c->Ybuf.p, 0, 3 * ycomp() * sizeof(float), s));
        pending_sum_timer.reset(new Timer(s));
        pending_sum_timer->start();
        if (cyclic_step12()) {
            run_step12_cyclic(P);
        } else {
            // Y is written component-major with the padded component stride
            launch_heat_sum(P, c->d_cbounds.p, c->d_crange.p, c->d_spos.p, c->d_swn.p, Ycomp(0), ycomp(), c->counters.p, s);
        }
        if (!coincident.empty()) {
            c->d_coinc.upload(coincident, s);
            k_poison_nodes<<<(unsigned)((coincident.size() + 127) / 128), 128, 0, s>>>(
                (int)coincident.size(), c->d_coinc.p, Ycomp(0), ycomp());
            SHM3D_LAUNCHED();
        }
        pending_sum_timer->stop();
        st.pairs_bruteforce = (int64_t)L0.n() * M;
    }
    std::unique_ptr<Timer> pending_sum_timer;

    // Slab-parallel Steps 1-2, load-balanced.  The cost of a node is the number of source clusters it cannot cull, and that
    // depends on where the node sits: at 1024^3 / 8 GPUs the outer slabs take 532 ms, the inner ones 407 (max / mean 1.14,
    // profiles/experiments/r02_ksum_slab_balance.jsonl), and the solve waits for the slowest.  So the summation is done on
    // z-chunks of kChunk planes (one tile layer of k_sum) dealt round-robin to the ranks -- every rank samples the whole
    // z range -- into a staging buffer, and one grouped round of ncclSend/ncclRecv then moves each chunk to the slab that
    // owns it (12 B/node over NVLink: ~1.4 GB per rank at 1024^3 / 8, a few ms against the ~65 ms the imbalance cost).
    static constexpr int kChunk = 8;
    // who computes what, who owns what: to[r] = the chunks this rank computes that slab r owns, from[r] = the chunks of this
    // rank's slab that rank r computes -- both in increasing chunk order, which is the order of the data inside the one
    // message per peer (checked for every rank pair by tests/test_dist.py through shm3d_debug_cyclic_plan)
    static void cyclic_plan(int nz, int W, int me, std::vector<std::vector<int>>& to, std::vector<std::vector<int>>& from) {
        const int n_chunks = nz / kChunk, per_slab = n_chunks / W;
        to.assign(W, {});
        from.assign(W, {});
        for (int ch = me; ch < n_chunks; ch += W) to[ch / per_slab].push_back(ch);
        for (int ch = me * per_slab; ch < (me + 1) * per_slab; ch++) from[ch % W].push_back(ch);
    }
    bool cyclic_step12() const {
        // (2 slabs of a centred object are mirror images: nothing to balance, and half of Y would travel for nothing)
        return c->dist && c->world > 2 && !(p->flags & SHM3D_FLAG_NO_CYCLIC_SUM) && G.nz % (kChunk * c->world) == 0;
    }
    void run_step12_cyclic(SumParams P) {
        const int W = c->world, me = c->rank;
        const int n_chunks = G.nz / kChunk, per_slab = n_chunks / W;  // (slabs are uniform and chunk-aligned here)
        const size_t pl = L0.plane(), chunk_comp = (size_t)kChunk * pl, chunk_all = 3 * chunk_comp;
        // my chunks (ch = me, me + W, ...), grouped by the slab that owns them: everything bound for one peer is contiguous in
        // the staging buffer, so the round is ONE send and ONE receive per peer (a first version with a send per chunk and
        // component spent ~0.3 ms per operation inside the NCCL group)
        std::vector<std::vector<int>> to, from;
        cyclic_plan(G.nz, W, me, to, from);
        size_t n_out = 0, n_in = 0;
        for (int r = 0; r < W; r++) {
            n_out += to[r].size();
            if (r != me) n_in += from[r].size();
        }
        c->Ystage.alloc(std::max<size_t>(1, n_out) * chunk_all);
        c->Yrecv.alloc(std::max<size_t>(1, n_in) * chunk_all);
        std::vector<Dist::P2P> sends, recvs;
        size_t slot = 0;
        for (int r = 0; r < W; r++) {
            if (r != me && !to[r].empty()) sends.push_back({c->Ystage.p + slot * chunk_all, to[r].size() * chunk_all, r});
            for (int ch : to[r]) {
                P.k0 = ch * kChunk;
                P.k1 = P.k0 + kChunk;
                float* out = c->Ystage.p + slot * chunk_all;
                launch_heat_sum(P, c->d_cbounds.p, c->d_crange.p, c->d_spos.p, c->d_swn.p, out, chunk_comp, c->counters.p, s);
                if (r == me)  // my own chunk: straight into place
                    for (int a = 0; a < 3; a++)
                        SHM3D_CUDA_CHECK(cudaMemcpyAsync(Ycomp(a) + (size_t)(ch * kChunk - G.k0) * pl, out + (size_t)a * chunk_comp,
                                                         chunk_comp * sizeof(float), cudaMemcpyDeviceToDevice, s));
                slot++;
            }
        }
        slot = 0;
        std::vector<std::pair<int, size_t>> placed;  // (chunk, slot in Yrecv)
        for (int r = 0; r < W; r++) {
            if (r == me || from[r].empty()) continue;
            recvs.push_back({c->Yrecv.p + slot * chunk_all, from[r].size() * chunk_all, r});
            for (int ch : from[r]) placed.emplace_back(ch, slot++);
        }
        c->dist->p2p_round(sends, recvs, s);
        for (const auto& pc : placed)
            for (int a = 0; a < 3; a++)
                SHM3D_CUDA_CHECK(cudaMemcpyAsync(Ycomp(a) + (size_t)(pc.first * kChunk - G.k0) * pl,
                                                 c->Yrecv.p + pc.second * chunk_all + (size_t)a * chunk_comp,
                                                 chunk_comp * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }

    void finish_step12_stats() {
        if (pending_sum_timer) {
            st.ms_sum = pending_sum_timer->ms();
            pending_sum_timer.reset();
            unsigned long long np = 0;
            SHM3D_CUDA_CHECK(cudaMemcpy(&np, c->counters.p, sizeof(np), cudaMemcpyDeviceToHost));
            st.pairs_evaluated = (int64_t)np;
        }
    }

    // ---------------------------------------------------------------- rhs
    void run_rhs(float* b) {
        Timer t(s);
        t.start();
        if (c->dist) c->dist->exchange_halo3(Ycomp(0), ycomp(), L0, s);
        SHM3D_CUDA_CHECK(cudaMemsetAsync(c->nonfinite.p, 0, sizeof(unsigned int), s));
        launch_div_rhs(L0, (float)G.cell, Ycomp(0), ycomp(), b, (p->flags & SHM3D_FLAG_SCRUB_NONFINITE) ? 1 : 0,
                       c->nonfinite.p, s);
        t.stop();
        st.ms_rhs = t.ms();
        unsigned int nf = 0;
        SHM3D_CUDA_CHECK(cudaMemcpy(&nf, c->nonfinite.p, sizeof(nf), cudaMemcpyDeviceToHost));
        if (c->dist) nf = c->dist->allreduce_max_host(nf);
        if (nf && !(p->flags & SHM3D_FLAG_SCRUB_NONFINITE))
            // the point-cloud overload does not scrub (reference :180) and solveSquare's checkFinite throws
            throw Error(SHM3D_ERR_NONFINITE, "right-hand side has " + std::to_string(nf) + " non-finite entries");
    }

    // ---------------------------------------------------------------- multigrid hierarchy + constraints
    void build_levels() {
        double t0 = now_ms();
        const bool dbg = getenv("SHM3D_DEBUG") != nullptr;
        std::vector<MGLevel>& lv = c->levels;
        // level geometry (buffers of a previous call with the same shape are reused)
        std::vector<LevelDims> dims;
        std::vector<std::array<double, 4>> geo;  // bmin xyz, cell
        dims.push_back(L0);
        geo.push_back({G.bmin[0], G.bmin[1], G.bmin[2], G.cell});
        std::vector<char> repl(1, 0);
        if (use_mg) {
            // Slab-parallel runs: a level stays z-partitioned while the slab boundaries coarsen cleanly and it is big enough
            // for the split to pay; from 128^3 down every rank holds the whole level ("replicated": the restricted
            // right-hand side is all-gathered once, everything further down needs no communication).  Kernels on
            // <= 128^3 grids are launch-latency-bound on one GPU already, so replicating them costs nothing while it
            // removes ~7 halo exchanges and 5 constraint all-reduces per level and V-cycle.
            bool replicated = false;
            while (true) {
                const LevelDims Lf = dims.back();
                if ((size_t)Lf.nx * Lf.ny * Lf.nz <= 64) break;  // small enough for the dense coarse solve
                if ((Lf.nx | Lf.ny | Lf.nz) & 1) break;
                if (Lf.nx / 2 < 4 || Lf.ny / 2 < 4 || Lf.nz / 2 < 4) break;
                LevelDims Lc{Lf.nx / 2, Lf.ny / 2, Lf.nz / 2, Lf.k0 / 2, Lf.k1 / 2};
                if (c->world > 1 && !replicated) {
                    const int per = Lc.nz / c->world;
                    const bool clean = !(Lf.k0 & 1) && !(Lf.k1 & 1) && Lc.nz % c->world == 0 && Lc.k0 == c->rank * per &&
                                       Lc.k1 == (c->rank + 1) * per;
                    if (!clean) break;  // uneven slabs: no hierarchy below this level
                    if (per < 2 || (size_t)Lc.nx * Lc.ny * Lc.nz <= kReplicateBelow) replicated = true;
                }
                if (c->world == 1 && ((Lf.k0 & 1) || (Lf.k1 & 1))) break;
                if (replicated) {
                    Lc.k0 = 0;
                    Lc.k1 = Lc.nz;
                }
                dims.push_back(Lc);
                repl.push_back(replicated ? 1 : 0);
                const std::array<double, 4> gf = geo.back();
                geo.push_back({gf[0] + 0.5 * gf[3], gf[1] + 0.5 * gf[3], gf[2] + 0.5 * gf[3], 2 * gf[3]});
            }
            const LevelDims Lc = dims.back();
            if (dims.size() == 1 || (size_t)Lc.nx * Lc.ny * Lc.nz > 512 || (c->world > 1 && !repl.back())) {
                // no usable hierarchy (odd sizes, uneven slabs) -> plain projected CG, which only small grids can afford
                if (c->world > 1 && (size_t)L0.nx * L0.ny * L0.nz > (size_t)128 * 128 * 128)
                    throw Error(SHM3D_ERR_INVALID_ARG,
                                "slab-partitioned solve: nz = " + std::to_string(L0.nz) + " does not split into " +
                                    std::to_string(c->world) + " z-slabs that coarsen cleanly (need nz divisible by "
                                    "ranks * 2^levels); without the multigrid hierarchy the constrained CG would not "
                                    "converge at this size -- use a rank count that divides nz / 8");
                dims.resize(1);
                geo.resize(1);
                repl.resize(1);
                use_mg = false;
            }
        }
        bool same = lv.size() == dims.size();
        for (size_t l = 0; same && l < dims.size(); l++)
            same = lv[l].L.nx == dims[l].nx && lv[l].L.ny == dims[l].ny && lv[l].L.nz == dims[l].nz &&
                   lv[l].L.k0 == dims[l].k0 && lv[l].L.k1 == dims[l].k1 && lv[l].x.buf.p != nullptr;
        if (!same) {
            lv.clear();
            lv.resize(dims.size());
            for (size_t l = 0; l < dims.size(); l++) {
                lv[l].L = dims[l];
                lv[l].replicated = repl[l] != 0;
                lv[l].x.alloc(dims[l], s);
                lv[l].tmp.alloc(dims[l], s);
                lv[l].r.alloc(dims[l], s);
                if (l > 0) lv[l].b.alloc(dims[l], s);
            }
        }
        double t1 = now_ms();
        // constraints per level
        for (size_t l = 0; l < lv.size(); l++) {
            MGLevel& Lv = lv[l];
            for (int a = 0; a < 3; a++) Lv.bmin[a] = geo[l][a];
            Lv.cell = geo[l][3];
            Lv.rows = ConstraintRows();
            if (l > 0 && (!constrained_mg || (int)l < cmg_from)) {
                Lv.proj.reset();
                continue;
            }
            build_constraint_rows(Lv.L.nx, Lv.L.ny, Lv.L.nz, Lv.bmin, Lv.cell, M, pos, l == 0, Lv.rows);
            bool last = use_mg && (l + 1 == lv.size());
            if (last && lv.size() > 1) Lv.proj.reset();
            if (!last || lv.size() == 1) {
                if (!Lv.proj) Lv.proj.reset(new Projector());  // kept across solves: its arenas are reused
                Lv.proj->build(Lv.rows, Lv.L, /*uniform=*/l == 0, s);
                if (c->dist && !Lv.replicated) c->dist->attach(*Lv.proj);
                else Lv.proj->reduce_hook_ = nullptr;
            }
        }
        double t2 = now_ms();
        st.m_constraints = lv[0].rows.m;
        if (use_mg) {
            std::vector<float> B;
            coarse_operator(lv.back().L, lv.back().rows, B);
            c->d_pinv.upload(B, s);
        }
        st.ms_constraints = now_ms() - t0;
        if (dbg)
            fprintf(stderr, "[shm3d] build_levels: buffers %.1f ms, constraints+factor+upload %.1f ms, coarse %.1f ms\n",
                    t1 - t0, t2 - t1, now_ms() - t2);
    }

    bool dist_level(int l) const { return c->dist && !c->levels[l].replicated; }
    bool level_projected(int l) const { return constrained_mg && l >= cmg_from && c->levels[l].proj; }

    // one (projected-)Jacobi sweep: xo = x + Pi w D^-1 (b - K x).  dot_acc: also r.z and sum z (last fine sweep).
    void smooth_sweep(int l, const float* b, const double* sum_b, double n_global, float om, double* dot_acc = nullptr,
                      bool exchange = true) {
        MGLevel& Lv = c->levels[l];
        if (dot_acc) launch_mg_smooth_dot(Lv.L, Lv.tmp.ip(), Lv.x.ip(), b, sum_b, n_global, om, dot_acc, s);
        else launch_mg_smooth(Lv.L, Lv.tmp.ip(), Lv.x.ip(), b, sum_b, n_global, om, s);
        if (level_projected(l)) Lv.proj->apply_update(Lv.tmp.ip(), Lv.x.ip(), s);
        std::swap(Lv.x, Lv.tmp);
        if (exchange && dist_level(l)) c->dist->exchange_halo(Lv.x.ip(), Lv.L, s);
    }

    // V-cycle: levels[l].x = V(b - mean)   (mean only at level 0, passed as device scalar).
    // dot_acc (level 0 only, when its smoothers are unprojected): the last sweep also produces r.z and sum z.
    // b needs valid ghost planes in slab-parallel runs (the fused first sweeps read its z-neighbours).
    void vcycle(int l, const float* b, const double* sum_b, double n_global, double* dot_acc = nullptr) {
        std::vector<MGLevel>& lv = c->levels;
        MGLevel& Lv = lv[l];
        if (l == tail_level) {  // everything from here down and back up: one launch (b == Lv.b, result in Lv.x)
            launch_cluster_program(c->tail_prog.p, tail_len, tail_ctas, nullptr, nullptr, nullptr, 1.0, s);
            return;
        }
        if (l + 1 == (int)lv.size()) {
            launch_mg_coarse_solve((int)Lv.L.n(), c->d_pinv.p, b, Lv.x.ip(), s);
            return;
        }
        const bool proj = level_projected(l);
        const int nul = (l >= cmg_from && nu_coarse > 0) ? nu_coarse : nu;
        int done = 1;
        if (!proj && nul >= 2) {
            // sweeps 1 and 2 in one pass over b
            launch_mg_smooth01(Lv.L, Lv.x.ip(), b, sum_b, n_global, sweep_omega(0, nul), sweep_omega(1, nul), s);
            done = 2;
        } else {
            launch_mg_smooth0(Lv.L, Lv.x.ip(), b, sum_b, n_global, sweep_omega(0, nul), s);
            if (proj) Lv.proj->apply(Lv.x.ip(), s);
        }
        const bool dl = dist_level(l);
        if (dl) c->dist->exchange_halo(Lv.x.ip(), Lv.L, s);
        for (int k = done; k < nul; k++) smooth_sweep(l, b, sum_b, n_global, sweep_omega(k, nul));
        MGLevel& Lc = lv[l + 1];
        launch_mg_residual(Lv.L, Lv.x.ip(), b, sum_b, n_global, Lv.r.ip(), s);
        if (dl) c->dist->exchange_halo(Lv.r.ip(), Lv.L, s);
        if (dl && Lc.replicated) {
            // this rank restricts its own planes into the full coarse vector, then the slabs are all-gathered
            const LevelDims Ls{Lc.L.nx, Lc.L.ny, Lc.L.nz, Lv.L.k0 / 2, Lv.L.k1 / 2};
            launch_mg_restrict(Lv.L, Ls, Lv.r.ip(), Lc.b.ip() + (size_t)Ls.k0 * Ls.plane(), s);
            c->dist->allgather(Lc.b.ip(), Ls.n(), s);
        } else {
            launch_mg_restrict(Lv.L, Lc.L, Lv.r.ip(), Lc.b.ip(), s);
            if (dl) c->dist->exchange_halo(Lc.b.ip(), Lc.L, s);
        }
        vcycle(l + 1, Lc.b.ip(), nullptr, 1.0);
        if (dist_level(l + 1)) c->dist->exchange_halo(Lc.x.ip(), Lc.L, s);
        if (dl) {
            // the correction of the two ghost planes follows from the coarse ghost planes: no exchange afterwards
            LevelDims Le = Lv.L;
            Le.k0 = std::max(0, Lv.L.k0 - 1);
            Le.k1 = std::min(Lv.L.nz, Lv.L.k1 + 1);
            launch_mg_prolong_add(Le, Lc.L, Lv.x.ip() - (size_t)(Lv.L.k0 - Le.k0) * Lv.L.plane(), Lc.x.ip(), s);
        } else {
            launch_mg_prolong_add(Lv.L, Lc.L, Lv.x.ip(), Lc.x.ip(), s);
        }
        // (the halo of the final iterate is exchanged by whoever reads it next: the parent level before prolongating,
        // the PCG after projecting z)
        for (int k = 0; k < nul; k++)
            smooth_sweep(l, b, sum_b, n_global, sweep_omega(nul - 1 - k, nul), (k + 1 == nul) ? dot_acc : nullptr,
                         /*exchange=*/k + 1 < nul);
    }

    // ---------------------------------------------------------------- V-cycle tail as one program launch (opt-in experiment)
    // From the first level with <= 16^3 nodes down, every operation of the V-cycle -- sweeps, transfers, the dense coarsest
    // solve and each tree level of the projected smoothers' multifrontal sweeps -- is a launch-latency-bound kernel of a
    // few microseconds.  record_tail() restates vcycle() for those levels as a program of TailOp (mg_tail.cuh) that one
    // CTA executes in a single launch.  Measured on B200 at 512^3 (profiles/experiments/r02_pcg_probe_*.jsonl), against the
    // same iteration replayed from a CUDA graph (4.04 ms): 16-CTA cluster from 64^3 down 4.78 ms, one CTA from 32^3 down
    // 4.32 ms, one CTA from 16^3 down 4.08 ms -- a graph node boundary (~3 us) is cheaper than an op of a 1- or 16-SM
    // program whose loads all miss L1.  Hence SHM3D_FLAG_TAIL_PROGRAM is off by default.
    bool use_tail = false, use_graph = true;
    int tail_level = -1, tail_len = 0;
    int tail_ctas = 1;                                 // one CTA (see mg_tail.cuh for the 16-CTA cluster measurement)
    size_t tail_max_nodes = (size_t)16 * 16 * 16;      // first level the tail program takes over (one quad per thread)
    std::vector<TailOp> tail_host;

    void record_tail() {
        tail_level = -1;
        tail_len = 0;
        std::vector<MGLevel>& lv = c->levels;
        const int nl = (int)lv.size();
        if (!use_tail || !use_mg || nl < 3) return;
        int lt = -1;
        for (int l = 1; l + 1 < nl && lt < 0; l++)
            if ((size_t)lv[l].L.nx * lv[l].L.ny * lv[l].L.nz <= tail_max_nodes && (c->world == 1 || lv[l].replicated)) lt = l;
        if (lt < 0) return;
        for (int l = lt; l < nl; l++) {  // the row bodies need nx % 4 == 0 and factor-2 transfers
            if (lv[l].L.nx % 4 || lv[l].L.nzl() != lv[l].L.nz) return;
            if (l + 1 < nl && lv[l].L.nx != 2 * lv[l + 1].L.nx) return;
            if (level_projected(l) && (lv[l].proj->m() > Projector::kClusterMaxRows || lv[l].proj->reduce_hook_)) return;
        }
        std::vector<TailOp>& ops = tail_host;  // (member: stays alive until the upload has run)
        ops.clear();
        std::vector<float*> X(nl), T(nl);
        for (int l = 0; l < nl; l++) {
            X[l] = lv[l].x.ip();
            T[l] = lv[l].tmp.ip();
        }
        auto mk = [](int code, const LevelDims& L) {
            TailOp op;
            memset(&op, 0, sizeof(op));
            op.code = code;
            op.L = L;
            op.Lc = L;
            return op;
        };
        std::function<void(int)> rec = [&](int l) {
            MGLevel& Lv = lv[l];
            float* rhs = Lv.b.ip();
            if (l + 1 == nl) {
                TailOp op = mk(kTCoarse, Lv.L);
                op.h = (int)Lv.L.n();
                op.a = c->d_pinv.p;
                op.b = rhs;
                op.o = X[l];
                ops.push_back(op);
                return;
            }
            const bool proj = level_projected(l);
            const int nul = (l >= cmg_from && nu_coarse > 0) ? nu_coarse : nu;
            auto sweep = [&](float om) {  // smooth_sweep(): T = X + om D^-1 (b - K X), projected update, swap
                TailOp op = mk(kTSmooth, Lv.L);
                op.omega = om;
                op.a = rhs;
                op.b = X[l];
                op.o = T[l];
                ops.push_back(op);
                if (proj) Lv.proj->record_apply(ops, T[l], X[l], false);
                std::swap(X[l], T[l]);
            };
            TailOp op = mk(kTSmooth0, Lv.L);
            op.omega = sweep_omega(0, nul);
            op.a = rhs;
            op.o = X[l];
            ops.push_back(op);
            if (proj) Lv.proj->record_apply(ops, X[l], nullptr, false);
            for (int k = 1; k < nul; k++) sweep(sweep_omega(k, nul));
            op = mk(kTResidual, Lv.L);
            op.a = rhs;
            op.b = X[l];
            op.o = Lv.r.ip();
            ops.push_back(op);
            op = mk(kTRestrict, Lv.L);
            op.Lc = lv[l + 1].L;
            op.a = Lv.r.ip();
            op.o = lv[l + 1].b.ip();
            ops.push_back(op);
            rec(l + 1);
            op = mk(kTProlong, Lv.L);
            op.Lc = lv[l + 1].L;
            op.a = X[l + 1];
            op.o = X[l];
            ops.push_back(op);
            for (int k = 0; k < nul; k++) sweep(sweep_omega(nul - 1 - k, nul));
        };
        rec(lt);
        if (X[lt] != lv[lt].x.ip()) {  // an odd number of sweeps: the parent level reads lv[lt].x
            TailOp op = mk(kTCopy, lv[lt].L);
            op.a = X[lt];
            op.o = lv[lt].x.ip();
            ops.push_back(op);
        }
        c->tail_prog.upload(ops, s);
        tail_level = lt;
        tail_len = (int)ops.size();
        st.tail_ops = tail_len;
    }

    // ---------------------------------------------------------------- constrained PCG
    // On entry c->vr holds b' = cell^2 D^T Y.  On exit c->vx holds phi (unshifted).
    //
    // Projected PCG in the null space of A:  g = Pi Q V Q r,  r <- r - alpha Pi K p.
    // Per iteration on the fine level: V-cycle (last sweep fused with r.z / sum z), Pi z, fused [p update + K p + p.q],
    // Pi q, fused [x, r update + sum r].  Convergence is checked on the host every kCheck iterations from a pinned
    // ring of rho values, so the stream never drains inside the loop.
    void run_pcg() {
        std::vector<MGLevel>& lv = c->levels;
        Projector& P = *lv[0].proj;
        double* sc = c->sc.p;
        const double Ng = (double)G.nglobal();
        const size_t n = L0.n();
        float *x = c->vx.ip(), *r = c->vr.ip(), *q = c->vq.ip();
        float* pbuf[2] = {c->vp.ip(), c->vp2.ip()};
        // relative preconditioned residual; the unpreconditioned fallback (odd grids) needs a tighter bar for the same
        // error in phi because its residual norm under-weights the smooth error components
        // default 1e-5.  Against the fp64 fixtures (tools/tol_probe.py, profiles/experiments/r02_tol_probe_after_drift_fix.jsonl)
        // the distance of phi to the oracle at 512^3 is the same 1e-6 .. 6e-6 for every tolerance from 3e-5 down to 5e-7 and
        // 8.6e-6 at 1e-4; on knot.obj @128^3 it is ~3e-6 for every tolerance <= 1e-5.  (Round 1 had tightened this to 3e-6
        // because of "run-to-run differences of 3e-5 at 512^3": that was the constant drift removed at the end of this
        // function, not the stopping point.)
        const double tol = p->cg_rel_tol > 0 ? p->cg_rel_tol : (use_mg ? 1e-5 : 3e-7);
        const int maxit = p->cg_max_iters > 0 ? p->cg_max_iters : 2000;
        const bool verbose = (p->flags & SHM3D_FLAG_VERBOSE) != 0;
        const int kCheck = verbose ? 1 : 4;  // (the event profiler is asynchronous: it does not need per-iteration syncs)
        const bool fuse_dot = use_mg && !level_projected(0) && lv.size() > 1 && nu >= 1;
        if (!c->h_rho) {
            SHM3D_CUDA_CHECK(cudaHostAlloc((void**)&c->h_rho, kRhoRing * sizeof(double), cudaHostAllocMapped));
            SHM3D_CUDA_CHECK(cudaHostGetDevicePointer((void**)&c->d_rho_ring, c->h_rho, 0));
        }
        Timer t(s);
        t.start();
        SHM3D_CUDA_CHECK(cudaMemsetAsync(sc, 0, kNumSc * sizeof(double), s));
        SHM3D_CUDA_CHECK(cudaMemsetAsync(x, 0, n * sizeof(float), s));
        // beta = rho / rho_old is 0 on the first iteration (rho_old = +inf, k_scalars): the old p must be finite there
        SHM3D_CUDA_CHECK(cudaMemsetAsync(pbuf[0] - L0.plane(), 0, (n + 2 * L0.plane()) * sizeof(float), s));
        P.apply(r, s);  // r~ = P b
        launch_vec_sum(r, n, sc + kSumR, s);
        if (c->dist) c->dist->allreduce(sc + kSumR, 1, s);

        // One iteration = part A (z = V(r - mean r), r.z, g = P (z - mean z), rho) and part B (p, q = K p, p.q, Pi q, x/r
        // update).  The loop runs A(0), then [B(it), A(it+1)] per iteration, so the convergence check sits between an A
        // and the following B exactly as in the textbook order.  All scalars (rho, beta, alpha, the iteration counter)
        // live on the device and buffers alternate with period 2, so [B, A] is captured ONCE per parity into a CUDA graph
        // and replayed: ~90 launches per iteration cost one graph launch on the host and back-to-back nodes on the GPU.
        auto part_a = [&](bool profile) {
            float* z = lv[0].x.ip();
            if (use_mg) {
                if (c->dist) c->dist->exchange_halo(r, L0, s);
                if (profile) prof.begin(s, kProfVcycle);
                vcycle(0, r, sc + kSumR, Ng, fuse_dot ? sc + kRZ : nullptr);
                if (profile) prof.end(s);
                z = lv[0].x.ip();  // (the sweeps swap x and tmp)
                if (!fuse_dot) launch_dot_rz(L0, r, z, sc + kRZ, s);  // writes kRZ, kSumZ
            } else {
                launch_copy(z, r, n, s);  // z = r (identity preconditioner); keep r intact
                launch_dot_rz(L0, r, z, sc + kRZ, s);
            }
            if (c->dist) c->dist->allreduce(sc + kRZ, 2, s);
            // g = P (z - mean z): z <- z - A^T (A A^T)^-1 A (z - mean z); the mean itself is removed in the p update
            if (profile) prof.begin(s, kProfProjector);
            P.apply_shifted(z, sc + kSumZ, Ng, s);
            if (profile) prof.end(s);
            k_scalars<<<1, 1, 0, s>>>(sc, Ng, c->d_rho_ring);
            SHM3D_LAUNCHED();
        };
        auto part_b = [&](int it, bool profile) {
            float* z = lv[0].x.ip();
            if (c->dist) c->dist->exchange_halo(z, L0, s);
            // p = (z - mean z) + beta p ; q = K p ; p.q   (one pass; p ping-pongs between two buffers)
            float* pn = pbuf[(it + 1) & 1];
            const float* po = pbuf[it & 1];
            if (profile) prof.begin(s, kProfStencil);
            launch_update_p_stencil(L0, pn, po, z, q, sc + kSumZ, Ng, sc + kRho, sc + kTmp, 0, sc + kPQ, s);
            if (profile) prof.end(s);
            if (c->dist) {
                // the ghost planes of the new p follow from the ghost planes of z and the old p: no extra exchange
                LevelDims G1{L0.nx, L0.ny, 1, 0, 1};
                const size_t pl = L0.plane();
                if (L0.k0 > 0) launch_update_p(G1, pn - pl, po - pl, z - pl, sc + kSumZ, Ng, sc + kRho, sc + kTmp, 0, s);
                if (L0.k1 < L0.nz) launch_update_p(G1, pn + n, po + n, z + n, sc + kSumZ, Ng, sc + kRho, sc + kTmp, 0, s);
                c->dist->allreduce(sc + kPQ, 1, s);
            }
            // alpha = rho / p.q ; x += alpha p ; r -= alpha P q
            if (profile) prof.begin(s, kProfProjector);
            P.apply(q, s);
            if (profile) prof.end(s);
            if (profile) prof.begin(s, kProfUpdate);
            launch_update_xr(L0, x, r, pn, q, sc + kRho, sc + kPQ, sc + kSumR, s);
            if (profile) prof.end(s);
            if (c->dist) c->dist->allreduce(sc + kSumR, 1, s);
        };
        // profiling (bench.py's per-kernel roofline numbers): the first iterations run eagerly with CUDA events around the
        // kernels of interest; the rest of the solve replays the graphs like an unprofiled one
        const int n_eager = prof.on ? 6 : 1;  // (the first pass also runs every kernel's one-time set-up outside a capture)
        bool captured[2] = {false, false};
        int64_t graph_nodes[2] = {0, 0};

        double rho0 = 0, rho = 0;
        int it = 0, bad = 0, checked = 0;
        double rel = 1.0;
        bool stop = false;
        part_a(prof.on);  // A(0)
        for (;;) {
            if ((it + 1) % kCheck == 0 || it == 0 || it >= maxit) {
                SHM3D_CUDA_CHECK(cudaStreamSynchronize(s));
                for (; checked <= it; checked++) {
                    rho = c->h_rho[checked % kRhoRing];
                    if (checked == 0) rho0 = rho;
                    if (!(rho0 > 0) || !std::isfinite(rho)) {
                        if (rho0 == 0) { rel = 0; stop = true; break; }  // zero right-hand side: phi = 0
                        throw Error(SHM3D_ERR_NONFINITE, "constrained PCG broke down (non-finite or non-positive r.z)");
                    }
                    rel = std::sqrt(std::fabs(rho) / rho0);
                    if (verbose) fprintf(stderr, "[shm3d] pcg it %d rel %.3e\n", checked, rel);
                    if (rel < tol) stop = true;           // x of iteration `checked` was good enough; the few extra
                    if (rho <= 0 && ++bad > 2) stop = true;  // iterations already in flight only improve it
                }
                if (stop || it >= maxit) break;
            }
            const int next = it + 1;
            if (!use_graph || next <= n_eager) {
                part_b(it, prof.on);
                part_a(prof.on);
            } else {
                const int par = next & 1;
                if (!captured[par]) {
                    const int64_t l0 = g_kernel_launches;
                    cudaGraph_t g = nullptr;
                    SHM3D_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
                    try {
                        part_b(it, false);
                        part_a(false);
                    } catch (...) {
                        cudaStreamEndCapture(s, &g);
                        if (g) cudaGraphDestroy(g);
                        throw;
                    }
                    SHM3D_CUDA_CHECK(cudaStreamEndCapture(s, &g));
                    graph_nodes[par] = g_kernel_launches - l0;
                    g_kernel_launches = l0;  // recorded, not launched: counted per replay below
                    // same topology as the previous solve on this context -> update the executable graph in place
                    bool ok = false;
                    if (c->pcg_graph[par]) {
                        cudaGraphExecUpdateResultInfo info;
                        ok = cudaGraphExecUpdate(c->pcg_graph[par], g, &info) == cudaSuccess;
                        if (!ok) {
                            cudaGetLastError();
                            cudaGraphExecDestroy(c->pcg_graph[par]);
                            c->pcg_graph[par] = nullptr;
                        }
                    }
                    if (!ok) SHM3D_CUDA_CHECK(cudaGraphInstantiate(&c->pcg_graph[par], g, 0));
                    cudaGraphDestroy(g);
                    captured[par] = true;
                }
                SHM3D_CUDA_CHECK(cudaGraphLaunch(c->pcg_graph[par], s));
                g_kernel_launches += graph_nodes[par];
                st.graph_replays++;
            }
            it = next;
        }
        // x = sum alpha_k p_k leaves null(A) by fp32 rounding, and it does so along the CONSTANTS: K annihilates them, so
        // nothing in the iteration sees or corrects a drift c * 1 (|A x| ~ 1e-4 after ~100 iterations at 512^3), while
        // A (c * 1) = c because the rows of A sum to one.  Measured against the fp64 oracle at 512^3
        // (profiles/experiments/r02_err_probe_constant_offset.jsonl): phi was off by a uniform 1e-5 .. 1e-4 (whatever tau or
        // the tolerance were) and by 1e-6 .. 5e-6 once that constant was removed -- the last projection alone only pulls
        // the pinned cells back (a local correction) and thereby also hides the offset from the source-average shift.
        // So: first subtract the mean constraint violation from the whole field, then project.
        if (P.m() > 0) {
            P.violation_sum(x, sc + kTmp, s);
            launch_axpy_const(x, n, sc + kTmp, (double)P.m(), -1.f, s);
        }
        P.apply(x, s);
        t.stop();
        st.ms_pcg = t.ms();
        {
            double ms[kNumProf] = {0, 0, 0, 0};
            int64_t cnt[kNumProf] = {0, 0, 0, 0};
            prof.resolve(ms, cnt);
            st.ms_pcg_stencil = ms[kProfStencil];
            st.pcg_stencil_launches = cnt[kProfStencil];
            st.ms_pcg_vcycle = ms[kProfVcycle];
            st.pcg_vcycles = cnt[kProfVcycle];
            st.ms_pcg_projector = ms[kProfProjector];
            st.pcg_projector_applies = cnt[kProfProjector];
            st.ms_pcg_update = ms[kProfUpdate];
        }
        st.cg_iters = it;
        st.cg_rel_residual = rel;
        if (it >= maxit && rel >= tol)
            throw Error(SHM3D_ERR_NO_CONVERGENCE, "constrained PCG did not converge in " + std::to_string(maxit) +
                                                      " iterations (rel " + std::to_string(rel) + ")");
    }

    // ---------------------------------------------------------------- fastIntegration (:77-78, :224-275)
    // phi (c->vx) from Y by the reference's greedy breadth-first integration, restated as prefix sums (grid_ops.cu).
    // Slab-parallel: the z prefix sums form a chain over the ranks -- each rank receives the plane below its slab from
    // rank-1, integrates its planes, and passes its top plane on.
    void run_fast_integration() {
        Timer t(s);
        t.start();
        float* phi = c->vx.ip();
        const size_t pl = L0.plane(), n = L0.n();
        if (c->dist) c->dist->exchange_halo3(Ycomp(0), ycomp(), L0, s);
        if (L0.k0 == 0) launch_fast_integrate_base(L0, (float)G.cell, Ycomp(0), ycomp(), phi, s);
        else c->dist->recv_plane(phi - pl, pl, c->rank - 1, s);
        launch_fast_integrate_z(L0, (float)G.cell, Ycomp(0), ycomp(), phi, s);
        if (c->dist && L0.k1 < L0.nz) c->dist->send_plane(phi + n - pl, pl, c->rank + 1, s);
        t.stop();
        st.ms_pcg = t.ms();
        st.cg_iters = 0;
    }

    // ---------------------------------------------------------------- shift + output
    void run_shift_and_output(double* phi_host, float* phi_dev) {
        double* sc = c->sc.p;
        Timer t(s);
        t.start();
        const int nb = 128;
        DevBuf<double>& part = c->shift_part;
        part.alloc(2 * nb);
        k_source_average<<<nb, 256, 0, s>>>(L0, G.bmin[0], G.bmin[1], G.bmin[2], G.cell, M, c->d_pos.p, c->d_area.p,
                                            c->vx.ip(), part.p);
        SHM3D_LAUNCHED();
        k_fold_pairs<<<1, 32, 0, s>>>(part.p, nb, sc + kShiftNum);
        SHM3D_LAUNCHED();
        if (c->dist) c->dist->allreduce(sc + kShiftNum, 1, s);  // numerator only: every rank sums all areas
        const size_t n = L0.n();
        if (phi_host) c->d_phi64.alloc(n);
        k_finish_phi<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, c->vx.ip(), sc, phi_host ? c->d_phi64.p : nullptr,
                                                                 phi_dev);
        SHM3D_LAUNCHED();
        t.stop();
        st.ms_shift = t.ms();
        double h[2];
        SHM3D_CUDA_CHECK(cudaMemcpy(h, sc + kShiftNum, 2 * sizeof(double), cudaMemcpyDeviceToHost));
        st.shift = h[0] / h[1];
        if (phi_host) {
            double t0 = now_ms();
            SHM3D_CUDA_CHECK(cudaMemcpy(phi_host, c->d_phi64.p, n * sizeof(double), cudaMemcpyDeviceToHost));
            st.ms_d2h = now_ms() - t0;
        }
    }

    void alloc_pcg_vectors() {
        c->vx.alloc(L0, s);
        c->vr.alloc(L0, s);
        c->vp.alloc(L0, s);
        c->vp2.alloc(L0, s);
        c->vq.alloc(L0, s);
    }
};

}  // namespace shm3d

// ================================================================================================
// C ABI
// ================================================================================================
#define SHM3D_API_BEGIN(ctx)                                   \
    if (!(ctx)) return SHM3D_ERR_INVALID_ARG;                  \
    try {                                                      \
        SHM3D_CUDA_CHECK(cudaSetDevice((ctx)->device));        \
        set_reduction_scratch((ctx)->red_partials.p, (ctx)->red_counter.p); \
        set_march_config(true, (ctx)->sm_count);
#define SHM3D_API_END(ctx)                                     \
    }                                                          \
    catch (const shm3d::Error& e) {                            \
        (ctx)->err = e.what();                                 \
        cudaGetLastError();                                    \
        return e.code;                                         \
    }                                                          \
    catch (const std::exception& e) {                          \
        (ctx)->err = e.what();                                 \
        return SHM3D_ERR_INVALID_ARG;                          \
    }                                                          \
    return SHM3D_OK;

extern "C" {

const char* shm3d_version(void) { return "shm3d-b200 0.1 sm_100a"; }

static int create_common(shm3d_ctx** out, int device, int rank, int world, const void* nccl_id) {
    if (!out) return SHM3D_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                         "); this library has no CPU fallback";
        cudaGetLastError();
        return SHM3D_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "device ordinal out of range";
        return SHM3D_ERR_INVALID_ARG;
    }
    std::unique_ptr<shm3d_ctx> c(new shm3d_ctx());
    c->device = device;
    c->rank = rank;
    c->world = world;
    try {
        SHM3D_CUDA_CHECK(cudaSetDevice(device));
        SHM3D_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        SHM3D_CUDA_CHECK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
        set_host_ranks_hint(world);
        c->red_partials.alloc(reduction_scratch_doubles());
        c->red_counter.alloc(1);
        SHM3D_CUDA_CHECK(cudaMemset(c->red_counter.p, 0, sizeof(unsigned int)));
        if (world > 1) c->dist.reset(new Dist(rank, world, nccl_id, c->stream));
    } catch (const shm3d::Error& ex) {
        g_create_error = ex.what();
        return ex.code;
    }
    *out = c.release();
    return SHM3D_OK;
}

int shm3d_ctx_create(shm3d_ctx** out, int device) { return create_common(out, device, 0, 1, nullptr); }

int shm3d_ctx_create_dist(shm3d_ctx** out, int device, int rank, int world, const void* nccl_id) {
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_id)) {
        g_create_error = "bad rank/world/nccl_id";
        return SHM3D_ERR_INVALID_ARG;
    }
    return create_common(out, device, rank, world, nccl_id);
}

int shm3d_nccl_unique_id(void* out128) { return Dist::unique_id(out128); }

void shm3d_ctx_destroy(shm3d_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    // the captured PCG iteration holds NCCL operations of this context's communicator: the executable graphs go first
    for (cudaGraphExec_t& g : ctx->pcg_graph)
        if (g) {
            cudaGraphExecDestroy(g);
            g = nullptr;
        }
    ctx->levels.clear();
    ctx->dist.reset();
    if (ctx->h_rho) cudaFreeHost(ctx->h_rho);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void* shm3d_ctx_stream(const shm3d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

void* shm3d_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void shm3d_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

const char* shm3d_last_error(const shm3d_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int shm3d_slab_range(int32_t rank, int32_t world, int32_t nz, int32_t* k0, int32_t* k1) {
    if (!k0 || !k1 || world < 1 || rank < 0 || rank >= world || nz < 1) return SHM3D_ERR_INVALID_ARG;
    int a, b;
    slab_range(rank, world, nz, a, b);
    *k0 = a;
    *k1 = b;
    return SHM3D_OK;
}

int shm3d_slab(const shm3d_ctx* ctx, int32_t nz, int32_t* k0, int32_t* k1) {
    if (!ctx || !k0 || !k1 || nz < 1) return SHM3D_ERR_INVALID_ARG;
    int a, b;
    slab_range(ctx->rank, ctx->world, nz, a, b);
    *k0 = a;
    *k1 = b;
    return SHM3D_OK;
}

static int solve_impl(shm3d_ctx* ctx, const shm3d_params* p, int64_t M, const double* pos, const double* nrm,
                      const double* area, double* phi_host, float* phi_dev, shm3d_stats* stats, bool on_device) {
    SHM3D_API_BEGIN(ctx)
    if (!phi_host && !phi_dev) throw Error(SHM3D_ERR_INVALID_ARG, "no output buffer");
    double t0 = now_ms();
    int64_t l0 = g_kernel_launches;
    Solver S(ctx, p);
    S.prepare_sources(M, pos, nrm, area, on_device);
    S.cluster_and_upload();
    if (p->flags & SHM3D_FLAG_FAST) {
        // SignedHeat3DOptions.fastIntegration: Steps 1-2, then the greedy integration instead of the constrained solve
        ctx->vx.alloc(S.L0, ctx->stream);
        S.run_step12();
        S.finish_step12_stats();
        S.run_fast_integration();
    } else {
        S.alloc_pcg_vectors();
        S.run_step12();       // asynchronous on the GPU ...
        S.build_levels();     // ... while the host builds and factorises the constraint systems
        S.record_tail();
        S.finish_step12_stats();
        S.run_rhs(ctx->vr.ip());
        S.run_pcg();
    }
    S.run_shift_and_output(phi_host, phi_dev);
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    S.st.kernel_launches = g_kernel_launches - l0;
    S.st.ms_total = now_ms() - t0;
    if (stats) *stats = S.st;
    SHM3D_API_END(ctx)
}

int shm3d_solve(shm3d_ctx* ctx, const shm3d_params* p, int64_t M, const double* pos, const double* nrm,
                const double* area, double* phi_out, shm3d_stats* stats) {
    return solve_impl(ctx, p, M, pos, nrm, area, phi_out, nullptr, stats, false);
}

int shm3d_solve_device(shm3d_ctx* ctx, const shm3d_params* p, int64_t M, const double* d_pos, const double* d_nrm,
                       const double* d_area, float* phi_dev, shm3d_stats* stats) {
    return solve_impl(ctx, p, M, d_pos, d_nrm, d_area, nullptr, phi_dev, stats, true);
}

int shm3d_step12(shm3d_ctx* ctx, const shm3d_params* p, int64_t M, const double* pos, const double* nrm,
                 const double* area, float* Y_out, shm3d_stats* stats) {
    SHM3D_API_BEGIN(ctx)
    if (!Y_out) throw Error(SHM3D_ERR_INVALID_ARG, "Y_out is NULL");
    double t0 = now_ms();
    int64_t l0 = g_kernel_launches;
    Solver S(ctx, p);
    S.prepare_sources(M, pos, nrm, area, false, /*inside_grid=*/false);
    S.cluster_and_upload();
    S.run_step12();
    S.finish_step12_stats();
    const size_t n = S.L0.n();
    for (int a = 0; a < 3; a++)
        SHM3D_CUDA_CHECK(cudaMemcpyAsync(Y_out + (size_t)a * n, S.Ycomp(a), n * sizeof(float), cudaMemcpyDeviceToHost,
                                         ctx->stream));
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    S.st.kernel_launches = g_kernel_launches - l0;
    S.st.ms_total = now_ms() - t0;
    if (stats) *stats = S.st;
    SHM3D_API_END(ctx)
}

int shm3d_step12_points(shm3d_ctx* ctx, double lambda, int64_t M, const double* pos, const double* nrm,
                        const double* area, int64_t nq, const double* query, float* Y_out) {
    SHM3D_API_BEGIN(ctx)
    if (!pos || !nrm || !area || !query || !Y_out || M <= 0 || nq <= 0) throw Error(SHM3D_ERR_INVALID_ARG, "NULL / empty input");
    if (!(lambda > 0) || !std::isfinite(lambda)) throw Error(SHM3D_ERR_INVALID_ARG, "lambda must be positive and finite");
    if (M > 0x7fffffff) throw Error(SHM3D_ERR_INVALID_ARG, "too many sources");
    cudaStream_t s = ctx->stream;
    // origin = centre of the sources' bounding box (fp32 positions are stored relative to it)
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < M; i++)
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], pos[3 * i + a]);
            hi[a] = std::max(hi[a], pos[3 * i + a]);
        }
    const double origin[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
    ClusteredSources cs;
    build_clusters(M, pos, nrm, area, origin, lambda, 4.0, cs);  // (also validates: non-finite input -> SHM3D_ERR_NONFINITE)
    ctx->d_spos.upload(cs.pos, s);
    ctx->d_swn.upload(cs.wn, s);
    std::vector<float4> q((size_t)nq);
    for (int64_t i = 0; i < nq; i++)
        q[i] = make_float4((float)(query[3 * i] - origin[0]), (float)(query[3 * i + 1] - origin[1]),
                           (float)(query[3 * i + 2] - origin[2]), 0.f);
    ctx->d_qpts.upload(q, s);
    ctx->d_qY.alloc((size_t)3 * nq);
    launch_heat_sum_points((int)M, ctx->d_spos.p, ctx->d_swn.p, (float)(lambda * 1.4426950408889634), nq, ctx->d_qpts.p,
                           ctx->d_qY.p, s);
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(Y_out, ctx->d_qY.p, (size_t)3 * nq * sizeof(float), cudaMemcpyDeviceToHost, s));
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(s));
    SHM3D_API_END(ctx)
}

int shm3d_rhs(shm3d_ctx* ctx, const shm3d_params* p, const float* Y, float* b_out) {
    SHM3D_API_BEGIN(ctx)
    if (!Y || !b_out) throw Error(SHM3D_ERR_INVALID_ARG, "NULL buffer");
    Solver S(ctx, p);
    const size_t n = S.L0.n();
    ctx->Ybuf.alloc(3 * S.ycomp());
    SHM3D_CUDA_CHECK(cudaMemsetAsync(ctx->Ybuf.p, 0, 3 * S.ycomp() * sizeof(float), ctx->stream));
    for (int a = 0; a < 3; a++)
        SHM3D_CUDA_CHECK(cudaMemcpyAsync(S.Ycomp(a), Y + (size_t)a * n, n * sizeof(float), cudaMemcpyHostToDevice,
                                         ctx->stream));
    ctx->vr.alloc(S.L0, ctx->stream);
    S.run_rhs(ctx->vr.ip());
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(b_out, ctx->vr.ip(), n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    SHM3D_API_END(ctx)
}

int shm3d_step3(shm3d_ctx* ctx, const shm3d_params* p, int64_t M, const double* pos, const double* area,
                const float* b, double* phi_out, shm3d_stats* stats) {
    SHM3D_API_BEGIN(ctx)
    if (!b || !phi_out) throw Error(SHM3D_ERR_INVALID_ARG, "NULL buffer");
    double t0 = now_ms();
    int64_t l0 = g_kernel_launches;
    Solver S(ctx, p);
    S.prepare_sources(M, pos, nullptr, area, false);
    S.alloc_pcg_vectors();
    S.build_levels();
    S.record_tail();
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(ctx->vr.ip(), b, S.L0.n() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    S.run_pcg();
    S.run_shift_and_output(phi_out, nullptr);
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    S.st.kernel_launches = g_kernel_launches - l0;
    S.st.ms_total = now_ms() - t0;
    if (stats) *stats = S.st;
    SHM3D_API_END(ctx)
}

// the exchange plan of the balanced Steps 1-2 (host logic; tests): chunks rank `rank` sends to / receives from `peer`, in
// message order.  Returns the two counts through n_to / n_from; -1 when the grid does not qualify (nz % (8 * world) != 0).
int shm3d_debug_cyclic_plan(int32_t nz, int32_t world, int32_t rank, int32_t peer, int32_t* to_out, int32_t* n_to,
                            int32_t* from_out, int32_t* n_from) {
    if (world < 1 || rank < 0 || rank >= world || peer < 0 || peer >= world || nz < 1 || !n_to || !n_from) return SHM3D_ERR_INVALID_ARG;
    if (nz % (Solver::kChunk * world) != 0) return -1;
    std::vector<std::vector<int>> to, from;
    Solver::cyclic_plan(nz, world, rank, to, from);
    *n_to = (int32_t)to[peer].size();
    *n_from = (int32_t)from[peer].size();
    if (to_out) std::copy(to[peer].begin(), to[peer].end(), to_out);
    if (from_out) std::copy(from[peer].begin(), from[peer].end(), from_out);
    return SHM3D_OK;
}

// ---- host-logic probes (no GPU needed): used by the CPU test-suite to check the constraint assembly and the
// nested-dissection factorisation against scipy.  They run no device code and are not part of the product path.
int shm3d_debug_constraints(const shm3d_params* p, int64_t M, const double* pos, int32_t* m_out, int64_t* node_out,
                            double* w_out, int64_t* src_out, int64_t capacity_rows) {
    if (!p || !pos || !m_out) return SHM3D_ERR_INVALID_ARG;
    try {
        ConstraintRows rows;
        build_constraint_rows(p->nx, p->ny, p->nz, p->bbox_min, p->cell, M, pos, true, rows);
        *m_out = rows.m;
        if (node_out && w_out && src_out) {
            if (capacity_rows < rows.m) return SHM3D_ERR_INVALID_ARG;
            memcpy(node_out, rows.node.data(), rows.node.size() * sizeof(int64_t));
            memcpy(w_out, rows.w.data(), rows.w.size() * sizeof(double));
            memcpy(src_out, rows.src.data(), rows.src.size() * sizeof(int64_t));
        }
    } catch (const shm3d::Error& e) {
        g_create_error = e.what();
        return e.code;
    }
    return SHM3D_OK;
}

// v (m doubles, constraint-row order) <- (A D^-1 A^T)^-1 v using the host image of the GPU factor
int shm3d_debug_factor_solve(const shm3d_params* p, int64_t M, const double* pos, int32_t uniform, double* v,
                             int32_t m_expected, double* factor_megabytes, int32_t* tree_height) {
    if (!p || !pos || !v) return SHM3D_ERR_INVALID_ARG;
    try {
        ConstraintRows rows;
        build_constraint_rows(p->nx, p->ny, p->nz, p->bbox_min, p->cell, M, pos, true, rows);
        if (rows.m != m_expected) return SHM3D_ERR_INVALID_ARG;
        HostFactor hf;
        factor_constraints(rows, p->nx, p->ny, p->nz, uniform != 0, hf);
        std::vector<double> pv(rows.m);
        for (int r = 0; r < rows.m; r++) pv[hf.perm[r]] = v[r];
        hf.solve_host(pv);
        for (int r = 0; r < rows.m; r++) v[r] = pv[hf.perm[r]];
        if (factor_megabytes) *factor_megabytes = hf.mat_size * sizeof(double) / 1048576.0;
        if (tree_height) *tree_height = (int)hf.by_height.size();
    } catch (const shm3d::Error& e) {
        g_create_error = e.what();
        return e.code;
    }
    return SHM3D_OK;
}

// One stencil operation of the PCG / V-cycle on caller-supplied PADDED vectors ((k1-k0) + 2 planes of nx*ny floats each,
// ghost planes first and last), once through the row-streaming kernels and once through the TMA-staged marching
// kernels: the GPU tests require identical fields.  op: 0 update_p_stencil (in0 = z, in1 = p_old; scal = mean, beta),
// 1 smooth (in0 = x, pw = b; scal = mean, omega), 2 smooth + dots, 3 residual, 4 smooth01 (in0 = b; scal = mean, omega, omega2).
int shm3d_debug_stencil_op(shm3d_ctx* ctx, int32_t op, int32_t nx, int32_t ny, int32_t nz, int32_t k0, int32_t k1,
                           const float* in0, const float* in1, const float* pw, const double* scal, int32_t use_tma,
                           float* out0, float* out1, double* red, int32_t reps, double* ms_per_launch) {
    SHM3D_API_BEGIN(ctx)
    const LevelDims L{nx, ny, nz, k0, k1};
    if (nx < 4 || ny < 4 || k1 <= k0 || k0 < 0 || k1 > nz || !in0 || !out0 || !scal) throw Error(SHM3D_ERR_INVALID_ARG, "bad argument");
    const size_t pl = L.plane(), tot = L.n() + 2 * pl;
    cudaStream_t s = ctx->stream;
    DevBuf<float> d0(tot), d1(tot), dp(tot), o0(tot), o1(tot);
    DevBuf<double> sc(8);
    SHM3D_CUDA_CHECK(cudaMemsetAsync(o0.p, 0, tot * 4, s));
    SHM3D_CUDA_CHECK(cudaMemsetAsync(o1.p, 0, tot * 4, s));
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(d0.p, in0, tot * 4, cudaMemcpyHostToDevice, s));
    if (in1) SHM3D_CUDA_CHECK(cudaMemcpyAsync(d1.p, in1, tot * 4, cudaMemcpyHostToDevice, s));
    if (pw) SHM3D_CUDA_CHECK(cudaMemcpyAsync(dp.p, pw, tot * 4, cudaMemcpyHostToDevice, s));
    // device scalars: [0] = sum (mean * N), [1] = rho_new (= beta), [2] = rho_old (= 1), [4..5] reductions
    const double ng = (double)nx * ny * nz;
    const double h[8] = {scal[0] * ng, scal[1], 1.0, 0, 0, 0, 0, 0};
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(sc.p, h, sizeof(h), cudaMemcpyHostToDevice, s));
    set_march_config(use_tma != 0, ctx->sm_count);
    float *a = d0.p + pl, *b = d1.p + pl, *w = dp.p + pl, *x0 = o0.p + pl, *x1 = o1.p + pl;
    auto run = [&]() {
        switch (op) {
            case 0: launch_update_p_stencil(L, x0, b, a, x1, sc.p, ng, sc.p + 1, sc.p + 2, 0, sc.p + 4, s); break;
            case 1: launch_mg_smooth(L, x0, a, w, sc.p, ng, (float)scal[1], s); break;
            case 2: launch_mg_smooth_dot(L, x0, a, w, sc.p, ng, (float)scal[1], sc.p + 4, s); break;
            case 3: launch_mg_residual(L, a, w, sc.p, ng, x0, s); break;
            case 4: launch_mg_smooth01(L, x0, a, sc.p, ng, (float)scal[1], (float)scal[2], s); break;
            default: throw Error(SHM3D_ERR_INVALID_ARG, "unknown op");
        }
    };
    run();
    if (reps > 0 && ms_per_launch) {  // device time per launch (CUDA events on the solver's stream)
        Timer t(s);
        t.start();
        for (int i = 0; i < reps; i++) run();
        t.stop();
        *ms_per_launch = t.ms() / reps;
    }
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(out0, o0.p, tot * 4, cudaMemcpyDeviceToHost, s));
    if (out1) SHM3D_CUDA_CHECK(cudaMemcpyAsync(out1, o1.p, tot * 4, cudaMemcpyDeviceToHost, s));
    if (red) SHM3D_CUDA_CHECK(cudaMemcpyAsync(red, sc.p + 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(s));
    SHM3D_API_END(ctx)
}

// ---- row N3: consumer of phi on the device (isosurface.cu) -------------------------------------------------------
// The field the consumer kernels read.  Single-GPU contexts: the caller's full field (staged / narrowed if it is on the
// host).  Slab contexts: every rank passes ITS slab as device float32 (what shm3d_solve_device left in phi_dev); the slabs
// are gathered over NVLink into rank 0's staging buffer (4.3 GB at 1024^3: ~10 ms, against 8.6 GB of doubles over PCIe to
// the host in the reference flow) and rank 0 does the consumer's work; the other ranks get nullptr = nothing to do.
static const float* n3_field(shm3d_ctx* ctx, const shm3d_params* p, const void* phi, int32_t kind, const char* who) {
    if (!p || !phi) throw Error(SHM3D_ERR_INVALID_ARG, std::string(who) + ": null argument");
    if (p->nx < 2 || p->ny < 2 || p->nz < 2) throw Error(SHM3D_ERR_INVALID_ARG, std::string(who) + ": grid too small");
    if (!ctx->iso) ctx->iso.reset(new IsoSurface());
    if (ctx->world == 1) return ctx->iso->stage_field(ctx->stream, (size_t)p->nx * p->ny * p->nz, phi, kind);
    if (kind != SHM3D_FIELD_DEVICE_F32)
        throw Error(SHM3D_ERR_INVALID_ARG, std::string(who) + ": on a z-slab context the field is this rank's slab as device "
                                                                  "float32 (SHM3D_FIELD_DEVICE_F32, phi_dev of shm3d_solve_device)");
    const size_t plane = (size_t)p->nx * p->ny;
    float* full = ctx->rank == 0 ? ctx->iso->field_buffer(plane * (size_t)p->nz) : nullptr;
    ctx->dist->gather_slabs(static_cast<const float*>(phi), full, plane, p->nz, 0, ctx->stream);
    return full;
}

int shm3d_isosurface(shm3d_ctx* ctx, const shm3d_params* p, const void* phi, int32_t field_kind, float isoval,
                     const float* bound_min, const float* bound_max, uint32_t iso_flags, shm3d_iso_stats* out) {
    SHM3D_API_BEGIN(ctx)
    const int64_t launches0 = g_kernel_launches;
    const float* d_field = n3_field(ctx, p, phi, field_kind, "shm3d_isosurface");
    float bmin[3], bmax[3];
    const int n[3] = {p->nx, p->ny, p->nz};
    for (int a = 0; a < 3; a++) {  // the narrowing of bboxMin / bboxMax to glm::vec3 (src/signed_heat_grid_solver.cpp:20-24)
        bmin[a] = bound_min ? bound_min[a] : (float)p->bbox_min[a];
        bmax[a] = bound_max ? bound_max[a] : (float)(p->bbox_min[a] + p->cell * (double)(n[a] - 1));
    }
    const bool lattice = (iso_flags & SHM3D_ISO_LATTICE) != 0;
    if (!d_field) {  // slab context, rank > 0: the mesh is rank 0's
        ctx->iso->clear_result();
        SHM3D_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // (the slab has left)
        if (out) memset(out, 0, sizeof(*out));
        return SHM3D_OK;
    }
    IsoResult r = ctx->iso->extract(ctx->stream, p->nx, p->ny, p->nz, d_field, isoval, lattice ? nullptr : bmin,
                                    lattice ? nullptr : bmax);
    if (out) {
        out->n_vertices = r.n_vertices;
        out->n_triangles = r.n_triangles;
        out->ms_device = r.ms_device;
        out->gpu_launches = g_kernel_launches - launches0;
    }
    SHM3D_API_END(ctx)
}

int shm3d_isosurface_fetch(shm3d_ctx* ctx, float* vertices_out, uint32_t* triangles_out) {
    SHM3D_API_BEGIN(ctx)
    if (!ctx->iso) throw Error(SHM3D_ERR_INVALID_ARG, "shm3d_isosurface_fetch: no isosurface has been extracted on this context");
    ctx->iso->fetch(ctx->stream, vertices_out, triangles_out);
    SHM3D_API_END(ctx)
}

int shm3d_isosurface_device(shm3d_ctx* ctx, const float** d_vertices, const uint32_t** d_triangles) {
    SHM3D_API_BEGIN(ctx)
    if (!ctx->iso) throw Error(SHM3D_ERR_INVALID_ARG, "shm3d_isosurface_device: no isosurface has been extracted on this context");
    if (d_vertices) *d_vertices = ctx->iso->d_vertices();
    if (d_triangles) *d_triangles = ctx->iso->d_triangles();
    SHM3D_API_END(ctx)
}

int shm3d_slice(shm3d_ctx* ctx, const shm3d_params* p, const void* phi, int32_t field_kind, const double* origin,
                const double* du, const double* dv, int32_t nu, int32_t nv, float* out) {
    SHM3D_API_BEGIN(ctx)
    if (!origin || !du || !dv || !out) throw Error(SHM3D_ERR_INVALID_ARG, "shm3d_slice: null argument");
    const float* d_field = n3_field(ctx, p, phi, field_kind, "shm3d_slice");
    if (!d_field) return SHM3D_OK;  // slab context, rank > 0: rank 0 samples the gathered field (`out` is left untouched)
    ctx->iso->slice(ctx->stream, p->nx, p->ny, p->nz, d_field, p->bbox_min, p->cell, origin, du, dv, nu, nv, out);
    SHM3D_API_END(ctx)
}

}  // extern "C"
