// projector.cu -- constraint rows, nested-dissection multifrontal Cholesky of A D^-1 A^T (host, fp64) and the
// device-side projector application.  See projector.cuh for the role of each piece.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>

#include <cooperative_groups.h>

#include "proj_dev.cuh"

namespace shm3d {

// ================================================================================================
// constraint rows (host)
// ================================================================================================
void build_constraint_rows(int nx, int ny, int nz, const double bmin[3], double cell, int64_t M, const double* pos,
                           bool strict, ConstraintRows& out) {
    out = ConstraintRows();
    std::unordered_map<int64_t, int> used;
    used.reserve((size_t)std::min<int64_t>(M, 1 << 22));
    for (int64_t s = 0; s < M; s++) {
        // trilinearCoefficients, src/signed_heat_grid_solver.cpp:433-464
        double d[3] = {pos[3 * s] - bmin[0], pos[3 * s + 1] - bmin[1], pos[3 * s + 2] - bmin[2]};
        double fi = std::floor(d[0] / cell), fj = std::floor(d[1] / cell), fk = std::floor(d[2] / cell);
        if (!(fi >= 0 && fj >= 0 && fk >= 0 && fi < nx - 1 && fj < ny - 1 && fk < nz - 1)) {
            if (strict) throw Error(SHM3D_ERR_INVALID_ARG, "source " + std::to_string(s) + " lies outside the grid");
            continue;
        }
        int i = (int)fi, j = (int)fj, k = (int)fk;
        int64_t cid = (int64_t)i + (int64_t)j * nx + (int64_t)k * nx * ny;
        if (!used.emplace(cid, out.m).second) continue;  // hasCellBeenUsed (:93)
        double tx = (pos[3 * s] - (bmin[0] + i * cell)) / cell;
        double ty = (pos[3 * s + 1] - (bmin[1] + j * cell)) / cell;
        double tz = (pos[3 * s + 2] - (bmin[2] + k * cell)) / cell;
        static const int cx[8] = {0, 1, 0, 0, 1, 1, 0, 1}, cy[8] = {0, 0, 1, 0, 1, 0, 1, 1},
                         cz[8] = {0, 0, 0, 1, 0, 1, 1, 1};
        for (int c = 0; c < 8; c++) {
            out.node.push_back(cid + cx[c] + (int64_t)cy[c] * nx + (int64_t)cz[c] * nx * ny);
            out.w.push_back((cx[c] ? tx : 1. - tx) * (cy[c] ? ty : 1. - ty) * (cz[c] ? tz : 1. - tz));
        }
        out.cell.push_back(i);
        out.cell.push_back(j);
        out.cell.push_back(k);
        out.src.push_back(s);
        out.m++;
    }
}

// ================================================================================================
// nested dissection + multifrontal Cholesky (host)
// ================================================================================================
namespace {

int g_ranks_on_node = 1;  // set through set_host_ranks_hint() before the pool is first used

// Small blocking thread pool for the host factorisation.  (OpenMP's spin-waiting workers made the many short
// parallel regions of the multifrontal sweep several times SLOWER on cgroup-limited and 128-thread hosts.)
class Pool {
  public:
    static Pool& get() {
        static Pool p;
        return p;
    }
    int size() const { return (int)workers_.size() + 1; }
    // run fn(i) for i in [0,n), dynamically scheduled over the pool + the calling thread
    template <typename F>
    void parallel_for(int n, const F& fn) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty() || busy_) {
            for (int i = 0; i < n; i++) fn(i);
            return;
        }
        busy_ = true;
        std::function<void(int)> f = fn;
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &f;
            n_ = n;
            next_.store(0);
            active_ = (int)workers_.size();
            gen_++;
        }
        cv_.notify_all();
        for (int i; (i = next_.fetch_add(1)) < n;) fn(i);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return active_ == 0; });
        job_ = nullptr;
        busy_ = false;
    }

  private:
    Pool() {
        // all hardware threads the ranks of this node can share without oversubscribing (a slab-parallel run has
        // `world` processes factorising at the same time), at most 16
        unsigned hw = std::thread::hardware_concurrency();
        int nt = (int)std::max(2u, std::min(16u, hw / (unsigned)std::max(1, g_ranks_on_node)));
        if (const char* e = getenv("SHM3D_HOST_THREADS")) nt = std::max(1, atoi(e));
        for (int i = 1; i < nt; i++) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void loop() {
        unsigned long seen = 0;
        for (;;) {
            std::function<void(int)>* job;
            int n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                job = job_;
                n = n_;
            }
            if (job)
                for (int i; (i = next_.fetch_add(1)) < n;) (*job)(i);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--active_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::function<void(int)>* job_ = nullptr;
    std::atomic<int> next_{0};
    int n_ = 0, active_ = 0;
    unsigned long gen_ = 0;
    bool stop_ = false;
    bool busy_ = false;  // nested use runs serially
};

double wall() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct TreeNode {
    int s0 = 0, s1 = 0;
    int left = -1, right = -1;
    int height = 0;
    std::vector<int> B;  // sorted permuted indices > s1
    long long fwd = 0, bwd = 0, bidx = 0;
};

struct ND {
    const int* cell;
    int leaf;
    std::vector<int> order;  // permuted index -> row
    std::vector<TreeNode> nodes;

    int build(std::vector<int>& rows) {
        const int n = (int)rows.size();
        TreeNode t;
        if (n <= leaf) {
            t.s0 = (int)order.size();
            order.insert(order.end(), rows.begin(), rows.end());
            t.s1 = (int)order.size();
            nodes.push_back(std::move(t));
            return (int)nodes.size() - 1;
        }
        int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-(1 << 30), -(1 << 30), -(1 << 30)};
        for (int r : rows)
            for (int a = 0; a < 3; a++) {
                lo[a] = std::min(lo[a], cell[3 * r + a]);
                hi[a] = std::max(hi[a], cell[3 * r + a]);
            }
        int ax = 0;
        for (int a = 1; a < 3; a++)
            if (hi[a] - lo[a] > hi[ax] - lo[ax]) ax = a;
        std::vector<int> coords(n);
        for (int i = 0; i < n; i++) coords[i] = cell[3 * rows[i] + ax];
        std::nth_element(coords.begin(), coords.begin() + n / 2, coords.end());
        const int c = coords[n / 2];
        std::vector<int> L, R, S;
        for (int r : rows) {
            int v = cell[3 * r + ax];
            (v < c ? L : (v > c ? R : S)).push_back(r);
        }
        std::vector<int>().swap(rows);
        int li = -1, ri = -1, h = 0;
        if (!L.empty()) { li = build(L); h = std::max(h, nodes[li].height + 1); }
        if (!R.empty()) { ri = build(R); h = std::max(h, nodes[ri].height + 1); }
        t.left = li;
        t.right = ri;
        t.height = h;
        t.s0 = (int)order.size();
        order.insert(order.end(), S.begin(), S.end());
        t.s1 = (int)order.size();
        nodes.push_back(std::move(t));
        return (int)nodes.size() - 1;
    }
};

}  // namespace
// AVX2 inner kernels (host_blas.cpp)
void host_transpose_panel(const double* F, int f, int k0, int nb, int k1, double* Bt, int ldb);
void host_syrk_rows(double* F, int f, int k0, int nb, int k1, int i0, int i1, const double* Bt, int ldb);
void host_tri_inverse_rows(const double* F, int f, int s, double* W, int c0, int c1);
namespace {

// blocked partial Cholesky of the leading s columns of the f x f symmetric matrix F (row-major, lower part used)
bool partial_cholesky(double* F, int f, int s, bool par) {
    const int NB = 48;
    std::vector<double> Bt;
    for (int k0 = 0; k0 < s; k0 += NB) {
        const int k1 = std::min(s, k0 + NB), nb = k1 - k0;
        // diagonal block
        for (int k = k0; k < k1; k++) {
            double d = F[(size_t)k * f + k];
            for (int t = k0; t < k; t++) d -= F[(size_t)k * f + t] * F[(size_t)k * f + t];
            if (!(d > 0.0)) return false;
            d = std::sqrt(d);
            F[(size_t)k * f + k] = d;
            for (int i = k + 1; i < k1; i++) {
                double v = F[(size_t)i * f + k];
                for (int t = k0; t < k; t++) v -= F[(size_t)i * f + t] * F[(size_t)k * f + t];
                F[(size_t)i * f + k] = v / d;
            }
        }
        // panel rows below: row_i[k0:k1] <- row_i[k0:k1] * Lkk^-T
        auto panel_rows = [&](int i0, int i1) {
            for (int i = i0; i < i1; i++) {
                double* ri = F + (size_t)i * f;
                for (int k = k0; k < k1; k++) {
                    double v = ri[k];
                    const double* rk = F + (size_t)k * f;
                    for (int t = k0; t < k; t++) v -= ri[t] * rk[t];
                    ri[k] = v / rk[k];
                }
            }
        };
        // trailing update (lower triangle): F[i][j] -= <row_i[k0:k1], row_j[k0:k1]> as C -= A * Bt (host_blas.cpp)
        const int rows = f - k1;
        if (rows <= 0) continue;
        const int ldb = (rows + 7) & ~7;
        auto trailing = [&](int i0, int i1) { host_syrk_rows(F, f, k0, nb, k1, i0, i1, Bt.data(), ldb); };
        if (par && rows > 256) {
            const int chunk = 16, nch = (rows + chunk - 1) / chunk;
            Pool::get().parallel_for(nch, [&](int c) { panel_rows(k1 + c * chunk, std::min(f, k1 + (c + 1) * chunk)); });
            Bt.resize((size_t)nb * ldb);
            host_transpose_panel(F, f, k0, nb, k1, Bt.data(), ldb);
            Pool::get().parallel_for(nch, [&](int c) { trailing(k1 + c * chunk, std::min(f, k1 + (c + 1) * chunk)); });
        } else {
            panel_rows(k1, f);
            Bt.resize((size_t)nb * ldb);
            host_transpose_panel(F, f, k0, nb, k1, Bt.data(), ldb);
            trailing(k1, f);
        }
    }
    return true;
}

}  // namespace

void set_host_ranks_hint(int ranks_on_node) { g_ranks_on_node = std::max(1, ranks_on_node); }

// programmatic dependent launch of the sweep kernels on/off for the calling host thread (SHM3D_FLAG_NO_PDL)
void set_projector_chained_launches(bool enabled);

// ================================================================================================
// device kernels
// ================================================================================================
namespace {

// The four kernels form a chain of ~2*height + 2 dependent launches of a few microseconds each.  They are launched with
// programmatic dependent launch (launch_chained below): every kernel lets its successor's CTAs be scheduled right away
// (griddepcontrol.launch_dependents) and then waits for its predecessor's writes (griddepcontrol.wait) before touching
// memory, so the launch latency of step k+1 hides behind step k instead of adding to it.
__device__ __forceinline__ void chain_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__global__ void k_proj_gather(ProjDev A, const float* v, const float* w, const double* shift_num, double shift_den) {
    chain_prologue();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.m) return;
    proj_gather_row(A, r, v, w, shift_num ? *shift_num / shift_den : 0.0);
}

// one warp per matrix row of the supernodes at one tree height (proj_dev.cuh)
__global__ void k_proj_fwd(ProjDev A, int n_rows, const int* row_node, const int* row_local) {
    chain_prologue();
    int R = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (R < n_rows) proj_fwd_row<32>(A, row_node, row_local, R, threadIdx.x & 31);
}

__global__ void k_proj_bwd(ProjDev A, int n_rows, const int* row_node, const int* row_local) {
    chain_prologue();
    int R = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (R < n_rows) proj_bwd_row<32>(A, row_node, row_local, R, threadIdx.x & 31);
}

__global__ void k_proj_scatter(ProjDev A, float* v) {
    chain_prologue();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < A.n_touched) proj_scatter_node(A, t, v);
}

// launch with the programmatic-stream-serialization attribute (the kernel must start with chain_prologue())
static thread_local bool g_chained_launches = true;
template <typename... KArgs, typename... Args>
void launch_chained(void (*kern)(KArgs...), unsigned grid, unsigned block, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = s;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = g_chained_launches ? 1 : 0;
    SHM3D_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
    SHM3D_LAUNCHED();
}

}  // namespace

void set_projector_chained_launches(bool enabled) { g_chained_launches = enabled; }

// ================================================================================================
// Projector
// ================================================================================================
Projector::~Projector() {
    if (h_arena_) cudaFreeHost(h_arena_);
    if (h_mat_) cudaFreeHost(h_mat_);
}

void factor_constraints(const ConstraintRows& rows, int nx_, int ny_, int nz_, bool uniform, HostFactor& out) {
    {
        std::function<double*(size_t)> keep = std::move(out.mat_alloc);
        out = HostFactor();
        out.mat_alloc = std::move(keep);
    }
    out.m = rows.m;
    if (rows.m == 0) return;
    const int m = rows.m;
    const int64_t nx = nx_, ny = ny_, nz = nz_;
    const int64_t pl = nx * ny;

    // ---- node diagonal (number of in-range neighbours) and the "all interior" property
    auto node_d = [&](int64_t n) -> double {
        int64_t k = n / pl, rem = n - k * pl, j = rem / nx, i = rem - j * nx;
        return (double)((i > 0) + (i < nx - 1) + (j > 0) + (j < ny - 1) + (k > 0) + (k < nz - 1));
    };
    std::vector<double>& dinv = out.dinv;
    dinv.resize((size_t)m * 8);
    for (size_t e = 0; e < (size_t)m * 8; e++) {
        double d = node_d(rows.node[e]);
        if (d != 6.0) out.all_interior = false;
        dinv[e] = uniform ? 1.0 : 1.0 / d;
    }

    // ---- adjacency through the cell lattice
    std::unordered_map<int64_t, int> cellrow;
    cellrow.reserve((size_t)m * 2);
    for (int r = 0; r < m; r++)
        cellrow[(int64_t)rows.cell[3 * r] + (int64_t)rows.cell[3 * r + 1] * nx + (int64_t)rows.cell[3 * r + 2] * pl] = r;
    std::vector<int> nb_ptr(m + 1, 0), nb;
    nb.reserve((size_t)m * 12);
    for (int r = 0; r < m; r++) {
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int ci = rows.cell[3 * r] + dx, cj = rows.cell[3 * r + 1] + dy, ck = rows.cell[3 * r + 2] + dz;
                    if (ci < 0 || cj < 0 || ck < 0 || ci >= nx || cj >= ny || ck >= nz) continue;
                    auto it = cellrow.find((int64_t)ci + (int64_t)cj * nx + (int64_t)ck * pl);
                    if (it != cellrow.end()) nb.push_back(it->second);
                }
        nb_ptr[r + 1] = (int)nb.size();
    }
    auto entry = [&](int r, int q) -> double {  // (A D^-1 A^T)_{rq}
        double a = 0;
        for (int c = 0; c < 8; c++)
            for (int e = 0; e < 8; e++)
                if (rows.node[(size_t)r * 8 + c] == rows.node[(size_t)q * 8 + e])
                    a += rows.w[(size_t)r * 8 + c] * rows.w[(size_t)q * 8 + e] * dinv[(size_t)r * 8 + c];
        return a;
    };

    const bool dbg = getenv("SHM3D_DEBUG") != nullptr;
    double tdbg = wall();
    // ---- nested dissection ordering
    ND nd;
    nd.cell = rows.cell.data();
    // leaves of ~128 rows: two fewer tree levels (= 4 fewer launches per application) than with 40, for +8 % host time
    nd.leaf = 128;
#ifdef SHM3D_TUNING_KNOBS
    if (const char* e = getenv("SHM3D_ND_LEAF")) nd.leaf = std::max(8, atoi(e));
#endif
    nd.order.reserve(m);
    {
        std::vector<int> all(m);
        for (int r = 0; r < m; r++) all[r] = r;
        nd.build(all);
    }
    std::vector<TreeNode>& T = nd.nodes;
    const int nT = (int)T.size();
    std::vector<int>& perm_ = out.perm;
    perm_.assign(m, 0);
    for (int p = 0; p < m; p++) perm_[nd.order[p]] = p;

    // ---- symbolic: boundary sets (children precede parents in T)
    for (int t = 0; t < nT; t++) {
        TreeNode& n = T[t];
        std::vector<int>& B = n.B;
        for (int p = n.s0; p < n.s1; p++) {
            int r = nd.order[p];
            for (int e = nb_ptr[r]; e < nb_ptr[r + 1]; e++) {
                int q = perm_[nb[e]];
                if (q >= n.s1) B.push_back(q);
            }
        }
        for (int ch : {n.left, n.right})
            if (ch >= 0)
                for (int q : T[ch].B)
                    if (q >= n.s1) B.push_back(q);
        std::sort(B.begin(), B.end());
        B.erase(std::unique(B.begin(), B.end()), B.end());
    }

    if (dbg) {
        fprintf(stderr, "[shm3d] ND: m=%d nodes=%d adjacency+order+symbolic %.3fs\n", m, nT, wall() - tdbg);
        tdbg = wall();
        std::vector<std::pair<long long, int>> big;
        for (int t = 0; t < nT; t++) big.emplace_back((long long)(T[t].s1 - T[t].s0 + (int)T[t].B.size()), t);
        std::sort(big.rbegin(), big.rend());
        for (int i = 0; i < std::min(8, nT); i++) {
            const TreeNode& n = T[big[i].second];
            fprintf(stderr, "[shm3d]   front %d: s=%d b=%d height=%d\n", big[i].second, n.s1 - n.s0, (int)n.B.size(), n.height);
        }
    }
    // ---- storage layout
    long long mat_total = 0, bidx_total = 0;
    int max_h = 0;
    for (int t = 0; t < nT; t++) {
        TreeNode& n = T[t];
        long long s = n.s1 - n.s0, b = (long long)n.B.size(), f = s + b;
        n.fwd = mat_total;
        mat_total += f * s;
        n.bwd = mat_total;
        mat_total += s * f;
        n.bidx = bidx_total;
        bidx_total += b;
        max_h = std::max(max_h, n.height);
    }
    if (out.mat_alloc) {
        out.mat = out.mat_alloc((size_t)mat_total);
    } else {
        out.mat_own.resize((size_t)mat_total);
        out.mat = out.mat_own.data();
    }
    out.mat_size = (size_t)mat_total;
    double* const mat = out.mat;  // every block is zero-filled by the task that builds it
    std::vector<int>& bidx = out.bidx;
    bidx.resize((size_t)bidx_total);
    for (int t = 0; t < nT; t++) std::copy(T[t].B.begin(), T[t].B.end(), bidx.begin() + T[t].bidx);

    // ---- numeric multifrontal factorisation, level by level (nodes of equal height are independent)
    std::vector<std::vector<int>>& by_h = out.by_height;
    by_h.assign(max_h + 1, {});
    for (int t = 0; t < nT; t++) by_h[T[t].height].push_back(t);
    std::vector<std::vector<double>> U(nT);  // update matrices (b x b, lower used)
    std::mutex prof_mu;
    double prof_t[6] = {0, 0, 0, 0, 0, 0};
    const int nthreads = Pool::get().size();
    bool ok = true;
    for (int h = 0; h <= max_h && ok; h++) {
        const std::vector<int>& lv = by_h[h];
        const bool outer = (int)lv.size() >= 2 * nthreads;
        double tl = wall();
        auto do_node = [&](int li) {
            if (!ok) return;
            const int t = lv[li];
            TreeNode& n = T[t];
            const int s = n.s1 - n.s0, b = (int)n.B.size(), f = s + b;
            const double tp0 = wall();
            std::vector<double> F((size_t)f * f, 0.0);
            auto loc = [&](int q) -> int {
                if (q < n.s1) return q - n.s0;
                return s + (int)(std::lower_bound(n.B.begin(), n.B.end(), q) - n.B.begin());
            };
            // original entries of the separator rows
            for (int p = n.s0; p < n.s1; p++) {
                int r = nd.order[p];
                for (int e = nb_ptr[r]; e < nb_ptr[r + 1]; e++) {
                    int q = perm_[nb[e]];
                    if (q < p) continue;
                    F[(size_t)loc(q) * f + (p - n.s0)] += entry(r, nb[e]);
                }
            }
            // extend-add the children's update matrices
            for (int ch : {n.left, n.right}) {
                if (ch < 0) continue;
                const std::vector<int>& CB = T[ch].B;
                const int cb = (int)CB.size();
                std::vector<int> map(cb);
                for (int i = 0; i < cb; i++) map[i] = loc(CB[i]);
                const std::vector<double>& Uc = U[ch];
                for (int i = 0; i < cb; i++)
                    for (int j = 0; j <= i; j++) F[(size_t)map[i] * f + map[j]] += Uc[(size_t)i * cb + j];
                std::vector<double>().swap(U[ch]);
            }
            const double tp1 = wall();
            if (!partial_cholesky(F.data(), f, s, !outer)) {
                ok = false;
                return;
            }
            const double tp2 = wall();
            // update matrix for the parent
            if (b > 0) {
                U[t].resize((size_t)b * b);
                for (int i = 0; i < b; i++)
                    for (int j = 0; j <= i; j++) U[t][(size_t)i * b + j] = F[(size_t)(s + i) * f + (s + j)];
            }
            // W = L11^-1 (lower), G = L21 W ; FWD = [W; G] (f x s), BWD = [W^T | -G^T] (s x f)
            double* FW = mat + n.fwd;
            double* BW = mat + n.bwd;
            std::fill(FW, FW + (size_t)f * s, 0.0);
            // W = L11^-1 (host_tri_inverse_rows, by rows in axpy form); G = L21 W
            auto g_row = [&](int i) {  // G[i, c] = sum_{k>=c} L21[i,k] W[k,c]
                const double* Li = F.data() + (size_t)(s + i) * f;
                double* Gi = FW + (size_t)(s + i) * s;
                for (int k = 0; k < s; k++) {
                    const double l = Li[k];
                    const double* Wk = FW + (size_t)k * s;
                    for (int c = 0; c <= k; c++) Gi[c] += l * Wk[c];
                }
            };
            const double tp3 = wall();
            if (!outer && s > 256) {
                const int nblk = std::min(nthreads * 2, (s + 63) / 64);  // independent column ranges
                Pool::get().parallel_for(nblk, [&](int q) {
                    host_tri_inverse_rows(F.data(), f, s, FW, (int)((long long)s * q / nblk), (int)((long long)s * (q + 1) / nblk));
                });
            } else {
                host_tri_inverse_rows(F.data(), f, s, FW, 0, s);
            }
            const double tp4 = wall();
            if (!outer && b > 256) Pool::get().parallel_for(b, g_row);
            else for (int i = 0; i < b; i++) g_row(i);
            const double tp5 = wall();
            for (int r = 0; r < s; r++) {
                for (int c = 0; c < s; c++) BW[(size_t)r * f + c] = FW[(size_t)c * s + r];
                for (int c = 0; c < b; c++) BW[(size_t)r * f + s + c] = -FW[(size_t)(s + c) * s + r];
            }
            if (dbg) {
                const double tp6 = wall();
                std::lock_guard<std::mutex> lk(prof_mu);
                prof_t[0] += tp1 - tp0; prof_t[1] += tp2 - tp1; prof_t[2] += tp3 - tp2; prof_t[3] += tp4 - tp3;
                prof_t[4] += tp5 - tp4; prof_t[5] += tp6 - tp5;
            }
        };
        if (outer) Pool::get().parallel_for((int)lv.size(), do_node);
        else for (int li = 0; li < (int)lv.size(); li++) do_node(li);
        if (dbg) fprintf(stderr, "[shm3d]   height %d: %d nodes outer=%d %.4fs\n", h, (int)lv.size(), (int)outer, wall() - tl);
    }
    (void)0;
    if (dbg) {
        fprintf(stderr, "[shm3d] numeric factorisation %.3fs, %.1f MB\n", wall() - tdbg, out.mat_size * 8e-6);
        fprintf(stderr, "[shm3d]   thread-seconds: assemble %.3f  cholesky %.3f  update-matrix %.3f  W %.3f  G %.3f  transpose %.3f\n",
                prof_t[0], prof_t[1], prof_t[2], prof_t[3], prof_t[4], prof_t[5]);
    }
    if (!ok)
        throw Error(SHM3D_ERR_FACTORIZATION,
                    "constraint system A A^T is not positive definite (coincident / dependent source constraints)");
    out.nodes.resize(nT);
    for (int t = 0; t < nT; t++) {
        out.nodes[t].s0 = T[t].s0;
        out.nodes[t].s = T[t].s1 - T[t].s0;
        out.nodes[t].b = (int)T[t].B.size();
        out.nodes[t].fwd = T[t].fwd;
        out.nodes[t].bwd = T[t].bwd;
        out.nodes[t].bidx = T[t].bidx;
    }
}

// Host mirror of k_proj_fwd / k_proj_bwd (tests only).
void HostFactor::solve_host(std::vector<double>& v) const {
    if (!m) return;
    std::vector<double> y(m, 0.0), x(m, 0.0);
    for (size_t h = 0; h < by_height.size(); h++)
        for (int t : by_height[h]) {
            const ProjNodeDesc& nd = nodes[t];
            const int f = nd.s + nd.b;
            std::vector<double> val(f, 0.0);
            for (int r = 0; r < f; r++) {
                const double* row = mat + nd.fwd + (long long)r * nd.s;
                double acc = 0;
                for (int c = 0; c < nd.s; c++) acc += row[c] * v[nd.s0 + c];
                val[r] = acc;
            }
            for (int r = 0; r < nd.s; r++) y[nd.s0 + r] = val[r];
            for (int r = 0; r < nd.b; r++) v[bidx[nd.bidx + r]] -= val[nd.s + r];
        }
    for (int h = (int)by_height.size() - 1; h >= 0; h--)
        for (int t : by_height[h]) {
            const ProjNodeDesc& nd = nodes[t];
            const int f = nd.s + nd.b;
            for (int r = 0; r < nd.s; r++) {
                const double* row = mat + nd.bwd + (long long)r * f;
                double acc = 0;
                for (int c = 0; c < nd.s; c++) acc += row[c] * y[nd.s0 + c];
                for (int c = 0; c < nd.b; c++) acc += row[nd.s + c] * x[bidx[nd.bidx + c]];
                x[nd.s0 + r] = acc;
            }
        }
    v = x;
}

namespace {
// bump allocator over a byte arena (256-byte aligned sub-buffers)
struct Bump {
    size_t off = 0;
    template <typename T>
    size_t take(size_t n) {
        size_t o = off;
        off += (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255;
        return o;
    }
};
}  // namespace

void Projector::build(const ConstraintRows& rows, const LevelDims& L, bool uniform, cudaStream_t stream) {
    m_ = rows.m;
    L_ = L;
    if (m_ == 0) return;
    const int m = m_;
    HostFactor hf;
    hf.mat_alloc = [this](size_t n) -> double* {
        if (n > h_mat_cap_) {
            if (h_mat_) cudaFreeHost(h_mat_);
            h_mat_ = nullptr;
            h_mat_cap_ = n + n / 4;
            SHM3D_CUDA_CHECK(cudaHostAlloc((void**)&h_mat_, h_mat_cap_ * sizeof(double), cudaHostAllocDefault));
        }
        return h_mat_;
    };
    const double tb0 = wall();
    factor_constraints(rows, L.nx, L.ny, L.nz, uniform, hf);
    const double tb1 = wall();
    all_interior_ = hf.all_interior;
    perm_ = hf.perm;
    factor_bytes_ = hf.mat_size * sizeof(double);
    const std::vector<ProjNodeDesc>& descs = hf.nodes;
    const std::vector<std::vector<int>>& by_h = hf.by_height;
    const int max_h = (int)by_h.size() - 1;
    const std::vector<double>& dinv = hf.dinv;
    const int64_t pl = (int64_t)L.nx * L.ny;

    // ---- the factor blocks go up first (the bulk of the bytes), straight from the page-locked buffer
    d_matbuf_.alloc(hf.mat_size * sizeof(double));
    d_mat_ = (double*)d_matbuf_.p;
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(d_mat_, hf.mat, hf.mat_size * sizeof(double), cudaMemcpyHostToDevice, stream));

    // ---- sizes of everything else
    size_t n_fwd_rows = 0, n_bwd_rows = 0;
    for (const ProjNodeDesc& d : descs) {
        n_fwd_rows += (size_t)d.s + d.b;
        n_bwd_rows += (size_t)d.s;
    }
    const size_t n_rowmaps = 2 * (n_fwd_rows + n_bwd_rows);
    // node-centric transpose over the nodes this rank owns
    const int64_t lo = (int64_t)L.k0 * pl, hi = (int64_t)L.k1 * pl;
    std::vector<std::pair<int64_t, int>> ent;  // (local node, e)
    ent.reserve((size_t)m * 8);
    for (size_t e = 0; e < (size_t)m * 8; e++) {
        int64_t n = rows.node[e];
        if (n >= lo && n < hi) ent.emplace_back(n - lo, (int)e);
    }
    std::sort(ent.begin(), ent.end());
    size_t n_touched = 0;
    for (size_t a = 0; a < ent.size(); a++)
        if (a == 0 || ent[a].first != ent[a - 1].first) n_touched++;
    n_touched_ = (int)n_touched;

    Bump B;
    const size_t o_rnode = B.take<int64_t>((size_t)m * 8), o_rw = B.take<double>((size_t)m * 8),
                 o_rperm = B.take<int>(m), o_tnode = B.take<int64_t>(n_touched), o_tptr = B.take<int>(n_touched + 1),
                 o_trow = B.take<int>(ent.size()), o_tw = B.take<double>(ent.size()),
                 o_nodes = B.take<ProjNodeDesc>(descs.size()), o_bidx = B.take<int>(hf.bidx.size()),
                 o_rowmaps = B.take<int>(n_rowmaps), o_levels = B.take<ProjLevelInfo>(max_h + 1),
                 o_self = B.take<ProjDev>(1);
    const size_t upload_bytes = B.off;
    const size_t o_rhs = B.take<double>(m), o_y = B.take<double>(m), o_sol = B.take<double>(m);
    if (upload_bytes > h_arena_cap_) {
        if (h_arena_) cudaFreeHost(h_arena_);
        h_arena_ = nullptr;
        h_arena_cap_ = upload_bytes + upload_bytes / 4;
        SHM3D_CUDA_CHECK(cudaHostAlloc((void**)&h_arena_, h_arena_cap_, cudaHostAllocDefault));
    }
    d_arena_.alloc(B.off + B.off / 4);
    unsigned char* H = h_arena_;
    unsigned char* D = d_arena_.p;

    // ---- fill the staging arena
    int64_t* h_rnode = (int64_t*)(H + o_rnode);
    for (size_t e = 0; e < (size_t)m * 8; e++) {
        int64_t n = rows.node[e];
        h_rnode[e] = (n >= lo && n < hi) ? n - lo : -1;
    }
    memcpy(H + o_rw, rows.w.data(), (size_t)m * 8 * sizeof(double));
    memcpy(H + o_rperm, perm_.data(), (size_t)m * sizeof(int));
    {
        int64_t* tnode = (int64_t*)(H + o_tnode);
        int* tptr = (int*)(H + o_tptr);
        int* trow = (int*)(H + o_trow);
        double* tw = (double*)(H + o_tw);
        size_t t = 0;
        for (size_t a = 0; a < ent.size(); a++) {
            if (a == 0 || ent[a].first != ent[a - 1].first) {
                tnode[t] = ent[a].first;
                tptr[t] = (int)a;
                t++;
            }
            const int e = ent[a].second;
            trow[a] = perm_[e / 8];
            tw[a] = rows.w[e] * dinv[e];
        }
        tptr[t] = (int)ent.size();
    }
    memcpy(H + o_nodes, descs.data(), descs.size() * sizeof(ProjNodeDesc));
    if (!hf.bidx.empty()) memcpy(H + o_bidx, hf.bidx.data(), hf.bidx.size() * sizeof(int));
    // launch batches per height: forward rows = f per node, backward rows = s per node
    fwd_levels_.assign(max_h + 1, LevelBatch());
    bwd_off_.assign(4 * (size_t)(max_h + 1), 0);
    {
        int* rm = (int*)(H + o_rowmaps);
        ProjLevelInfo* li = (ProjLevelInfo*)(H + o_levels);
        size_t w = 0;
        for (int h = 0; h <= max_h; h++) {
            const size_t o_fn = w;
            for (int t : by_h[h])
                for (int r = 0; r < descs[t].s + descs[t].b; r++) rm[w++] = t;
            const size_t o_fl = w;
            for (int t : by_h[h])
                for (int r = 0; r < descs[t].s + descs[t].b; r++) rm[w++] = r;
            const size_t o_bn = w;
            for (int t : by_h[h])
                for (int r = 0; r < descs[t].s; r++) rm[w++] = t;
            const size_t o_bl = w;
            for (int t : by_h[h])
                for (int r = 0; r < descs[t].s; r++) rm[w++] = r;
            LevelBatch& lb = fwd_levels_[h];
            lb.n_rows = (int)(o_fl - o_fn);
            lb.n_nodes = (int)(o_bl - o_bn);  // = backward rows at this height
            lb.row_node = (int*)(D + o_rowmaps) + o_fn;
            lb.row_local = (int*)(D + o_rowmaps) + o_fl;
            bwd_off_[4 * h] = o_fn;
            bwd_off_[4 * h + 1] = o_fl;
            bwd_off_[4 * h + 2] = o_bn;
            bwd_off_[4 * h + 3] = o_bl;
            li[h].n_fwd = lb.n_rows;
            li[h].n_bwd = lb.n_nodes;
            li[h].fwd_node = (long long)o_fn;
            li[h].fwd_local = (long long)o_fl;
            li[h].bwd_node = (long long)o_bn;
            li[h].bwd_local = (long long)o_bl;
        }
    }
    n_levels_ = max_h + 1;
    d_rnode_ = (int64_t*)(D + o_rnode);
    d_rw_ = (double*)(D + o_rw);
    d_rperm_ = (int*)(D + o_rperm);
    d_tnode_ = (int64_t*)(D + o_tnode);
    d_tptr_ = (int*)(D + o_tptr);
    d_trow_ = (int*)(D + o_trow);
    d_tw_ = (double*)(D + o_tw);
    d_nodes_ = (ProjNodeDesc*)(D + o_nodes);
    d_bidx_ = (int*)(D + o_bidx);
    d_rowmaps_ = (int*)(D + o_rowmaps);
    d_levels_ = (ProjLevelInfo*)(D + o_levels);
    d_rhs_ = (double*)(D + o_rhs);
    d_y_ = (double*)(D + o_y);
    d_sol_ = (double*)(D + o_sol);
    d_self_ = (ProjDev*)(D + o_self);
    *(ProjDev*)(H + o_self) = dev_view();
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(D, H, upload_bytes, cudaMemcpyHostToDevice, stream));
    if (getenv("SHM3D_DEBUG"))
        fprintf(stderr, "[shm3d] projector m=%d: factor %.1f ms, maps+upload issue %.1f ms\n", m, (tb1 - tb0) * 1e3,
                (wall() - tb1) * 1e3);
    // the staging buffers are reused by the next build() of this object; every solve ends with a stream
    // synchronisation, so the copies issued here have completed by then
}

ProjDev Projector::dev_view() const {
    ProjDev A;
    A.m = m_;
    A.n_touched = n_touched_;
    A.n_levels = n_levels_;
    A.rnode = d_rnode_;
    A.rw = d_rw_;
    A.rperm = d_rperm_;
    A.levels = d_levels_;
    A.rowmaps = d_rowmaps_;
    A.nodes = d_nodes_;
    A.mat = d_mat_;
    A.bidx = d_bidx_;
    A.tnode = d_tnode_;
    A.tptr = d_tptr_;
    A.trow = d_trow_;
    A.tw = d_tw_;
    A.rhs = d_rhs_;
    A.y = d_y_;
    A.sol = d_sol_;
    return A;
}

void Projector::gather(const float* v, const float* w, const double* shift_num, double shift_den,
                       cudaStream_t s) const {
    if (!m_) return;
    launch_chained(k_proj_gather, (m_ + 127) / 128, 128, s, dev_view(), v, w, shift_num, shift_den);
    if (reduce_hook_) reduce_hook_(d_rhs_, m_, s);
}

void Projector::solve(cudaStream_t s) const {
    if (!m_) return;
    const ProjDev A = dev_view();
    const int H = (int)fwd_levels_.size();
    for (int h = 0; h < H; h++) {
        const LevelBatch& lb = fwd_levels_[h];
        if (!lb.n_rows) continue;
        launch_chained(k_proj_fwd, (lb.n_rows * 32 + 255) / 256, 256, s, A, lb.n_rows, (const int*)lb.row_node,
                       (const int*)lb.row_local);
    }
    for (int h = H - 1; h >= 0; h--) {
        const LevelBatch& lb = fwd_levels_[h];
        const int nr = lb.n_nodes;
        if (!nr) continue;
        launch_chained(k_proj_bwd, (nr * 32 + 255) / 256, 256, s, A, nr, (const int*)(d_rowmaps_ + bwd_off_[4 * h + 2]),
                       (const int*)(d_rowmaps_ + bwd_off_[4 * h + 3]));
    }
}

void Projector::scatter_sub(float* v, cudaStream_t s) const {
    if (!m_ || !n_touched_) return;
    launch_chained(k_proj_scatter, (n_touched_ + 127) / 128, 128, s, dev_view(), v);
}

void Projector::record_apply(std::vector<TailOp>& ops, float* v, const float* w, bool shifted) const {
    if (!m_) return;
    TailOp op;
    memset(&op, 0, sizeof(op));
    op.proj = d_self_;
    op.code = kTGather;
    op.a = v;
    op.b = w;
    op.h = shifted ? 1 : 0;
    ops.push_back(op);
    op.a = op.b = nullptr;
    const int H = (int)fwd_levels_.size();
    for (int h = 0; h < H; h++) {
        if (!fwd_levels_[h].n_rows) continue;
        op.code = kTFwd;
        op.h = h;
        ops.push_back(op);
    }
    for (int h = H - 1; h >= 0; h--) {
        if (!fwd_levels_[h].n_nodes) continue;
        op.code = kTBwd;
        op.h = h;
        ops.push_back(op);
    }
    if (n_touched_) {
        op.code = kTScatter;
        op.h = 0;
        op.o = v;
        ops.push_back(op);
    }
}

void Projector::apply(float* v, cudaStream_t s) const {
    if (!m_) return;
    gather(v, nullptr, nullptr, 1.0, s);
    solve(s);
    scatter_sub(v, s);
}

void Projector::apply_update(float* v, const float* w, cudaStream_t s) const {
    if (!m_) return;
    gather(v, w, nullptr, 1.0, s);
    solve(s);
    scatter_sub(v, s);
}

void Projector::apply_shifted(float* v, const double* shift_num, double shift_den, cudaStream_t s) const {
    if (!m_) return;
    gather(v, nullptr, shift_num, shift_den, s);
    solve(s);
    scatter_sub(v, s);
}

namespace {
__global__ void k_sum_doubles(const double* __restrict__ v, int m, double* out) {  // one CTA, fixed order
    __shared__ double sh[256];
    double a = 0;
    for (int i = threadIdx.x; i < m; i += 256) a += v[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}
}  // namespace

void Projector::violation_sum(const float* v, double* out_sum, cudaStream_t s) const {
    if (!m_) return;
    gather(v, nullptr, nullptr, 1.0, s);  // (with the all-reduce of the slab-parallel runs)
    k_sum_doubles<<<1, 256, 0, s>>>(d_rhs_, m_, out_sum);
    SHM3D_LAUNCHED();
}

void Projector::multipliers(const float* v, std::vector<double>& lam_host, cudaStream_t s) const {
    lam_host.assign(m_, 0.0);
    if (!m_) return;
    gather(v, nullptr, nullptr, 1.0, s);
    solve(s);
    std::vector<double> tmp(m_);
    SHM3D_CUDA_CHECK(cudaMemcpyAsync(tmp.data(), d_sol_, m_ * sizeof(double), cudaMemcpyDeviceToHost, s));
    SHM3D_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int r = 0; r < m_; r++) lam_host[r] = tmp[perm_[r]];
}

}  // namespace shm3d
