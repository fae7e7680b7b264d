// projector.cuh -- the zero-level-set constraints of Step 3 and the projector onto their null space.
//
// Reference: src/signed_heat_grid_solver.cpp:80-100 builds A (m x N, one trilinear row per occupied grid
// cell, first source per cell wins, weights from trilinearCoefficients :433-464) and solves the KKT system
// [[L, A^T],[A, 0]] by sparse LU (:101-108).  Here the same constraint set drives a null-space method:
//   Pi = I - D^-1 A^T (A D^-1 A^T)^-1 A        (D = diag of K'; D-orthogonal projector onto null(A))
// applied O(100) times per solve.  A D^-1 A^T is m x m, sparse (cells sharing a node), SPD but badly
// conditioned, so it is factorised directly in fp64 by a geometric nested-dissection
// multifrontal Cholesky on the host; the factor is stored as dense per-supernode blocks with explicitly inverted
// diagonal blocks, so that each triangular solve on the GPU is one batched GEMV launch per tree level.
#pragma once
#include <functional>
#include <memory>

#include "kernels.cuh"

namespace shm3d {

// Host-side constraint rows of one grid level.
struct ConstraintRows {
    int m = 0;
    std::vector<int64_t> node;  // [m*8] global node index i + j*nx + k*nx*ny, corner order 000,100,010,001,110,101,011,111
    std::vector<double> w;      // [m*8] trilinear weights
    std::vector<int> cell;      // [m*3] cell coordinates
    std::vector<int64_t> src;   // [m] index of the source that pins the cell
};

// first-source-per-cell constraint rows (src/signed_heat_grid_solver.cpp:86-98).  `strict`: throw when a
// source lies outside the node lattice (fine level); otherwise skip it (coarse multigrid levels).
void build_constraint_rows(int nx, int ny, int nz, const double bmin[3], double cell, int64_t M, const double* pos,
                           bool strict, ConstraintRows& out);

// one supernode of the factor (device-visible)
struct ProjNodeDesc {
    int s0, s, b;           // first permuted index, separator size, boundary size
    long long fwd, bwd;     // offsets (in doubles) of the [f x s] forward and [s x f] backward blocks
    long long bidx;         // offset into the boundary index list
};

// Host image of the factor of A D^-1 A^T (what gets uploaded).  Kept separate so that the host logic can be
// checked without a GPU (tests/test_host_logic.py through shm3d_debug_factor_solve).
struct HostFactor {
    int m = 0;
    bool all_interior = true;
    std::vector<int> perm;                 // row -> permuted index
    std::vector<ProjNodeDesc> nodes;       // supernodes, children before parents
    std::vector<std::vector<int>> by_height;  // node ids per tree height
    double* mat = nullptr;                 // per node: FWD [f x s] then BWD [s x f]
    size_t mat_size = 0;
    std::vector<double> mat_own;           // backing store when no external allocator is given
    std::function<double*(size_t)> mat_alloc;  // optional: storage for `mat` (e.g. page-locked staging memory)
    std::vector<int> bidx;                 // boundary index lists
    std::vector<double> dinv;              // [m*8] 1/d of each corner node (1 when uniform)
    void solve_host(std::vector<double>& v) const;  // v (permuted order) <- (A D^-1 A^T)^-1 v, same algorithm as the GPU
};
// how many ranks share this node's host cores (sizes the factorisation thread pool; call before the first solve)
void set_host_ranks_hint(int ranks_on_node);
void set_projector_chained_launches(bool enabled);  // per host thread, like set_march_config
void factor_constraints(const ConstraintRows& rows, int nx, int ny, int nz, bool uniform, HostFactor& out);

// per tree height: row maps of the forward (f rows per supernode) and backward (s rows) sweeps
struct ProjLevelInfo {
    int n_fwd, n_bwd;
    long long fwd_node, fwd_local, bwd_node, bwd_local;  // offsets into the row-map array
};

// Device-resident projector for one level.
class Projector {
  public:
    Projector() {}
    ~Projector();
    Projector(const Projector&) = delete;
    Projector& operator=(const Projector&) = delete;

    // Factorise A D^-1 A^T (D = number of in-range neighbours of each node; uniform = true uses D = I, the
    // Euclidean projector) and upload everything.  L describes the local slab of this level.
    void build(const ConstraintRows& rows, const LevelDims& L, bool uniform, cudaStream_t s);

    int m() const { return m_; }
    bool all_interior() const { return all_interior_; }
    size_t factor_bytes() const { return factor_bytes_; }
    int tree_height() const { return (int)fwd_levels_.size(); }

    // v <- Pi v  (in place; v is an interior pointer of a local slab vector)
    void apply(float* v, cudaStream_t s) const;
    // v <- v - D^-1 A^T (A D^-1 A^T)^-1 A (v - w): projects the UPDATE v - w (w = previous iterate)
    void apply_update(float* v, const float* w, cudaStream_t s) const;
    // v <- v - D^-1 A^T (A D^-1 A^T)^-1 A (v - shift), shift = *shift_num / shift_den (rows of A sum to one)
    void apply_shifted(float* v, const double* shift_num, double shift_den, cudaStream_t s) const;
    // *out_sum (device) = sum of the m entries of A v: the rows of A sum to one, so sum / m is the constant by which v
    // violates the constraints on average (Solver::run_pcg removes it from the whole field before the last projection)
    void violation_sum(const float* v, double* out_sum, cudaStream_t s) const;
    // lam <- (A D^-1 A^T)^-1 (A v): multipliers only (diagnostics / tests); lam_host has m entries in row order
    void multipliers(const float* v, std::vector<double>& lam_host, cudaStream_t s) const;
    // weighted source average helper is elsewhere (solver.cu)

    // raw pieces used by the fused paths
    // rhs_ = A (v - w - shift), w optional, shift = *shift_num / shift_den optional (rows of A sum to 1)
    void gather(const float* v, const float* w, const double* shift_num, double shift_den, cudaStream_t s) const;
    void solve(cudaStream_t s) const;                         // sol_ = (A D^-1 A^T)^-1 rhs_
    void scatter_sub(float* v, cudaStream_t s) const;         // v -= D^-1 A^T sol_

    // Cluster programs (mg_tail.cuh): the V-cycle tail runs whole projector applications as ops of one launch.
    static constexpr int kClusterMaxRows = 4096;  // largest system a single-CTA tail program takes on
    ProjDev dev_view() const;
    const ProjDev* dev_ptr() const { return d_self_; }
    // appends the ops of one application  v <- v - D^-1 A^T (A D^-1 A^T)^-1 A (v - w)  (w may be null) to a program
    void record_apply(std::vector<TailOp>& ops, float* v, const float* w, bool shifted) const;

  private:
    struct LevelBatch {
        int n_rows = 0;          // total matrix rows handled by this launch
        int* row_node = nullptr;   // [n_rows] -> node slot in the level
        int* row_local = nullptr;  // [n_rows] local row within the node block
        int first_node = 0, n_nodes = 0;
    };
    int m_ = 0;
    bool all_interior_ = true;
    size_t factor_bytes_ = 0;
    LevelDims L_{};
    // All device arrays live in two grow-only arenas (factor blocks / everything else) filled from two page-locked
    // staging buffers by one cudaMemcpyAsync each: no cudaMalloc / cudaFree (device-synchronising) in a steady-state
    // solve, so the host-side factorisation really overlaps the summation kernel running on the stream.
    DevBuf<unsigned char> d_arena_, d_matbuf_;
    unsigned char* h_arena_ = nullptr;  // cudaHostAlloc
    size_t h_arena_cap_ = 0;
    double* h_mat_ = nullptr;           // cudaHostAlloc
    size_t h_mat_cap_ = 0;
    // constraint rows on device (row-major, permuted row order)
    int64_t* d_rnode_ = nullptr;  // [m*8] LOCAL node index (interior-relative), or -1 if not on this rank
    double* d_rw_ = nullptr;      // [m*8]
    int* d_rperm_ = nullptr;      // [m] row -> permuted index
    // node-centric transpose (touched nodes)
    int n_touched_ = 0;
    int64_t* d_tnode_ = nullptr;  // [n_touched] local node index
    int* d_tptr_ = nullptr;       // [n_touched+1]
    int* d_trow_ = nullptr;       // [nnz] permuted row
    double* d_tw_ = nullptr;      // [nnz] weight / d_node
    // factor
    ProjNodeDesc* d_nodes_ = nullptr;
    double* d_mat_ = nullptr;
    int* d_bidx_ = nullptr;
    std::vector<LevelBatch> fwd_levels_;  // ascending height
    int* d_rowmaps_ = nullptr;
    mutable double* d_rhs_ = nullptr;  // [m] work vectors (permuted order)
    mutable double* d_y_ = nullptr;
    mutable double* d_sol_ = nullptr;
    std::vector<int> perm_;  // row -> permuted index
    std::vector<size_t> bwd_off_;
    ProjLevelInfo* d_levels_ = nullptr;  // [tree height] row maps per height, for the cluster programs
    int n_levels_ = 0;
    ProjDev* d_self_ = nullptr;          // dev_view() in device memory (what the ops of a cluster program point to)

  public:
    // multi-GPU: called on the gathered partial sums A v (m doubles on device) before the solve (allreduce)
    std::function<void(double*, int, cudaStream_t)> reduce_hook_;
};

}  // namespace shm3d
