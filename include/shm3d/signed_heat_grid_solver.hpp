// signed_heat_grid_solver.hpp -- dependency-free C++ mirror of the reference's grid-solver surface, on top of the
// C ABI (include/shm3d_grid.h).  Header-only; link with -lshm3d_grid.
//
// Mirrors (names, argument meaning, defaults, error behaviour):
//   struct SignedHeat3DOptions            <- include/signed_heat_3d.h:20-28
//   class  SignedHeatGridSolver           <- include/signed_heat_grid_solver.h:11-47
//     bool VERBOSE                        <- :22
//     computeDistance(mesh, options)      <- :16-17, src/signed_heat_grid_solver.cpp:5-114
//     computeDistance(points, options)    <- :19-20, src/signed_heat_grid_solver.cpp:116-222
//     isosurface(phi, isoval)             <- what src/main.cpp:116-128 asks polyscope for (registerIsosurfaceAsMesh,
//                                            deps/polyscope/src/volume_grid_scalar_quantity.cpp:209-228), on the GPU
// What differs, and why: the reference's overloads take geometry-central objects (VertexPositionGeometry&,
// PointPositionNormalGeometry&) and return Eigen::VectorXd; neither library can be a dependency of this repository
// (Eigen is not vendored anywhere on this image), so the inputs here are flat arrays carrying exactly what the
// reference reads from those objects.  INTEGRATION.md holds the thin geometry-central adapter that restores the
// original signatures wherever those headers exist.
// Errors are C++ exceptions like geometry-central's: std::invalid_argument (bad input / factorisation failed),
// std::logic_error (non-finite right-hand side: checkFinite, square_solvers.cpp:123-125), std::runtime_error
// (CUDA / NCCL / no convergence).  There is no CPU fallback: constructing a solver without a GPU throws.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../shm3d_grid.h"

namespace shm3d {

// include/signed_heat_3d.h:20-28.  levelSetConstraint / useCrouzeixRaviart are tet-solver options; the grid solver
// ignores them (src/signed_heat_grid_solver.cpp:75) and so does this mirror.
enum class LevelSetConstraint { None, ZeroSet, Multiple };
struct SignedHeat3DOptions {
    LevelSetConstraint levelSetConstraint = LevelSetConstraint::ZeroSet;
    double tCoef = 1.0;
    double hCoef = 0.0;
    bool rebuild = true;
    double scale = 2.0;
    bool useCrouzeixRaviart = true;
    bool fastIntegration = false;
};

// What computeDistance(VertexPositionGeometry&) reads from the geometry: vertex positions and the polygon soup.
struct PolygonMesh {
    std::vector<double> vertexPositions;  // [nV][3]
    std::vector<int64_t> faceVertices;    // concatenated vertex indices
    std::vector<int64_t> faceOffsets;     // [nF+1]
    int64_t nVertices() const { return (int64_t)vertexPositions.size() / 3; }
    int64_t nFaces() const { return faceOffsets.empty() ? 0 : (int64_t)faceOffsets.size() - 1; }
};

// What computeDistance(PointPositionNormalGeometry&) reads: positions, normals, and the tufted-triangulation
// quantities geometry-central derives from them (vertex dual areas, mean intrinsic edge length;
// src/signed_heat_grid_solver.cpp:149-151) -- supplied by the caller.
struct OrientedPointCloud {
    std::vector<double> positions;  // [nP][3]
    std::vector<double> normals;    // [nP][3]
    std::vector<double> areas;      // [nP]
    double meanEdgeLength = 0.0;
    int64_t nPoints() const { return (int64_t)positions.size() / 3; }
    // Fill areas / meanEdgeLength from positions + normals with shm3d_point_weights: geometry-central's pipeline restated
    // (kNN 30, local Delaunay stars, triangle soup, mollification, tufted cover, intrinsic Delaunay flips).
    void computeWeights() {
        areas.assign((size_t)nPoints(), 0.0);
        int rc = shm3d_point_weights(positions.data(), normals.data(), nPoints(), 30, areas.data(), &meanEdgeLength, nullptr, nullptr,
                                     nullptr, nullptr);
        if (rc != SHM3D_OK) throw std::invalid_argument("OrientedPointCloud::computeWeights: need > 30 finite points with usable normals");
    }
};

class SignedHeatGridSolver {
  public:
    explicit SignedHeatGridSolver(int device = 0) {
        int rc = shm3d_ctx_create(&ctx_, device);
        if (rc != SHM3D_OK) throw std::runtime_error(std::string("SignedHeatGridSolver: ") + shm3d_last_error(nullptr));
    }
    ~SignedHeatGridSolver() { shm3d_ctx_destroy(ctx_); }
    SignedHeatGridSolver(const SignedHeatGridSolver&) = delete;
    SignedHeatGridSolver& operator=(const SignedHeatGridSolver&) = delete;

    bool VERBOSE = true;
    // Reproduce the reference's double-precision underflow of X.norm() at far nodes (SHM3D_FLAG_FP64_UNDERFLOW,
    // include/shm3d_grid.h): on, because a drop-in returns what the reference returns (validated on the B200 against the
    // reference's own output and the fp64 oracle, tests/test_gpu_baseline_configs.py).  false = Steps 1-2 finite everywhere.
    bool referenceUnderflow = true;

    // Mesh overload.  Returns phi at the nx*ny*nz nodes, index i + j*nx + k*nx*ny (x fastest).
    std::vector<double> computeDistance(const PolygonMesh& mesh, const SignedHeat3DOptions& options = SignedHeat3DOptions()) {
        const int64_t nF = mesh.nFaces();
        if (nF <= 0 || mesh.nVertices() <= 0) throw std::invalid_argument("computeDistance: empty mesh");
        shm3d_params p;
        std::vector<double> pos(3 * nF), nrm(3 * nF), area(nF);
        double h = 0;
        const bool rebuild = options.rebuild || !haveGrid_;  // src/signed_heat_grid_solver.cpp:8
        int rc = shm3d_prepare_mesh(mesh.vertexPositions.data(), mesh.nVertices(), mesh.faceVertices.data(),
                                    mesh.faceOffsets.data(), nF, options.tCoef, rebuild ? options.hCoef : 0.0,
                                    options.scale, &p, pos.data(), nrm.data(), area.data(), &h);
        if (rc != SHM3D_OK) throw std::invalid_argument("computeDistance: invalid mesh (face of degree < 3 or vertex index out of range)");
        if (!rebuild) keepCachedGrid(p);
        return finish(p, nF, pos, nrm, area, options);
    }

    // Point-cloud overload.  The reference rebuilds the grid on every call of this overload (its poissonSolver
    // stays null, :119) and does not scrub non-finite right-hand-side entries (:180).
    std::vector<double> computeDistance(const OrientedPointCloud& cloud, const SignedHeat3DOptions& options = SignedHeat3DOptions()) {
        const int64_t nP = cloud.nPoints();
        if (nP <= 0 || (int64_t)cloud.normals.size() != 3 * nP || (int64_t)cloud.areas.size() != nP)
            throw std::invalid_argument("computeDistance: inconsistent point cloud arrays");
        shm3d_params p;
        int rc = shm3d_prepare_points(cloud.positions.data(), nP, cloud.meanEdgeLength, options.tCoef, options.hCoef,
                                      options.scale, &p);
        if (rc != SHM3D_OK) throw std::invalid_argument("computeDistance: invalid point cloud (mean edge length must be > 0)");
        return finish(p, nP, cloud.positions, cloud.normals, cloud.areas, options);
    }

    // The grid of the last solve -- what the reference hands to polyscope::registerVolumeGrid("domain", ...)
    // (src/signed_heat_grid_solver.cpp:35,143).
    size_t nx() const { return (size_t)grid_.nx; }
    size_t ny() const { return (size_t)grid_.ny; }
    size_t nz() const { return (size_t)grid_.nz; }
    double cellSize() const { return grid_.cell; }
    void bbox(double bmin[3], double bmax[3]) const {
        const int n[3] = {grid_.nx, grid_.ny, grid_.nz};
        for (int a = 0; a < 3; a++) {
            bmin[a] = grid_.bbox_min[a];
            bmax[a] = grid_.bbox_min[a] + grid_.cell * (n[a] - 1);
        }
    }
    const shm3d_stats& lastStats() const { return stats_; }

    // The level set {phi = isoval} of a field on the grid of the last solve, as the indexed mesh polyscope's
    // registerIsosurfaceAsMesh would produce from the float32-narrowed values: same vertex coordinates (world), same
    // vertex numbering, same triangle order.
    struct IsoMesh {
        std::vector<float> vertices;      // [nV][3]
        std::vector<uint32_t> triangles;  // [nT][3]
        shm3d_iso_stats stats{};
    };
    IsoMesh isosurface(const std::vector<double>& phi, float isoval = 0.f) {
        if (!haveGrid_) throw std::invalid_argument("isosurface: no grid yet (call computeDistance first)");
        if (phi.size() != (size_t)grid_.nx * grid_.ny * grid_.nz) throw std::invalid_argument("isosurface: field size does not match the grid");
        IsoMesh m;
        int rc = shm3d_isosurface(ctx_, &grid_, phi.data(), SHM3D_FIELD_HOST_F64, isoval, nullptr, nullptr, 0u, &m.stats);
        if (rc == SHM3D_OK) {
            m.vertices.resize((size_t)m.stats.n_vertices * 3);
            m.triangles.resize((size_t)m.stats.n_triangles * 3);
            rc = shm3d_isosurface_fetch(ctx_, m.vertices.data(), m.triangles.data());
        }
        if (rc != SHM3D_OK) {
            const std::string msg = shm3d_last_error(ctx_);
            if (rc == SHM3D_ERR_INVALID_ARG) throw std::invalid_argument(msg);
            throw std::runtime_error(msg);
        }
        return m;
    }
    shm3d_params solverParams{};  // optional overrides: cull_tau, cg_rel_tol, cg_max_iters, mg_smooth (0 = defaults)

  private:
    shm3d_ctx* ctx_ = nullptr;
    bool haveGrid_ = false;
    shm3d_params grid_{};
    shm3d_stats stats_{};

    void keepCachedGrid(shm3d_params& p) const {
        const double lambda = p.lambda;
        const uint32_t flags = p.flags;
        p = grid_;
        p.lambda = lambda;
        p.flags = flags;
    }

    std::vector<double> finish(shm3d_params& p, int64_t n, const std::vector<double>& pos, const std::vector<double>& nrm,
                               const std::vector<double>& area, const SignedHeat3DOptions& options) {
        if (VERBOSE) p.flags |= SHM3D_FLAG_VERBOSE;
        if (options.fastIntegration) p.flags |= SHM3D_FLAG_FAST;
        if (referenceUnderflow) p.flags |= SHM3D_FLAG_FP64_UNDERFLOW;
        else p.flags &= ~SHM3D_FLAG_FP64_UNDERFLOW;
        p.cull_tau = solverParams.cull_tau;
        p.cg_rel_tol = solverParams.cg_rel_tol;
        p.cg_max_iters = solverParams.cg_max_iters;
        p.mg_smooth = solverParams.mg_smooth;
        p.mg_constrained_from = solverParams.mg_constrained_from;
        std::vector<double> phi((size_t)p.nx * p.ny * p.nz);
        if (VERBOSE) std::fprintf(stderr, "nx: %d\tny: %d\tnz: %d\n", p.nx, p.ny, p.nz);  // :27
        int rc = shm3d_solve(ctx_, &p, n, pos.data(), nrm.data(), area.data(), phi.data(), &stats_);
        if (rc != SHM3D_OK) {
            const std::string msg = shm3d_last_error(ctx_);
            switch (rc) {
                case SHM3D_ERR_INVALID_ARG: throw std::invalid_argument(msg);
                case SHM3D_ERR_NONFINITE: throw std::logic_error(msg);       // checkFinite (square_solvers.cpp:123-125)
                case SHM3D_ERR_FACTORIZATION: throw std::invalid_argument(msg);  // "Solver factorization failed"
                default: throw std::runtime_error(msg);
            }
        }
        grid_ = p;
        haveGrid_ = true;
        if (VERBOSE)
            std::fprintf(stderr, "Solve time (s): %.4f  [sum %.1f ms | constraints %.1f ms | pcg %.1f ms, %d its | m = %d]\n",
                         stats_.ms_total * 1e-3, stats_.ms_sum, stats_.ms_constraints, stats_.ms_pcg, stats_.cg_iters,
                         stats_.m_constraints);
        return phi;
    }
};

}  // namespace shm3d
