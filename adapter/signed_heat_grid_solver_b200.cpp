// signed_heat_grid_solver_b200.cpp -- the translation unit a maintainer of nzfeng/signed-heat-3d drops in INSTEAD OF
// src/signed_heat_grid_solver.cpp to run the grid solver on a B200 through libshm3d_grid.so (include/shm3d_grid.h).
//
// It is compiled against the reference's own, unchanged headers (include/signed_heat_grid_solver.h,
// include/signed_heat_3d.h) and keeps everything src/main.cpp relies on: the class, both computeDistance overloads,
// VERBOSE, the `rebuild` caching rule (:8, :119), the polyscope::registerVolumeGrid("domain", ...) side effect (:35,
// :143; main.cpp:95 looks the grid up by that name), exceptions of the types geometry-central's solvers throw.
// Host-side geometry (centroid, radius, meanEdgeLength, setFaceVectorAreas, barycentres) stays with geometry-central /
// signed_heat_3d.cpp exactly as in the reference; only Steps 1-3 + shift go to the GPU.
//
// In this repository the file is compile-checked against the reference's headers with the oracle's shim standing in
// for geometry-central / Eigen / polyscope (tests/test_reference_build.py::test_adapter_compiles_against_the_reference_headers).
#include "signed_heat_grid_solver.h"

#include <stdexcept>
#include <string>

#include "shm3d_grid.h"

namespace {
shm3d_ctx* b200_context() {  // one context per process, created on first use (device 0)
    static shm3d_ctx* ctx = [] {
        shm3d_ctx* c = nullptr;
        if (shm3d_ctx_create(&c, 0) != SHM3D_OK) throw std::runtime_error(shm3d_last_error(nullptr));
        return c;
    }();
    return ctx;
}
[[noreturn]] void raise_like_geometry_central(int rc) {
    const std::string msg = shm3d_last_error(b200_context());
    if (rc == SHM3D_ERR_NONFINITE) throw std::logic_error(msg);  // checkFinite (square_solvers.cpp:123-125,164-169)
    if (rc == SHM3D_ERR_INVALID_ARG || rc == SHM3D_ERR_FACTORIZATION) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
}  // namespace

SignedHeatGridSolver::SignedHeatGridSolver() {}

// barycenter() is declared in the reference's header and used below (reference: :498-503)
Vector3 SignedHeatGridSolver::barycenter(VertexPositionGeometry& geometry, const Face& f) const {
    Vector3 c = {0, 0, 0};
    for (Vertex v : f.adjacentVertices()) c += geometry.vertexPositions[v];
    c /= f.degree();
    return c;
}

Vector<double> SignedHeatGridSolver::computeDistance(VertexPositionGeometry& geometry, const SignedHeat3DOptions& options) {
    if (options.rebuild || nx == 0) {  // reference :8-36 without the Laplacian it factorises and never uses (:30)
        Vector3 c = centroid(geometry);
        double r = radius(geometry, c);
        double s = r * options.scale;
        bboxMin = {-s, -s, -s};
        bboxMax = {s, s, s};
        bboxMin += c;
        bboxMax += c;
        glm::vec3 boundMin, boundMax;
        for (int i = 0; i < 3; i++) {
            boundMin[i] = bboxMin[i];
            boundMax[i] = bboxMax[i];
        }
        nx = 2 * std::pow(2, options.hCoef + 3);
        ny = nx;
        nz = nx;
        cellSize = 2. * s / (nx - 1);
        polyscope::registerVolumeGrid("domain", {nx, ny, nz}, boundMin, boundMax);
    }
    SurfaceMesh& mesh = geometry.mesh;
    double h = meanEdgeLength(geometry);  // :42-44
    shortTime = options.tCoef * h * h;

    shm3d_params p = shm3d_params();
    p.nx = (int32_t)nx;
    p.ny = (int32_t)ny;
    p.nz = (int32_t)nz;
    for (int a = 0; a < 3; a++) p.bbox_min[a] = bboxMin[a];
    p.cell = cellSize;
    p.lambda = std::sqrt(1. / shortTime);
    p.flags = SHM3D_FLAG_SCRUB_NONFINITE | SHM3D_FLAG_FP64_UNDERFLOW | (VERBOSE ? SHM3D_FLAG_VERBOSE : 0u) | (options.fastIntegration ? SHM3D_FLAG_FAST : 0u);

    setFaceVectorAreas(geometry, faceAreas, faceNormals);  // :47
    const size_t F = mesh.nFaces();
    std::vector<double> pos(3 * F), nrm(3 * F), area(F);
    size_t i = 0;
    for (Face f : mesh.faces()) {  // face order = constraint order ("first face per cell wins", :86-98)
        Vector3 y = barycenter(geometry, f);
        Vector3 n = faceNormals[f];
        for (int a = 0; a < 3; a++) {
            pos[3 * i + a] = y[a];
            nrm[3 * i + a] = n[a];
        }
        area[i++] = faceAreas[f];
    }
    Vector<double> phi = Vector<double>::Zero(nx * ny * nz);
    int rc = shm3d_solve(b200_context(), &p, (int64_t)F, pos.data(), nrm.data(), area.data(), &phi[0], nullptr);
    if (rc != SHM3D_OK) raise_like_geometry_central(rc);
    return phi;
}

Vector<double> SignedHeatGridSolver::computeDistance(pointcloud::PointPositionNormalGeometry& pointGeom,
                                                     const SignedHeat3DOptions& options) {
    {  // the reference rebuilds the grid on every call of this overload (its poissonSolver stays null, :119)
        Vector3 c = centroid(pointGeom);
        double r = radius(pointGeom, c);
        double s = r * options.scale;
        bboxMin = {-s, -s, -s};
        bboxMax = {s, s, s};
        bboxMin += c;
        bboxMax += c;
        glm::vec3 boundMin, boundMax;
        for (int i = 0; i < 3; i++) {
            boundMin[i] = bboxMin[i];
            boundMax[i] = bboxMax[i];
        }
        nx = 2 * std::pow(2, options.hCoef + 3);
        ny = nx;
        nz = nx;
        cellSize = 2. * s / (nx - 1);
        polyscope::registerVolumeGrid("domain", {nx, ny, nz}, boundMin, boundMax);
    }
    pointGeom.requireTuftedTriangulation();  // :149-151
    pointGeom.tuftedGeom->requireVertexDualAreas();
    double h = meanEdgeLength(*(pointGeom.tuftedGeom));
    shortTime = options.tCoef * h * h;

    shm3d_params p = shm3d_params();
    p.nx = (int32_t)nx;
    p.ny = (int32_t)ny;
    p.nz = (int32_t)nz;
    for (int a = 0; a < 3; a++) p.bbox_min[a] = bboxMin[a];
    p.cell = cellSize;
    p.lambda = std::sqrt(1. / shortTime);
    p.flags = SHM3D_FLAG_FP64_UNDERFLOW | (VERBOSE ? SHM3D_FLAG_VERBOSE : 0u) | (options.fastIntegration ? SHM3D_FLAG_FAST : 0u);  // no scrub (:180)

    const size_t P = pointGeom.cloud.nPoints();
    std::vector<double> pos(3 * P), nrm(3 * P), area(P);
    for (size_t i = 0; i < P; i++) {
        Vector3 y = pointGeom.positions[i];
        Vector3 n = pointGeom.normals[i];
        for (int a = 0; a < 3; a++) {
            pos[3 * i + a] = y[a];
            nrm[3 * i + a] = n[a];
        }
        area[i] = pointGeom.tuftedGeom->vertexDualAreas[i];  // :165
    }
    Vector<double> phi = Vector<double>::Zero(nx * ny * nz);
    int rc = shm3d_solve(b200_context(), &p, (int64_t)P, pos.data(), nrm.data(), area.data(), &phi[0], nullptr);
    pointGeom.unrequireTuftedTriangulation();  // :219-220
    pointGeom.tuftedGeom->unrequireVertexDualAreas();
    if (rc != SHM3D_OK) raise_like_geometry_central(rc);
    return phi;
}
