// gc_adapter_harness.cpp -- TEST INFRASTRUCTURE.  A SignedHeatGridSolver translation unit compiled against the
// reference's unchanged headers AND the real geometry-central headers / sources from the reference tree (Eigen: stub,
// see ref_shim/eigen_stub; polyscope: registerVolumeGrid stub over the real glm), driven the way src/main.cpp drives the
// class.  Linked twice (oracle/Makefile): with the PRODUCT's drop-in TU adapter/signed_heat_grid_solver_b200.cpp ->
// _ref/libshm_adapter_gc.so (needs a GPU), and with the REFERENCE's own src/signed_heat_grid_solver.cpp ->
// _ref/libshm_ref_gc.so (CPU; the KKT solve goes to the callback set with gcad_set_solver).  Inputs: real SurfaceMesh + VertexPositionGeometry built with
// makeSurfaceMeshAndGeometry (main.cpp:269-271 via readSurfaceMesh), real PointCloud + PointPositionNormalGeometry
// (main.cpp:277-285, geometry-central's own tufted-cover weights).  oracle/Makefile -> _ref/libshm_adapter_gc.so.
#include <cstdint>
#include <cstring>
#include <string>

#include "geometrycentral/surface/surface_mesh_factories.h"
#include "signed_heat_grid_solver.h"

namespace {
std::string g_err;
SignedHeat3DOptions make_opts(double tCoef, double hCoef, double scale, int fast) {
    SignedHeat3DOptions o;
    o.tCoef = tCoef;
    o.hCoef = hCoef;
    o.scale = scale;
    o.fastIntegration = fast != 0;
    o.rebuild = true;
    return o;
}
int finish(const Vector<double>& phi, double* phi_out, int64_t capacity, int64_t* dims_out, float* bbox_out) {
    if ((int64_t)phi.size() > capacity) {
        g_err = "output buffer too small";
        return 2;
    }
    for (Eigen::Index i = 0; i < phi.size(); i++) phi_out[i] = phi[i];
    const polyscope::VolumeGrid& g = polyscope::shim_last_grid();
    if (dims_out)
        for (int a = 0; a < 3; a++) dims_out[a] = (int64_t)g.dim[a];
    if (bbox_out)
        for (int a = 0; a < 3; a++) {
            bbox_out[a] = g.bmin[a];
            bbox_out[3 + a] = g.bmax[a];
        }
    return 0;
}
}  // namespace

extern "C" {

const char* gcad_last_error(void) { return g_err.c_str(); }

// Used when this harness wraps the REFERENCE's own src/signed_heat_grid_solver.cpp (oracle/_ref/libshm_ref_gc.so): the
// routine the Eigen stub's SparseLU hands the assembled KKT system to (scipy SuperLU in the tests).  Unused by the adapter.
void gcad_set_solver(Eigen::shm_stub_solve_fn fn) { Eigen::shm_stub_solver() = fn; }

int gcad_compute_distance_mesh(const double* V, int64_t nV, const int64_t* face_vertices, const int64_t* face_offsets,
                               int64_t nF, double tCoef, double hCoef, double scale, int fast, double* phi_out,
                               int64_t capacity, int64_t* dims_out, float* bbox_out) {
    try {
        std::vector<std::vector<size_t>> polygons((size_t)nF);
        for (int64_t f = 0; f < nF; f++)
            polygons[(size_t)f].assign(face_vertices + face_offsets[f], face_vertices + face_offsets[f + 1]);
        std::vector<Vector3> pos((size_t)nV);
        for (int64_t i = 0; i < nV; i++) pos[(size_t)i] = Vector3{V[3 * i], V[3 * i + 1], V[3 * i + 2]};
        std::unique_ptr<SurfaceMesh> mesh;
        std::unique_ptr<VertexPositionGeometry> geometry;
        std::tie(mesh, geometry) = makeSurfaceMeshAndGeometry(polygons, pos);
        SignedHeatGridSolver s;
        s.VERBOSE = false;
        Vector<double> phi = s.computeDistance(*geometry, make_opts(tCoef, hCoef, scale, fast));
        return finish(phi, phi_out, capacity, dims_out, bbox_out);
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

int gcad_compute_distance_points(const double* P, const double* N, int64_t nP, double tCoef, double hCoef, double scale,
                                 int fast, double* phi_out, int64_t capacity, int64_t* dims_out, float* bbox_out) {
    try {
        pointcloud::PointCloud cloud((size_t)nP);
        pointcloud::PointData<Vector3> pointPositions(cloud), pointNormals(cloud);
        for (int64_t i = 0; i < nP; i++) {
            pointPositions[(size_t)i] = Vector3{P[3 * i], P[3 * i + 1], P[3 * i + 2]};
            pointNormals[(size_t)i] = Vector3{N[3 * i], N[3 * i + 1], N[3 * i + 2]};
        }
        pointcloud::PointPositionNormalGeometry pointGeom(cloud, pointPositions, pointNormals);
        SignedHeatGridSolver s;
        s.VERBOSE = false;
        Vector<double> phi = s.computeDistance(pointGeom, make_opts(tCoef, hCoef, scale, fast));
        return finish(phi, phi_out, capacity, dims_out, bbox_out);
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
}
