// harness.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's own SignedHeatGridSolver, compiled from
// /root/reference/src/{signed_heat_grid_solver,signed_heat_3d}.cpp against the shim headers (include/shm_ref_shim.h).
// Loaded by oracle/reference_build.py; used only by tests/ (and, when present, as the CPU reference in bench.py).
#include <cstring>

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
    const polyscope::VolumeGrid& g = polyscope::shim_last_grid();  // what registerVolumeGrid("domain", ...) received
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

const char* ref_last_error(void) { return g_err.c_str(); }

// computeDistance(VertexPositionGeometry&, options) -- reference src/signed_heat_grid_solver.cpp:5-114
int ref_compute_distance_mesh(const double* V, int64_t nV, const int64_t* face_vertices, const int64_t* face_offsets,
                              int64_t nF, double tCoef, double hCoef, double scale, int fast,
                              geometrycentral::shim_solve_fn solver, double* phi_out, int64_t capacity, int64_t* dims_out,
                              float* bbox_out) {
    try {
        geometrycentral::shim_solver() = solver;
        std::vector<size_t> fv(face_vertices, face_vertices + face_offsets[nF]), fo(face_offsets, face_offsets + nF + 1);
        SurfaceMesh mesh((size_t)nV, fv, fo);
        std::vector<Vector3> pos((size_t)nV);
        for (int64_t i = 0; i < nV; i++) pos[i] = Vector3{V[3 * i], V[3 * i + 1], V[3 * i + 2]};
        VertexPositionGeometry geometry(mesh, pos);
        SignedHeatGridSolver s;
        s.VERBOSE = false;
        Vector<double> phi = s.computeDistance(geometry, make_opts(tCoef, hCoef, scale, fast));
        return finish(phi, phi_out, capacity, dims_out, bbox_out);
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// computeDistance(PointPositionNormalGeometry&, options) -- :116-222.  The tufted-triangulation quantities the
// reference reads (vertexDualAreas, mean edge length) are supplied by the caller.
int ref_compute_distance_points(const double* P, const double* N, const double* areas, int64_t nP, double mean_edge_length,
                                double tCoef, double hCoef, double scale, int fast, geometrycentral::shim_solve_fn solver,
                                double* phi_out, int64_t capacity, int64_t* dims_out, float* bbox_out) {
    try {
        geometrycentral::shim_solver() = solver;
        pointcloud::PointCloud cloud((size_t)nP);
        pointcloud::PointPositionNormalGeometry geom(cloud);
        for (int64_t i = 0; i < nP; i++) {
            geom.positions[i] = Vector3{P[3 * i], P[3 * i + 1], P[3 * i + 2]};
            geom.normals[i] = Vector3{N[3 * i], N[3 * i + 1], N[3 * i + 2]};
        }
        // a stand-in "tufted mesh": nP vertices, one triangle whose three edges all have the given length, so that
        // meanEdgeLength() = mean_edge_length and vertexDualAreas[p] = areas[p]
        if (nP < 3) throw std::invalid_argument("need at least 3 points");
        SurfaceMesh tmesh((size_t)nP, std::vector<size_t>{0, 1, 2}, std::vector<size_t>{0, 3});
        geom.tuftedGeom.reset(new EdgeLengthGeometry(tmesh));
        geom.tuftedGeom->edgeLengths = EdgeData<double>(tmesh);
        for (size_t e = 0; e < tmesh.nEdges(); e++) geom.tuftedGeom->edgeLengths[e] = mean_edge_length;
        geom.tuftedGeom->vertexDualAreas = VertexData<double>(tmesh);
        for (int64_t i = 0; i < nP; i++) geom.tuftedGeom->vertexDualAreas[(size_t)i] = areas[i];
        SignedHeatGridSolver s;
        s.VERBOSE = false;
        Vector<double> phi = s.computeDistance(geom, make_opts(tCoef, hCoef, scale, fast));
        return finish(phi, phi_out, capacity, dims_out, bbox_out);
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// the free helpers of signed_heat_3d.cpp on a mesh: centroid, radius, meanEdgeLength, per-face area / normal
int ref_mesh_scalars(const double* V, int64_t nV, const int64_t* face_vertices, const int64_t* face_offsets, int64_t nF,
                     double* centroid_out, double* radius_out, double* h_out, double* area_out, double* normal_out) {
    try {
        std::vector<size_t> fv(face_vertices, face_vertices + face_offsets[nF]), fo(face_offsets, face_offsets + nF + 1);
        SurfaceMesh mesh((size_t)nV, fv, fo);
        std::vector<Vector3> pos((size_t)nV);
        for (int64_t i = 0; i < nV; i++) pos[i] = Vector3{V[3 * i], V[3 * i + 1], V[3 * i + 2]};
        VertexPositionGeometry geometry(mesh, pos);
        Vector3 c = centroid(geometry);
        for (int a = 0; a < 3; a++) centroid_out[a] = c[a];
        *radius_out = radius(geometry, c);
        *h_out = meanEdgeLength(geometry);
        FaceData<double> areas;
        FaceData<Vector3> normals;
        setFaceVectorAreas(geometry, areas, normals);
        for (int64_t f = 0; f < nF; f++) {
            area_out[f] = areas[(size_t)f];
            for (int a = 0; a < 3; a++) normal_out[3 * f + a] = normals[(size_t)f][a];
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

double ref_yukawa(const double* x, const double* y, double lambda) {
    return yukawaPotential(Vector3{x[0], x[1], x[2]}, Vector3{y[0], y[1], y[2]}, lambda);
}
}
