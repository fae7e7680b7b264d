// gc_harness.cpp -- TEST INFRASTRUCTURE.  C entry point around geometry-central's OWN point-cloud pipeline, compiled from
// where its sources lie under /root/reference/deps/geometry-central (recipe: oracle/Makefile -> oracle/_ref/libshm_gc_ref.so)
// against oracle/ref_shim/eigen_stub, a stand-in for Eigen's *interface* only (Eigen itself is fetched by
// geometry-central's configure step and is absent from this image).  The calls below are the ones the reference makes
// for the point-cloud overload (src/main.cpp:277-285, src/signed_heat_grid_solver.cpp:149-151,165): none of the code
// they execute touches a matrix, so every geometry-central line on the path -- kNN (nanoflann), tangent coordinates,
// buildLocalTriangulations, the triangle-soup SurfaceMesh, mollifyIntrinsic, buildIntrinsicTuftedCover, flipToDelaunay,
// vertexDualAreas -- and the reference's own meanEdgeLength (src/signed_heat_3d.cpp:51-60) run as written.
// Pins row N1 (shm3d_point_weights).
#include <cstdint>
#include <string>

#include "geometrycentral/pointcloud/local_triangulation.h"
#include "geometrycentral/surface/meshio.h"
#include "geometrycentral/surface/surface_mesh_factories.h"
#include "signed_heat_3d.h"

namespace {
std::string g_err;
}

extern "C" {

const char* gcref_last_error(void) { return g_err.c_str(); }

int gcref_point_weights(const double* P, const double* N, int64_t nP, double* areas_out, double* h_out,
                        int64_t* n_faces_out, int64_t* n_edges_out) {
    try {
        pointcloud::PointCloud cloud((size_t)nP);                       // main.cpp:277
        pointcloud::PointData<Vector3> pointPositions(cloud), pointNormals(cloud);
        for (int64_t i = 0; i < nP; i++) {
            pointPositions[(size_t)i] = Vector3{P[3 * i], P[3 * i + 1], P[3 * i + 2]};
            pointNormals[(size_t)i] = Vector3{N[3 * i], N[3 * i + 1], N[3 * i + 2]};
        }
        pointcloud::PointPositionNormalGeometry pointGeom(cloud, pointPositions, pointNormals);  // main.cpp:284-285
        pointGeom.requireTuftedTriangulation();                          // src/signed_heat_grid_solver.cpp:149
        pointGeom.tuftedGeom->requireVertexDualAreas();                  // :150
        *h_out = meanEdgeLength(*(pointGeom.tuftedGeom));                // :151
        for (int64_t i = 0; i < nP; i++) areas_out[i] = pointGeom.tuftedGeom->vertexDualAreas[(size_t)i];  // :165
        if (n_faces_out) *n_faces_out = (int64_t)pointGeom.tuftedMesh->nFaces();
        if (n_edges_out) *n_edges_out = (int64_t)pointGeom.tuftedMesh->nEdges();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// The mesh-side host quantities of the grid solver through the REAL geometry-central containers (rows a4-a6):
// centroid / radius / meanEdgeLength / setFaceVectorAreas of the reference's src/signed_heat_3d.cpp, and the face
// barycentres in mesh.faces() order with f.adjacentVertices() (src/signed_heat_grid_solver.cpp:498-503) -- the order that
// decides which source pins a cell.  Mesh built as src/main.cpp does for OBJ input (makeSurfaceMeshAndGeometry on the
// polygon soup: general, possibly non-manifold SurfaceMesh).
int gcref_mesh_sources(const double* V, int64_t nV, const int64_t* face_vertices, const int64_t* face_offsets, int64_t nF,
                       double* centroid_out, double* radius_out, double* h_out, double* area_out, double* normal_out,
                       double* bary_out, int64_t* n_edges_out) {
    try {
        std::vector<std::vector<size_t>> polygons((size_t)nF);
        for (int64_t f = 0; f < nF; f++)
            polygons[(size_t)f].assign(face_vertices + face_offsets[f], face_vertices + face_offsets[f + 1]);
        std::vector<Vector3> pos((size_t)nV);
        for (int64_t i = 0; i < nV; i++) pos[(size_t)i] = Vector3{V[3 * i], V[3 * i + 1], V[3 * i + 2]};
        std::unique_ptr<SurfaceMesh> mesh;
        std::unique_ptr<VertexPositionGeometry> geometry;
        std::tie(mesh, geometry) = makeSurfaceMeshAndGeometry(polygons, pos);
        Vector3 c = centroid(*geometry);
        for (int a = 0; a < 3; a++) centroid_out[a] = c[a];
        *radius_out = radius(*geometry, c);
        *h_out = meanEdgeLength(*geometry);
        FaceData<double> areas;
        FaceData<Vector3> normals;
        setFaceVectorAreas(*geometry, areas, normals);
        size_t i = 0;
        for (Face f : mesh->faces()) {
            Vector3 b = {0, 0, 0};
            for (Vertex v : f.adjacentVertices()) b += geometry->vertexPositions[v];
            b /= f.degree();
            area_out[i] = areas[f];
            for (int a = 0; a < 3; a++) {
                normal_out[3 * i + a] = normals[f][a];
                bary_out[3 * i + a] = b[a];
            }
            i++;
        }
        if (n_edges_out) *n_edges_out = (int64_t)mesh->nEdges();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// The input side of the path: geometry-central's own mesh reader as src/main.cpp:269 calls it (readSurfaceMesh ->
// SimplePolygonMesh OBJ/PLY/... parser, stripUnusedVertices), then the same host quantities as gcref_mesh_sources.
// Call with the output arrays NULL to get the counts.
int gcref_read_mesh(const char* path, int64_t* nV_out, int64_t* nF_out, double* centroid_out, double* radius_out,
                    double* h_out, double* area_out, double* normal_out, double* bary_out) {
    try {
        std::unique_ptr<SurfaceMesh> mesh;
        std::unique_ptr<VertexPositionGeometry> geometry;
        std::tie(mesh, geometry) = readSurfaceMesh(std::string(path));
        *nV_out = (int64_t)mesh->nVertices();
        *nF_out = (int64_t)mesh->nFaces();
        Vector3 c = centroid(*geometry);
        for (int a = 0; a < 3; a++) centroid_out[a] = c[a];
        *radius_out = radius(*geometry, c);
        *h_out = meanEdgeLength(*geometry);
        if (!area_out) return 0;
        FaceData<double> areas;
        FaceData<Vector3> normals;
        setFaceVectorAreas(*geometry, areas, normals);
        size_t i = 0;
        for (Face f : mesh->faces()) {
            Vector3 b = {0, 0, 0};
            for (Vertex v : f.adjacentVertices()) b += geometry->vertexPositions[v];
            b /= f.degree();
            area_out[i] = areas[f];
            for (int a = 0; a < 3; a++) {
                normal_out[3 * i + a] = normals[f][a];
                bary_out[3 * i + a] = b[a];
            }
            i++;
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// Intermediate results of the same pipeline, for diagnosing differences: the k nearest neighbours of every point
// (nanoflann order) and the triangle soup of all local triangulations (point_position_geometry.cpp:166-169).
// neighbors_out: int64[nP][k]; tris_out: int64[capacity][3]; returns the number of soup triangles in *n_tris_out.
int gcref_local_triangulations(const double* P, const double* N, int64_t nP, int64_t k, int64_t* neighbors_out,
                               int64_t* tris_out, int64_t capacity, int64_t* n_tris_out) {
    try {
        pointcloud::PointCloud cloud((size_t)nP);
        pointcloud::PointData<Vector3> pointPositions(cloud), pointNormals(cloud);
        for (int64_t i = 0; i < nP; i++) {
            pointPositions[(size_t)i] = Vector3{P[3 * i], P[3 * i + 1], P[3 * i + 2]};
            pointNormals[(size_t)i] = Vector3{N[3 * i], N[3 * i + 1], N[3 * i + 2]};
        }
        pointcloud::PointPositionNormalGeometry geom(cloud, pointPositions, pointNormals);
        geom.requireNeighbors();
        for (int64_t i = 0; i < nP; i++) {
            const std::vector<pointcloud::Point>& nb = geom.neighbors->neighbors[(size_t)i];
            for (int64_t j = 0; j < k; j++) neighbors_out[i * k + j] = j < (int64_t)nb.size() ? (int64_t)nb[(size_t)j].getIndex() : -1;
        }
        geom.requireTangentCoordinates();
        pointcloud::PointData<std::vector<std::array<pointcloud::Point, 3>>> local = pointcloud::buildLocalTriangulations(cloud, geom, true);
        std::vector<std::vector<size_t>> all = pointcloud::handleToFlatInds(cloud, local);
        *n_tris_out = (int64_t)all.size();
        if ((int64_t)all.size() > capacity) return 2;
        for (size_t t = 0; t < all.size(); t++)
            for (int a = 0; a < 3; a++) tris_out[3 * t + a] = (int64_t)all[t][a];
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
}
