// host_api.cu -- host-side half of the reference interface for the grid path,
// dependency-free (no geometry-central / Eigen / polyscope), compiled into libshm3d_grid.so.
//
//   centroid / radius           src/signed_heat_3d.cpp:3-43
//   grid set-up                 src/signed_heat_grid_solver.cpp:13-26 (mesh) / :124-137 (points)
//   meanEdgeLength -> lambda    src/signed_heat_3d.cpp:51-60, src/signed_heat_grid_solver.cpp:42-44
//   setFaceVectorAreas          src/signed_heat_3d.cpp:62-89  (shoelace vector area; polygons allowed)
//   barycenter                  src/signed_heat_grid_solver.cpp:498-503
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

using namespace shm3d;

extern "C" {

int shm3d_prepare_mesh(const double* V, int64_t nV, const int64_t* face_vertices, const int64_t* face_offsets,
                       int64_t nF, double tCoef, double hCoef, double scale, shm3d_params* out, double* pos_out,
                       double* nrm_out, double* area_out, double* h_out) {
    if (!V || !face_vertices || !face_offsets || !out || nV <= 0 || nF <= 0) return SHM3D_ERR_INVALID_ARG;
    // centroid over mesh vertices, radius = max distance to it
    double c[3] = {0, 0, 0};
    for (int64_t v = 0; v < nV; v++)
        for (int a = 0; a < 3; a++) c[a] += V[3 * v + a];
    for (int a = 0; a < 3; a++) c[a] /= (double)nV;
    double r = 0;
    for (int64_t v = 0; v < nV; v++) {
        double d0 = c[0] - V[3 * v], d1 = c[1] - V[3 * v + 1], d2 = c[2] - V[3 * v + 2];
        r = std::max(r, std::sqrt(d0 * d0 + d1 * d1 + d2 * d2));
    }
    // mean length of the unique (unordered) vertex-pair edges, visited in (min vertex, max vertex) order: counting sort on
    // the smaller vertex, then each vertex's handful of partners sorted and deduplicated (a global sort of the 3e5 pairs of a
    // 1e5-triangle mesh was 30 of the 40 ms this function took)
    std::vector<int64_t> start((size_t)nV + 1, 0);
    for (int64_t f = 0; f < nF; f++) {
        int64_t b = face_offsets[f], e = face_offsets[f + 1], d = e - b;
        if (d < 3) return SHM3D_ERR_INVALID_ARG;
        for (int64_t t = 0; t < d; t++) {
            int64_t va = face_vertices[b + t], vb = face_vertices[b + (t + 1) % d];
            if (va < 0 || vb < 0 || va >= nV || vb >= nV) return SHM3D_ERR_INVALID_ARG;
            start[(size_t)std::min(va, vb) + 1]++;
        }
    }
    for (int64_t v = 0; v < nV; v++) start[(size_t)v + 1] += start[(size_t)v];
    std::vector<int64_t> partner((size_t)start[(size_t)nV]);
    {
        std::vector<int64_t> fill(start.begin(), start.end() - 1);
        for (int64_t f = 0; f < nF; f++) {
            int64_t b = face_offsets[f], d = face_offsets[f + 1] - b;
            for (int64_t t = 0; t < d; t++) {
                int64_t va = face_vertices[b + t], vb = face_vertices[b + (t + 1) % d];
                partner[(size_t)fill[(size_t)std::min(va, vb)]++] = std::max(va, vb);
            }
        }
    }
    double hsum = 0;
    int64_t n_edges = 0;
    for (int64_t v = 0; v < nV; v++) {
        int64_t* pb = partner.data() + start[(size_t)v];
        int64_t* pe = partner.data() + start[(size_t)v + 1];
        std::sort(pb, pe);
        pe = std::unique(pb, pe);
        const double* a = V + 3 * v;
        for (int64_t* q = pb; q < pe; q++) {
            const double* b = V + 3 * *q;
            hsum += std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
            n_edges++;
        }
    }
    const double h = hsum / (double)n_edges;
    if (h_out) *h_out = h;

    const double s = r * scale;
    memset(out, 0, sizeof(*out));
    const int nx = (int)(size_t)(2 * std::pow(2.0, hCoef + 3));  // size_t truncation like the reference
    out->nx = out->ny = out->nz = nx;
    for (int a = 0; a < 3; a++) out->bbox_min[a] = c[a] - s;
    out->cell = 2. * s / (nx - 1);
    out->lambda = std::sqrt(1. / (tCoef * h * h));
    // the mesh overload scrubs non-finite rhs entries (:72-74); Step 2 follows the reference's double-precision
    // X /= X.norm() where it underflows (:61), so the drop-in returns what the reference returns
    out->flags = SHM3D_FLAG_SCRUB_NONFINITE | SHM3D_FLAG_FP64_UNDERFLOW;

    if (pos_out && nrm_out && area_out) {
        for (int64_t f = 0; f < nF; f++) {
            int64_t b = face_offsets[f], d = face_offsets[f + 1] - b;
            double N[3] = {0, 0, 0}, y[3] = {0, 0, 0};
            for (int64_t t = 0; t < d; t++) {
                const double* pa = V + 3 * face_vertices[b + t];
                const double* pb = V + 3 * face_vertices[b + (t + 1) % d];
                N[0] += pa[1] * pb[2] - pa[2] * pb[1];
                N[1] += pa[2] * pb[0] - pa[0] * pb[2];
                N[2] += pa[0] * pb[1] - pa[1] * pb[0];
                for (int a = 0; a < 3; a++) y[a] += pa[a];
            }
            for (int a = 0; a < 3; a++) N[a] *= 0.5;
            double A = std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
            area_out[f] = A;
            for (int a = 0; a < 3; a++) {
                nrm_out[3 * f + a] = N[a] / A;
                pos_out[3 * f + a] = y[a] / (double)d;
            }
        }
    }
    return SHM3D_OK;
}

int shm3d_prepare_points(const double* P, int64_t nP, double h, double tCoef, double hCoef, double scale,
                         shm3d_params* out) {
    if (!P || !out || nP <= 0 || !(h > 0)) return SHM3D_ERR_INVALID_ARG;
    double c[3] = {0, 0, 0};
    for (int64_t v = 0; v < nP; v++)
        for (int a = 0; a < 3; a++) c[a] += P[3 * v + a];
    for (int a = 0; a < 3; a++) c[a] /= (double)nP;
    double r = 0;
    for (int64_t v = 0; v < nP; v++) {
        double d0 = c[0] - P[3 * v], d1 = c[1] - P[3 * v + 1], d2 = c[2] - P[3 * v + 2];
        r = std::max(r, std::sqrt(d0 * d0 + d1 * d1 + d2 * d2));
    }
    const double s = r * scale;
    memset(out, 0, sizeof(*out));
    const int nx = (int)(size_t)(2 * std::pow(2.0, hCoef + 3));
    out->nx = out->ny = out->nz = nx;
    for (int a = 0; a < 3; a++) out->bbox_min[a] = c[a] - s;
    out->cell = 2. * s / (nx - 1);
    out->lambda = std::sqrt(1. / (tCoef * h * h));
    // the point overload does not scrub (src/signed_heat_grid_solver.cpp:180): where the reference's X.norm() underflows
    // (:171) its solve throws on the non-finite right-hand side, and so does this one
    out->flags = SHM3D_FLAG_FP64_UNDERFLOW;
    return SHM3D_OK;
}

}  // extern "C"
