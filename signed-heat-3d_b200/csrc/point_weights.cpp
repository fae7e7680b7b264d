// point_weights.cpp -- host-side source weights for the point-cloud overload (partial).
//
// The reference reads two things from geometry-central's tufted triangulation of the cloud
// (src/signed_heat_grid_solver.cpp:149-151,165): per-point vertex dual areas and the mean edge length h.
// geometry-central builds them as (deps/geometry-central/src/pointcloud/point_position_geometry.cpp:160-190):
//   kNN(30) -> tangent-plane coordinates -> local Delaunay 1-ring of every point (local_triangulation.cpp:10-210)
//   -> the union of all local triangles as a triangle soup -> intrinsic mollification (1e-5) -> tufted cover
//   -> intrinsic edge flips to Delaunay -> vertexDualAreas = sum of incident face areas / 3, mean intrinsic edge length.
// Restated here: everything up to and including the mollification, and the two sheets of the cover (a factor 2).
// NOT restated: the intrinsic flips on the cover -- they preserve the total area and move little of it between
// neighbouring points when the local triangulations agree with each other (well-sampled surfaces); the result is
// therefore an approximation of geometry-central's weights, stated as such in DESIGN.md / INTEGRATION.md.
// Any positive rescaling of all areas cancels in Steps 1-2 (normalisation) and in the shift (a weighted mean).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../../include/shm3d_grid.h"

namespace {

struct V2 {
    double x, y;
};
inline double cross2(const V2& a, const V2& b) { return a.x * b.y - a.y * b.x; }
inline double dot2(const V2& a, const V2& b) { return a.x * b.x + a.y * b.y; }
inline double norm2v(const V2& a) { return a.x * a.x + a.y * a.y; }

double det3(double a, double b, double c, double d, double e, double f, double g, double h, double i) {
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}
// deps/geometry-central/src/utilities/elementary_geometry.cpp:8-18: det of rows (x, y, |p|^2, 1) > 0
bool in_circle(const V2& A, const V2& B, const V2& C, const V2& T) {
    const double a2 = norm2v(A), b2 = norm2v(B), c2 = norm2v(C), t2 = norm2v(T);
    // expand along the last column (all ones)
    const double d = -det3(B.x, B.y, b2, C.x, C.y, c2, T.x, T.y, t2) + det3(A.x, A.y, a2, C.x, C.y, c2, T.x, T.y, t2) -
                     det3(A.x, A.y, a2, B.x, B.y, b2, T.x, T.y, t2) + det3(A.x, A.y, a2, B.x, B.y, b2, C.x, C.y, c2);
    return d > 0.;
}

const size_t kInvalid = (size_t)-1;

// local_triangulation.cpp:10-210 for ONE point: `pts` are the tangent-plane coordinates of its neighbours (the point
// itself sits at the origin).  Returns the surviving neighbours in counter-clockwise order and, per consecutive pair,
// whether a triangle (origin, ring[i], ring[i+1]) is emitted.
void local_ring(std::vector<V2> pts, std::vector<size_t>& ring, std::vector<char>& tri_after) {
    ring.clear();
    tri_after.clear();
    const size_t n = pts.size();
    const double THRESH = 1e-7;
    double len2 = 0;
    for (const V2& p : pts) len2 = std::fmax(len2, norm2v(p));
    const double lenScale = std::sqrt(len2);
    if (!std::isfinite(lenScale) || lenScale <= 0) return;  // hopelessly degenerate neighbourhood: no triangles
    for (size_t i = 0; i < n; i++) {                         // perturb points (nearly) on top of the centre (:52-70)
        V2& q = pts[i];
        const double dist = std::sqrt(norm2v(q));
        if (dist < lenScale * THRESH) {
            V2 dir{q.x / dist, q.y / dist};
            if (!std::isfinite(dir.x) || !std::isfinite(dir.y)) {
                const double th = (2. * M_PI * (double)i) / (double)n;
                dir = V2{std::cos(th), std::sin(th)};
            }
            const double len = (1. + (double)i / (double)n) * lenScale * THRESH * 10;
            q = V2{len * dir.x, len * dir.y};
        }
    }
    std::vector<size_t> sortInds(n);
    std::vector<double> ang(n);
    const double BAD = -777;
    for (size_t i = 0; i < n; i++) {
        const double l = std::sqrt(norm2v(pts[i]));
        double a = std::atan2(pts[i].y / l, pts[i].x / l);
        if (!std::isfinite(a)) a = BAD;
        sortInds[i] = i;
        ang[i] = a;
    }
    std::sort(sortInds.begin(), sortInds.end(), [&](size_t a, size_t b) { return ang[a] < ang[b]; });
    for (size_t i = 0; i < n; i++)
        if (ang[sortInds[i]] == BAD) sortInds[i] = kInvalid;
    auto is_boundary = [&](size_t a, size_t b) { return cross2(pts[a], pts[b]) <= 0.; };
    const V2 origin{0., 0.};
    bool changed = true;
    while (changed) {  // discard (= flip away) points until the star of the centre is Delaunay (:118-176)
        changed = false;
        for (size_t iM = 0; iM < n; iM++) {
            if (sortInds[iM] == kInvalid) continue;
            size_t iP = iM, iN = iM;
            size_t guard = 0;
            do { iP = (iP + n - 1) % n; } while (sortInds[iP] == kInvalid && ++guard < n);
            guard = 0;
            do { iN = (iN + 1) % n; } while (sortInds[iN] == kInvalid && ++guard < n);
            const size_t prev = sortInds[iP], curr = sortInds[iM], next = sortInds[iN];
            if (prev == kInvalid || next == kInvalid) continue;
            if (curr == prev || curr == next || prev == next) continue;
            {   // collinear points: keep only the closest (:146-161)
                const double lp = std::sqrt(norm2v(pts[prev])), lc = std::sqrt(norm2v(pts[curr])), ln = std::sqrt(norm2v(pts[next]));
                const bool colP = std::fabs(cross2(pts[curr], pts[prev])) < (lp * lc) * THRESH && dot2(pts[curr], pts[prev]) > 0;
                const bool colN = std::fabs(cross2(pts[curr], pts[next])) < (ln * lc) * THRESH && dot2(pts[curr], pts[next]) > 0;
                if ((colN && lc > ln) || (colP && lc > lp)) {
                    sortInds[iM] = kInvalid;
                    changed = true;
                    continue;
                }
            }
            if (is_boundary(prev, curr) || is_boundary(curr, next)) continue;
            if (!in_circle(origin, pts[prev], pts[next], pts[curr])) {
                sortInds[iM] = kInvalid;
                changed = true;
            }
        }
    }
    for (size_t i = 0; i < n; i++)
        if (sortInds[i] != kInvalid) ring.push_back(sortInds[i]);
    tri_after.assign(ring.size(), 0);
    if (ring.size() < 2) return;
    for (size_t i = 0; i < ring.size(); i++) {
        const size_t a = ring[i], b = ring[(i + 1) % ring.size()];
        if (a != b && !is_boundary(a, b)) tri_after[i] = 1;
    }
}

// exact k nearest neighbours (self excluded), ascending distance then index -- a uniform hash grid
void knn_all(const double* P, int64_t n, int k, std::vector<int64_t>& nbr) {
    nbr.assign((size_t)n * k, -1);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], P[3 * i + a]);
            hi[a] = std::max(hi[a], P[3 * i + a]);
        }
    const double ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-300});
    const int g = std::max(1, (int)std::cbrt((double)n / 4.0));
    const double cs = ext / g * (1 + 1e-12);
    auto cell_of = [&](const double* q, int c[3]) {
        for (int a = 0; a < 3; a++) c[a] = std::min(g - 1, std::max(0, (int)((q[a] - lo[a]) / cs)));
    };
    std::vector<std::vector<int64_t>> bucket((size_t)g * g * g);
    for (int64_t i = 0; i < n; i++) {
        int c[3];
        cell_of(P + 3 * i, c);
        bucket[(size_t)c[0] + (size_t)c[1] * g + (size_t)c[2] * g * g].push_back(i);
    }
    std::vector<std::pair<double, int64_t>> cand;
    for (int64_t i = 0; i < n; i++) {
        int c[3];
        cell_of(P + 3 * i, c);
        cand.clear();
        for (int ring = 0; ring <= g; ring++) {
            for (int dz = -ring; dz <= ring; dz++)
                for (int dy = -ring; dy <= ring; dy++)
                    for (int dx = -ring; dx <= ring; dx++) {
                        if (std::max({std::abs(dx), std::abs(dy), std::abs(dz)}) != ring) continue;
                        const int x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
                        if (x < 0 || y < 0 || z < 0 || x >= g || y >= g || z >= g) continue;
                        for (int64_t j : bucket[(size_t)x + (size_t)y * g + (size_t)z * g * g]) {
                            if (j == i) continue;
                            double d = 0;
                            for (int a = 0; a < 3; a++) {
                                const double t = P[3 * i + a] - P[3 * j + a];
                                d += t * t;
                            }
                            cand.emplace_back(d, j);
                        }
                    }
            if ((int)cand.size() >= k) {
                std::nth_element(cand.begin(), cand.begin() + (k - 1), cand.end());
                // everything in the rings searched so far that is closer than ring*cs is final
                if (std::sqrt(cand[k - 1].first) <= ring * cs) break;
            }
        }
        std::sort(cand.begin(), cand.end());
        for (int t = 0; t < k; t++) nbr[(size_t)i * k + t] = cand[t].second;
    }
}

}  // namespace

extern "C" {

int shm3d_debug_local_ring(const double* coords2d, int32_t n, int32_t* ring_out, int32_t* tri_after_out) {
    if (!coords2d || n < 0 || !ring_out) return -1;
    std::vector<V2> pts((size_t)n);
    for (int i = 0; i < n; i++) pts[i] = V2{coords2d[2 * i], coords2d[2 * i + 1]};
    std::vector<size_t> ring;
    std::vector<char> tri;
    local_ring(pts, ring, tri);
    for (size_t i = 0; i < ring.size(); i++) {
        ring_out[i] = (int32_t)ring[i];
        if (tri_after_out) tri_after_out[i] = tri[i];
    }
    return (int)ring.size();
}

int shm3d_point_weights(const double* P, const double* N, int64_t nP, int32_t k_neighbors, double* areas_out,
                        double* h_out, int64_t* n_triangles_out) {
    if (!P || !N || !areas_out || !h_out || nP <= 0) return SHM3D_ERR_INVALID_ARG;
    const int k = k_neighbors > 0 ? k_neighbors : 30;  // PointPositionGeometry::kNeighborSize
    if ((int64_t)k + 1 > nP) return SHM3D_ERR_INVALID_ARG;  // "k+1 is greater than number of points" (knn.cpp:53)
    for (int64_t i = 0; i < 3 * nP; i++)
        if (!std::isfinite(P[i]) || !std::isfinite(N[i])) return SHM3D_ERR_NONFINITE;
    std::vector<int64_t> nbr;
    knn_all(P, nP, k, nbr);

    // soup triangles (p, a, b) from every point's local triangulation
    std::vector<int64_t> tris;
    std::vector<V2> pts((size_t)k);
    std::vector<size_t> ring;
    std::vector<char> tri_after;
    for (int64_t p = 0; p < nP; p++) {
        const double* c = P + 3 * p;
        const double nl = std::sqrt(N[3 * p] * N[3 * p] + N[3 * p + 1] * N[3 * p + 1] + N[3 * p + 2] * N[3 * p + 2]);
        const double nrm[3] = {N[3 * p], N[3 * p + 1], N[3 * p + 2]};
        const double u[3] = {nrm[0] / nl, nrm[1] / nl, nrm[2] / nl};
        // Vector3::buildTangentBasis (utilities/vector3.ipp:148-159)
        double t[3] = {1., 0., 0.};
        if (std::fabs(u[0]) > 0.9) { t[0] = 0.; t[1] = 1.; }
        double bx[3] = {t[1] * u[2] - t[2] * u[1], t[2] * u[0] - t[0] * u[2], t[0] * u[1] - t[1] * u[0]};
        double l = std::sqrt(bx[0] * bx[0] + bx[1] * bx[1] + bx[2] * bx[2]);
        for (double& v : bx) v /= l;
        double by[3] = {u[1] * bx[2] - u[2] * bx[1], u[2] * bx[0] - u[0] * bx[2], u[0] * bx[1] - u[1] * bx[0]};
        l = std::sqrt(by[0] * by[0] + by[1] * by[1] + by[2] * by[2]);
        for (double& v : by) v /= l;
        for (int j = 0; j < k; j++) {  // tangent coordinates (point_position_geometry.cpp:113-134)
            const double* q = P + 3 * nbr[(size_t)p * k + j];
            double v[3] = {q[0] - c[0], q[1] - c[1], q[2] - c[2]};
            const double dn = nrm[0] * v[0] + nrm[1] * v[1] + nrm[2] * v[2];  // removeComponent(normal), normal as given
            for (int a = 0; a < 3; a++) v[a] -= nrm[a] * dn;
            pts[j] = V2{bx[0] * v[0] + bx[1] * v[1] + bx[2] * v[2], by[0] * v[0] + by[1] * v[1] + by[2] * v[2]};
        }
        local_ring(pts, ring, tri_after);
        for (size_t i = 0; i < ring.size(); i++)
            if (tri_after[i]) {
                tris.push_back(p);
                tris.push_back(nbr[(size_t)p * k + ring[i]]);
                tris.push_back(nbr[(size_t)p * k + ring[(i + 1) % ring.size()]]);
            }
    }
    const int64_t T = (int64_t)tris.size() / 3;
    if (n_triangles_out) *n_triangles_out = T;
    for (int64_t i = 0; i < nP; i++) areas_out[i] = 0;
    if (T == 0) {
        *h_out = 0;
        return SHM3D_ERR_INVALID_ARG;
    }
    // 3-D edge lengths per soup triangle, intrinsic mollification (intrinsic_mollification.cpp:7-38): every length
    // grows by eps = max(0, max over corners of lC - lA - lB + delta), delta = 1e-5 * mean edge length
    std::vector<double> len((size_t)3 * T);
    double perim = 0;
    for (int64_t f = 0; f < T; f++)
        for (int e = 0; e < 3; e++) {
            const double* a = P + 3 * tris[3 * f + e];
            const double* b = P + 3 * tris[3 * f + (e + 1) % 3];
            const double d = std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
            len[3 * f + e] = d;
            perim += d;
        }
    const double delta = (perim / (3.0 * T)) * 1e-5;
    double eps = 0;
    for (int64_t f = 0; f < T; f++)
        for (int e = 0; e < 3; e++)
            eps = std::fmax(eps, len[3 * f + (e + 2) % 3] - len[3 * f + e] - len[3 * f + (e + 1) % 3] + delta);
    double hsum = 0;
    for (int64_t f = 0; f < T; f++) {
        const double a = len[3 * f] + eps, b = len[3 * f + 1] + eps, c = len[3 * f + 2] + eps;
        const double s = 0.5 * (a + b + c);
        const double area = std::sqrt(std::max(0.0, s * (s - a) * (s - b) * (s - c)));  // Heron on the intrinsic lengths
        for (int e = 0; e < 3; e++) areas_out[tris[3 * f + e]] += 2.0 * area / 3.0;     // two sheets of the tufted cover
        hsum += a + b + c;
    }
    *h_out = hsum / (3.0 * T);  // cover edges: 3T, each carrying the length of the triangle side it came from
    return SHM3D_OK;
}

}  // extern "C"
