// point_weights.cpp -- host-side source weights for the point-cloud overload.
//
// The reference reads two things from geometry-central's tufted triangulation of the cloud
// (src/signed_heat_grid_solver.cpp:149-151,165): per-point vertex dual areas and the mean edge length h.
// geometry-central builds them as (deps/geometry-central/src/pointcloud/point_position_geometry.cpp:160-190):
//   kNN(30) -> tangent-plane coordinates -> local Delaunay 1-ring of every point (local_triangulation.cpp:10-210)
//   -> the union of all local triangles as a triangle soup -> intrinsic mollification (1e-5) -> tufted cover
//   -> intrinsic edge flips to Delaunay -> vertexDualAreas = sum of incident face areas / 3, mean intrinsic edge length.
// All of it is restated here from the cited sources: the soup mesh's edge / sibling order (surface_mesh.cpp:60-205), the
// mollification (intrinsic_mollification.cpp:7-38), the gluing rule of the cover (tufted_laplacian.cpp:39-121, the
// position-free "natural ordering" branch the point-cloud path takes), the Euclidean intrinsic flips (simple_idt.cpp:11-188,
// SurfaceMesh::flip :848-928), and -- because structured inputs make every discrete decision a tie -- also the things that
// decide ties: nanoflann's kd-tree (visiting order = which of several equidistant points is the 30th neighbour),
// geometry-central's multiply-by-reciprocal normalisations, Eigen 3.3's 4x4 determinant expression in the in-circle test,
// the numbering of the cover's edges (= order of the flip queue).  This file is compiled WITHOUT floating-point contraction
// (csrc/build.sh) so that those expressions round the same on every host.
// Checked against geometry-central's own sources (compiled from the reference tree against an Eigen stub,
// oracle/_ref/libshm_gc_ref.so): areas and h equal to <= 1e-12 on all of the reference's sample clouds and on lattice /
// duplicated / quantised clouds (tests/test_point_weights.py); independent checks: every local star against scipy's
// Delaunay triangulation, total area preserved by the flips, the final cover intrinsically Delaunay, closed forms on
// sampled spheres.  Any positive rescaling of all areas cancels in Steps 1-2 (normalisation) and in the shift.
// kNN queries and local triangulations run on the host threads (chunks assembled in point order: thread-count independent).
//
// THIS FILE IS A PORT, not a re-design: `class KdTree` follows nanoflann (BSD 2-clause; Copyright 2008-2009 Marius Muja,
// David G. Lowe; 2011-2016 Jose Luis Blanco) and `local_ring` follows geometry-central's local_triangulation.cpp (MIT;
// Copyright 2017-2019 Nicholas Sharp and the geometry-central contributors) statement by statement; licence texts in
// THIRD_PARTY_NOTICES.md at the repository root.  Why a port: measured with an independent kNN (knn_all_cells below,
// ties by point index instead of the kd-tree's visiting order) four of the reference's five sample clouds give
// bit-identical weights, but data/SprayBottle.pc -- the vertices of a structured mesh, full of exactly equidistant
// neighbours -- changes h by 6e-5, single areas by up to 88 %, and phi by 1.8e-4 relative L2 (fp64 oracle, 32^3 and
// 64^3): above the 1e-4 parity bar.  Identity with the reference therefore needs nanoflann's tie order itself.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

#include "../../include/shm3d_grid.h"

namespace {

struct V2 {
    double x, y;
};
inline double cross2(const V2& a, const V2& b) { return a.x * b.y - a.y * b.x; }
inline double dot2(const V2& a, const V2& b) { return a.x * b.x + a.y * b.y; }
inline double norm2v(const V2& a) { return a.x * a.x + a.y * a.y; }

// deps/geometry-central/src/utilities/elementary_geometry.cpp:8-18: sign of the 4x4 determinant with rows
// (x, y, |p|^2, 1).  For (nearly) cocircular points -- vertices of structured meshes -- that sign is decided by rounding,
// so the determinant is evaluated by the very expression Eigen 3.3 (geometry-central's pin) uses for fixed 4x4 matrices
// (Eigen/src/LU/Determinant.h, bruteforce_det4_helper: 2x2 minors of columns 0-1 times 2x2 minors of columns 2-3), with
// floating-point contraction off so that the host compiler's FMA choices cannot change it.
bool in_circle(const V2& A, const V2& B, const V2& C, const V2& T) {
    const double m[4][4] = {{A.x, A.y, norm2v(A), 1.}, {B.x, B.y, norm2v(B), 1.}, {C.x, C.y, norm2v(C), 1.}, {T.x, T.y, norm2v(T), 1.}};
    auto h = [&m](int j, int k, int a, int b) {
        return (m[j][0] * m[k][1] - m[k][0] * m[j][1]) * (m[a][2] * m[b][3] - m[b][2] * m[a][3]);
    };
    const double d = h(0, 1, 2, 3) - h(0, 2, 1, 3) + h(0, 3, 1, 2) + h(1, 2, 0, 3) - h(1, 3, 0, 2) + h(2, 3, 0, 1);
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
            const double rq = 1. / std::sqrt(q.x * q.x + q.y * q.y);  // Vector2::normalize multiplies by the reciprocal
            V2 dir{q.x * rq, q.y * rq};
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
        const double r = 1. / std::sqrt(pts[i].x * pts[i].x + pts[i].y * pts[i].y);  // arg(unit(p)), vector2.ipp:72-75,126
        double a = std::atan2(pts[i].y * r, pts[i].x * r);
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

// contiguous chunks of [0,n) on the host threads; body(begin, end, chunk_index).  Results are assembled in chunk order,
// so the output does not depend on the number of threads.
template <typename Body>
int parallel_chunks(int64_t n, Body body) {
    unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (n < 4096) T = 1;
    const int64_t chunk = (n + T - 1) / T;
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; t++)
        th.emplace_back([=, &body] { body(std::min(n, (int64_t)t * chunk), std::min(n, (int64_t)(t + 1) * chunk), (int)t); });
    body(0, std::min(n, chunk), 0);
    for (std::thread& x : th) x.join();
    return (int)T;
}
int chunk_count(int64_t n) {
    unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    return n < 4096 ? 1 : (int)T;
}

// k nearest neighbours of every point, self excluded, EXACTLY as geometry-central gets them
// (NearestNeighborFinder::kNearestNeighbors, src/utilities/knn.cpp:49-72: a (k+1)-search around the point itself, then the
// point is erased from the list).  Point sets with exactly equidistant neighbours -- the vertices of structured meshes --
// make the k-th neighbour a matter of tie-breaking, and nanoflann (the library behind it; vendored by the reference,
// deps/geometry-central/deps/nanoflann/include/nanoflann.hpp) breaks ties by the order in which its kd-tree visits
// points.  So the tree and the search are restated here step by step:
//   build   KDTreeBaseClass::divideTree / middleSplit_ / planeSplit (:858-1003), leaf size 10, root box = data bounds
//   search  KDTreeSingleIndexAdaptor::searchLevel (:1347-1411) with KNNResultSet::addPoint (:175-202): nearer child first,
//           the other one if its box distance does not exceed the current worst; inside a leaf the worst distance is read
//           once; a candidate equal to a stored distance goes behind it, and is dropped when the set is full
//   metric  L2_Simple_Adaptor (:432-445): sum of squared differences, x then y then z
class KdTree {
  public:
    KdTree(const double* P, int64_t n) : P_(P), n_((size_t)n), vind_((size_t)n) {
        for (size_t i = 0; i < n_; i++) vind_[i] = i;
        for (int a = 0; a < 3; a++) root_[a].low = root_[a].high = pt(0, a);  // computeBoundingBox (:1313-1335)
        for (size_t k = 1; k < n_; k++)
            for (int a = 0; a < 3; a++) {
                if (pt(k, a) < root_[a].low) root_[a].low = pt(k, a);
                if (pt(k, a) > root_[a].high) root_[a].high = pt(k, a);
            }
        nodes_.reserve(n_ / 4 + 16);
        root_node_ = divide(0, n_, root_);
    }
    // indices of the `count` nearest points to q (ascending distance, ties in visiting order), like knnSearch (:1251-1258)
    void knn(const double* q, size_t count, size_t* idx, double* dist) const {
        Result r{idx, dist, count, 0};
        dist[count - 1] = std::numeric_limits<double>::max();
        double dists[3] = {0., 0., 0.};
        double distsq = 0;  // computeInitialDistances (:1006-1023)
        for (int a = 0; a < 3; a++) {
            if (q[a] < root_[a].low) {
                dists[a] = (q[a] - root_[a].low) * (q[a] - root_[a].low);
                distsq += dists[a];
            }
            if (q[a] > root_[a].high) {
                dists[a] = (q[a] - root_[a].high) * (q[a] - root_[a].high);
                distsq += dists[a];
            }
        }
        search(r, q, root_node_, distsq, dists);
    }

  private:
    struct Interval {
        double low, high;
    };
    typedef std::array<Interval, 3> Box;
    struct Node {
        int64_t child1 = -1, child2 = -1;  // -1, -1: leaf
        size_t left = 0, right = 0;        // leaf: vind_[left .. right)
        int divfeat = 0;
        double divlow = 0, divhigh = 0;
    };
    struct Result {
        size_t* indices;
        double* dists;
        size_t capacity, count;
        double worst() const { return dists[capacity - 1]; }
        void add(double dist, size_t index) {
            size_t i;
            for (i = count; i > 0; --i) {
                if (dists[i - 1] > dist) {
                    if (i < capacity) {
                        dists[i] = dists[i - 1];
                        indices[i] = indices[i - 1];
                    }
                } else
                    break;
            }
            if (i < capacity) {
                dists[i] = dist;
                indices[i] = index;
            }
            if (count < capacity) count++;
        }
    };
    const double* P_;
    size_t n_;
    std::vector<size_t> vind_;
    std::vector<Node> nodes_;
    Box root_;
    int64_t root_node_ = -1;
    static constexpr size_t kLeafMax = 10;

    double pt(size_t i, int a) const { return P_[3 * i + a]; }
    void min_max(const size_t* ind, size_t count, int a, double& lo, double& hi) const {
        lo = hi = pt(ind[0], a);
        for (size_t i = 1; i < count; ++i) {
            const double v = pt(ind[i], a);
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
    }
    int64_t divide(size_t left, size_t right, Box& bbox) {
        const int64_t me = (int64_t)nodes_.size();
        nodes_.emplace_back();
        if (right - left <= kLeafMax) {
            nodes_[(size_t)me].left = left;
            nodes_[(size_t)me].right = right;
            for (int a = 0; a < 3; a++) bbox[a].low = bbox[a].high = pt(vind_[left], a);
            for (size_t k = left + 1; k < right; ++k)
                for (int a = 0; a < 3; a++) {
                    if (bbox[a].low > pt(vind_[k], a)) bbox[a].low = pt(vind_[k], a);
                    if (bbox[a].high < pt(vind_[k], a)) bbox[a].high = pt(vind_[k], a);
                }
            return me;
        }
        size_t idx;
        int cutfeat;
        double cutval;
        middle_split(&vind_[0] + left, right - left, idx, cutfeat, cutval, bbox);
        Box lb(bbox);
        lb[cutfeat].high = cutval;
        const int64_t c1 = divide(left, left + idx, lb);
        Box rb(bbox);
        rb[cutfeat].low = cutval;
        const int64_t c2 = divide(left + idx, right, rb);
        Node& nd = nodes_[(size_t)me];
        nd.child1 = c1;
        nd.child2 = c2;
        nd.divfeat = cutfeat;
        nd.divlow = lb[cutfeat].high;
        nd.divhigh = rb[cutfeat].low;
        for (int a = 0; a < 3; a++) {
            bbox[a].low = std::min(lb[a].low, rb[a].low);
            bbox[a].high = std::max(lb[a].high, rb[a].high);
        }
        return me;
    }
    void middle_split(size_t* ind, size_t count, size_t& index, int& cutfeat, double& cutval, const Box& bbox) const {
        const double EPS = static_cast<double>(0.00001);
        double max_span = bbox[0].high - bbox[0].low;
        for (int a = 1; a < 3; a++) {
            const double span = bbox[a].high - bbox[a].low;
            if (span > max_span) max_span = span;
        }
        double max_spread = -1;
        cutfeat = 0;
        for (int a = 0; a < 3; a++) {
            const double span = bbox[a].high - bbox[a].low;
            if (span > (1 - EPS) * max_span) {
                double lo, hi;
                min_max(ind, count, a, lo, hi);
                const double spread = hi - lo;
                if (spread > max_spread) {
                    cutfeat = a;
                    max_spread = spread;
                }
            }
        }
        const double split_val = (bbox[cutfeat].low + bbox[cutfeat].high) / 2;
        double lo, hi;
        min_max(ind, count, cutfeat, lo, hi);
        if (split_val < lo) cutval = lo;
        else if (split_val > hi) cutval = hi;
        else cutval = split_val;
        size_t lim1, lim2;
        plane_split(ind, count, cutfeat, cutval, lim1, lim2);
        if (lim1 > count / 2) index = lim1;
        else if (lim2 < count / 2) index = lim2;
        else index = count / 2;
    }
    // unsigned index arithmetic on purpose (the "!right" exits are part of the library's behaviour)
    void plane_split(size_t* ind, const size_t count, int cutfeat, double& cutval, size_t& lim1, size_t& lim2) const {
        size_t left = 0;
        size_t right = count - 1;
        for (;;) {
            while (left <= right && pt(ind[left], cutfeat) < cutval) ++left;
            while (right && left <= right && pt(ind[right], cutfeat) >= cutval) --right;
            if (left > right || !right) break;
            std::swap(ind[left], ind[right]);
            ++left;
            --right;
        }
        lim1 = left;
        right = count - 1;
        for (;;) {
            while (left <= right && pt(ind[left], cutfeat) <= cutval) ++left;
            while (right && left <= right && pt(ind[right], cutfeat) > cutval) --right;
            if (left > right || !right) break;
            std::swap(ind[left], ind[right]);
            ++left;
            --right;
        }
        lim2 = left;
    }
    void search(Result& r, const double* q, int64_t node, double mindistsq, double* dists) const {
        const Node& nd = nodes_[(size_t)node];
        if (nd.child1 < 0 && nd.child2 < 0) {
            const double worst = r.worst();
            for (size_t i = nd.left; i < nd.right; ++i) {
                const size_t index = vind_[i];
                double dist = 0;
                for (int a = 0; a < 3; a++) {
                    const double diff = q[a] - pt(index, a);
                    dist += diff * diff;
                }
                if (dist < worst) r.add(dist, index);
            }
            return;
        }
        const int idx = nd.divfeat;
        const double val = q[idx];
        const double diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int64_t best, other;
        double cut_dist;
        if ((diff1 + diff2) < 0) {
            best = nd.child1;
            other = nd.child2;
            cut_dist = (val - nd.divhigh) * (val - nd.divhigh);
        } else {
            best = nd.child2;
            other = nd.child1;
            cut_dist = (val - nd.divlow) * (val - nd.divlow);
        }
        search(r, q, best, mindistsq, dists);
        const double dst = dists[idx];
        mindistsq = mindistsq + cut_dist - dst;
        dists[idx] = cut_dist;
        if (mindistsq * 1.0f <= r.worst()) search(r, q, other, mindistsq, dists);
        dists[idx] = dst;
    }
};

void knn_all(const double* P, int64_t n, int k, std::vector<int64_t>& nbr) {
    nbr.assign((size_t)n * k, -1);
    const KdTree tree(P, n);
    parallel_chunks(n, [&](int64_t i_begin, int64_t i_end, int) {
        std::vector<size_t> idx((size_t)k + 1);
        std::vector<double> dist((size_t)k + 1);
        for (int64_t i = i_begin; i < i_end; i++) {
            tree.knn(P + 3 * i, (size_t)k + 1, idx.data(), dist.data());
            // remove the source from the list; if it did not appear, drop the last entry (knn.cpp:56-71)
            size_t self = (size_t)k;
            for (size_t t = 0; t <= (size_t)k; t++)
                if (idx[t] == (size_t)i) {
                    self = t;
                    break;
                }
            for (size_t t = 0, o = 0; t <= (size_t)k; t++)
                if (t != self) nbr[(size_t)i * k + o++] = (int64_t)idx[t];
        }
    });
}

// Own kNN: a uniform cell list over the cloud's bounding box, exact squared distances (x, y, z terms summed in that order),
// neighbours ordered by (distance, index) -- i.e. ties are broken by point index, not by a search tree's visiting order.
// Shells of cells around the query's cell are scanned until the (k+1)-th best distance is strictly inside the scanned
// block.  Same result as knn_all wherever no two candidates are exactly equidistant from the query.
void knn_all_cells(const double* P, int64_t n, int k, std::vector<int64_t>& nbr) {
    nbr.assign((size_t)n * k, -1);
    double lo[3], hi[3];
    for (int a = 0; a < 3; a++) lo[a] = hi[a] = P[a];
    for (int64_t i = 1; i < n; i++)
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], P[3 * i + a]);
            hi[a] = std::max(hi[a], P[3 * i + a]);
        }
    double ext = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    if (!(ext > 0)) ext = 1;
    const int G = (int)std::max(1.0, std::min(512.0, std::floor(std::cbrt((double)n))));
    const double cs = ext / G * (1 + 1e-12);
    int dims[3];
    for (int a = 0; a < 3; a++) dims[a] = std::max(1, std::min(G, (int)std::floor((hi[a] - lo[a]) / cs) + 1));
    auto cell_of = [&](const double* q, int* c) {
        for (int a = 0; a < 3; a++) c[a] = std::max(0, std::min(dims[a] - 1, (int)std::floor((q[a] - lo[a]) / cs)));
    };
    const size_t ncell = (size_t)dims[0] * dims[1] * dims[2];
    std::vector<int64_t> start(ncell + 1, 0), order((size_t)n);
    std::vector<int> cid((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        int c[3];
        cell_of(P + 3 * i, c);
        cid[(size_t)i] = c[0] + dims[0] * (c[1] + dims[1] * c[2]);
        start[(size_t)cid[(size_t)i] + 1]++;
    }
    for (size_t c = 0; c < ncell; c++) start[c + 1] += start[c];
    {
        std::vector<int64_t> fill(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < n; i++) order[(size_t)fill[(size_t)cid[(size_t)i]]++] = i;  // ascending index inside a cell
    }
    parallel_chunks(n, [&](int64_t i_begin, int64_t i_end, int) {
        typedef std::pair<double, int64_t> Cand;  // (squared distance, index): lexicographic order = the tie rule
        std::vector<Cand> best;                   // max-heap of the k+1 best so far
        for (int64_t i = i_begin; i < i_end; i++) {
            const double* q = P + 3 * i;
            int c[3];
            cell_of(q, c);
            best.clear();
            for (int R = 0;; R++) {
                bool any_cell = false;
                for (int z = c[2] - R; z <= c[2] + R; z++) {
                    if (z < 0 || z >= dims[2]) continue;
                    for (int y = c[1] - R; y <= c[1] + R; y++) {
                        if (y < 0 || y >= dims[1]) continue;
                        const bool inner_yz = std::abs(z - c[2]) < R && std::abs(y - c[1]) < R;
                        for (int x = c[0] - R; x <= c[0] + R; x += (inner_yz && x == c[0] - R && R > 0) ? 2 * R : 1) {
                            if (x < 0 || x >= dims[0]) continue;
                            any_cell = true;
                            const size_t cc = (size_t)x + (size_t)dims[0] * ((size_t)y + (size_t)dims[1] * (size_t)z);
                            for (int64_t t = start[cc]; t < start[cc + 1]; t++) {
                                const int64_t j = order[(size_t)t];
                                double d = 0;
                                for (int a = 0; a < 3; a++) {
                                    const double diff = q[a] - P[3 * j + a];
                                    d += diff * diff;
                                }
                                const Cand cd(d, j);
                                if ((int)best.size() < k + 1) {
                                    best.push_back(cd);
                                    std::push_heap(best.begin(), best.end());
                                } else if (cd < best.front()) {
                                    std::pop_heap(best.begin(), best.end());
                                    best.back() = cd;
                                    std::push_heap(best.begin(), best.end());
                                }
                            }
                        }
                    }
                }
                if ((int)best.size() == k + 1) {
                    // every point outside the scanned block is farther than the block's nearest face
                    double dmin = 1e300;
                    for (int a = 0; a < 3; a++) {
                        if (c[a] - R > 0) dmin = std::min(dmin, q[a] - (lo[a] + cs * (c[a] - R)));
                        if (c[a] + R < dims[a] - 1) dmin = std::min(dmin, (lo[a] + cs * (c[a] + R + 1)) - q[a]);
                    }
                    if (dmin == 1e300 || best.front().first < dmin * dmin) break;
                }
                if (!any_cell && R > std::max(dims[0], std::max(dims[1], dims[2]))) break;
            }
            std::sort(best.begin(), best.end());
            // the point itself leaves the list; if it is not in it (k+1 or more exact duplicates), the last entry does
            size_t self = best.size() - 1;
            for (size_t t = 0; t < best.size(); t++)
                if (best[t].second == i) {
                    self = t;
                    break;
                }
            for (size_t t = 0, o = 0; t < best.size() && o < (size_t)k; t++)
                if (t != self) nbr[(size_t)i * k + o++] = best[t].second;
        }
    });
}

int g_knn_mode = 0;  // 0: knn_all (kd-tree order), 1: knn_all_cells (ties by index)

// ---------------------------------------------------------------- tufted cover + intrinsic Delaunay flips
struct CoverStats {
    int64_t flips = 0;
    double min_cotan = 0;    // smallest edge cotan weight after the flips (>= -1e-6 when intrinsically Delaunay)
    double area_before = 0;  // total cover area before the flips (the flips must preserve it)
};

inline double tri_area(double a, double b, double c) {  // utilities/elementary_geometry.ipp:7-12
    const double s = (a + b + c) / 2.0;
    return std::sqrt(std::max(0., s * (s - a) * (s - b) * (s - c)));
}
// third vertex of a triangle laid out in the plane (elementary_geometry.ipp:18-33)
inline V2 layout_vertex(const V2& pA, const V2& pB, double lBC, double lCA) {
    const double lAB = std::sqrt((pB.x - pA.x) * (pB.x - pA.x) + (pB.y - pA.y) * (pB.y - pA.y));
    const double h = 2. * tri_area(lAB, lBC, lCA) / lAB;
    const double w = (lAB * lAB - lBC * lBC + lCA * lCA) / (2. * lAB);
    const double rAB = 1. / lAB;  // Vector2::operator/(double) multiplies by the reciprocal (vector2.ipp:17-20)
    const V2 n{(pB.x - pA.x) * rAB, (pB.y - pA.y) * rAB};
    return V2{pA.x + w * n.x - h * n.y, pA.y + w * n.y + h * n.x};
}

void tufted_cover_weights(const double* P, int64_t nP, const std::vector<int64_t>& tris, double* areas_out, double* h_out,
                          CoverStats& st) {
    const int64_t T = (int64_t)tris.size() / 3;
    // ---- soup mesh: unique edges in creation order, their front halfedges in creation order (surface_mesh.cpp:145-185)
    struct SoupEdge {
        int64_t v0;                 // tail of the first halfedge created on the edge (defines orientation = true)
        double len;
        size_t begin, count;        // its front halfedges 3f + s in creation order h_1 .. h_n: sorted[begin .. begin + count)
    };
    std::vector<SoupEdge> edges;
    std::vector<int64_t> he_edge((size_t)3 * T);
    // halfedges ordered by (min vertex, max vertex, halfedge index): counting sort on the min vertex (ascending halfedge
    // index inside a bucket by construction), then each small bucket by (max vertex, index)
    std::vector<std::pair<std::pair<int64_t, int64_t>, int64_t>> sorted((size_t)3 * T);
    {
        std::vector<int64_t> start((size_t)nP + 1, 0);
        for (int64_t h = 0; h < 3 * T; h++) {
            const int64_t f = h / 3, sl = h % 3, a = tris[3 * f + sl], b = tris[3 * f + (sl + 1) % 3];
            start[(size_t)std::min(a, b) + 1]++;
        }
        for (int64_t v = 0; v < nP; v++) start[(size_t)v + 1] += start[(size_t)v];
        {
            std::vector<int64_t> fill(start.begin(), start.end() - 1);
            for (int64_t h = 0; h < 3 * T; h++) {
                const int64_t f = h / 3, sl = h % 3, a = tris[3 * f + sl], b = tris[3 * f + (sl + 1) % 3];
                sorted[(size_t)fill[(size_t)std::min(a, b)]++] = {{std::min(a, b), std::max(a, b)}, h};
            }
        }
        parallel_chunks(nP, [&](int64_t v_begin, int64_t v_end, int) {
            for (int64_t v = v_begin; v < v_end; v++)
                std::sort(sorted.begin() + start[(size_t)v], sorted.begin() + start[(size_t)v + 1]);
        });
        // creation order of edges = order of their first (lowest) halfedge: mark it, then walk the halfedges in order
        std::vector<int64_t> group_of_first((size_t)3 * T, -1);
        for (size_t i = 0; i < sorted.size();) {
            size_t j = i;
            while (j < sorted.size() && sorted[j].first == sorted[i].first) j++;
            group_of_first[(size_t)sorted[i].second] = (int64_t)i;
            i = j;
        }
        std::vector<int64_t> first_of_group;
        first_of_group.reserve((size_t)3 * T / 2 + 1);
        for (int64_t h = 0; h < 3 * T; h++)
            if (group_of_first[(size_t)h] >= 0) first_of_group.push_back(group_of_first[(size_t)h]);
        edges.reserve(first_of_group.size());
        for (int64_t gi : first_of_group) {
            SoupEdge e;
            const int64_t h1 = sorted[gi].second;
            e.v0 = tris[3 * (h1 / 3) + h1 % 3];
            const double* a = P + 3 * sorted[gi].first.first;
            const double* b = P + 3 * sorted[gi].first.second;
            e.len = std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
            e.begin = (size_t)gi;
            e.count = 0;
            for (size_t i = (size_t)gi; i < sorted.size() && sorted[i].first == sorted[gi].first; i++) {
                e.count++;
                he_edge[sorted[i].second] = (int64_t)edges.size();
            }
            edges.push_back(e);
        }
    }
    // ---- intrinsic mollification (intrinsic_mollification.cpp:7-38)
    {
        double sum = 0;
        for (const SoupEdge& e : edges) sum += e.len;
        const double delta = (sum / (double)edges.size()) * 1e-5;
        double eps = 0;
        for (int64_t f = 0; f < T; f++)
            for (int sl = 0; sl < 3; sl++) {
                const double lA = edges[he_edge[3 * f + sl]].len, lB = edges[he_edge[3 * f + (sl + 1) % 3]].len,
                             lC = edges[he_edge[3 * f + (sl + 2) % 3]].len;
                eps = std::fmax(eps, lC - lA - lB + delta);
            }
        for (SoupEdge& e : edges) e.len += eps;
    }
    // ---- the cover: front face f -> halfedges 6f + s (v_s -> v_{s+1}), back face (orientation inverted) -> halfedges
    // 6f + 3 + s running v_{s+1} -> v_s; otherSheet pairs 6f + s with 6f + 3 + s
    const int64_t H = 6 * T;
    std::vector<int64_t> next(H), twin(H, -1), vert(H), hedge(H, -1);
    std::vector<double> elen;  // per cover edge
    for (int64_t f = 0; f < T; f++)
        for (int sl = 0; sl < 3; sl++) {
            next[6 * f + sl] = 6 * f + (sl + 1) % 3;
            vert[6 * f + sl] = tris[3 * f + sl];
            next[6 * f + 3 + sl] = 6 * f + 3 + (sl + 2) % 3;
            vert[6 * f + 3 + sl] = tris[3 * f + (sl + 1) % 3];
        }
    auto front_of = [&](int64_t soup_he) { return 6 * (soup_he / 3) + soup_he % 3; };
    auto other = [&](int64_t h) { return (h % 6) < 3 ? h + 3 : h - 3; };
    std::vector<int64_t> F;
    const int64_t n_soup_edges = (int64_t)edges.size();
    {
        size_t total = 0;
        for (const SoupEdge& e : edges) total += e.count;
        elen.assign(total, 0.);
    }
    std::vector<int64_t> edge_he(elen.size(), -1);
    int64_t n_new = 0, soup_index = 0;
    for (const SoupEdge& e : edges) {
        // e.adjacentHalfedges(): h_1, then the sibling chain h_n, h_{n-1}, ..., h_2 (surface_mesh.cpp:187-203)
        const size_t n = e.count;
        F.resize(n);
        F[0] = front_of(sorted[e.begin].second);
        for (size_t i = 1; i < n; i++) F[i] = front_of(sorted[e.begin + n - i].second);
        // orientation(): front halfedge = (tail == tail of h_1); the inverted back copy has the opposite flag
        auto orient = [&](int64_t h) { return vert[h] == e.v0; };
        // tufted_laplacian.cpp:103-113
        int64_t curr = F[0];
        if (orient(curr)) curr = other(curr);
        for (size_t i = 0; i < n; i++) {
            int64_t nxt = F[(i + 1) % n];
            if (orient(curr) == orient(nxt)) nxt = other(nxt);
            twin[curr] = nxt;
            twin[nxt] = curr;
            // SurfaceMesh::separateToNewEdge (surface_mesh.cpp:930-964): every pair but the last one moves to a NEW edge
            // (appended behind all existing ones, with the first halfedge of the pair as its halfedge); the last pair
            // keeps the original edge.  Edge order = order of the flip queue and of the final sums.
            const int64_t id = (i + 1 < n) ? n_soup_edges + n_new++ : soup_index;
            hedge[curr] = hedge[nxt] = id;
            elen[(size_t)id] = e.len;
            edge_he[(size_t)id] = curr;
            curr = other(nxt);
        }
        soup_index++;
    }
    const int64_t E = (int64_t)elen.size();
    auto face_area = [&](int64_t h) { return tri_area(elen[hedge[h]], elen[hedge[next[h]]], elen[hedge[next[next[h]]]]); };
    for (int64_t f = 0; f < 2 * T; f++) st.area_before += face_area(3 * f);
    // ---- intrinsic flips to Delaunay (simple_idt.cpp:11-188, Euclidean, eps 1e-6)
    auto he_cotan = [&](int64_t h) {
        const double lij = elen[hedge[h]], ljk = elen[hedge[next[h]]], lki = elen[hedge[next[next[h]]]];
        return (-lij * lij + ljk * ljk + lki * lki) / (4. * tri_area(lij, ljk, lki)) / 2;
    };
    auto edge_cotan = [&](int64_t e) { return he_cotan(edge_he[e]) + he_cotan(twin[edge_he[e]]); };
    std::vector<int64_t> queue(E);
    std::iota(queue.begin(), queue.end(), 0);
    std::vector<char> in_queue(E, 1);
    size_t head = 0;
    while (head < queue.size()) {
        const int64_t e = queue[head++];
        in_queue[e] = 0;
        if (!(edge_cotan(e) < -1e-6)) continue;
        const int64_t ha1 = edge_he[e], ha2 = next[ha1], ha3 = next[ha2];
        const int64_t hb1 = twin[ha1], hb2 = next[hb1], hb3 = next[hb2];
        if (hb1 < 0 || ha2 == hb1 || hb2 == ha1) continue;  // SurfaceMesh::flip: incident on a degree-1 vertex
        // new length by laying out the two triangles (simple_idt.cpp:15-52)
        const double l01 = elen[hedge[ha2]], l12 = elen[hedge[ha3]], l23 = elen[hedge[hb2]], l30 = elen[hedge[hb3]],
                     l02 = elen[e];
        const V2 p3{0., 0.}, p0{l30, 0.};
        const V2 p2 = layout_vertex(p3, p0, l02, l23);
        const V2 p1 = layout_vertex(p2, p0, l01, l12);
        const double nl = std::sqrt((p1.x - p3.x) * (p1.x - p3.x) + (p1.y - p3.y) * (p1.y - p3.y));
        if (!std::isfinite(nl)) continue;
        // combinatorial flip (surface_mesh.cpp:895-919)
        const int64_t vc = vert[ha3], vd = vert[hb3];
        next[ha1] = hb3; next[hb3] = ha2; next[ha2] = ha1;
        next[hb1] = ha3; next[ha3] = hb2; next[hb2] = hb1;
        vert[ha1] = vc;
        vert[hb1] = vd;
        elen[e] = nl;
        st.flips++;
        const int64_t nb[4] = {hedge[next[ha1]], hedge[next[next[ha1]]], hedge[next[hb1]], hedge[next[next[hb1]]]};
        for (int64_t ne : nb)
            if (!in_queue[ne]) {
                queue.push_back(ne);
                in_queue[ne] = 1;
            }
    }
    // ---- vertex dual areas (intrinsic_geometry_interface.cpp:90-101) and mean edge length (src/signed_heat_3d.cpp:51-60)
    for (int64_t i = 0; i < nP; i++) areas_out[i] = 0;
    std::vector<char> seen(H, 0);
    st.min_cotan = 1e300;
    for (int64_t h = 0; h < H; h++) {
        if (seen[h]) continue;
        const int64_t h1 = next[h], h2 = next[h1];
        seen[h] = seen[h1] = seen[h2] = 1;
        const double A = face_area(h);
        areas_out[vert[h]] += A / 3.0;
        areas_out[vert[h1]] += A / 3.0;
        areas_out[vert[h2]] += A / 3.0;
    }
    double hs = 0;
    for (int64_t e = 0; e < E; e++) {
        hs += elen[e];
        st.min_cotan = std::min(st.min_cotan, edge_cotan(e));
    }
    *h_out = hs / (double)E;
}

// soup triangles (p, a, b) from every point's local triangulation, ordered by centre point like the reference's
// (point_position_geometry.cpp:113-134 for the tangent coordinates, local_triangulation.cpp for the rings)
void build_soup(const double* P, const double* N, int64_t nP, int k, const std::vector<int64_t>& nbr, std::vector<int64_t>& tris_out) {
    std::vector<std::vector<int64_t>> tris_of_chunk((size_t)chunk_count(nP));
    parallel_chunks(nP, [&](int64_t p_begin, int64_t p_end, int chunk) {
    std::vector<int64_t>& tris = tris_of_chunk[(size_t)chunk];
    std::vector<V2> pts((size_t)k);
    std::vector<size_t> ring;
    std::vector<char> tri_after;
    for (int64_t p = p_begin; p < p_end; p++) {
        const double* c = P + 3 * p;
        // geometry-central normalises by multiplying with the reciprocal length (vector3.ipp:132-135); the last bit matters
        // for the degenerate decisions below, so the same is done here
        const double rn = 1. / std::sqrt(N[3 * p] * N[3 * p] + N[3 * p + 1] * N[3 * p + 1] + N[3 * p + 2] * N[3 * p + 2]);
        const double nrm[3] = {N[3 * p], N[3 * p + 1], N[3 * p + 2]};
        const double u[3] = {nrm[0] * rn, nrm[1] * rn, nrm[2] * rn};
        // Vector3::buildTangentBasis (utilities/vector3.ipp:148-159)
        double t[3] = {1., 0., 0.};
        if (std::fabs(u[0]) > 0.9) { t[0] = 0.; t[1] = 1.; }
        double bx[3] = {t[1] * u[2] - t[2] * u[1], t[2] * u[0] - t[0] * u[2], t[0] * u[1] - t[1] * u[0]};
        double l = 1. / std::sqrt(bx[0] * bx[0] + bx[1] * bx[1] + bx[2] * bx[2]);
        for (double& v : bx) v *= l;
        double by[3] = {u[1] * bx[2] - u[2] * bx[1], u[2] * bx[0] - u[0] * bx[2], u[0] * bx[1] - u[1] * bx[0]};
        l = 1. / std::sqrt(by[0] * by[0] + by[1] * by[1] + by[2] * by[2]);
        for (double& v : by) v *= l;
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
    });
    tris_out.clear();  // chunks in point order
    for (const std::vector<int64_t>& t : tris_of_chunk) tris_out.insert(tris_out.end(), t.begin(), t.end());
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

// probe for the tests: tufted cover + intrinsic Delaunay flips of a GIVEN triangle soup (tris[T][3], vertex indices)
int shm3d_debug_tufted_weights(const double* P, int64_t nP, const int64_t* tris, int64_t T, double* areas_out, double* h_out,
                               int64_t* n_flips_out, double* min_cotan_out, double* area_before_out) {
    if (!P || !tris || !areas_out || !h_out || nP <= 0 || T <= 0) return SHM3D_ERR_INVALID_ARG;
    for (int64_t i = 0; i < 3 * T; i++)
        if (tris[i] < 0 || tris[i] >= nP) return SHM3D_ERR_INVALID_ARG;
    std::vector<int64_t> t(tris, tris + 3 * T);
    CoverStats cs;
    tufted_cover_weights(P, nP, t, areas_out, h_out, cs);
    if (n_flips_out) *n_flips_out = cs.flips;
    if (min_cotan_out) *min_cotan_out = cs.min_cotan;
    if (area_before_out) *area_before_out = cs.area_before;
    return SHM3D_OK;
}

void shm3d_debug_knn_mode(int32_t mode) { g_knn_mode = mode; }

int shm3d_point_weights(const double* P, const double* N, int64_t nP, int32_t k_neighbors, double* areas_out,
                        double* h_out, int64_t* n_triangles_out, int64_t* n_flips_out, double* min_cotan_out,
                        double* area_before_out) {
    if (!P || !N || !areas_out || !h_out || nP <= 0) return SHM3D_ERR_INVALID_ARG;
    const int k = k_neighbors > 0 ? k_neighbors : 30;  // PointPositionGeometry::kNeighborSize
    if ((int64_t)k + 1 > nP) return SHM3D_ERR_INVALID_ARG;  // "k+1 is greater than number of points" (knn.cpp:53)
    for (int64_t i = 0; i < 3 * nP; i++)
        if (!std::isfinite(P[i]) || !std::isfinite(N[i])) return SHM3D_ERR_NONFINITE;
    const bool dbg = std::getenv("SHM3D_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    std::vector<int64_t> nbr;
    if (g_knn_mode == 1) knn_all_cells(P, nP, k, nbr);
    else knn_all(P, nP, k, nbr);
    if (dbg) std::fprintf(stderr, "[shm3d] point_weights: kNN %.3fs", now() - t0), t0 = now();

    std::vector<int64_t> tris;
    build_soup(P, N, nP, k, nbr, tris);
    const int64_t T = (int64_t)tris.size() / 3;
    if (dbg) std::fprintf(stderr, "  local triangulations %.3fs", now() - t0), t0 = now();
    if (n_triangles_out) *n_triangles_out = T;
    for (int64_t i = 0; i < nP; i++) areas_out[i] = 0;
    if (T == 0) {
        *h_out = 0;
        return SHM3D_ERR_INVALID_ARG;
    }
    CoverStats cs;
    tufted_cover_weights(P, nP, tris, areas_out, h_out, cs);
    if (dbg) std::fprintf(stderr, "  cover + flips %.3fs (%lld flips)\n", now() - t0, (long long)cs.flips);
    if (n_flips_out) *n_flips_out = cs.flips;
    if (min_cotan_out) *min_cotan_out = cs.min_cotan;
    if (area_before_out) *area_before_out = cs.area_before;
    return SHM3D_OK;
}

}  // extern "C"
