// shm_ref_shim.h -- TEST INFRASTRUCTURE.  A minimal stand-in for the parts of geometry-central, Eigen and polyscope that
// the reference's grid-solver translation units (src/signed_heat_grid_solver.cpp, src/signed_heat_3d.cpp) touch, so that
// those two files can be compiled UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/libshm_ref.so
// (recipe: oracle/Makefile).  The real libraries cannot be used here: Eigen is not vendored and there is no network
// (SURVEY.md section 0 D7).
//
// What this buys: every line of first-party reference code on the hot path -- the Step 1-2 loops, laplacian(),
// gradient(), the constraint selection, trilinearCoefficients, evaluateFunction, the shift, integrateGreedily -- runs
// as written, and the oracle (oracle/shm_oracle.py) is checked against it (tests/test_reference_build.py).
// What it does not cover: the sparse LU itself (Eigen::SparseLU behind geometry-central's solveSquare) -- the shim's
// solveSquare hands the assembled KKT matrix to a callback (scipy SuperLU in the tests) -- and geometry-central's
// own mesh / point-cloud machinery, whose observable behaviour for this path is restated below with citations.
//
// Container semantics restated (geometry-central @ the reference's submodule):
//   faces iterate in input order, a face's vertices / halfedges in input order (surface_mesh.cpp:95-96);
//   he.vertex() is the halfedge's tail, he.next() the next halfedge of the same face;
//   edges are the unique unordered vertex pairs (surface_mesh.cpp:145,162-166), edge length = Euclidean distance;
//   horizontalStack / verticalStack concatenate blocks (numerical/linear_algebra_utilities.ipp:22-88).
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <queue>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

// ------------------------------------------------------------------------------------------------ Eigen (tiny subset)
namespace Eigen {
typedef std::ptrdiff_t Index;
const int Dynamic = -1;

template <typename T>
class VecX;

template <typename T>
struct HeadProxy {  // v.head(n): assignable view of the first n entries
    VecX<T>& v;
    Index n;
    HeadProxy& operator=(const VecX<T>& o);
    operator VecX<T>() const;
    VecX<T> operator-() const;
};

template <typename T>
class VecX {
  public:
    VecX() {}
    explicit VecX(Index n) : d_(n) {}
    static VecX Zero(Index n) {
        VecX r(n);
        for (Cell& c : r.d_) c.v = T(0);
        return r;
    }
    static VecX Ones(Index n) {
        VecX r(n);
        for (Cell& c : r.d_) c.v = T(1);
        return r;
    }
    Index size() const { return (Index)d_.size(); }
    T& operator()(Index i) { return d_[i].v; }
    const T& operator()(Index i) const { return d_[i].v; }
    T& operator[](Index i) { return d_[i].v; }
    const T& operator[](Index i) const { return d_[i].v; }
    HeadProxy<T> head(Index n) { return HeadProxy<T>{*this, n}; }
    VecX head(Index n) const {
        VecX r(n);
        for (Index i = 0; i < n; i++) r[i] = (*this)[i];
        return r;
    }
    VecX operator-() const {
        VecX r(size());
        for (Index i = 0; i < size(); i++) r[i] = -(*this)[i];
        return r;
    }
    VecX& operator-=(const VecX& o) {
        for (Index i = 0; i < size(); i++) (*this)[i] -= o[i];
        return *this;
    }
    VecX& operator+=(const VecX& o) {
        for (Index i = 0; i < size(); i++) (*this)[i] += o[i];
        return *this;
    }
    const T* data() const { return &d_[0].v; }
    T* data() { return &d_[0].v; }

  private:
    struct Cell { T v; };  // (a plain struct so that VecX<bool> does not become the bit-packed std::vector<bool>)
    std::vector<Cell> d_;
};
template <typename T>
HeadProxy<T>& HeadProxy<T>::operator=(const VecX<T>& o) {
    for (Index i = 0; i < n; i++) v[i] = o[i];
    return *this;
}
template <typename T>
HeadProxy<T>::operator VecX<T>() const {
    return static_cast<const VecX<T>&>(v).head(n);
}
template <typename T>
VecX<T> HeadProxy<T>::operator-() const {
    return -static_cast<const VecX<T>&>(v).head(n);
}
template <typename T>
VecX<T> operator*(T s, const VecX<T>& v) {
    VecX<T> r(v.size());
    for (Index i = 0; i < v.size(); i++) r[i] = s * v[i];
    return r;
}
typedef VecX<double> VectorXd;

struct Vector3d {
    double v[3];
    Vector3d() : v{0, 0, 0} {}
    Vector3d(double a, double b, double c) : v{a, b, c} {}
    double& operator[](int i) { return v[i]; }
    const double& operator[](int i) const { return v[i]; }
    Vector3d operator+(const Vector3d& o) const { return Vector3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vector3d& operator/=(double s) {
        v[0] /= s;
        v[1] /= s;
        v[2] /= s;
        return *this;
    }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};

template <typename T>
struct Triplet {
    Index r, c;
    T val;
    Triplet(Index r_, Index c_, T v_) : r(r_), c(c_), val(v_) {}
    Index row() const { return r; }
    Index col() const { return c; }
    T value() const { return val; }
};

// column-compressed like Eigen's default; duplicates summed by setFromTriplets
template <typename T>
class SparseMatrix {
  public:
    SparseMatrix() {}
    SparseMatrix(Index r, Index c) : rows_(r), cols_(c), outer_(c + 1, 0) {}
    void resize(Index r, Index c) {
        rows_ = r;
        cols_ = c;
        outer_.assign(c + 1, 0);
        inner_.clear();
        val_.clear();
    }
    Index rows() const { return rows_; }
    Index cols() const { return cols_; }
    Index nonZeros() const { return (Index)val_.size(); }
    template <typename It>
    void setFromTriplets(It b, It e) {
        std::vector<std::pair<std::pair<Index, Index>, T>> t;  // (col, row) -> value
        for (It it = b; it != e; ++it) t.push_back({{it->col(), it->row()}, it->value()});
        std::stable_sort(t.begin(), t.end(), [](const std::pair<std::pair<Index, Index>, T>& x,
                                                const std::pair<std::pair<Index, Index>, T>& y) { return x.first < y.first; });
        outer_.assign(cols_ + 1, 0);
        inner_.clear();
        val_.clear();
        for (size_t i = 0; i < t.size();) {
            size_t j = i;
            T s = 0;
            while (j < t.size() && t[j].first == t[i].first) s += t[j++].second;
            if (t[i].first.first < 0 || t[i].first.first >= cols_ || t[i].first.second < 0 || t[i].first.second >= rows_)
                throw std::out_of_range("shim SparseMatrix: triplet index out of range");
            inner_.push_back(t[i].first.second);
            val_.push_back(s);
            outer_[t[i].first.first + 1]++;
            i = j;
        }
        for (Index c = 0; c < cols_; c++) outer_[c + 1] += outer_[c];
    }
    SparseMatrix transpose() const {
        std::vector<Triplet<T>> t;
        for (Index c = 0; c < cols_; c++)
            for (Index p = outer_[c]; p < outer_[c + 1]; p++) t.emplace_back(c, inner_[p], val_[p]);
        SparseMatrix r(cols_, rows_);
        r.setFromTriplets(t.begin(), t.end());
        return r;
    }
    VecX<T> operator*(const VecX<T>& x) const {
        VecX<T> y = VecX<T>::Zero(rows_);
        for (Index c = 0; c < cols_; c++)
            for (Index p = outer_[c]; p < outer_[c + 1]; p++) y[inner_[p]] += val_[p] * x[c];
        return y;
    }
    SparseMatrix operator/(T s) const {
        SparseMatrix r = *this;
        for (T& v : r.val_) v /= s;
        return r;
    }
    const std::vector<Index>& outer() const { return outer_; }
    const std::vector<Index>& inner() const { return inner_; }
    const std::vector<T>& values() const { return val_; }

  private:
    Index rows_ = 0, cols_ = 0;
    std::vector<Index> outer_, inner_;
    std::vector<T> val_;
};
}  // namespace Eigen

// ------------------------------------------------------------------------------------------------ geometry-central
namespace geometrycentral {

template <typename T>
using Vector = Eigen::VecX<T>;
template <typename T>
using SparseMatrix = Eigen::SparseMatrix<T>;

struct Vector3 {
    double x, y, z;
    double& operator[](int i) { return (&x)[i]; }
    const double& operator[](int i) const { return (&x)[i]; }
    Vector3 operator+(const Vector3& o) const { return Vector3{x + o.x, y + o.y, z + o.z}; }
    Vector3 operator-(const Vector3& o) const { return Vector3{x - o.x, y - o.y, z - o.z}; }
    Vector3 operator*(double s) const { return Vector3{x * s, y * s, z * s}; }
    Vector3 operator/(double s) const { return Vector3{x / s, y / s, z / s}; }
    Vector3 operator-() const { return Vector3{-x, -y, -z}; }
    Vector3& operator+=(const Vector3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    Vector3& operator-=(const Vector3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    Vector3& operator*=(double s) { x *= s; y *= s; z *= s; return *this; }
    Vector3& operator/=(double s) { x /= s; y /= s; z /= s; return *this; }
    double norm() const { return std::sqrt(x * x + y * y + z * z); }
    double norm2() const { return x * x + y * y + z * z; }
};
inline Vector3 operator*(double s, const Vector3& v) { return v * s; }
inline double dot(const Vector3& a, const Vector3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vector3 cross(const Vector3& a, const Vector3& b) {
    return Vector3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// numerical/linear_algebra_utilities.ipp:22-88
template <typename T>
SparseMatrix<T> horizontalStack(const std::vector<SparseMatrix<T>>& mats) {
    Eigen::Index rows = mats.empty() ? 0 : mats[0].rows(), cols = 0;
    std::vector<Eigen::Triplet<T>> t;
    for (const SparseMatrix<T>& M : mats) {
        if (M.rows() != rows) throw std::logic_error("horizontalStack: row mismatch");
        for (Eigen::Index c = 0; c < M.cols(); c++)
            for (Eigen::Index p = M.outer()[c]; p < M.outer()[c + 1]; p++) t.emplace_back(M.inner()[p], cols + c, M.values()[p]);
        cols += M.cols();
    }
    SparseMatrix<T> R(rows, cols);
    R.setFromTriplets(t.begin(), t.end());
    return R;
}
template <typename T>
SparseMatrix<T> verticalStack(const std::vector<SparseMatrix<T>>& mats) {
    Eigen::Index cols = mats.empty() ? 0 : mats[0].cols(), rows = 0;
    std::vector<Eigen::Triplet<T>> t;
    for (const SparseMatrix<T>& M : mats) {
        if (M.cols() != cols) throw std::logic_error("verticalStack: column mismatch");
        for (Eigen::Index c = 0; c < M.cols(); c++)
            for (Eigen::Index p = M.outer()[c]; p < M.outer()[c + 1]; p++) t.emplace_back(rows + M.inner()[p], c, M.values()[p]);
        rows += M.rows();
    }
    SparseMatrix<T> R(rows, cols);
    R.setFromTriplets(t.begin(), t.end());
    return R;
}

// solveSquare (src/numerical/square_solvers.cpp:190-193): "Matrix must be square" / non-finite checks, then the LU --
// delegated to the harness's callback (column-compressed arrays).
typedef void (*shim_solve_fn)(int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowidx, const double* val,
                              const double* rhs, double* x);
shim_solve_fn& shim_solver();
inline Vector<double> solveSquare(SparseMatrix<double>& A, const Vector<double>& rhs) {
    if (A.rows() != A.cols()) throw std::logic_error("Matrix must be square");
    for (double v : A.values())
        if (!std::isfinite(v)) throw std::logic_error("Matrix has non-finite entries");
    for (Eigen::Index i = 0; i < rhs.size(); i++)
        if (!std::isfinite(rhs[i])) throw std::logic_error("right-hand side has non-finite entries");  // checkFinite (:164-169)
    if (!shim_solver()) throw std::runtime_error("shim: no linear solver callback installed");
    std::vector<int64_t> cp(A.outer().begin(), A.outer().end()), ri(A.inner().begin(), A.inner().end());
    Vector<double> x = Vector<double>::Zero(rhs.size());
    shim_solver()((int64_t)A.rows(), (int64_t)A.nonZeros(), cp.data(), ri.data(), A.values().data(), rhs.data(), x.data());
    return x;
}
// The reference factorises the Laplacian with this and never solves with it (src/signed_heat_grid_solver.cpp:30).
template <typename T>
class PositiveDefiniteSolver {
  public:
    explicit PositiveDefiniteSolver(SparseMatrix<T>&) {}
};

namespace surface {
enum class LevelSetConstraint { None = 0, ZeroSet, Multiple };

class SurfaceMesh;
struct Element {
    const SurfaceMesh* mesh = nullptr;
    size_t ind = 0;
    size_t getIndex() const { return ind; }
};
struct Vertex : Element {};
struct Edge : Element {};
struct Face;
struct Halfedge : Element {  // ind = position in the flattened face-vertex list
    Vertex vertex() const;
    Halfedge next() const;
};
template <typename E>
struct Range {
    std::vector<E> items;
    typename std::vector<E>::const_iterator begin() const { return items.begin(); }
    typename std::vector<E>::const_iterator end() const { return items.end(); }
};
// allocation-free ranges over the corners of one face (the reference calls these once per (node, face) pair)
template <typename E, bool kDeref>
struct CornerIter {
    const SurfaceMesh* mesh;
    size_t pos;
    bool operator!=(const CornerIter& o) const { return pos != o.pos; }
    void operator++() { ++pos; }
    E operator*() const;
};
template <typename E, bool kDeref>
struct CornerRange {
    const SurfaceMesh* mesh;
    size_t b, e;
    CornerIter<E, kDeref> begin() const { return CornerIter<E, kDeref>{mesh, b}; }
    CornerIter<E, kDeref> end() const { return CornerIter<E, kDeref>{mesh, e}; }
};
struct Face : Element {
    size_t degree() const;
    CornerRange<Vertex, true> adjacentVertices() const;
    CornerRange<Halfedge, false> adjacentHalfedges() const;
};

class SurfaceMesh {
  public:
    SurfaceMesh(size_t nV, const std::vector<size_t>& faceVertices, const std::vector<size_t>& faceOffsets)
        : nV_(nV), fv_(faceVertices), fo_(faceOffsets) {
        std::map<std::pair<size_t, size_t>, size_t> seen;
        for (size_t f = 0; f + 1 < fo_.size(); f++) {
            const size_t d = fo_[f + 1] - fo_[f];
            for (size_t t = 0; t < d; t++) {
                size_t a = fv_[fo_[f] + t], b = fv_[fo_[f] + (t + 1) % d];
                std::pair<size_t, size_t> key(std::min(a, b), std::max(a, b));
                if (seen.emplace(key, edges_.size()).second) edges_.push_back(key);
            }
        }
    }
    size_t nVertices() const { return nV_; }
    size_t nFaces() const { return fo_.size() - 1; }
    size_t nEdges() const { return edges_.size(); }
    bool isTriangular() const {
        for (size_t f = 0; f + 1 < fo_.size(); f++)
            if (fo_[f + 1] - fo_[f] != 3) return false;
        return true;
    }
    const Range<Vertex>& vertices() const {
        if (vr_.items.size() != nV_) vr_ = make<Vertex>(nV_);
        return vr_;
    }
    const Range<Face>& faces() const {
        if (fr_.items.size() != nFaces()) fr_ = make<Face>(nFaces());
        return fr_;
    }
    const Range<Edge>& edges() const {
        if (er_.items.size() != nEdges()) er_ = make<Edge>(nEdges());
        return er_;
    }
    const std::vector<size_t>& faceVertices() const { return fv_; }
    const std::vector<size_t>& faceOffsets() const { return fo_; }
    const std::vector<std::pair<size_t, size_t>>& edgeList() const { return edges_; }

  private:
    template <typename E>
    Range<E> make(size_t n) const {
        Range<E> r;
        r.items.resize(n);
        for (size_t i = 0; i < n; i++) {
            r.items[i].mesh = this;
            r.items[i].ind = i;
        }
        return r;
    }
    size_t nV_;
    std::vector<size_t> fv_, fo_;
    std::vector<std::pair<size_t, size_t>> edges_;
    mutable Range<Vertex> vr_;
    mutable Range<Face> fr_;
    mutable Range<Edge> er_;
};
inline size_t face_of_halfedge(const SurfaceMesh& m, size_t he) {
    const std::vector<size_t>& fo = m.faceOffsets();
    return (size_t)(std::upper_bound(fo.begin(), fo.end(), he) - fo.begin()) - 1;
}
inline Vertex Halfedge::vertex() const {
    Vertex v;
    v.mesh = mesh;
    v.ind = mesh->faceVertices()[ind];
    return v;
}
inline Halfedge Halfedge::next() const {
    const size_t f = face_of_halfedge(*mesh, ind), b = mesh->faceOffsets()[f], d = mesh->faceOffsets()[f + 1] - b;
    Halfedge h;
    h.mesh = mesh;
    h.ind = b + (ind - b + 1) % d;
    return h;
}
inline size_t Face::degree() const { return mesh->faceOffsets()[ind + 1] - mesh->faceOffsets()[ind]; }
template <>
inline Vertex CornerIter<Vertex, true>::operator*() const {
    Vertex v;
    v.mesh = mesh;
    v.ind = mesh->faceVertices()[pos];
    return v;
}
template <>
inline Halfedge CornerIter<Halfedge, false>::operator*() const {
    Halfedge h;
    h.mesh = mesh;
    h.ind = pos;
    return h;
}
inline CornerRange<Vertex, true> Face::adjacentVertices() const {
    return CornerRange<Vertex, true>{mesh, mesh->faceOffsets()[ind], mesh->faceOffsets()[ind + 1]};
}
inline CornerRange<Halfedge, false> Face::adjacentHalfedges() const {
    return CornerRange<Halfedge, false>{mesh, mesh->faceOffsets()[ind], mesh->faceOffsets()[ind + 1]};
}

template <typename E, typename T>
class MeshData {
  public:
    MeshData() {}
    MeshData(const SurfaceMesh&, size_t n) : d_(n) {}
    T& operator[](const E& e) { return d_[e.ind]; }
    const T& operator[](const E& e) const { return d_[e.ind]; }
    T& operator[](size_t i) { return d_[i]; }
    const T& operator[](size_t i) const { return d_[i]; }
    size_t size() const { return d_.size(); }

  protected:
    std::vector<T> d_;
};
template <typename T>
struct VertexData : MeshData<Vertex, T> {
    VertexData() {}
    explicit VertexData(const SurfaceMesh& m) : MeshData<Vertex, T>(m, m.nVertices()) {}
};
template <typename T>
struct FaceData : MeshData<Face, T> {
    FaceData() {}
    explicit FaceData(const SurfaceMesh& m) : MeshData<Face, T>(m, m.nFaces()) {}
};
template <typename T>
struct EdgeData : MeshData<Edge, T> {
    EdgeData() {}
    explicit EdgeData(const SurfaceMesh& m) : MeshData<Edge, T>(m, m.nEdges()) {}
};

class IntrinsicGeometryInterface {
  public:
    explicit IntrinsicGeometryInterface(SurfaceMesh& m) : mesh(m) {}
    virtual ~IntrinsicGeometryInterface() {}
    SurfaceMesh& mesh;
    EdgeData<double> edgeLengths;
    VertexData<double> vertexDualAreas;
    virtual void requireEdgeLengths() {}
    void unrequireEdgeLengths() {}
    void requireVertexDualAreas() {}
    void unrequireVertexDualAreas() {}
};
// intrinsic geometry given by edge lengths (the tufted triangulation of a point cloud); the harness fills it
class EdgeLengthGeometry : public IntrinsicGeometryInterface {
  public:
    explicit EdgeLengthGeometry(SurfaceMesh& m) : IntrinsicGeometryInterface(m) {}
};

class VertexPositionGeometry : public IntrinsicGeometryInterface {
  public:
    VertexPositionGeometry(SurfaceMesh& m, const std::vector<Vector3>& pos) : IntrinsicGeometryInterface(m), vertexPositions(m) {
        for (size_t i = 0; i < pos.size(); i++) vertexPositions[i] = pos[i];
    }
    VertexData<Vector3> vertexPositions;
    FaceData<double> faceAreas;
    FaceData<Vector3> faceNormals;
    void requireEdgeLengths() override {
        edgeLengths = EdgeData<double>(mesh);
        for (size_t e = 0; e < mesh.nEdges(); e++)
            edgeLengths[e] = (vertexPositions[mesh.edgeList()[e].first] - vertexPositions[mesh.edgeList()[e].second]).norm();
    }
    // triangle areas / normals (only reached through setFaceVectorAreas' triangular branch, whose result the reference
    // then overwrites with the shoelace values -- src/signed_heat_3d.cpp:65-88)
    void requireFaceAreas() { computeFaces(); }
    void requireFaceNormals() { computeFaces(); }
    void unrequireFaceAreas() {}
    void unrequireFaceNormals() {}

  private:
    void computeFaces() {
        faceAreas = FaceData<double>(mesh);
        faceNormals = FaceData<Vector3>(mesh);
        for (size_t f = 0; f < mesh.nFaces(); f++) {
            const size_t b = mesh.faceOffsets()[f];
            const Vector3 p0 = vertexPositions[mesh.faceVertices()[b]], p1 = vertexPositions[mesh.faceVertices()[b + 1]],
                          p2 = vertexPositions[mesh.faceVertices()[b + 2]];
            const Vector3 n = cross(p1 - p0, p2 - p0);
            faceAreas[f] = 0.5 * n.norm();
            faceNormals[f] = n / n.norm();
        }
    }
};
}  // namespace surface

namespace pointcloud {
class PointCloud {
  public:
    explicit PointCloud(size_t n) : n_(n) {}
    size_t nPoints() const { return n_; }

  private:
    size_t n_;
};
template <typename T>
class PointData {
  public:
    PointData() {}
    explicit PointData(size_t n) : d_(n) {}
    T& operator[](size_t i) { return d_[i]; }
    const T& operator[](size_t i) const { return d_[i]; }

  private:
    std::vector<T> d_;
};
class PointPositionGeometry {
  public:
    explicit PointPositionGeometry(PointCloud& c) : cloud(c), positions(c.nPoints()) {}
    virtual ~PointPositionGeometry() {}
    PointCloud& cloud;
    PointData<Vector3> positions;
    std::unique_ptr<surface::EdgeLengthGeometry> tuftedGeom;  // supplied by the harness (tufted cover = row N1)
    void requireTuftedTriangulation() {}
    void unrequireTuftedTriangulation() {}
};
class PointPositionNormalGeometry : public PointPositionGeometry {
  public:
    explicit PointPositionNormalGeometry(PointCloud& c) : PointPositionGeometry(c), normals(c.nPoints()) {}
    PointData<Vector3> normals;
};
}  // namespace pointcloud
}  // namespace geometrycentral

// ------------------------------------------------------------------------------------------------ polyscope / glm
namespace glm {
struct vec3 {
    float v[3];
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};
struct uvec3 {
    size_t v[3];
    uvec3(size_t a, size_t b, size_t c) : v{a, b, c} {}
};
}  // namespace glm
namespace polyscope {
struct VolumeGrid {
    std::string name;
    size_t dim[3];
    float bmin[3], bmax[3];
};
// the side effect src/main.cpp:95 relies on: recorded so the harness can report it
VolumeGrid& shim_last_grid();
inline VolumeGrid* registerVolumeGrid(const std::string& name, glm::uvec3 dim, glm::vec3 bmin, glm::vec3 bmax) {
    VolumeGrid& g = shim_last_grid();
    g.name = name;
    for (int i = 0; i < 3; i++) {
        g.dim[i] = dim.v[i];
        g.bmin[i] = bmin[i];
        g.bmax[i] = bmax[i];
    }
    return &g;
}
}  // namespace polyscope
