// shm_eigen_stub.h -- TEST INFRASTRUCTURE.  A stand-in for the slice of Eigen that geometry-central's sources and the
// reference's grid solver mention, so that those sources compile from where they lie under /root/reference (the real
// Eigen 3.3.8 is fetched by geometry-central's configure step and is absent from this image).
//
// What works, because the code paths the oracle exercises need it:
//   * Matrix<T, ...> as a CONTAINER (size, resize, element access, fill, copy) -- geometry-central keeps all of its
//     per-element data (MeshData, utilities/mesh_data.h:195) in Eigen vectors;
//   * the vector arithmetic of src/signed_heat_grid_solver.cpp (Zero / Ones, head(n) as l- and r-value, unary minus,
//     scalar * vector, -=, 3-vector + and norm) and element-wise helpers;
//   * SparseMatrix as a column-compressed matrix: setFromTriplets (duplicates summed, like Eigen), InnerIterator, coeff,
//     transpose, matrix * vector, matrix / scalar -- what laplacian(), gradient(), the constraint assembly and
//     geometry-central's horizontalStack / verticalStack / checkFinite / checkHermitian use;
//   * the 4x4 determinant of geometry-central's in-circle test, written as Eigen 3.3's fixed-size kernel evaluates it;
//   * SparseLU::compute / solve: NOT an LU -- the assembled system is handed to a callback installed by the harness
//     (scipy's SuperLU in the tests; Eigen::SparseLU descends from the same SuperLU code);
//   * SimplicialLDLT::compute: a no-op reporting success -- the reference factorises its Laplacian with it and never
//     solves with the factor (src/signed_heat_grid_solver.cpp:30).
// Everything else is declared so that the sources compile and ABORTS if it is ever reached.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace Eigen {

[[noreturn]] inline void shm_stub_unreachable(const char* what) {
    std::fprintf(stderr, "Eigen stub: %s is not implemented (oracle/ref_shim/eigen_stub)\n", what);
    std::abort();
}

typedef std::ptrdiff_t Index;
const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { Unaligned = 0, Aligned8 = 8, Aligned16 = 16, Aligned32 = 32, Aligned64 = 64, AlignedMax = 64 };
enum { ComputeFullU = 4, ComputeThinU = 8, ComputeFullV = 16, ComputeThinV = 32 };
enum { Lower = 1, Upper = 2 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

template <typename T>
struct aligned_allocator : public std::allocator<T> {
    template <typename U>
    struct rebind {
        typedef aligned_allocator<U> other;
    };
};

template <typename T>
struct NumTraits {
    typedef T Real;
};
template <typename T>
struct NumTraits<std::complex<T>> {
    typedef T Real;
};

template <typename Scalar_, int Rows_ = Dynamic, int Cols_ = Dynamic, int Options_ = 0, int MaxRows_ = Rows_, int MaxCols_ = Cols_>
class Matrix;

// v.head(n): assignable view of the first n entries of a column vector
template <typename M>
struct HeadProxy {
    M& m;
    Index n;
    HeadProxy& operator=(const M& o) {
        for (Index i = 0; i < n; i++) m[i] = o[i];
        return *this;
    }
    operator M() const { return static_cast<const M&>(m).head(n); }
    M operator-() const { return -static_cast<const M&>(m).head(n); }
};

template <typename Scalar_, int Rows_, int Cols_, int Options_, int MaxRows_, int MaxCols_>
class Matrix {
  public:
    typedef Scalar_ Scalar;
    typedef typename NumTraits<Scalar_>::Real RealScalar;
    enum { RowsAtCompileTime = Rows_, ColsAtCompileTime = Cols_ };

    // ---- container
    Matrix() { init(Rows_ == Dynamic ? 0 : Rows_, Cols_ == Dynamic ? 0 : Cols_); }
    explicit Matrix(Index n) { Cols_ == 1 || Cols_ == Dynamic ? init(n, 1) : init(1, n); }
    Matrix(Index r, Index c) { init(r, c); }
    Matrix(const Scalar& a, const Scalar& b, const Scalar& c) {  // fixed 3-vectors: Eigen::Vector3d v = {x, y, z}
        if (Rows_ * Cols_ != 3) shm_stub_unreachable("Matrix(x, y, z) on a non-3-vector");
        init(Rows_, Cols_);
        at(0) = a;
        at(1) = b;
        at(2) = c;
    }
    Index rows() const { return rows_; }
    Index cols() const { return cols_; }
    Index size() const { return rows_ * cols_; }
    Scalar& operator()(Index i) { return at(i); }
    const Scalar& operator()(Index i) const { return at(i); }
    Scalar& operator()(Index i, Index j) { return at(i + j * rows_); }
    const Scalar& operator()(Index i, Index j) const { return at(i + j * rows_); }
    Scalar& operator[](Index i) { return at(i); }
    const Scalar& operator[](Index i) const { return at(i); }
    Scalar* data() { return store_.empty() ? nullptr : &at(0); }
    const Scalar* data() const { return store_.empty() ? nullptr : &at(0); }
    void resize(Index n) { Cols_ == 1 || Cols_ == Dynamic ? init(n, 1) : init(1, n); }
    void resize(Index r, Index c) { init(r, c); }
    void conservativeResize(Index n) {
        store_.resize((size_t)n);
        if (Cols_ == 1 || Cols_ == Dynamic) rows_ = n, cols_ = 1;
        else rows_ = 1, cols_ = n;
    }
    void setZero() { fill(Scalar()); }
    void setZero(Index n) { resize(n); }
    void setZero(Index r, Index c) { resize(r, c); }
    void setConstant(const Scalar& v) { fill(v); }
    void fill(const Scalar& v) {
        for (auto& b : store_) b.v = v;
    }
    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(Index n) { return Matrix(n); }
    static Matrix Zero(Index r, Index c) { return Matrix(r, c); }
    static Matrix Constant(Index n, const Scalar& v) {
        Matrix m(n);
        m.fill(v);
        return m;
    }
    static Matrix Constant(Index r, Index c, const Scalar& v) {
        Matrix m(r, c);
        m.fill(v);
        return m;
    }
    static Matrix Ones(Index n) { return Constant(n, Scalar(1)); }
    static Matrix Ones(Index r, Index c) { return Constant(r, c, Scalar(1)); }

    // ---- element-wise arithmetic (same shape)
    Matrix operator-() const {
        Matrix r(*this);
        for (Index i = 0; i < size(); i++) r.at(i) = -at(i);
        return r;
    }
    Matrix& operator+=(const Matrix& o) {
        same_shape(o);
        for (Index i = 0; i < size(); i++) at(i) += o.at(i);
        return *this;
    }
    Matrix& operator-=(const Matrix& o) {
        same_shape(o);
        for (Index i = 0; i < size(); i++) at(i) -= o.at(i);
        return *this;
    }
    Matrix& operator*=(const Scalar& s) {
        for (Index i = 0; i < size(); i++) at(i) *= s;
        return *this;
    }
    Matrix& operator/=(const Scalar& s) {
        for (Index i = 0; i < size(); i++) at(i) /= s;
        return *this;
    }
    Matrix operator+(const Matrix& o) const {
        Matrix r(*this);
        r += o;
        return r;
    }
    Matrix operator-(const Matrix& o) const {
        Matrix r(*this);
        r -= o;
        return r;
    }
    Matrix operator*(const Scalar& s) const {
        Matrix r(*this);
        r *= s;
        return r;
    }
    Matrix operator/(const Scalar& s) const {
        Matrix r(*this);
        r /= s;
        return r;
    }
    RealScalar squaredNorm() const {
        RealScalar s = RealScalar();
        for (Index i = 0; i < size(); i++) s += std::abs(at(i)) * std::abs(at(i));
        return s;
    }
    RealScalar norm() const {  // real scalars: sqrt of the running sum of squares, first to last
        RealScalar s = RealScalar();
        for (Index i = 0; i < size(); i++) s += std::abs(at(i)) * std::abs(at(i));
        return std::sqrt(s);
    }
    Scalar sum() const {
        Scalar s = Scalar();
        for (Index i = 0; i < size(); i++) s += at(i);
        return s;
    }
    bool allFinite() const {
        for (Index i = 0; i < size(); i++)
            if (!std::isfinite(std::abs(at(i)))) return false;
        return true;
    }
    template <typename NewScalar>
    Matrix<NewScalar, Rows_, Cols_> cast() const {
        Matrix<NewScalar, Rows_, Cols_> r(rows_, cols_);
        for (Index i = 0; i < size(); i++) r[i] = static_cast<NewScalar>(at(i));
        return r;
    }
    HeadProxy<Matrix> head(Index n) { return HeadProxy<Matrix>{*this, n}; }
    Matrix head(Index n) const {
        Matrix r(n);
        for (Index i = 0; i < n; i++) r.at(i) = at(i);
        return r;
    }

    // The one dense kernel on the exercised path: inCircleTest (src/utilities/elementary_geometry.cpp:8-19) takes the sign
    // of a 4x4 determinant.  Restated as Eigen 3.3's fixed-size 4x4 kernel evaluates it (Eigen/src/LU/Determinant.h,
    // bruteforce_det4_helper: products of 2x2 minors of columns 0-1 and 2-3), so that rounding -- which can only matter
    // for nearly cocircular points -- follows the same expression.
    Scalar determinant() const {
        if (rows_ != 4 || cols_ != 4) shm_stub_unreachable("Matrix::determinant (only 4x4)");
        const Matrix& m = *this;
        auto h = [&m](int j, int k, int a, int b) {
            return (m(j, 0) * m(k, 1) - m(k, 0) * m(j, 1)) * (m(a, 2) * m(b, 3) - m(b, 2) * m(a, 3));
        };
        return h(0, 1, 2, 3) - h(0, 2, 1, 3) + h(0, 3, 1, 2) + h(1, 2, 0, 3) - h(1, 3, 0, 2) + h(2, 3, 0, 1);
    }
    struct CommaInit {  // A << a, b, c, ...;  fills row by row, like Eigen's
        Matrix* m;
        Index k;
        CommaInit& operator,(const Scalar& v) {
            (*m)(k / m->cols(), k % m->cols()) = v;
            k++;
            return *this;
        }
    };
    CommaInit operator<<(const Scalar& v) {
        CommaInit c{this, 0};
        return (c, v);
    }

    // ---- declared only: abort when reached
    void setOnes() { fill(Scalar(1)); }
    static Matrix Identity() { shm_stub_unreachable("Matrix::Identity"); }
    static Matrix Identity(Index, Index) { shm_stub_unreachable("Matrix::Identity"); }
    static Matrix Random(Index) { shm_stub_unreachable("Matrix::Random"); }
    static Matrix Random(Index, Index) { shm_stub_unreachable("Matrix::Random"); }
    Matrix<Scalar, Dynamic, 1> col(Index) const { shm_stub_unreachable("Matrix::col"); }
    Matrix<Scalar, 1, Dynamic> row(Index) const { shm_stub_unreachable("Matrix::row"); }
    Matrix<Scalar, Dynamic, Dynamic> transpose() const { shm_stub_unreachable("Matrix::transpose"); }
    Matrix<Scalar, Dynamic, Dynamic> adjoint() const { shm_stub_unreachable("Matrix::adjoint"); }
    Matrix<Scalar, Dynamic, Dynamic> inverse() const { shm_stub_unreachable("Matrix::inverse"); }
    Matrix<Scalar, Dynamic, Dynamic> asDiagonal() const { shm_stub_unreachable("Matrix::asDiagonal"); }
    Matrix conjugate() const { shm_stub_unreachable("Matrix::conjugate"); }
    Matrix cwiseAbs() const { shm_stub_unreachable("Matrix::cwiseAbs"); }
    Matrix cwiseInverse() const { shm_stub_unreachable("Matrix::cwiseInverse"); }
    Matrix array() const { shm_stub_unreachable("Matrix::array"); }
    Matrix matrix() const { shm_stub_unreachable("Matrix::matrix"); }
    Matrix tail(Index) const { shm_stub_unreachable("Matrix::tail"); }
    Matrix segment(Index, Index) const { shm_stub_unreachable("Matrix::segment"); }
    Matrix block(Index, Index, Index, Index) const { shm_stub_unreachable("Matrix::block"); }
    Matrix normalized() const { shm_stub_unreachable("Matrix::normalized"); }
    Scalar mean() const { shm_stub_unreachable("Matrix::mean"); }
    Scalar maxCoeff() const { shm_stub_unreachable("Matrix::maxCoeff"); }
    Scalar minCoeff() const { shm_stub_unreachable("Matrix::minCoeff"); }
    template <typename O>
    Scalar dot(const O&) const { shm_stub_unreachable("Matrix::dot"); }
    bool hasNaN() const { shm_stub_unreachable("Matrix::hasNaN"); }
    Matrix<RealScalar, Rows_, Cols_> real() const { shm_stub_unreachable("Matrix::real"); }
    Matrix<RealScalar, Rows_, Cols_> imag() const { shm_stub_unreachable("Matrix::imag"); }
    struct SolverStub {
        template <typename B>
        Matrix<Scalar, Dynamic, Dynamic> solve(const B&) const { shm_stub_unreachable("dense solve"); }
    };
    SolverStub colPivHouseholderQr() const { shm_stub_unreachable("colPivHouseholderQr"); }
    SolverStub householderQr() const { shm_stub_unreachable("householderQr"); }
    SolverStub ldlt() const { shm_stub_unreachable("ldlt"); }
    SolverStub llt() const { shm_stub_unreachable("llt"); }
    // products and shape-changing conversions between different matrix types: expression results the exercised paths
    // never form
    template <typename S2, int R2, int C2, int O2, int MR2, int MC2>
    Matrix(const Matrix<S2, R2, C2, O2, MR2, MC2>&) { shm_stub_unreachable("conversion between matrix types"); }
    template <typename S2, int R2, int C2, int O2, int MR2, int MC2>
    Matrix<Scalar, Dynamic, Dynamic> operator*(const Matrix<S2, R2, C2, O2, MR2, MC2>&) const { shm_stub_unreachable("dense product"); }

  private:
    struct Box {  // std::vector<bool> has no bool&; one uniform representation for every Scalar
        Scalar v;
    };
    std::vector<Box> store_;
    Index rows_ = 0, cols_ = 0;
    void init(Index r, Index c) {
        rows_ = r;
        cols_ = c;
        store_.assign((size_t)(r * c), Box{Scalar()});
    }
    void same_shape(const Matrix& o) const {
        if (o.rows_ != rows_ || o.cols_ != cols_) shm_stub_unreachable("element-wise operation on different shapes");
    }
    Scalar& at(Index i) { return store_[(size_t)i].v; }
    const Scalar& at(Index i) const { return store_[(size_t)i].v; }
};

template <typename S, int R, int C, int O, int MR, int MC>
Matrix<S, R, C, O, MR, MC> operator*(const S& s, const Matrix<S, R, C, O, MR, MC>& m) {
    return m * s;
}

template <typename Derived>
class MatrixBase {  // appears in template signatures of geometry-central headers only
  public:
    typedef double Scalar;
    Index rows() const { shm_stub_unreachable("MatrixBase"); }
    Index cols() const { shm_stub_unreachable("MatrixBase"); }
    Index size() const { shm_stub_unreachable("MatrixBase"); }
    Scalar operator()(Index) const { shm_stub_unreachable("MatrixBase"); }
    Scalar operator()(Index, Index) const { shm_stub_unreachable("MatrixBase"); }
    template <typename NewScalar>
    Matrix<NewScalar, Dynamic, Dynamic> cast() const { shm_stub_unreachable("MatrixBase"); }
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
};
template <typename Derived>
class DenseBase : public MatrixBase<Derived> {};
template <typename Derived>
class EigenBase : public MatrixBase<Derived> {};
template <typename Derived>
class SparseMatrixBase : public MatrixBase<Derived> {};

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<int, Dynamic, Dynamic> MatrixXi;
typedef Matrix<std::complex<double>, Dynamic, Dynamic> MatrixXcd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<std::complex<double>, Dynamic, 1> VectorXcd;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 3, 3> Matrix3f;

template <typename PlainObjectType, int MapOptions = Unaligned, typename StrideType = void>
class Map : public PlainObjectType {
  public:
    typedef typename PlainObjectType::Scalar Scalar;
    Map(const Scalar*) { shm_stub_unreachable("Map"); }
    Map(const Scalar*, Index) { shm_stub_unreachable("Map"); }
    Map(const Scalar*, Index, Index) { shm_stub_unreachable("Map"); }
    template <typename O>
    Map& operator=(const O&) { shm_stub_unreachable("Map"); }
};
template <typename PlainObjectType, int MapOptions, typename StrideType>
class Map<const PlainObjectType, MapOptions, StrideType> : public PlainObjectType {
  public:
    typedef typename PlainObjectType::Scalar Scalar;
    Map(const Scalar*) { shm_stub_unreachable("Map"); }
    Map(const Scalar*, Index) { shm_stub_unreachable("Map"); }
    Map(const Scalar*, Index, Index) { shm_stub_unreachable("Map"); }
};

template <typename Scalar_, typename StorageIndex_ = int>
class Triplet {
  public:
    Triplet() : r_(0), c_(0), v_() {}
    Triplet(const StorageIndex_& i, const StorageIndex_& j, const Scalar_& v = Scalar_()) : r_(i), c_(j), v_(v) {}
    const StorageIndex_& row() const { return r_; }
    const StorageIndex_& col() const { return c_; }
    const Scalar_& value() const { return v_; }

  private:
    StorageIndex_ r_, c_;
    Scalar_ v_;
};

// Column-compressed like Eigen's default.  setFromTriplets sums duplicates (Eigen's documented behaviour) and leaves the
// entries of a column sorted by row.
template <typename Scalar_, int Options_ = 0, typename StorageIndex_ = int>
class SparseMatrix {
  public:
    typedef Scalar_ Scalar;
    typedef typename NumTraits<Scalar_>::Real RealScalar;
    typedef StorageIndex_ StorageIndex;
    SparseMatrix() : outer_(1, 0) {}
    SparseMatrix(Index r, Index c) : rows_(r), cols_(c), outer_((size_t)c + 1, 0) {}
    Index rows() const { return rows_; }
    Index cols() const { return cols_; }
    Index nonZeros() const { return (Index)val_.size(); }
    Index outerSize() const { return cols_; }
    Index innerSize() const { return rows_; }
    void resize(Index r, Index c) {
        rows_ = r;
        cols_ = c;
        outer_.assign((size_t)c + 1, 0);
        inner_.clear();
        val_.clear();
    }
    void reserve(Index) {}
    template <typename V>
    void reserve(const V&) {}
    void setZero() { resize(rows_, cols_); }
    void makeCompressed() {}
    bool isCompressed() const { return true; }
    template <typename It>
    void setFromTriplets(It b, It e) {
        typedef std::pair<std::pair<Index, Index>, Scalar> Entry;  // (col, row) -> value
        std::vector<Entry> t;
        for (It it = b; it != e; ++it) t.push_back(Entry(std::make_pair((Index)it->col(), (Index)it->row()), it->value()));
        std::stable_sort(t.begin(), t.end(), [](const Entry& x, const Entry& y) { return x.first < y.first; });
        outer_.assign((size_t)cols_ + 1, 0);
        inner_.clear();
        val_.clear();
        for (size_t i = 0; i < t.size();) {
            size_t j = i;
            Scalar s = Scalar();
            while (j < t.size() && t[j].first == t[i].first) s += t[j++].second;
            if (t[i].first.first < 0 || t[i].first.first >= cols_ || t[i].first.second < 0 || t[i].first.second >= rows_)
                throw std::out_of_range("Eigen stub SparseMatrix: triplet index out of range");
            inner_.push_back((StorageIndex)t[i].first.second);
            val_.push_back(s);
            outer_[(size_t)t[i].first.first + 1]++;
            i = j;
        }
        for (Index c = 0; c < cols_; c++) outer_[(size_t)c + 1] += outer_[(size_t)c];
    }
    Scalar coeff(Index r, Index c) const {
        for (StorageIndex p = outer_[(size_t)c]; p < outer_[(size_t)c + 1]; p++)
            if (inner_[(size_t)p] == (StorageIndex)r) return val_[(size_t)p];
        return Scalar();
    }
    SparseMatrix transpose() const {
        std::vector<Triplet<Scalar, Index>> t;
        for (Index c = 0; c < cols_; c++)
            for (StorageIndex p = outer_[(size_t)c]; p < outer_[(size_t)c + 1]; p++)
                t.emplace_back(c, (Index)inner_[(size_t)p], val_[(size_t)p]);
        SparseMatrix r(cols_, rows_);
        r.setFromTriplets(t.begin(), t.end());
        return r;
    }
    Matrix<Scalar, Dynamic, 1> operator*(const Matrix<Scalar, Dynamic, 1>& x) const {
        if (x.size() != cols_) shm_stub_unreachable("sparse * vector with mismatched sizes");
        Matrix<Scalar, Dynamic, 1> y(rows_);
        for (Index c = 0; c < cols_; c++)
            for (StorageIndex p = outer_[(size_t)c]; p < outer_[(size_t)c + 1]; p++) y[inner_[(size_t)p]] += val_[(size_t)p] * x[c];
        return y;
    }
    SparseMatrix operator/(const Scalar& s) const {
        SparseMatrix r(*this);
        for (Scalar& v : r.val_) v /= s;
        return r;
    }
    SparseMatrix operator*(const Scalar& s) const {
        SparseMatrix r(*this);
        for (Scalar& v : r.val_) v *= s;
        return r;
    }
    const StorageIndex* outerIndexPtr() const { return outer_.data(); }
    const StorageIndex* innerIndexPtr() const { return inner_.data(); }
    const Scalar* valuePtr() const { return val_.data(); }
    StorageIndex* outerIndexPtr() { return outer_.data(); }
    StorageIndex* innerIndexPtr() { return inner_.data(); }
    Scalar* valuePtr() { return val_.data(); }
    class InnerIterator {
      public:
        InnerIterator(const SparseMatrix& m, Index outer)
            : m_(&m), c_(outer), p_(m.outer_[(size_t)outer]), end_(m.outer_[(size_t)outer + 1]) {}
        InnerIterator& operator++() {
            ++p_;
            return *this;
        }
        operator bool() const { return p_ < end_; }
        Scalar value() const { return m_->val_[(size_t)p_]; }
        Scalar& valueRef() { return const_cast<SparseMatrix*>(m_)->val_[(size_t)p_]; }
        Index row() const { return (Index)m_->inner_[(size_t)p_]; }
        Index col() const { return c_; }
        Index index() const { return row(); }

      private:
        const SparseMatrix* m_;
        Index c_;
        StorageIndex p_, end_;
    };

    // ---- declared only: abort when reached
    template <typename S2, int R2, int C2, int O2, int MR2, int MC2>
    SparseMatrix(const Matrix<S2, R2, C2, O2, MR2, MC2>&) { shm_stub_unreachable("dense -> sparse"); }
    void setIdentity() { shm_stub_unreachable("SparseMatrix::setIdentity"); }
    Scalar& insert(Index, Index) { shm_stub_unreachable("SparseMatrix::insert"); }
    Scalar& coeffRef(Index, Index) { shm_stub_unreachable("SparseMatrix::coeffRef"); }
    SparseMatrix adjoint() const { shm_stub_unreachable("SparseMatrix::adjoint"); }
    SparseMatrix conjugate() const { shm_stub_unreachable("SparseMatrix::conjugate"); }
    SparseMatrix pruned() const { shm_stub_unreachable("SparseMatrix::pruned"); }
    SparseMatrix pruned(const RealScalar&) const { shm_stub_unreachable("SparseMatrix::pruned"); }
    SparseMatrix cwiseAbs() const { shm_stub_unreachable("SparseMatrix::cwiseAbs"); }
    Matrix<Scalar, Dynamic, 1> diagonal() const { shm_stub_unreachable("SparseMatrix::diagonal"); }
    Matrix<Scalar, Dynamic, Dynamic> toDense() const { shm_stub_unreachable("SparseMatrix::toDense"); }
    SparseMatrix block(Index, Index, Index, Index) const { shm_stub_unreachable("SparseMatrix::block"); }
    RealScalar norm() const { shm_stub_unreachable("SparseMatrix::norm"); }
    RealScalar squaredNorm() const { shm_stub_unreachable("SparseMatrix::squaredNorm"); }
    Scalar sum() const { shm_stub_unreachable("SparseMatrix::sum"); }
    template <typename NewScalar>
    SparseMatrix<NewScalar, Options_, StorageIndex_> cast() const { shm_stub_unreachable("SparseMatrix::cast"); }
    SparseMatrix<RealScalar, Options_, StorageIndex_> real() const { shm_stub_unreachable("SparseMatrix::real"); }
    SparseMatrix<RealScalar, Options_, StorageIndex_> imag() const { shm_stub_unreachable("SparseMatrix::imag"); }
    template <typename O>
    SparseMatrix& operator+=(const O&) { shm_stub_unreachable("SparseMatrix::operator+="); }
    template <typename O>
    SparseMatrix& operator-=(const O&) { shm_stub_unreachable("SparseMatrix::operator-="); }
    template <typename O>
    SparseMatrix& operator*=(const O&) { shm_stub_unreachable("SparseMatrix::operator*="); }
    SparseMatrix operator*(const SparseMatrix&) const { shm_stub_unreachable("sparse * sparse"); }
    template <int R2, int C2, int O2, int MR2, int MC2>
    Matrix<Scalar, Dynamic, C2> operator*(const Matrix<Scalar, R2, C2, O2, MR2, MC2>&) const { shm_stub_unreachable("sparse * dense"); }
    SparseMatrix operator+(const SparseMatrix&) const { shm_stub_unreachable("sparse + sparse"); }
    SparseMatrix operator-(const SparseMatrix&) const { shm_stub_unreachable("sparse - sparse"); }
    SparseMatrix operator-() const { shm_stub_unreachable("-sparse"); }

  private:
    Index rows_ = 0, cols_ = 0;
    std::vector<StorageIndex> outer_, inner_;
    std::vector<Scalar> val_;
};
template <typename S, int O, typename I>
SparseMatrix<S, O, I> operator*(const S& s, const SparseMatrix<S, O, I>& m) {
    return m * s;
}

// ---- sparse direct solvers --------------------------------------------------------------------------------------------
template <typename T>
struct COLAMDOrdering {};
template <typename T>
struct AMDOrdering {};
template <typename T>
struct NaturalOrdering {};

// The harness installs the routine that actually solves A x = b (column-compressed arrays, double).
typedef void (*shm_stub_solve_fn)(int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowidx, const double* val,
                                  const double* rhs, double* x);
inline shm_stub_solve_fn& shm_stub_solver() {
    static shm_stub_solve_fn fn = nullptr;
    return fn;
}

template <typename MatrixType>
class SparseSolverStub {
  public:
    typedef typename MatrixType::Scalar Scalar;
    SparseSolverStub() {}
    ComputationInfo info() const { return Success; }
    std::string lastErrorMessage() const { return std::string(); }
    Index rank() const { shm_stub_unreachable("rank"); }
    void setPivotThreshold(double) {}
    void analyzePattern(const MatrixType&) { shm_stub_unreachable("analyzePattern"); }
    void factorize(const MatrixType&) { shm_stub_unreachable("factorize"); }
};

// compute(): accepted and ignored; solve(): abort.  (The reference's only use: src/signed_heat_grid_solver.cpp:30.)
template <typename MatrixType, int UpLo = Lower, typename Ordering = AMDOrdering<int>>
class SimplicialLDLT : public SparseSolverStub<MatrixType> {
  public:
    void compute(const MatrixType&) {}
    template <typename B>
    Matrix<typename MatrixType::Scalar, Dynamic, 1> solve(const B&) const { shm_stub_unreachable("SimplicialLDLT::solve"); }
};
template <typename MatrixType, int UpLo = Lower, typename Ordering = AMDOrdering<int>>
class SimplicialLLT : public SimplicialLDLT<MatrixType, UpLo, Ordering> {};

namespace shm_stub_detail {
template <typename S>
struct SolveThroughCallback {
    template <typename M>
    static Matrix<S, Dynamic, 1> run(const M&, const Matrix<S, Dynamic, 1>&) { shm_stub_unreachable("SparseLU::solve for this scalar type"); }
};
template <>
struct SolveThroughCallback<double> {
    template <typename M>
    static Matrix<double, Dynamic, 1> run(const M& A, const Matrix<double, Dynamic, 1>& rhs) {
        if (!shm_stub_solver()) throw std::runtime_error("Eigen stub: no linear solver callback installed");
        const Index n = A.rows(), nnz = A.nonZeros();
        std::vector<int64_t> cp(A.outerIndexPtr(), A.outerIndexPtr() + A.cols() + 1), ri(A.innerIndexPtr(), A.innerIndexPtr() + nnz);
        Matrix<double, Dynamic, 1> x(n);
        shm_stub_solver()((int64_t)n, (int64_t)nnz, cp.data(), ri.data(), A.valuePtr(), rhs.data(), x.data());
        return x;
    }
};
}  // namespace shm_stub_detail

// compute() keeps a copy of the matrix, solve() hands matrix and right-hand side to the installed callback.
template <typename MatrixType, typename Ordering = COLAMDOrdering<int>>
class SparseLU : public SparseSolverStub<MatrixType> {
  public:
    void compute(const MatrixType& m) { mat_ = m; }
    Matrix<typename MatrixType::Scalar, Dynamic, 1> solve(const Matrix<typename MatrixType::Scalar, Dynamic, 1>& rhs) const {
        return shm_stub_detail::SolveThroughCallback<typename MatrixType::Scalar>::run(mat_, rhs);
    }

  private:
    MatrixType mat_;
};
template <typename MatrixType, typename Ordering = COLAMDOrdering<int>>
class SparseQR : public SparseSolverStub<MatrixType> {
  public:
    void compute(const MatrixType&) { shm_stub_unreachable("SparseQR"); }
    template <typename B>
    Matrix<typename MatrixType::Scalar, Dynamic, 1> solve(const B&) const { shm_stub_unreachable("SparseQR::solve"); }
};

template <typename MatrixType>
class JacobiSVD {
  public:
    JacobiSVD() {}
    JacobiSVD(const MatrixType&, unsigned int = 0) { shm_stub_unreachable("JacobiSVD"); }
    MatrixType matrixU() const { shm_stub_unreachable("JacobiSVD"); }
    MatrixType matrixV() const { shm_stub_unreachable("JacobiSVD"); }
    Matrix<typename MatrixType::Scalar, Dynamic, 1> singularValues() const { shm_stub_unreachable("JacobiSVD"); }
};

template <typename MatrixType>
class SelfAdjointEigenSolver {
  public:
    SelfAdjointEigenSolver() {}
    explicit SelfAdjointEigenSolver(const MatrixType&) { shm_stub_unreachable("SelfAdjointEigenSolver"); }
    MatrixType eigenvectors() const { shm_stub_unreachable("SelfAdjointEigenSolver"); }
    Matrix<typename MatrixType::Scalar, Dynamic, 1> eigenvalues() const { shm_stub_unreachable("SelfAdjointEigenSolver"); }
    ComputationInfo info() const { return Success; }
};

}  // namespace Eigen
