// shm_eigen_stub.h -- TEST INFRASTRUCTURE.  A stand-in for the slice of Eigen's *interface* that geometry-central's
// point-cloud / tufted-cover sources mention, so that those sources compile from where they lie under
// /root/reference/deps/geometry-central (the real Eigen is fetched by geometry-central's configure step and is absent
// from this image).  Matrix<T, ...> works as a CONTAINER (size, resize, element access, fill, copy) because
// geometry-central keeps all of its per-element data (MeshData, utilities/mesh_data.h:195) in Eigen vectors; nothing
// numerical is implemented: the code path exercised by the oracle (kNN -> local Delaunay triangulations -> triangle-soup
// mesh -> mollification -> tufted cover -> intrinsic Delaunay flips -> vertex dual areas, mean edge length) never does
// linear algebra, and every such operation below aborts if it is ever reached.
#pragma once
#include <complex>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <memory>
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
class Matrix {
  public:
    typedef Scalar_ Scalar;
    typedef typename NumTraits<Scalar_>::Real RealScalar;
    enum { RowsAtCompileTime = Rows_, ColsAtCompileTime = Cols_ };
    Matrix() { init(Rows_ == Dynamic ? 0 : Rows_, Cols_ == Dynamic ? 0 : Cols_); }
    explicit Matrix(Index n) { Cols_ == 1 || Cols_ == Dynamic ? init(n, 1) : init(1, n); }
    Matrix(Index r, Index c) { init(r, c); }
    template <typename A, typename B, typename C>
    Matrix(const A&, const B&, const C&) { init(Rows_ == Dynamic ? 0 : Rows_, Cols_ == Dynamic ? 0 : Cols_); }
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
    void setZero(Index n) { resize(n); fill(Scalar()); }
    void setZero(Index r, Index c) { resize(r, c); fill(Scalar()); }
    void setConstant(const Scalar& v) { fill(v); }
    void fill(const Scalar& v) { for (auto& b : store_) b.v = v; }
    void setOnes() { shm_stub_unreachable("Matrix::setOnes"); }
    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(Index n) { Matrix m(n); return m; }
    static Matrix Zero(Index r, Index c) { Matrix m(r, c); return m; }
    static Matrix Ones(Index) { return Matrix(); }
    static Matrix Ones(Index, Index) { return Matrix(); }
    static Matrix Identity() { return Matrix(); }
    static Matrix Identity(Index, Index) { return Matrix(); }
    static Matrix Constant(Index n, const Scalar& v) { Matrix m(n); m.fill(v); return m; }
    static Matrix Constant(Index r, Index c, const Scalar& v) { Matrix m(r, c); m.fill(v); return m; }
    static Matrix Random(Index) { return Matrix(); }
    static Matrix Random(Index, Index) { return Matrix(); }
    Matrix<Scalar, Dynamic, 1> col(Index) const { return Matrix<Scalar, Dynamic, 1>(); }
    Matrix<Scalar, 1, Dynamic> row(Index) const { return Matrix<Scalar, 1, Dynamic>(); }
    Matrix<Scalar, Dynamic, Dynamic> transpose() const { return Matrix<Scalar, Dynamic, Dynamic>(); }
    Matrix<Scalar, Dynamic, Dynamic> adjoint() const { return Matrix<Scalar, Dynamic, Dynamic>(); }
    Matrix<Scalar, Dynamic, Dynamic> inverse() const { return Matrix<Scalar, Dynamic, Dynamic>(); }
    Matrix<Scalar, Dynamic, Dynamic> asDiagonal() const { return Matrix<Scalar, Dynamic, Dynamic>(); }
    Matrix conjugate() const { return Matrix(); }
    Matrix cwiseAbs() const { return Matrix(); }
    Matrix cwiseInverse() const { return Matrix(); }
    Matrix array() const { return Matrix(); }
    Matrix matrix() const { return Matrix(); }
    Matrix head(Index) const { return Matrix(); }
    Matrix tail(Index) const { return Matrix(); }
    Matrix segment(Index, Index) const { return Matrix(); }
    Matrix block(Index, Index, Index, Index) const { return Matrix(); }
    Matrix normalized() const { return Matrix(); }
    RealScalar norm() const { shm_stub_unreachable("Matrix::norm"); }
    RealScalar squaredNorm() const { shm_stub_unreachable("Matrix::squaredNorm"); }
    Scalar sum() const { shm_stub_unreachable("Matrix::sum"); }
    Scalar mean() const { shm_stub_unreachable("Matrix::mean"); }
    // The one numerical routine on the exercised path: inCircleTest (src/utilities/elementary_geometry.cpp:8-19) takes the
    // sign of a 4x4 determinant.  Restated as Eigen 3.3's fixed-size 4x4 kernel evaluates it (Eigen/src/LU/Determinant.h,
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
    Scalar maxCoeff() const { shm_stub_unreachable("Matrix::maxCoeff"); }
    Scalar minCoeff() const { shm_stub_unreachable("Matrix::minCoeff"); }
    template <typename O>
    Scalar dot(const O&) const { shm_stub_unreachable("Matrix::dot"); }
    struct SolverStub {
        template <typename B>
        Matrix<Scalar, Dynamic, Dynamic> solve(const B&) const { shm_stub_unreachable("solve"); }
    };
    SolverStub colPivHouseholderQr() const { return SolverStub(); }
    SolverStub householderQr() const { return SolverStub(); }
    SolverStub ldlt() const { return SolverStub(); }
    SolverStub llt() const { return SolverStub(); }
    bool allFinite() const { return true; }
    bool hasNaN() const { return false; }
    template <typename NewScalar>
    Matrix<NewScalar, Rows_, Cols_> cast() const { return Matrix<NewScalar, Rows_, Cols_>(); }
    Matrix<RealScalar, Rows_, Cols_> real() const { return Matrix<RealScalar, Rows_, Cols_>(); }
    Matrix<RealScalar, Rows_, Cols_> imag() const { return Matrix<RealScalar, Rows_, Cols_>(); }
    // comma initialiser:  A << a, b, c;
    struct CommaInit {  // fills row by row, like Eigen's
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
    template <typename O>
    Matrix& operator+=(const O&) { return *this; }
    template <typename O>
    Matrix& operator-=(const O&) { return *this; }
    template <typename O>
    Matrix& operator*=(const O&) { return *this; }
    template <typename O>
    Matrix& operator/=(const O&) { return *this; }
    // any other matrix type converts (expression templates collapse to plain matrices here; contents are not carried)
    template <typename S2, int R2, int C2, int O2, int MR2, int MC2>
    Matrix(const Matrix<S2, R2, C2, O2, MR2, MC2>&) { init(Rows_ == Dynamic ? 0 : Rows_, Cols_ == Dynamic ? 0 : Cols_); }

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
    Scalar& at(Index i) { return store_[(size_t)i].v; }
    const Scalar& at(Index i) const { return store_[(size_t)i].v; }
};

template <typename S, int R, int C, int O, int MR, int MC, typename Rhs>
Matrix<S, Dynamic, Dynamic> operator*(const Matrix<S, R, C, O, MR, MC>&, const Rhs&) { return Matrix<S, Dynamic, Dynamic>(); }
template <typename S, int R, int C, int O, int MR, int MC>
Matrix<S, R, C> operator*(const S&, const Matrix<S, R, C, O, MR, MC>&) { return Matrix<S, R, C>(); }
template <typename S, int R, int C, int O, int MR, int MC, typename Rhs>
Matrix<S, R, C> operator+(const Matrix<S, R, C, O, MR, MC>&, const Rhs&) { return Matrix<S, R, C>(); }
template <typename S, int R, int C, int O, int MR, int MC, typename Rhs>
Matrix<S, R, C> operator-(const Matrix<S, R, C, O, MR, MC>&, const Rhs&) { return Matrix<S, R, C>(); }
template <typename S, int R, int C, int O, int MR, int MC>
Matrix<S, R, C> operator-(const Matrix<S, R, C, O, MR, MC>&) { return Matrix<S, R, C>(); }
template <typename S, int R, int C, int O, int MR, int MC>
Matrix<S, R, C> operator/(const Matrix<S, R, C, O, MR, MC>&, const S&) { return Matrix<S, R, C>(); }

template <typename Derived>
class MatrixBase {
  public:
    typedef double Scalar;
    Index rows() const { return 0; }
    Index cols() const { return 0; }
    Index size() const { return 0; }
    Scalar operator()(Index) const { shm_stub_unreachable("MatrixBase::operator()"); }
    Scalar operator()(Index, Index) const { shm_stub_unreachable("MatrixBase::operator()"); }
    template <typename NewScalar>
    Matrix<NewScalar, Dynamic, Dynamic> cast() const { return Matrix<NewScalar, Dynamic, Dynamic>(); }
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
    Map(const Scalar*) {}
    Map(const Scalar*, Index) {}
    Map(const Scalar*, Index, Index) {}
    template <typename O>
    Map& operator=(const O&) { return *this; }
};
template <typename PlainObjectType, int MapOptions, typename StrideType>
class Map<const PlainObjectType, MapOptions, StrideType> : public PlainObjectType {
  public:
    typedef typename PlainObjectType::Scalar Scalar;
    Map(const Scalar*) {}
    Map(const Scalar*, Index) {}
    Map(const Scalar*, Index, Index) {}
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

template <typename Scalar_, int Options_ = 0, typename StorageIndex_ = int>
class SparseMatrix {
  public:
    typedef Scalar_ Scalar;
    typedef typename NumTraits<Scalar_>::Real RealScalar;
    typedef StorageIndex_ StorageIndex;
    SparseMatrix() {}
    SparseMatrix(Index, Index) {}
    template <typename S2, int R2, int C2, int O2, int MR2, int MC2>
    SparseMatrix(const Matrix<S2, R2, C2, O2, MR2, MC2>&) {}
    Index rows() const { return 0; }
    Index cols() const { return 0; }
    Index nonZeros() const { return 0; }
    Index outerSize() const { return 0; }
    Index innerSize() const { return 0; }
    void resize(Index, Index) {}
    void reserve(Index) {}
    template <typename V>
    void reserve(const V&) {}
    void setZero() {}
    void setIdentity() {}
    void makeCompressed() {}
    bool isCompressed() const { return true; }
    template <typename It>
    void setFromTriplets(It, It) { shm_stub_unreachable("SparseMatrix::setFromTriplets"); }
    Scalar& insert(Index, Index) { shm_stub_unreachable("SparseMatrix::insert"); }
    Scalar& coeffRef(Index, Index) { shm_stub_unreachable("SparseMatrix::coeffRef"); }
    Scalar coeff(Index, Index) const { shm_stub_unreachable("SparseMatrix::coeff"); }
    SparseMatrix transpose() const { return SparseMatrix(); }
    SparseMatrix adjoint() const { return SparseMatrix(); }
    SparseMatrix conjugate() const { return SparseMatrix(); }
    SparseMatrix pruned() const { return SparseMatrix(); }
    SparseMatrix pruned(const RealScalar&) const { return SparseMatrix(); }
    SparseMatrix cwiseAbs() const { return SparseMatrix(); }
    Matrix<Scalar, Dynamic, 1> diagonal() const { return Matrix<Scalar, Dynamic, 1>(); }
    Matrix<Scalar, Dynamic, Dynamic> toDense() const { return Matrix<Scalar, Dynamic, Dynamic>(); }
    SparseMatrix block(Index, Index, Index, Index) const { return SparseMatrix(); }
    RealScalar norm() const { shm_stub_unreachable("SparseMatrix::norm"); }
    RealScalar squaredNorm() const { shm_stub_unreachable("SparseMatrix::squaredNorm"); }
    Scalar sum() const { shm_stub_unreachable("SparseMatrix::sum"); }
    const StorageIndex* outerIndexPtr() const { return nullptr; }
    const StorageIndex* innerIndexPtr() const { return nullptr; }
    const Scalar* valuePtr() const { return nullptr; }
    StorageIndex* outerIndexPtr() { return nullptr; }
    StorageIndex* innerIndexPtr() { return nullptr; }
    Scalar* valuePtr() { return nullptr; }
    template <typename NewScalar>
    SparseMatrix<NewScalar, Options_, StorageIndex_> cast() const { return SparseMatrix<NewScalar, Options_, StorageIndex_>(); }
    SparseMatrix<RealScalar, Options_, StorageIndex_> real() const { return SparseMatrix<RealScalar, Options_, StorageIndex_>(); }
    SparseMatrix<RealScalar, Options_, StorageIndex_> imag() const { return SparseMatrix<RealScalar, Options_, StorageIndex_>(); }
    class InnerIterator {
      public:
        InnerIterator(const SparseMatrix&, Index) {}
        InnerIterator& operator++() { return *this; }
        operator bool() const { return false; }
        Scalar value() const { return Scalar(); }
        Scalar& valueRef() { shm_stub_unreachable("InnerIterator::valueRef"); }
        Index row() const { return 0; }
        Index col() const { return 0; }
        Index index() const { return 0; }
    };
    template <typename O>
    SparseMatrix& operator+=(const O&) { return *this; }
    template <typename O>
    SparseMatrix& operator-=(const O&) { return *this; }
    template <typename O>
    SparseMatrix& operator*=(const O&) { return *this; }
};

template <typename S, int O, typename I>
SparseMatrix<S, O, I> operator*(const SparseMatrix<S, O, I>&, const SparseMatrix<S, O, I>&) { return SparseMatrix<S, O, I>(); }
template <typename S, int O, typename I, int R, int C, int O2, int MR, int MC>
Matrix<S, Dynamic, C> operator*(const SparseMatrix<S, O, I>&, const Matrix<S, R, C, O2, MR, MC>&) { return Matrix<S, Dynamic, C>(); }
template <typename S, int O, typename I>
SparseMatrix<S, O, I> operator*(const S&, const SparseMatrix<S, O, I>&) { return SparseMatrix<S, O, I>(); }
template <typename S, int O, typename I>
SparseMatrix<S, O, I> operator*(const SparseMatrix<S, O, I>&, const S&) { return SparseMatrix<S, O, I>(); }
template <typename S, int O, typename I>
SparseMatrix<S, O, I> operator+(const SparseMatrix<S, O, I>&, const SparseMatrix<S, O, I>&) { return SparseMatrix<S, O, I>(); }
template <typename S, int O, typename I>
SparseMatrix<S, O, I> operator-(const SparseMatrix<S, O, I>&, const SparseMatrix<S, O, I>&) { return SparseMatrix<S, O, I>(); }
template <typename S, int O, typename I>
SparseMatrix<S, O, I> operator-(const SparseMatrix<S, O, I>&) { return SparseMatrix<S, O, I>(); }

// sparse direct solvers: declared so that geometry-central's solver wrappers compile; never run on the oracle's path
template <typename T>
struct COLAMDOrdering {};
template <typename T>
struct AMDOrdering {};
template <typename T>
struct NaturalOrdering {};
template <typename MatrixType>
class SparseSolverStub {
  public:
    typedef typename MatrixType::Scalar Scalar;
    SparseSolverStub() {}
    explicit SparseSolverStub(const MatrixType&) { shm_stub_unreachable("sparse factorisation"); }
    void compute(const MatrixType&) { shm_stub_unreachable("sparse factorisation"); }
    void analyzePattern(const MatrixType&) { shm_stub_unreachable("sparse factorisation"); }
    void factorize(const MatrixType&) { shm_stub_unreachable("sparse factorisation"); }
    template <typename B>
    Matrix<Scalar, Dynamic, 1> solve(const B&) const { shm_stub_unreachable("sparse solve"); }
    ComputationInfo info() const { return Success; }
    Index rank() const { return 0; }
    void setPivotThreshold(double) {}
};
template <typename MatrixType, int UpLo = Lower, typename Ordering = AMDOrdering<int>>
class SimplicialLDLT : public SparseSolverStub<MatrixType> {};
template <typename MatrixType, int UpLo = Lower, typename Ordering = AMDOrdering<int>>
class SimplicialLLT : public SparseSolverStub<MatrixType> {};
template <typename MatrixType, typename Ordering = COLAMDOrdering<int>>
class SparseLU : public SparseSolverStub<MatrixType> {};
template <typename MatrixType, typename Ordering = COLAMDOrdering<int>>
class SparseQR : public SparseSolverStub<MatrixType> {};

template <typename MatrixType>
class JacobiSVD {
  public:
    JacobiSVD() {}
    JacobiSVD(const MatrixType&, unsigned int = 0) {}
    MatrixType matrixU() const { return MatrixType(); }
    MatrixType matrixV() const { return MatrixType(); }
    Matrix<typename MatrixType::Scalar, Dynamic, 1> singularValues() const { return Matrix<typename MatrixType::Scalar, Dynamic, 1>(); }
};

template <typename MatrixType>
class SelfAdjointEigenSolver {
  public:
    SelfAdjointEigenSolver() {}
    explicit SelfAdjointEigenSolver(const MatrixType&) {}
    MatrixType eigenvectors() const { return MatrixType(); }
    Matrix<typename MatrixType::Scalar, Dynamic, 1> eigenvalues() const { return Matrix<typename MatrixType::Scalar, Dynamic, 1>(); }
    ComputationInfo info() const { return Success; }
};

}  // namespace Eigen
