// mini_eigen.hpp - TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Not part of the product, never shipped.
//
// A small, eager (no expression templates) stand-in for the subset of Eigen 3.4 that the reference's own sources on the
// MPC path use, so that those sources compile UNMODIFIED from /root/reference (Eigen is not installed in this image and
// there is no network):
//   ws/src/controllers/src/mit_controller/{mpc,gait,simple_gait_sequencer,adaptive_gait_sequencer,
//                                          mpc_trajectory_planner,raibert_foot_step_planner}.cpp
//   ws/src/common/include/common/{filters,quaternion_operations,eigen_util,model_interface,state_interface}.hpp
// Written from the published Eigen API documentation (class names, member names, argument order, storage order,
// Quaternion coefficient order x,y,z,w); the formulas are the textbook ones (quaternion product, Rodrigues rotation
// matrix, cofactor 3x3 inverse, slerp).  Element-wise results agree with Eigen up to rounding of the last bits; the
// bit-exact parity claims of this repo (contact masks, bounds) only involve reference code that uses std::array + fmod
// and no Eigen arithmetic at all.
#pragma once
// Eigen/Core pulls these standard headers in; the reference relies on that (std::unique_ptr, std::min_element, ...)
#include <algorithm>
#include <array>
#include <cassert>
#include <complex>
#include <cstring>
#include <functional>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <type_traits>

#define EIGEN_STATIC_ASSERT_VECTOR_SPECIFIC_SIZE(TYPE, SIZE) \
  static_assert(TYPE::SizeAtCompileTime == SIZE, "vector of this size expected")
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

constexpr int Dynamic = -1;
enum StorageOptions { ColMajor = 0, RowMajor = 1 };
enum DecompositionOptions { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };
using Index = std::ptrdiff_t;
constexpr int kMaxDynamic = 16;  // inline capacity of a dynamic dimension (only .segment() results are dynamic here)

template <class T> struct traits;
template <class S, int R, int C, int Opt = ColMajor> class Matrix;
template <class X, int BR, int BC> class Block;
template <class X> class Transpose;
template <class X> class Diagonal;
template <class P> class Ref;
template <class P> class Map;
template <class S> class Quaternion;
template <class S> class AngleAxis;

constexpr int pick_dim(int a, int b) { return a != Dynamic ? a : b; }

// ---------------------------------------------------------------------------------------------------------------------
template <class D>
class MatrixBase {
 public:
  using Scalar = typename traits<D>::Scalar;
  enum {
    RowsAtCompileTime = traits<D>::Rows,
    ColsAtCompileTime = traits<D>::Cols,
    SizeAtCompileTime = (traits<D>::Rows == Dynamic || traits<D>::Cols == Dynamic) ? Dynamic
                                                                                    : traits<D>::Rows * traits<D>::Cols,
    IsVectorAtCompileTime = (traits<D>::Rows == 1 || traits<D>::Cols == 1)
  };
  using PlainObject = Matrix<Scalar, traits<D>::Rows, traits<D>::Cols>;

  const D &derived() const { return *static_cast<const D *>(this); }
  D &derived() { return *static_cast<D *>(this); }
  Index rows() const { return derived().rows_(); }
  Index cols() const { return derived().cols_(); }
  Index size() const { return rows() * cols(); }

  Scalar coeff(Index i, Index j) const { return derived().coeff_(i, j); }
  Scalar coeff(Index i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
  Scalar operator()(Index i, Index j) const { return coeff(i, j); }
  Scalar operator()(Index i) const { return coeff(i); }
  Scalar operator[](Index i) const { return coeff(i); }
  Scalar x() const { return coeff(0); }
  Scalar y() const { return coeff(1); }
  Scalar z() const { return coeff(2); }
  Scalar w() const { return coeff(3); }

  PlainObject eval() const { return PlainObject(*this); }
  Scalar squaredNorm() const {
    Scalar s = 0;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) s += coeff(i, j) * coeff(i, j);
    return s;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  Scalar sum() const {
    Scalar s = 0;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) s += coeff(i, j);
    return s;
  }
  PlainObject normalized() const {
    PlainObject r(*this);
    const Scalar n = norm();
    if (n > Scalar(0)) r /= n;
    return r;
  }
  bool hasNaN() const {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i)
        if (std::isnan(coeff(i, j))) return true;
    return false;
  }
  template <class O>
  Scalar dot(const MatrixBase<O> &o) const {
    Scalar s = 0;
    for (Index i = 0; i < size(); ++i) s += coeff(i) * o.coeff(i);
    return s;
  }
  template <class O>
  Matrix<Scalar, 3, 1> cross(const MatrixBase<O> &o) const {
    return Matrix<Scalar, 3, 1>(coeff(1) * o.coeff(2) - coeff(2) * o.coeff(1), coeff(2) * o.coeff(0) - coeff(0) * o.coeff(2),
                                coeff(0) * o.coeff(1) - coeff(1) * o.coeff(0));
  }
  Transpose<const D> transpose() const { return Transpose<const D>(derived()); }
  template <int BR, int BC>
  Block<const D, BR, BC> block(Index i, Index j) const {
    return Block<const D, BR, BC>(derived(), i, j);
  }
  Matrix<Scalar, Dynamic, 1> segment(Index start, Index n) const;
  Matrix<Scalar, traits<D>::Rows, 1> col(Index j) const {
    Matrix<Scalar, traits<D>::Rows, 1> r;
    r.resize_(rows(), 1);
    for (Index i = 0; i < rows(); ++i) r.coeffRef(i, 0) = coeff(i, j);
    return r;
  }
  Matrix<Scalar, pick_dim(traits<D>::Rows, traits<D>::Cols), pick_dim(traits<D>::Rows, traits<D>::Cols)> asDiagonal() const {
    Matrix<Scalar, pick_dim(traits<D>::Rows, traits<D>::Cols), pick_dim(traits<D>::Rows, traits<D>::Cols)> r;
    r.setZero();
    for (Index i = 0; i < size(); ++i) r.coeffRef(i, i) = coeff(i);
    return r;
  }
  Scalar determinant() const {
    static_assert(traits<D>::Rows == 3 && traits<D>::Cols == 3, "mini_eigen: determinant() is 3x3 only");
    const D &a = derived();
    return a.coeff_(0, 0) * (a.coeff_(1, 1) * a.coeff_(2, 2) - a.coeff_(1, 2) * a.coeff_(2, 1)) -
           a.coeff_(0, 1) * (a.coeff_(1, 0) * a.coeff_(2, 2) - a.coeff_(1, 2) * a.coeff_(2, 0)) +
           a.coeff_(0, 2) * (a.coeff_(1, 0) * a.coeff_(2, 1) - a.coeff_(1, 1) * a.coeff_(2, 0));
  }
  // cofactor inverse (what Eigen does for fixed 3x3 as well)
  Matrix<Scalar, 3, 3> inverse() const {
    static_assert(traits<D>::Rows == 3 && traits<D>::Cols == 3, "mini_eigen: inverse() is 3x3 only");
    const D &a = derived();
    Matrix<Scalar, 3, 3> r;
    const Scalar c00 = a.coeff_(1, 1) * a.coeff_(2, 2) - a.coeff_(1, 2) * a.coeff_(2, 1);
    const Scalar c10 = a.coeff_(1, 2) * a.coeff_(2, 0) - a.coeff_(1, 0) * a.coeff_(2, 2);
    const Scalar c20 = a.coeff_(1, 0) * a.coeff_(2, 1) - a.coeff_(1, 1) * a.coeff_(2, 0);
    const Scalar invdet = Scalar(1) / (a.coeff_(0, 0) * c00 + a.coeff_(0, 1) * c10 + a.coeff_(0, 2) * c20);
    r(0, 0) = c00 * invdet;
    r(1, 0) = c10 * invdet;
    r(2, 0) = c20 * invdet;
    r(0, 1) = (a.coeff_(0, 2) * a.coeff_(2, 1) - a.coeff_(0, 1) * a.coeff_(2, 2)) * invdet;
    r(1, 1) = (a.coeff_(0, 0) * a.coeff_(2, 2) - a.coeff_(0, 2) * a.coeff_(2, 0)) * invdet;
    r(2, 1) = (a.coeff_(0, 1) * a.coeff_(2, 0) - a.coeff_(0, 0) * a.coeff_(2, 1)) * invdet;
    r(0, 2) = (a.coeff_(0, 1) * a.coeff_(1, 2) - a.coeff_(0, 2) * a.coeff_(1, 1)) * invdet;
    r(1, 2) = (a.coeff_(0, 2) * a.coeff_(1, 0) - a.coeff_(0, 0) * a.coeff_(1, 2)) * invdet;
    r(2, 2) = (a.coeff_(0, 0) * a.coeff_(1, 1) - a.coeff_(0, 1) * a.coeff_(1, 0)) * invdet;
    return r;
  }
  // decompositions the reference only reaches with options this harness never enables (z_on_plane, quaternion filter)
  struct UnsupportedDecomposition {
    PlainObject pseudoInverse() const { throw std::logic_error("mini_eigen: pseudoInverse() not available"); }
  };
  UnsupportedDecomposition completeOrthogonalDecomposition() const { return UnsupportedDecomposition(); }
};

// writable expressions
template <class D>
class DenseBase : public MatrixBase<D> {
 public:
  using Base = MatrixBase<D>;
  using Scalar = typename Base::Scalar;
  using Base::cols;
  using Base::derived;
  using Base::rows;
  using Base::size;
  using Base::operator();
  using Base::operator[];
  using Base::block;
  using Base::x;
  using Base::y;
  using Base::z;
  using Base::w;

  Scalar &coeffRef(Index i, Index j) { return derived().coeffRef_(i, j); }
  Scalar &coeffRef(Index i) { return cols() == 1 ? coeffRef(i, 0) : coeffRef(0, i); }
  Scalar &operator()(Index i, Index j) { return coeffRef(i, j); }
  Scalar &operator()(Index i) { return coeffRef(i); }
  Scalar &operator[](Index i) { return coeffRef(i); }
  Scalar &x() { return coeffRef(0); }
  Scalar &y() { return coeffRef(1); }
  Scalar &z() { return coeffRef(2); }
  Scalar &w() { return coeffRef(3); }

  D &setConstant(Scalar v) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = v;
    return derived();
  }
  D &setZero() { return setConstant(Scalar(0)); }
  D &setOnes() { return setConstant(Scalar(1)); }
  D &setIdentity() {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = (i == j) ? Scalar(1) : Scalar(0);
    return derived();
  }
  D &noalias() { return derived(); }
  void normalize() {
    const Scalar n = this->norm();
    if (n > Scalar(0)) *this /= n;
  }
  template <class O>
  D &assign(const MatrixBase<O> &o) {
    assert(o.rows() == rows() && o.cols() == cols());
    // evaluate first: the source may alias the destination (all operators are eager, views are not)
    typename MatrixBase<O>::PlainObject tmp(o);
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = tmp.coeff_(i, j);
    return derived();
  }
  template <class O>
  D &operator+=(const MatrixBase<O> &o) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) coeffRef(i, j) += o.coeff(i, j);
    return derived();
  }
  template <class O>
  D &operator-=(const MatrixBase<O> &o) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) coeffRef(i, j) -= o.coeff(i, j);
    return derived();
  }
  D &operator*=(Scalar s) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) coeffRef(i, j) *= s;
    return derived();
  }
  D &operator/=(Scalar s) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) coeffRef(i, j) /= s;
    return derived();
  }
  template <int BR, int BC>
  Block<D, BR, BC> block(Index i, Index j) {
    return Block<D, BR, BC>(derived(), i, j);
  }
  Diagonal<D> diagonal() { return Diagonal<D>(derived()); }

  // comma initialiser (row-major fill order)
  class CommaInitializer {
   public:
    CommaInitializer(D &m, Scalar first) : m_(m), k_(0) { put(first); }
    CommaInitializer &operator,(Scalar v) {
      put(v);
      return *this;
    }
    D &finished() { return m_; }

   private:
    void put(Scalar v) {
      const Index c = m_.cols();
      m_.coeffRef(k_ / c, k_ % c) = v;
      ++k_;
    }
    D &m_;
    Index k_;
  };
  CommaInitializer operator<<(Scalar first) { return CommaInitializer(derived(), first); }
};

// ---------------------------------------------------------------------------------------------------------------------
template <class S, int R, int C, int Opt>
struct traits<Matrix<S, R, C, Opt>> {
  using Scalar = S;
  enum { Rows = R, Cols = C };
};

// storage: a fixed-size matrix is exactly R*C scalars (the reference reinterpret_casts arrays of Vector3d to double*,
// mpc.cpp:458); only the dynamic column vector produced by segment() carries its length
template <class S, int R, int C>
struct DenseStorage {
  static_assert(R != Dynamic && C != Dynamic, "mini_eigen: only Matrix<S, Dynamic, 1> may be dynamic");
  S m_[R * C];
  Index rows() const { return R; }
  Index cols() const { return C; }
  void resize(Index r, Index c) {
    (void)r;
    (void)c;
    assert(r == R && c == C);
  }
};
template <class S>
struct DenseStorage<S, Dynamic, 1> {
  S m_[kMaxDynamic];
  int n_;
  Index rows() const { return n_; }
  Index cols() const { return 1; }
  void resize(Index r, Index c) {
    (void)c;
    assert(c == 1 && r <= kMaxDynamic);
    n_ = (int)r;
  }
};

template <class S, int R, int C, int Opt>
class Matrix : public DenseBase<Matrix<S, R, C, Opt>> {
 public:
  using Base = DenseBase<Matrix>;
  using Scalar = S;

  Matrix() : st_{} {}
  Matrix(const Matrix &) = default;
  Matrix &operator=(const Matrix &) = default;
  // fixed-size vector constructors
  template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 2, int>::type = 0>
  Matrix(S a, S b) : Matrix() {
    st_.m_[0] = a;
    st_.m_[1] = b;
  }
  template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 3, int>::type = 0>
  Matrix(S a, S b, S c) : Matrix() {
    st_.m_[0] = a;
    st_.m_[1] = b;
    st_.m_[2] = c;
  }
  template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 4, int>::type = 0>
  Matrix(S a, S b, S c, S d) : Matrix() {
    st_.m_[0] = a;
    st_.m_[1] = b;
    st_.m_[2] = c;
    st_.m_[3] = d;
  }
  template <class O>
  Matrix(const MatrixBase<O> &o) : Matrix() {
    resize_(o.rows(), o.cols());
    for (Index j = 0; j < o.cols(); ++j)
      for (Index i = 0; i < o.rows(); ++i) coeffRef_(i, j) = o.coeff(i, j);
  }
  template <class O>
  Matrix &operator=(const MatrixBase<O> &o) {
    Matrix tmp(o);
    *this = tmp;
    return *this;
  }

  static Matrix Constant(S v) {
    Matrix m;
    m.setConstant(v);
    return m;
  }
  static Matrix Zero() { return Constant(S(0)); }
  static Matrix Ones() { return Constant(S(1)); }
  static Matrix Identity() {
    Matrix m;
    m.setIdentity();
    return m;
  }
  static Matrix Unit(Index k) {
    Matrix m = Zero();
    m.coeffRef(k) = S(1);
    return m;
  }
  static Matrix UnitX() { return Unit(0); }
  static Matrix UnitY() { return Unit(1); }
  static Matrix UnitZ() { return Unit(2); }
  static Matrix UnitW() { return Unit(3); }

  S *data() { return st_.m_; }
  const S *data() const { return st_.m_; }
  Index innerStride() const { return 1; }
  Index outerStride() const { return (Opt & RowMajor) ? st_.cols() : st_.rows(); }

  // implementation hooks used by the CRTP bases
  Index rows_() const { return st_.rows(); }
  Index cols_() const { return st_.cols(); }
  S coeff_(Index i, Index j) const { return st_.m_[(Opt & RowMajor) ? i * st_.cols() + j : j * st_.rows() + i]; }
  S &coeffRef_(Index i, Index j) { return st_.m_[(Opt & RowMajor) ? i * st_.cols() + j : j * st_.rows() + i]; }
  void resize_(Index r, Index c) { st_.resize(r, c); }

 private:
  DenseStorage<S, R, C> st_;
};

template <class S, int N> using Vector = Matrix<S, N, 1>;
template <class S, int N> using RowVector = Matrix<S, 1, N>;
using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;
using Matrix2d = Matrix<double, 2, 2>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
using VectorXd = Matrix<double, Dynamic, 1>;
using Vector3f = Matrix<float, 3, 1>;
using Matrix3f = Matrix<float, 3, 3>;

template <class D>
Matrix<typename MatrixBase<D>::Scalar, Dynamic, 1> MatrixBase<D>::segment(Index start, Index n) const {
  Matrix<Scalar, Dynamic, 1> r;
  r.resize_(n, 1);
  for (Index i = 0; i < n; ++i) r.coeffRef(i, 0) = coeff(start + i);
  return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// views
template <class X, int BR, int BC>
struct traits<Block<X, BR, BC>> {
  using Scalar = typename traits<typename std::remove_const<X>::type>::Scalar;
  enum { Rows = BR, Cols = BC };
};
template <class X, int BR, int BC>
class Block : public DenseBase<Block<X, BR, BC>> {
 public:
  using Scalar = typename traits<Block>::Scalar;
  Block(X &x, Index i0, Index j0) : x_(x), i0_(i0), j0_(j0) {}
  Block(const Block &) = default;
  Block &operator=(const Block &o) { return this->assign(o); }
  template <class O>
  Block &operator=(const MatrixBase<O> &o) {
    return this->assign(o);
  }
  Index rows_() const { return BR; }
  Index cols_() const { return BC; }
  Scalar coeff_(Index i, Index j) const { return x_.coeff(i0_ + i, j0_ + j); }
  Scalar &coeffRef_(Index i, Index j) { return x_.coeffRef(i0_ + i, j0_ + j); }
  // direct access (only meaningful over column-major direct-access expressions)
  Scalar *data() { return &x_.coeffRef(i0_, j0_); }
  Index innerStride() const { return 1; }
  Index outerStride() const { return x_.outerStride(); }

 private:
  X &x_;
  Index i0_, j0_;
};

template <class X>
struct traits<Transpose<X>> {
  using Scalar = typename traits<typename std::remove_const<X>::type>::Scalar;
  enum { Rows = traits<typename std::remove_const<X>::type>::Cols, Cols = traits<typename std::remove_const<X>::type>::Rows };
};
template <class X>
class Transpose : public MatrixBase<Transpose<X>> {
 public:
  using Scalar = typename traits<Transpose>::Scalar;
  explicit Transpose(X &x) : x_(x) {}
  Index rows_() const { return x_.cols(); }
  Index cols_() const { return x_.rows(); }
  Scalar coeff_(Index i, Index j) const { return x_.coeff(j, i); }

 private:
  X &x_;
};

template <class X>
struct traits<Diagonal<X>> {
  using Scalar = typename traits<typename std::remove_const<X>::type>::Scalar;
  enum { Rows = traits<typename std::remove_const<X>::type>::Rows, Cols = 1 };
};
template <class X>
class Diagonal : public DenseBase<Diagonal<X>> {
 public:
  using Scalar = typename traits<Diagonal>::Scalar;
  explicit Diagonal(X &x) : x_(x) {}
  Diagonal(const Diagonal &) = default;
  template <class O>
  Diagonal &operator=(const MatrixBase<O> &o) {
    return this->assign(o);
  }
  Index rows_() const { return x_.rows() < x_.cols() ? x_.rows() : x_.cols(); }
  Index cols_() const { return 1; }
  Scalar coeff_(Index i, Index) const { return x_.coeff(i, i); }
  Scalar &coeffRef_(Index i, Index) { return x_.coeffRef(i, i); }

 private:
  X &x_;
};

// Map: column-major view of caller memory
template <class P>
struct traits<Map<P>> {
  using Scalar = typename traits<typename std::remove_const<P>::type>::Scalar;
  enum { Rows = traits<typename std::remove_const<P>::type>::Rows, Cols = traits<typename std::remove_const<P>::type>::Cols };
};
template <class P>
class Map : public DenseBase<Map<P>> {
 public:
  using Scalar = typename traits<Map>::Scalar;
  using Ptr = typename std::conditional<std::is_const<P>::value, const Scalar *, Scalar *>::type;
  explicit Map(Ptr p) : p_(const_cast<Scalar *>(p)) {}
  template <class O>
  Map &operator=(const MatrixBase<O> &o) {
    return this->assign(o);
  }
  Index rows_() const { return traits<Map>::Rows; }
  Index cols_() const { return traits<Map>::Cols; }
  Scalar coeff_(Index i, Index j) const { return p_[j * traits<Map>::Rows + i]; }
  Scalar &coeffRef_(Index i, Index j) { return p_[j * traits<Map>::Rows + i]; }
  Scalar *data() { return p_; }
  Index innerStride() const { return 1; }
  Index outerStride() const { return traits<Map>::Rows; }

 private:
  Scalar *p_;
};

// Ref<T>: writable view (pointer + outer stride) of a column-major direct-access expression.
template <class P>
struct traits<Ref<P>> {
  using Scalar = typename traits<typename std::remove_const<P>::type>::Scalar;
  enum { Rows = traits<typename std::remove_const<P>::type>::Rows, Cols = traits<typename std::remove_const<P>::type>::Cols };
};
template <class P>
class Ref : public DenseBase<Ref<P>> {
 public:
  using Scalar = typename traits<Ref>::Scalar;
  template <class X, typename = decltype(std::declval<X &>().data()), typename = decltype(std::declval<X &>().outerStride())>
  Ref(X &x) : p_(x.data()), os_(x.outerStride()) {
    static_assert(int(traits<X>::Rows) == int(traits<Ref>::Rows) && int(traits<X>::Cols) == int(traits<Ref>::Cols), "Ref size");
  }
  template <class X, int BR, int BC>
  Ref(Block<X, BR, BC> &&b) : p_(b.data()), os_(b.outerStride()) {
    static_assert(BR == int(traits<Ref>::Rows) && BC == int(traits<Ref>::Cols), "Ref size");
  }
  Ref(const Ref &) = default;
  template <class O>
  Ref &operator=(const MatrixBase<O> &o) {
    return this->assign(o);
  }
  Ref &operator=(const Ref &o) { return this->assign(o); }
  Index rows_() const { return traits<Ref>::Rows; }
  Index cols_() const { return traits<Ref>::Cols; }
  Scalar coeff_(Index i, Index j) const { return p_[j * os_ + i]; }
  Scalar &coeffRef_(Index i, Index j) { return p_[j * os_ + i]; }
  Scalar *data() { return p_; }
  Index innerStride() const { return 1; }
  Index outerStride() const { return os_; }

 private:
  Scalar *p_;
  Index os_;
};
// Ref<const T>: read-only; binds to any expression (a private copy is taken, as Eigen does for non-referable inputs)
template <class P>
class Ref<const P> : public MatrixBase<Ref<const P>> {
 public:
  using Scalar = typename traits<P>::Scalar;
  template <class O>
  Ref(const MatrixBase<O> &o) : v_(o) {}
  Ref(const Ref &) = default;
  Index rows_() const { return v_.rows(); }
  Index cols_() const { return v_.cols(); }
  Scalar coeff_(Index i, Index j) const { return v_.coeff_(i, j); }
  const Scalar *data() const { return v_.data(); }

 private:
  P v_;
};
template <class P>
struct traits<Ref<const P>> {
  using Scalar = typename traits<P>::Scalar;
  enum { Rows = traits<P>::Rows, Cols = traits<P>::Cols };
};

// ---------------------------------------------------------------------------------------------------------------------
// eager arithmetic
template <class A, class B>
using SumType = Matrix<typename traits<A>::Scalar, pick_dim(traits<A>::Rows, traits<B>::Rows), pick_dim(traits<A>::Cols, traits<B>::Cols)>;

template <class A, class B>
SumType<A, B> operator+(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  SumType<A, B> r;
  r.resize_(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef_(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <class A, class B>
SumType<A, B> operator-(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  SumType<A, B> r;
  r.resize_(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef_(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
template <class A>
typename MatrixBase<A>::PlainObject operator-(const MatrixBase<A> &a) {
  typename MatrixBase<A>::PlainObject r;
  r.resize_(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef_(i, j) = -a.coeff(i, j);
  return r;
}
template <class A, class T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A> &a, T s) {
  typename MatrixBase<A>::PlainObject r;
  r.resize_(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef_(i, j) = a.coeff(i, j) * s;
  return r;
}
template <class A, class T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
typename MatrixBase<A>::PlainObject operator*(T s, const MatrixBase<A> &a) {
  typename MatrixBase<A>::PlainObject r;
  r.resize_(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef_(i, j) = s * a.coeff(i, j);
  return r;
}
template <class A, class T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A> &a, T s) {
  typename MatrixBase<A>::PlainObject r;
  r.resize_(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef_(i, j) = a.coeff(i, j) / s;
  return r;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> r;
  r.resize_(a.rows(), b.cols());
  assert(a.cols() == b.rows());
  for (Index j = 0; j < b.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) {
      typename traits<A>::Scalar s = 0;
      for (Index k = 0; k < a.cols(); ++k) s += a.coeff(i, k) * b.coeff(k, j);
      r.coeffRef_(i, j) = s;
    }
  return r;
}
template <class A>
std::ostream &operator<<(std::ostream &os, const MatrixBase<A> &a) {
  for (Index i = 0; i < a.rows(); ++i) {
    for (Index j = 0; j < a.cols(); ++j) os << (j ? " " : "") << a.coeff(i, j);
    if (i + 1 < a.rows()) os << "\n";
  }
  return os;
}

// ---------------------------------------------------------------------------------------------------------------------
// geometry
template <class S>
class AngleAxis {
 public:
  AngleAxis(S angle, const Matrix<S, 3, 1> &axis) : angle_(angle), axis_(axis) {}
  S angle() const { return angle_; }
  const Matrix<S, 3, 1> &axis() const { return axis_; }

 private:
  S angle_;
  Matrix<S, 3, 1> axis_;
};

template <class S>
class Quaternion {
 public:
  using Scalar = S;
  using Vector3 = Matrix<S, 3, 1>;
  using Matrix3 = Matrix<S, 3, 3>;
  Quaternion() : c_(S(0), S(0), S(0), S(1)) {}
  Quaternion(S w, S x, S y, S z) : c_(x, y, z, w) {}
  Quaternion(const AngleAxis<S> &aa) {
    const S ha = S(0.5) * aa.angle();
    const S s = std::sin(ha);
    c_ = Matrix<S, 4, 1>(s * aa.axis().x(), s * aa.axis().y(), s * aa.axis().z(), std::cos(ha));
  }
  template <class O, typename = typename std::enable_if<int(traits<O>::Rows) == 4 && int(traits<O>::Cols) == 1>::type>
  Quaternion(const MatrixBase<O> &xyzw) : c_(xyzw) {}
  static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }

  S x() const { return c_.x(); }
  S y() const { return c_.y(); }
  S z() const { return c_.z(); }
  S w() const { return c_.w(); }
  S &x() { return c_.x(); }
  S &y() { return c_.y(); }
  S &z() { return c_.z(); }
  S &w() { return c_.w(); }
  const Matrix<S, 4, 1> &coeffs() const { return c_; }
  Matrix<S, 4, 1> &coeffs() { return c_; }
  Vector3 vec() const { return Vector3(x(), y(), z()); }

  S squaredNorm() const { return c_.squaredNorm(); }
  S norm() const { return c_.norm(); }
  void normalize() { c_.normalize(); }
  Quaternion normalized() const {
    Quaternion q(*this);
    q.normalize();
    return q;
  }
  Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
  Quaternion inverse() const {
    const S n2 = squaredNorm();
    return Quaternion(w() / n2, -x() / n2, -y() / n2, -z() / n2);
  }
  S dot(const Quaternion &o) const { return c_.dot(o.c_); }

  Quaternion operator*(const Quaternion &b) const {
    const Quaternion &a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Quaternion &operator*=(const Quaternion &b) {
    *this = *this * b;
    return *this;
  }
  Quaternion operator*(const AngleAxis<S> &aa) const { return *this * Quaternion(aa); }
  // rotation of a vector: v + 2 w (q x v) + 2 q x (q x v)
  template <class O, typename = typename std::enable_if<int(traits<O>::Rows) == 3 && int(traits<O>::Cols) == 1>::type>
  Vector3 operator*(const MatrixBase<O> &v) const {
    const Vector3 q = vec(), vv(v);
    const Vector3 uv = q.cross(vv) * S(2);
    return vv + uv * w() + q.cross(uv);
  }
  Matrix3 toRotationMatrix() const {
    Matrix3 R;
    const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
    const S twx = tx * w(), twy = ty * w(), twz = tz * w();
    const S txx = tx * x(), txy = ty * x(), txz = tz * x();
    const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    R(0, 0) = S(1) - (tyy + tzz);
    R(0, 1) = txy - twz;
    R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;
    R(1, 1) = S(1) - (txx + tzz);
    R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;
    R(2, 1) = tyz + twx;
    R(2, 2) = S(1) - (txx + tyy);
    return R;
  }
  Quaternion slerp(S t, const Quaternion &other) const {
    const S one = S(1) - std::numeric_limits<S>::epsilon();
    const S d = dot(other);
    const S ad = std::abs(d);
    S s0, s1;
    if (ad >= one) {
      s0 = S(1) - t;
      s1 = t;
    } else {
      const S theta = std::acos(ad);
      const S st = std::sin(theta);
      s0 = std::sin((S(1) - t) * theta) / st;
      s1 = std::sin(t * theta) / st;
    }
    if (d < S(0)) s1 = -s1;
    return Quaternion(s0 * w() + s1 * other.w(), s0 * x() + s1 * other.x(), s0 * y() + s1 * other.y(),
                      s0 * z() + s1 * other.z());
  }

 private:
  Matrix<S, 4, 1> c_;
};
template <class S>
Quaternion<S> operator*(const AngleAxis<S> &a, const AngleAxis<S> &b) {
  return Quaternion<S>(a) * Quaternion<S>(b);
}
template <class S>
Quaternion<S> operator*(const AngleAxis<S> &a, const Quaternion<S> &b) {
  return Quaternion<S>(a) * b;
}
template <class S>
std::ostream &operator<<(std::ostream &os, const Quaternion<S> &q) {
  return os << q.x() << "i + " << q.y() << "j + " << q.z() << "k + " << q.w();
}
using Quaterniond = Quaternion<double>;
using Quaternionf = Quaternion<float>;
using AngleAxisd = AngleAxis<double>;

template <class S, int N>
class Translation {
 public:
  Translation() {}
  explicit Translation(const Matrix<S, N, 1> &v) : v_(v) {}
  template <int NN = N, typename std::enable_if<NN == 3, int>::type = 0>
  Translation(S x, S y, S z) : v_(x, y, z) {}
  const Matrix<S, N, 1> &translation() const { return v_; }
  Matrix<S, N, 1> &translation() { return v_; }
  S x() const { return v_.x(); }
  S y() const { return v_.y(); }
  S z() const { return v_.z(); }
  static Translation Identity() { return Translation(Matrix<S, N, 1>::Zero()); }

 private:
  Matrix<S, N, 1> v_;
};
using Translation3d = Translation<double, 3>;

// only has to compile (filters.hpp specialises MovingAverage<Quaterniond> with it; never instantiated on this path)
template <class M>
class JacobiSVD {
 public:
  JacobiSVD(const M &, unsigned int = 0) {}
  M matrixU() const { throw std::logic_error("mini_eigen: JacobiSVD not available"); }
};

}  // namespace Eigen
