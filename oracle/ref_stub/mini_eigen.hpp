// Stand-in for the subset of Eigen3 (Dense / Sparse / Geometry) that the reference's estimator sources use.
//
// TEST INFRASTRUCTURE ONLY.  Eigen3 is not installed in this image, so the reference's own, UNMODIFIED sources
// (src/orien_est/src/orien_ekf.cpp, src/decentral_legged_est/src/{DecentralEst,MheSrb,EstSub}.cpp,
// src/decentral_legged_est/src/Spline/Bezier_simple.cpp, src/go1_example/src/go1Sub.cpp and the FROST expressions)
// are compiled against this header into oracle/_ref/ (oracle/Makefile target `ref`) to pin the C restatement in
// oracle/*.c against the reference's real control flow, indexing and formulas.  Never shipped, never on the product
// path.  What this is NOT: it is not Eigen -- every operation is evaluated eagerly on one dynamically sized,
// column-major, dense double matrix; "sparse" matrices are dense underneath (InnerIterator visits the non-zero
// entries of a column).  Results agree with real Eigen up to floating-point rounding of the linear-algebra kernels
// (inverse by partial-pivot LU, LLT by plain Cholesky), which is what the 1e-9 / 1e-6 parity tolerances absorb.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace Eigen {

typedef std::ptrdiff_t Index;
enum { Dynamic = -1 };
enum { Unaligned = 0, Aligned = 16 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

class Mat;
class View;
template <typename S, int O, typename I> class SparseMatrix;

// ------------------------------------------------------------------------------------------------ comma initialiser
// Eigen semantics: items fill the target left to right in block rows (a scalar is a 1x1 block).
class CommaInit {
 public:
  CommaInit(Mat *m, int i0, int j0, int r, int c, bool diag) : m_(m), i0_(i0), j0_(j0), r_(r), c_(c), diag_(diag) {}
  CommaInit &operator,(double v) { put_block(&v, 1, 1); return *this; }
  CommaInit &operator,(const Mat &b);
  CommaInit &operator,(const View &b);
  void put_block(const double *colmajor, int br, int bc);

 private:
  Mat *m_;
  int i0_, j0_, r_, c_;
  bool diag_;
  int row_ = 0, col_ = 0, cur_rows_ = 0;
};

// ------------------------------------------------------------------------------------------------ views (block / segment / diagonal)
class View {
 public:
  View(Mat *m, int i0, int j0, int r, int c, bool diag = false) : m_(m), i0_(i0), j0_(j0), r_(r), c_(c), diag_(diag) {}
  View(const View &) = default;
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  inline double &at(int i, int j) const;
  double &operator()(int i, int j) const { return at(i, j); }
  double &operator()(int i) const { return c_ == 1 ? at(i, 0) : at(0, i); }
  View &operator=(const Mat &o);
  View &operator=(const View &o);
  CommaInit operator<<(double v) { CommaInit ci(m_, i0_, j0_, r_, c_, diag_); ci, v; return ci; }
  CommaInit operator<<(const Mat &b) { CommaInit ci(m_, i0_, j0_, r_, c_, diag_); ci, b; return ci; }
  CommaInit operator<<(const View &b) { CommaInit ci(m_, i0_, j0_, r_, c_, diag_); ci, b; return ci; }
  void setZero() { for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) at(i, j) = 0.0; }
  inline Mat eval() const;
  inline Mat transpose() const;
  inline Mat inverse() const;
  inline Mat operator-() const;
  inline double norm() const;
  inline Mat cross(const Mat &o) const;
  View &operator+=(const Mat &o);
  View &operator-=(const Mat &o);

 private:
  Mat *m_;
  int i0_, j0_, r_, c_;
  bool diag_;
};

// ------------------------------------------------------------------------------------------------ the one dense matrix
class Mat {
 public:
  typedef double Scalar;
  Mat() : r_(0), c_(0) {}
  Mat(int r, int c) : r_(r), c_(c), v_((size_t)r * (size_t)c, 0.0) {}
  Mat(const View &b) : r_(b.rows()), c_(b.cols()), v_((size_t)b.rows() * (size_t)b.cols()) {
    for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) v_[(size_t)j * r_ + i] = b.at(i, j);
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  double &operator()(int i, int j) { return v_[(size_t)j * r_ + i]; }
  const double &operator()(int i, int j) const { return v_[(size_t)j * r_ + i]; }
  double &operator()(int i) { return v_[(size_t)i]; }
  const double &operator()(int i) const { return v_[(size_t)i]; }
  double &operator[](int i) { return v_[(size_t)i]; }
  const double &operator[](int i) const { return v_[(size_t)i]; }
  double &x() { return v_[0]; }
  double &y() { return v_[1]; }
  double &z() { return v_[2]; }
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double z() const { return v_[2]; }

  void setZero() { std::fill(v_.begin(), v_.end(), 0.0); }
  void setOnes() { std::fill(v_.begin(), v_.end(), 1.0); }
  void setConstant(double a) { std::fill(v_.begin(), v_.end(), a); }
  void setIdentity() { setZero(); for (int i = 0; i < std::min(r_, c_); ++i) (*this)(i, i) = 1.0; }
  void setIdentity(int r, int c) { resize(r, c); setIdentity(); }
  // Eigen: resize() does not keep the values (left uninitialised there; zero here)
  void resize(int r, int c) { r_ = r; c_ = c; v_.assign((size_t)r * (size_t)c, 0.0); }
  void resize(int n) { if (c_ == 1 || (r_ == 0 && c_ == 0)) resize(n, 1); else if (r_ == 1) resize(1, n); else resize(n, 1); }
  // keeps the top-left overlap, new entries zero (Eigen leaves new dense entries uninitialised; the reference zeroes
  // them itself where it matters: EigenUtils::vectorResize)
  void conservativeResize(int r, int c) {
    Mat t(r, c);
    for (int j = 0; j < std::min(c, c_); ++j) for (int i = 0; i < std::min(r, r_); ++i) t(i, j) = (*this)(i, j);
    r_ = r; c_ = c; v_.swap(t.v_);
  }
  void conservativeResize(int n) { if (r_ == 1 && c_ != 1) conservativeResize(1, n); else conservativeResize(n, 1); }

  View block(int i, int j, int r, int c) const { return View(const_cast<Mat *>(this), i, j, r, c); }
  template <int R, int C> View block(int i, int j) const { return View(const_cast<Mat *>(this), i, j, R, C); }
  View segment(int i, int n) const { return c_ == 1 ? View(const_cast<Mat *>(this), i, 0, n, 1) : View(const_cast<Mat *>(this), 0, i, 1, n); }
  template <int N> View segment(int i) const { return segment(i, N); }
  View head(int n) const { return segment(0, n); }
  template <int N> View head() const { return segment(0, N); }
  View tail(int n) const { return segment(size() - n, n); }
  template <int N> View tail() const { return segment(size() - N, N); }
  View col(int j) const { return View(const_cast<Mat *>(this), 0, j, r_, 1); }
  View row(int i) const { return View(const_cast<Mat *>(this), i, 0, 1, c_); }
  View diagonal() const { return View(const_cast<Mat *>(this), 0, 0, std::min(r_, c_), 1, true); }
  View topLeftCorner(int r, int c) const { return block(0, 0, r, c); }

  CommaInit operator<<(double v) { CommaInit ci(this, 0, 0, r_, c_, false); ci, v; return ci; }
  CommaInit operator<<(const Mat &b) { CommaInit ci(this, 0, 0, r_, c_, false); ci, b; return ci; }
  CommaInit operator<<(const View &b) { CommaInit ci(this, 0, 0, r_, c_, false); ci, b; return ci; }

  Mat transpose() const {
    Mat t(c_, r_);
    for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
    return t;
  }
  Mat eval() const { return *this; }
  template <typename T> Mat cast() const { return *this; }
  double squaredNorm() const { double s = 0; for (double a : v_) s += a * a; return s; }
  double norm() const { return std::sqrt(squaredNorm()); }
  double sum() const { double s = 0; for (double a : v_) s += a; return s; }
  double trace() const { double s = 0; for (int i = 0; i < std::min(r_, c_); ++i) s += (*this)(i, i); return s; }
  double maxCoeff() const { return *std::max_element(v_.begin(), v_.end()); }
  double minCoeff() const { return *std::min_element(v_.begin(), v_.end()); }
  Mat cwiseAbs() const { Mat t(*this); for (double &a : t.v_) a = std::fabs(a); return t; }
  double dot(const Mat &o) const { double s = 0; for (int i = 0; i < size(); ++i) s += v_[i] * o.v_[i]; return s; }
  Mat normalized() const { Mat t(*this); double n = norm(); for (double &a : t.v_) a /= n; return t; }
  void normalize() { double n = norm(); for (double &a : v_) a /= n; }
  Mat cross(const Mat &o) const {
    Mat t(3, 1);
    t(0) = v_[1] * o.v_[2] - v_[2] * o.v_[1];
    t(1) = v_[2] * o.v_[0] - v_[0] * o.v_[2];
    t(2) = v_[0] * o.v_[1] - v_[1] * o.v_[0];
    return t;
  }
  // general inverse: Gauss-Jordan on [A | I] with partial pivoting (Eigen: PartialPivLU for dynamic sizes)
  Mat inverse() const {
    int n = r_;
    Mat a(*this), inv(n, n);
    inv.setIdentity();
    for (int k = 0; k < n; ++k) {
      int p = k;
      double best = std::fabs(a(k, k));
      for (int i = k + 1; i < n; ++i) if (std::fabs(a(i, k)) > best) { best = std::fabs(a(i, k)); p = i; }
      if (p != k) for (int j = 0; j < n; ++j) { std::swap(a(k, j), a(p, j)); std::swap(inv(k, j), inv(p, j)); }
      double d = a(k, k);
      for (int j = 0; j < n; ++j) { a(k, j) /= d; inv(k, j) /= d; }
      for (int i = 0; i < n; ++i) {
        if (i == k) continue;
        double f = a(i, k);
        if (f == 0.0) continue;
        for (int j = 0; j < n; ++j) { a(i, j) -= f * a(k, j); inv(i, j) -= f * inv(k, j); }
      }
    }
    return inv;
  }
  double determinant() const {
    int n = r_;
    Mat a(*this);
    double det = 1.0;
    for (int k = 0; k < n; ++k) {
      int p = k;
      for (int i = k + 1; i < n; ++i) if (std::fabs(a(i, k)) > std::fabs(a(p, k))) p = i;
      if (a(p, k) == 0.0) return 0.0;
      if (p != k) { for (int j = 0; j < n; ++j) std::swap(a(k, j), a(p, j)); det = -det; }
      det *= a(k, k);
      for (int i = k + 1; i < n; ++i) { double f = a(i, k) / a(k, k); for (int j = k; j < n; ++j) a(i, j) -= f * a(k, j); }
    }
    return det;
  }
  inline SparseMatrix<double, 0, int> sparseView() const;

  Mat &operator+=(const Mat &o) { assert(size() == o.size()); for (size_t i = 0; i < v_.size(); ++i) v_[i] += o.v_[i]; return *this; }
  Mat &operator-=(const Mat &o) { assert(size() == o.size()); for (size_t i = 0; i < v_.size(); ++i) v_[i] -= o.v_[i]; return *this; }
  Mat &operator*=(double s) { for (double &a : v_) a *= s; return *this; }
  Mat &operator/=(double s) { for (double &a : v_) a /= s; return *this; }
  Mat operator-() const { Mat t(*this); for (double &a : t.v_) a = -a; return t; }

 protected:
  int r_, c_;
  std::vector<double> v_;
};

inline double &View::at(int i, int j) const { return diag_ ? (*m_)(i0_ + i, i0_ + i) : (*m_)(i0_ + i, j0_ + j); }
inline Mat View::eval() const { return Mat(*this); }
inline Mat View::transpose() const { return Mat(*this).transpose(); }
inline Mat View::inverse() const { return Mat(*this).inverse(); }
inline Mat View::operator-() const { return -Mat(*this); }
inline double View::norm() const { return Mat(*this).norm(); }
inline Mat View::cross(const Mat &o) const { return Mat(*this).cross(o); }
inline View &View::operator=(const Mat &o) {
  assert(o.size() == size());
  if (o.rows() == r_) { for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) at(i, j) = o(i, j); }
  else { int k = 0; for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) at(i, j) = o(k++); }  // vector of the other orientation
  return *this;
}
inline View &View::operator=(const View &o) { Mat t(o); return (*this = t); }
inline View &View::operator+=(const Mat &o) { for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) at(i, j) += o(i, j); return *this; }
inline View &View::operator-=(const Mat &o) { for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) at(i, j) -= o(i, j); return *this; }

inline void CommaInit::put_block(const double *cm, int br, int bc) {
  if (col_ >= c_) { row_ += cur_rows_; col_ = 0; cur_rows_ = 0; }
  if (cur_rows_ == 0) cur_rows_ = br;
  assert(row_ + br <= r_ && col_ + bc <= c_);
  View v(m_, i0_, j0_, r_, c_, diag_);
  for (int j = 0; j < bc; ++j) for (int i = 0; i < br; ++i) v.at(row_ + i, col_ + j) = cm[(size_t)j * br + i];
  col_ += bc;
}
inline CommaInit &CommaInit::operator,(const Mat &b) {
  // a vector of the other orientation filling a vector target is accepted (Eigen transposes vectors on assignment)
  if ((r_ == 1 && b.cols() == 1 && b.rows() > 1)) { Mat t = b.transpose(); put_block(t.data(), t.rows(), t.cols()); }
  else if ((c_ == 1 && b.rows() == 1 && b.cols() > 1)) { Mat t = b.transpose(); put_block(t.data(), t.rows(), t.cols()); }
  else put_block(b.data(), b.rows(), b.cols());
  return *this;
}
inline CommaInit &CommaInit::operator,(const View &b) { Mat t(b); return (*this, t); }

// ------------------------------------------------------------------------------------------------ arithmetic (eager)
inline Mat operator+(const Mat &a, const Mat &b) { assert(a.rows() == b.rows() && a.cols() == b.cols()); Mat t(a); t += b; return t; }
inline Mat operator-(const Mat &a, const Mat &b) { assert(a.rows() == b.rows() && a.cols() == b.cols()); Mat t(a); t -= b; return t; }
inline Mat operator*(const Mat &a, double s) { Mat t(a); t *= s; return t; }
inline Mat operator*(double s, const Mat &a) { Mat t(a); t *= s; return t; }
inline Mat operator/(const Mat &a, double s) { Mat t(a); t /= s; return t; }
inline Mat operator*(const Mat &a, const Mat &b) {
  assert(a.cols() == b.rows());
  Mat t(a.rows(), b.cols());
  for (int j = 0; j < b.cols(); ++j)
    for (int k = 0; k < a.cols(); ++k) {
      double bkj = b(k, j);
      if (bkj == 0.0) continue;
      for (int i = 0; i < a.rows(); ++i) t(i, j) += a(i, k) * bkj;
    }
  return t;
}
inline std::ostream &operator<<(std::ostream &os, const Mat &m) {
  for (int i = 0; i < m.rows(); ++i) {
    for (int j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j);
    if (i + 1 < m.rows()) os << "\n";
  }
  return os;
}

// ------------------------------------------------------------------------------------------------ named types
template <int R, int C>
class FMat : public Mat {
 public:
  FMat() : Mat(R, C) {}
  FMat(const Mat &m) : Mat(m) {}
  FMat(const View &v) : Mat(v) {}
  FMat(double a, double b) : Mat(R, C) { v_[0] = a; v_[1] = b; }
  FMat(double a, double b, double c) : Mat(R, C) { v_[0] = a; v_[1] = b; v_[2] = c; }
  FMat(double a, double b, double c, double d) : Mat(R, C) { v_[0] = a; v_[1] = b; v_[2] = c; v_[3] = d; }
  FMat &operator=(const Mat &m) { Mat::operator=(m); return *this; }
  FMat &operator=(const View &v) { Mat::operator=(Mat(v)); return *this; }
  static FMat Zero() { return FMat(); }
  static FMat Zero(int, int) { return FMat(); }
  static FMat Ones() { FMat m; m.setOnes(); return m; }
  static FMat Identity() { FMat m; m.setIdentity(); return m; }
  static FMat Identity(int, int) { return Identity(); }
  static FMat Constant(double a) { FMat m; m.setConstant(a); return m; }
};
typedef FMat<2, 1> Vector2d;
typedef FMat<3, 1> Vector3d;
typedef FMat<4, 1> Vector4d;
typedef FMat<2, 2> Matrix2d;
typedef FMat<3, 3> Matrix3d;
typedef FMat<4, 4> Matrix4d;

class MatrixXd : public Mat {
 public:
  MatrixXd() {}
  MatrixXd(int r, int c) : Mat(r, c) {}
  MatrixXd(const Mat &m) : Mat(m) {}
  MatrixXd(const View &v) : Mat(v) {}
  MatrixXd &operator=(const Mat &m) { Mat::operator=(m); return *this; }
  MatrixXd &operator=(const View &v) { Mat::operator=(Mat(v)); return *this; }
  static MatrixXd Zero(int r, int c) { return MatrixXd(r, c); }
  static MatrixXd Ones(int r, int c) { MatrixXd m(r, c); m.setOnes(); return m; }
  static MatrixXd Identity(int r, int c) { MatrixXd m(r, c); m.setIdentity(); return m; }
  static MatrixXd Constant(int r, int c, double a) { MatrixXd m(r, c); m.setConstant(a); return m; }
};

class VectorXd : public Mat {
 public:
  VectorXd() : Mat(0, 1) {}
  explicit VectorXd(int n) : Mat(n, 1) {}
  VectorXd(const Mat &m) : Mat(m) {}
  VectorXd(const View &v) : Mat(v) {}
  VectorXd &operator=(const Mat &m) { Mat::operator=(m); return *this; }
  VectorXd &operator=(const View &v) { Mat::operator=(Mat(v)); return *this; }
  static VectorXd Zero(int n) { return VectorXd(n); }
  static VectorXd Ones(int n) { VectorXd m(n); m.setOnes(); return m; }
  static VectorXd Constant(int n, double a) { VectorXd m(n); m.setConstant(a); return m; }
};

// float / int vectors (only the reference's data logger touches them)
template <typename T>
class VecT {
 public:
  typedef T Scalar;
  VecT() {}
  explicit VecT(int n) : v_((size_t)n, T(0)) {}
  int size() const { return (int)v_.size(); }
  int rows() const { return (int)v_.size(); }
  T *data() { return v_.data(); }
  const T *data() const { return v_.data(); }
  T &operator()(int i) { return v_[(size_t)i]; }
  const T &operator()(int i) const { return v_[(size_t)i]; }
  VecT &operator<<(const VecT &o) { v_ = o.v_; return *this; }
  template <typename U> VecT<U> cast() const { VecT<U> t(size()); for (int i = 0; i < size(); ++i) t(i) = (U)v_[(size_t)i]; return t; }
 protected:
  std::vector<T> v_;
};
typedef VecT<float> VectorXf;
typedef VecT<int> VectorXi;

template <typename V, int Options = 0>
class Map : public V {
 public:
  Map(typename V::Scalar *p, Index n) : V((int)n) { for (Index i = 0; i < n; ++i) this->data()[i] = p[i]; }
  Map(const typename V::Scalar *p, Index n) : V((int)n) { for (Index i = 0; i < n; ++i) this->data()[i] = p[i]; }
};

// ------------------------------------------------------------------------------------------------ "sparse"
template <typename S = double, int Options = 0, typename StorageIndex = int>
class SparseMatrix : public Mat {
 public:
  SparseMatrix() {}
  SparseMatrix(int r, int c) : Mat(r, c) {}
  SparseMatrix(const Mat &m) : Mat(m) {}
  SparseMatrix(const View &v) : Mat(v) {}
  SparseMatrix &operator=(const Mat &m) { Mat::operator=(m); return *this; }
  SparseMatrix &operator=(const View &v) { Mat::operator=(Mat(v)); return *this; }
  // Eigen: insert() requires the entry not to exist yet; the reference respects that (setZero()/setIdentity() before)
  double &insert(int i, int j) { return (*this)(i, j); }
  double &coeffRef(int i, int j) { return (*this)(i, j); }
  double coeff(int i, int j) const { return (*this)(i, j); }
  int outerSize() const { return c_; }
  int innerSize() const { return r_; }
  int nonZeros() const { int k = 0; for (double a : v_) k += (a != 0.0); return k; }
  void makeCompressed() {}
  void reserve(int) {}
  class InnerIterator {
   public:
    InnerIterator(const SparseMatrix &m, int outer) : m_(m), j_(outer), i_(-1) { ++(*this); }
    InnerIterator &operator++() { do { ++i_; } while (i_ < m_.rows() && m_(i_, j_) == 0.0); return *this; }
    operator bool() const { return i_ < m_.rows(); }
    int row() const { return i_; }
    int col() const { return j_; }
    int index() const { return i_; }
    double value() const { return m_(i_, j_); }
   private:
    const SparseMatrix &m_;
    int j_, i_;
  };
};
inline SparseMatrix<double, 0, int> Mat::sparseView() const { return SparseMatrix<double, 0, int>(*this); }

// SimplicialLLT stand-in: dense Cholesky A = L L^T, solve by two triangular sweeps
template <typename MatrixType>
class SimplicialLLT {
 public:
  SimplicialLLT() {}
  void compute(const Mat &A) {
    int n = A.rows();
    L_ = Mat(n, n);
    info_ = Success;
    for (int j = 0; j < n; ++j) {
      double d = A(j, j);
      for (int k = 0; k < j; ++k) d -= L_(j, k) * L_(j, k);
      if (!(d > 0.0)) { info_ = NumericalIssue; d = std::fabs(d) > 0 ? std::fabs(d) : 1e-300; }
      double l = std::sqrt(d);
      L_(j, j) = l;
      for (int i = j + 1; i < n; ++i) {
        double s = A(i, j);
        for (int k = 0; k < j; ++k) s -= L_(i, k) * L_(j, k);
        L_(i, j) = s / l;
      }
    }
  }
  Mat solve(const Mat &B) const {
    int n = L_.rows();
    Mat X(B);
    for (int c = 0; c < X.cols(); ++c) {
      for (int i = 0; i < n; ++i) { double s = X(i, c); for (int k = 0; k < i; ++k) s -= L_(i, k) * X(k, c); X(i, c) = s / L_(i, i); }
      for (int i = n - 1; i >= 0; --i) { double s = X(i, c); for (int k = i + 1; k < n; ++k) s -= L_(k, i) * X(k, c); X(i, c) = s / L_(i, i); }
    }
    return X;
  }
  ComputationInfo info() const { return info_; }
 private:
  Mat L_;
  ComputationInfo info_ = Success;
};

// ------------------------------------------------------------------------------------------------ geometry
class Quaterniond {
 public:
  Quaterniond() : w_(0), x_(0), y_(0), z_(0) {}  // Eigen leaves it uninitialised
  Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
  // Quaterniond(const Matrix3d &): Eigen's quaternionbase_assign_impl<Other, 3, 3> (Shepperd's method, same branches and order)
  explicit Quaterniond(const Mat &m) {
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.0) {
      t = std::sqrt(t + 1.0);
      w_ = 0.5 * t;
      t = 0.5 / t;
      x_ = (m(2, 1) - m(1, 2)) * t;
      y_ = (m(0, 2) - m(2, 0)) * t;
      z_ = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
      double v[3];
      v[i] = 0.5 * t;
      t = 0.5 / t;
      w_ = (m(k, j) - m(j, k)) * t;
      v[j] = (m(j, i) + m(i, j)) * t;
      v[k] = (m(k, i) + m(i, k)) * t;
      x_ = v[0]; y_ = v[1]; z_ = v[2];
    }
  }
  double &w() { return w_; }
  double &x() { return x_; }
  double &y() { return y_; }
  double &z() { return z_; }
  double w() const { return w_; }
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  double squaredNorm() const { return w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; }
  double norm() const { return std::sqrt(squaredNorm()); }
  Quaterniond normalized() const { double n = norm(); return Quaterniond(w_ / n, x_ / n, y_ / n, z_ / n); }
  void normalize() { *this = normalized(); }
  Quaterniond conjugate() const { return Quaterniond(w_, -x_, -y_, -z_); }
  Quaterniond inverse() const { double n2 = squaredNorm(); return Quaterniond(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2); }
  Quaterniond operator*(const Quaterniond &b) const {
    return Quaterniond(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                       w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  // same operation order as Eigen's QuaternionBase::toRotationMatrix (tx = 2x, twx = tx*w, ...)
  Matrix3d toRotationMatrix() const {
    Matrix3d R;
    const double tx = 2.0 * x_, ty = 2.0 * y_, tz = 2.0 * z_;
    const double twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const double txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const double tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    R(0, 0) = 1.0 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1.0 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1.0 - (txx + tyy);
    return R;
  }
  Matrix3d matrix() const { return toRotationMatrix(); }
 private:
  double w_, x_, y_, z_;
};

}  // namespace Eigen
