// Fixed-size 3x3 / 3-vector helpers for the per-instance estimator math.  Everything is fully
// unrolled straight-line code so that one estimator instance lives in the registers of one thread.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define DEKF_HD __host__ __device__ __forceinline__
#else
#define DEKF_HD inline
#endif

namespace dekf {

template <typename T>
struct V3 {
  T v[3];
  DEKF_HD T &operator[](int i) { return v[i]; }
  DEKF_HD const T &operator[](int i) const { return v[i]; }
};

// general 3x3, row-major
template <typename T>
struct M3 {
  T a[9];
  DEKF_HD T &operator()(int r, int c) { return a[r * 3 + c]; }
  DEKF_HD const T &operator()(int r, int c) const { return a[r * 3 + c]; }
};

// symmetric 3x3: (00,01,02,11,12,22)
template <typename T>
struct S3 {
  T a[6];
  DEKF_HD static constexpr int idx(int r, int c) {
    return r <= c ? (r == 0 ? c : (r == 1 ? 2 + c : 5)) : (c == 0 ? r : (c == 1 ? 2 + r : 5));
  }
  DEKF_HD T operator()(int r, int c) const { return a[idx(r, c)]; }
};

template <typename T>
DEKF_HD V3<T> v3(T x, T y, T z) {
  V3<T> r;
  r[0] = x;
  r[1] = y;
  r[2] = z;
  return r;
}
template <typename T>
DEKF_HD V3<T> add(const V3<T> &a, const V3<T> &b) { return v3<T>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template <typename T>
DEKF_HD V3<T> sub(const V3<T> &a, const V3<T> &b) { return v3<T>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template <typename T>
DEKF_HD V3<T> scale(T s, const V3<T> &a) { return v3<T>(s * a[0], s * a[1], s * a[2]); }
template <typename T>
DEKF_HD V3<T> cross(const V3<T> &a, const V3<T> &b) {
  return v3<T>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

template <typename T>
DEKF_HD M3<T> to_m3(const S3<T> &s) {
  M3<T> m;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) m(r, c) = s(r, c);
  return m;
}
// symmetric part taken from the upper triangle of a general matrix
template <typename T>
DEKF_HD S3<T> upper(const M3<T> &m) {
  S3<T> s;
  s.a[0] = m(0, 0);
  s.a[1] = m(0, 1);
  s.a[2] = m(0, 2);
  s.a[3] = m(1, 1);
  s.a[4] = m(1, 2);
  s.a[5] = m(2, 2);
  return s;
}
template <typename T>
DEKF_HD M3<T> transpose(const M3<T> &m) {
  M3<T> t;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) t(r, c) = m(c, r);
  return t;
}

// C = A * B
template <typename T>
DEKF_HD M3<T> mul(const M3<T> &A, const M3<T> &B) {
  M3<T> C;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C(r, c) = A(r, 0) * B(0, c) + A(r, 1) * B(1, c) + A(r, 2) * B(2, c);
  return C;
}
// C = A * B^T
template <typename T>
DEKF_HD M3<T> mul_nt(const M3<T> &A, const M3<T> &B) {
  M3<T> C;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C(r, c) = A(r, 0) * B(c, 0) + A(r, 1) * B(c, 1) + A(r, 2) * B(c, 2);
  return C;
}
// C = A^T * B
template <typename T>
DEKF_HD M3<T> mul_tn(const M3<T> &A, const M3<T> &B) {
  M3<T> C;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C(r, c) = A(0, r) * B(0, c) + A(1, r) * B(1, c) + A(2, r) * B(2, c);
  return C;
}
template <typename T>
DEKF_HD M3<T> mul(const M3<T> &A, const S3<T> &B) { return mul(A, to_m3(B)); }
template <typename T>
DEKF_HD M3<T> mul(const S3<T> &A, const M3<T> &B) { return mul(to_m3(A), B); }
template <typename T>
DEKF_HD V3<T> mul(const M3<T> &A, const V3<T> &x) {
  return v3<T>(A(0, 0) * x[0] + A(0, 1) * x[1] + A(0, 2) * x[2], A(1, 0) * x[0] + A(1, 1) * x[1] + A(1, 2) * x[2],
               A(2, 0) * x[0] + A(2, 1) * x[1] + A(2, 2) * x[2]);
}
template <typename T>
DEKF_HD V3<T> mul_t(const M3<T> &A, const V3<T> &x) {  // A^T x
  return v3<T>(A(0, 0) * x[0] + A(1, 0) * x[1] + A(2, 0) * x[2], A(0, 1) * x[0] + A(1, 1) * x[1] + A(2, 1) * x[2],
               A(0, 2) * x[0] + A(1, 2) * x[1] + A(2, 2) * x[2]);
}
template <typename T>
DEKF_HD V3<T> mul(const S3<T> &A, const V3<T> &x) {
  return v3<T>(A.a[0] * x[0] + A.a[1] * x[1] + A.a[2] * x[2], A.a[1] * x[0] + A.a[3] * x[1] + A.a[4] * x[2],
               A.a[2] * x[0] + A.a[4] * x[1] + A.a[5] * x[2]);
}

// upper triangle of A * B^T when the product is known to be symmetric (or only its symmetric
// part is wanted)
template <typename T>
DEKF_HD S3<T> mul_nt_sym(const M3<T> &A, const M3<T> &B) {
  S3<T> s;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = r; c < 3; ++c) s.a[S3<T>::idx(r, c)] = A(r, 0) * B(c, 0) + A(r, 1) * B(c, 1) + A(r, 2) * B(c, 2);
  return s;
}
// R * diag(d) * R^T
template <typename T>
DEKF_HD S3<T> rdrt(const M3<T> &R, const V3<T> &d) {
  S3<T> s;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = r; c < 3; ++c)
      s.a[S3<T>::idx(r, c)] = d[0] * R(r, 0) * R(c, 0) + d[1] * R(r, 1) * R(c, 1) + d[2] * R(r, 2) * R(c, 2);
  return s;
}
// R * S * R^T with S symmetric
template <typename T>
DEKF_HD S3<T> rsrt(const M3<T> &R, const S3<T> &S) {
  M3<T> RS = mul(R, S);
  return mul_nt_sym(RS, R);
}
template <typename T>
DEKF_HD S3<T> add(const S3<T> &a, const S3<T> &b) {
  S3<T> s;
#pragma unroll
  for (int i = 0; i < 6; ++i) s.a[i] = a.a[i] + b.a[i];
  return s;
}
template <typename T>
DEKF_HD M3<T> skew(const V3<T> &v) {  // EigenUtils.hpp:91-97
  M3<T> m;
  m(0, 0) = T(0);
  m(0, 1) = -v[2];
  m(0, 2) = v[1];
  m(1, 0) = v[2];
  m(1, 1) = T(0);
  m(1, 2) = -v[0];
  m(2, 0) = -v[1];
  m(2, 1) = v[0];
  m(2, 2) = T(0);
  return m;
}

// inverse of a symmetric 3x3 through the adjugate
template <typename T>
DEKF_HD S3<T> inverse(const S3<T> &s) {
  const T a = s.a[0], b = s.a[1], c = s.a[2], d = s.a[3], e = s.a[4], f = s.a[5];
  const T c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const T det = a * c00 + b * c01 + c * c02;
  const T id = T(1) / det;
  S3<T> r;
  r.a[0] = c00 * id;
  r.a[1] = c01 * id;
  r.a[2] = c02 * id;
  r.a[3] = (a * f - c * c) * id;
  r.a[4] = (b * c - a * e) * id;
  r.a[5] = (a * d - b * b) * id;
  return r;
}
// inverse of a general 3x3 through the adjugate
template <typename T>
DEKF_HD M3<T> inverse(const M3<T> &m) {
  const T c00 = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1);
  const T c01 = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2);
  const T c02 = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
  const T det = m(0, 0) * c00 + m(0, 1) * c01 + m(0, 2) * c02;
  const T id = T(1) / det;
  M3<T> r;
  r(0, 0) = c00 * id;
  r(1, 0) = c01 * id;
  r(2, 0) = c02 * id;
  r(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) * id;
  r(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) * id;
  r(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) * id;
  r(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) * id;
  r(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) * id;
  r(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) * id;
  return r;
}


// ---- accumulate-in-place products: every term is one fused multiply-add on the destination ----------------
// P -= A * B
template <typename T>
DEKF_HD void sub_mul(M3<T> &P, const M3<T> &A, const M3<T> &B) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T v = P(r, c);
      v -= A(r, 0) * B(0, c);
      v -= A(r, 1) * B(1, c);
      v -= A(r, 2) * B(2, c);
      P(r, c) = v;
    }
}
// upper(P) -= upper(A * B)      (A * B symmetric)
template <typename T>
DEKF_HD void sub_mul_sym(S3<T> &P, const M3<T> &A, const M3<T> &B) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = r; c < 3; ++c) {
      T v = P.a[S3<T>::idx(r, c)];
      v -= A(r, 0) * B(0, c);
      v -= A(r, 1) * B(1, c);
      v -= A(r, 2) * B(2, c);
      P.a[S3<T>::idx(r, c)] = v;
    }
}
// upper(P) -= upper(A * B^T)
template <typename T>
DEKF_HD void sub_mul_nt_sym(S3<T> &P, const M3<T> &A, const M3<T> &B) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = r; c < 3; ++c) {
      T v = P.a[S3<T>::idx(r, c)];
      v -= A(r, 0) * B(c, 0);
      v -= A(r, 1) * B(c, 1);
      v -= A(r, 2) * B(c, 2);
      P.a[S3<T>::idx(r, c)] = v;
    }
}
// P -= A * B^T
template <typename T>
DEKF_HD void sub_mul_nt(M3<T> &P, const M3<T> &A, const M3<T> &B) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T v = P(r, c);
      v -= A(r, 0) * B(c, 0);
      v -= A(r, 1) * B(c, 1);
      v -= A(r, 2) * B(c, 2);
      P(r, c) = v;
    }
}
// upper(P) -= upper(A^T * B)
template <typename T>
DEKF_HD void sub_mul_tn_sym(S3<T> &P, const M3<T> &A, const M3<T> &B) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = r; c < 3; ++c) {
      T v = P.a[S3<T>::idx(r, c)];
      v -= A(0, r) * B(0, c);
      v -= A(1, r) * B(1, c);
      v -= A(2, r) * B(2, c);
      P.a[S3<T>::idx(r, c)] = v;
    }
}
// x += A * t,  x += A^T * t
template <typename T>
DEKF_HD void add_mul(V3<T> &x, const M3<T> &A, const V3<T> &t) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    T v = x[r];
    v += A(r, 0) * t[0];
    v += A(r, 1) * t[1];
    v += A(r, 2) * t[2];
    x[r] = v;
  }
}
template <typename T>
DEKF_HD void add_mul_t(V3<T> &x, const M3<T> &A, const V3<T> &t) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    T v = x[r];
    v += A(0, r) * t[0];
    v += A(1, r) * t[1];
    v += A(2, r) * t[2];
    x[r] = v;
  }
}
// R diag(d) R^T = d0 I + (d1-d0) r1 r1^T + (d2-d0) r2 r2^T with r1, r2 the 2nd/3rd columns of the rotation R (the
// columns are orthonormal): o1 = r1 r1^T, o2 = r2 r2^T are shared between all the noise blocks of a stage.
template <typename T>
struct RotOuter {
  S3<T> o1, o2;
};
template <typename T>
DEKF_HD RotOuter<T> rot_outer(const M3<T> &R) {
  RotOuter<T> o;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = r; c < 3; ++c) {
      o.o1.a[S3<T>::idx(r, c)] = R(r, 1) * R(c, 1);
      o.o2.a[S3<T>::idx(r, c)] = R(r, 2) * R(c, 2);
    }
  return o;
}
// e = (d0, d1-d0, d2-d0)
template <typename T>
DEKF_HD S3<T> rdrt2(const RotOuter<T> &o, const T e[3]) {
  S3<T> s;
#pragma unroll
  for (int i = 0; i < 6; ++i) s.a[i] = e[1] * o.o1.a[i] + e[2] * o.o2.a[i];
  s.a[0] += e[0];
  s.a[3] += e[0];
  s.a[5] += e[0];
  return s;
}

// unscaled inverses: adj(m) and det(m), inverse = adj / det.  Lets the caller run independent products on the adjugate
// while the reciprocal of the determinant (a long dependent chain in fp64) is in flight.
template <typename T>
DEKF_HD M3<T> adjugate(const M3<T> &m, T &det) {
  M3<T> r;
  r(0, 0) = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1);
  r(1, 0) = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2);
  r(2, 0) = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
  det = m(0, 0) * r(0, 0) + m(0, 1) * r(1, 0) + m(0, 2) * r(2, 0);
  r(0, 1) = m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2);
  r(1, 1) = m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0);
  r(2, 1) = m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1);
  r(0, 2) = m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1);
  r(1, 2) = m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2);
  r(2, 2) = m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
  return r;
}
template <typename T>
DEKF_HD S3<T> adjugate(const S3<T> &s, T &det) {
  const T a = s.a[0], b = s.a[1], c = s.a[2], d = s.a[3], e = s.a[4], f = s.a[5];
  S3<T> r;
  r.a[0] = d * f - e * e;
  r.a[1] = c * e - b * f;
  r.a[2] = b * e - c * d;
  det = a * r.a[0] + b * r.a[1] + c * r.a[2];
  r.a[3] = a * f - c * c;
  r.a[4] = b * c - a * e;
  r.a[5] = a * d - b * b;
  return r;
}
template <typename T>
DEKF_HD M3<T> scale(T s, const M3<T> &m) {
  M3<T> r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.a[i] = s * m.a[i];
  return r;
}

// Eigen Quaterniond(q).normalized().toRotationMatrix(), q = [w,x,y,z]
// (orien_ekf.cpp:296-305, DecentralEst.cpp:867)
template <typename T>
DEKF_HD M3<T> quat_to_rot(T qw, T qx, T qy, T qz) {
  const T nrm = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  const T w = qw / nrm, x = qx / nrm, y = qy / nrm, z = qz / nrm;
  const T tx = T(2) * x, ty = T(2) * y, tz = T(2) * z;
  const T twx = tx * w, twy = ty * w, twz = tz * w;
  const T txx = tx * x, txy = ty * x, txz = tz * x;
  const T tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M3<T> R;
  R(0, 0) = T(1) - (tyy + tzz);
  R(0, 1) = txy - twz;
  R(0, 2) = txz + twy;
  R(1, 0) = txy + twz;
  R(1, 1) = T(1) - (txx + tzz);
  R(1, 2) = tyz - twx;
  R(2, 0) = txz - twy;
  R(2, 1) = tyz + twx;
  R(2, 2) = T(1) - (txx + tyy);
  return R;
}

}  // namespace dekf
