// Leg kinematics as __device__ functions.
//
// Go1: closed form of the reference's FROST-generated expressions
//   src/go1_example/src/Expressions/{FR,FL,RR,RL}_foot.cc:15-202, J_{FR,FL,RR,RL}.cc:15-~610
// evaluated the way the adapter calls them (go1Sub.cpp:68-75: six floating-base coordinates = 0,
// fixed foot joint = 0, Jacobian columns 6+4*leg..8+4*leg, go1Sub.cpp:89-121).  With those zeros
// the generated straight-line code collapses to a hip-roll / thigh-pitch / calf-pitch chain; the
// equality is pinned to 1e-15 against the reference sources themselves (tests/golden/
// go1_kin_golden.npz, generated from oracle/_ref).  3 sincos per leg instead of 14 trig calls.
//
// Cassie / PogoX: builder-defined serial chains (the reference ships no such model); identical to
// oracle/kin.c and synth.ROBOTS.
#pragma once
#include "smallmat.cuh"

namespace dekf {

template <typename T>
DEKF_HD void sincos_t(T x, T *s, T *c);
template <>
DEKF_HD void sincos_t<double>(double x, double *s, double *c) {
#if defined(__CUDA_ARCH__)
  sincos(x, s, c);
#else
  *s = sin(x);
  *c = cos(x);
#endif
}
template <>
DEKF_HD void sincos_t<float>(float x, float *s, float *c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, s, c);
#else
  *s = sinf(x);
  *c = cosf(x);
#endif
}

enum RobotId { ROBOT_GO1 = 0, ROBOT_CASSIE = 1, ROBOT_POGOX = 2 };

template <typename T>
struct Go1Model {
  static constexpr int NLEG = 4;
  static constexpr int NJ = 3;
  // foot position p (kinematics base frame, p_ib not added) and 3x3 joint Jacobian (row-major)
  DEKF_HD static void leg_fk(int leg, const T *q, V3<T> &p, T *J) {
    const T HX = T(0.1881), HY = T(0.04675), TY = T(0.08), L = T(0.213);
    const T sx = (leg < 2) ? T(1) : T(-1);
    const T sy = ((leg & 1) == 0) ? T(-1) : T(1);
    T s1, c1, s2, c2, s23, c23;
    sincos_t<T>(q[0], &s1, &c1);
    sincos_t<T>(q[1], &s2, &c2);
    sincos_t<T>(q[1] + q[2], &s23, &c23);
    const T xl = -L * (s2 + s23);
    const T zl = -L * (c2 + c23);
    const T yl = sy * TY;
    p[0] = sx * HX + xl;
    p[1] = sy * HY + c1 * yl - s1 * zl;
    p[2] = s1 * yl + c1 * zl;
    const T dz2 = L * (s2 + s23), dz3 = L * s23;
    J[0] = T(0);
    J[1] = -L * (c2 + c23);
    J[2] = -L * c23;
    J[3] = -s1 * yl - c1 * zl;
    J[4] = -s1 * dz2;
    J[5] = -s1 * dz3;
    J[6] = c1 * yl - s1 * zl;
    J[7] = c1 * dz2;
    J[8] = c1 * dz3;
  }
};

// Generic serial chain (revolute / prismatic joints about constant axes).
template <typename T, int NJ_>
struct ChainEval {
  DEKF_HD static M3<T> rot_axis(const T *a, T th) {
    T s, c;
    sincos_t<T>(th, &s, &c);
    const T v = T(1) - c;
    M3<T> R;
    R(0, 0) = c + a[0] * a[0] * v;
    R(0, 1) = a[0] * a[1] * v - a[2] * s;
    R(0, 2) = a[0] * a[2] * v + a[1] * s;
    R(1, 0) = a[1] * a[0] * v + a[2] * s;
    R(1, 1) = c + a[1] * a[1] * v;
    R(1, 2) = a[1] * a[2] * v - a[0] * s;
    R(2, 0) = a[2] * a[0] * v - a[1] * s;
    R(2, 1) = a[2] * a[1] * v + a[0] * s;
    R(2, 2) = c + a[2] * a[2] * v;
    return R;
  }
  // type[j]: 0 revolute, 1 prismatic; axis/off: [NJ][3]; tool[3]
  DEKF_HD static void fk(const int *type, const T (*axis)[3], const T (*off)[3], const T *tool, const T *q,
                         V3<T> &p, T *J) {
    M3<T> Rw;
#pragma unroll
    for (int i = 0; i < 9; ++i) Rw.a[i] = (i % 4 == 0) ? T(1) : T(0);
    V3<T> o = v3<T>(T(0), T(0), T(0));
    V3<T> origin[NJ_], axw[NJ_];
#pragma unroll
    for (int j = 0; j < NJ_; ++j) {
      o = add(o, mul(Rw, v3<T>(off[j][0], off[j][1], off[j][2])));
      axw[j] = mul(Rw, v3<T>(axis[j][0], axis[j][1], axis[j][2]));
      origin[j] = o;
      if (type[j] == 0)
        Rw = mul(Rw, rot_axis(axis[j], q[j]));
      else
        o = add(o, scale(q[j], axw[j]));
    }
    p = add(o, mul(Rw, v3<T>(tool[0], tool[1], tool[2])));
#pragma unroll
    for (int j = 0; j < NJ_; ++j) {
      V3<T> col = (type[j] == 0) ? cross(axw[j], sub(p, origin[j])) : axw[j];
      J[0 * NJ_ + j] = col[0];
      J[1 * NJ_ + j] = col[1];
      J[2 * NJ_ + j] = col[2];
    }
  }
};

template <typename T>
struct CassieModel {
  static constexpr int NLEG = 2;
  static constexpr int NJ = 5;
  DEKF_HD static void leg_fk(int leg, const T *q, V3<T> &p, T *J) {
    const T sy = (leg == 0) ? T(1) : T(-1);
    const int type[5] = {0, 0, 0, 0, 0};
    const T axis[5][3] = {{T(1), T(0), T(0)}, {T(0), T(0), T(1)}, {T(0), T(1), T(0)}, {T(0), T(1), T(0)}, {T(0), T(1), T(0)}};
    const T off[5][3] = {{T(0.021), sy * T(0.135), T(0)}, {T(0), T(0), T(-0.07)}, {T(0), T(0), T(-0.09)},
                         {T(0.12), T(0), T(-0.4896)}, {T(0.06), T(0), T(-0.5)}};
    const T tool[3] = {T(0.02), T(0), T(-0.05)};
    ChainEval<T, 5>::fk(type, axis, off, tool, q, p, J);
  }
};

template <typename T>
struct PogoXModel {
  static constexpr int NLEG = 1;
  static constexpr int NJ = 3;
  DEKF_HD static void leg_fk(int leg, const T *q, V3<T> &p, T *J) {
    (void)leg;
    const int type[3] = {0, 0, 1};
    const T axis[3][3] = {{T(1), T(0), T(0)}, {T(0), T(1), T(0)}, {T(0), T(0), T(-1)}};
    const T off[3][3] = {{T(0), T(0), T(-0.05)}, {T(0), T(0), T(0)}, {T(0), T(0), T(-0.25)}};
    const T tool[3] = {T(0), T(0), T(-0.05)};
    ChainEval<T, 3>::fk(type, axis, off, tool, q, p, J);
  }
};

}  // namespace dekf
