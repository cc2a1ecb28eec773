// Host-side parameter derivation: dekf_config (== robot_params + orien_ekf parameters) -> the
// constant blocks the kernels take by value.  Plain C++ (no CUDA) so the CPU math-debug harness in
// tests/hostsim can share it.
#pragma once
#include <algorithm>
#include <cmath>
#include <cmath>
#include <cstring>

#include "../../include/dekf_b200.h"
#include "estimator_core.cuh"
#include "box_solve.cuh"
#include "footstate.cuh"

namespace dekf {

inline void fill_go1_defaults(dekf_config *c) {
  // go1_example/config/parameters_go1.yaml:5-50 (est_sub) and :68-75 (orien_sub)
  std::memset(c, 0, sizeof(*c));
  c->abi_version = DEKF_ABI_VERSION;
  c->n_instances = 1;
  c->precision = DEKF_FP64;
  c->robot = DEKF_ROBOT_GO1;
  c->ekf_hist_depth = 64;
  for (int i = 0; i < 3; ++i) {
    c->p_init_std[i] = 0.001;
    c->v_init_std[i] = 0.001;
    c->foot_init_std[i] = 0.001;
    c->accel_bias_init_std[i] = 0.0001;
    c->p_process_std[i] = 0.001;
    c->gyro_input_std[i] = 0.03;
    c->foot_slide_std[i] = 0.003;
    c->foot_swing_std[i] = 10000000.0;
    c->vo_p_std[i] = 0.000015;
    c->ekf_process_std[i] = 0.1;
    c->ekf_gravity_meas_std[i] = 4.0;
  }
  for (int i = 0; i < 8; ++i) {
    c->joint_position_std[i] = 0.04;
    c->joint_velocity_std[i] = 0.22;
  }
  c->accel_input_std[0] = 0.025;
  c->accel_input_std[1] = 0.025;
  c->accel_input_std[2] = 0.02;
  c->accel_bias_std[0] = 0.07;
  c->accel_bias_std[1] = 0.02;
  c->accel_bias_std[2] = 0.03;
  c->quaternion_ib[0] = 1.0;
  c->p_ib[0] = 0.01592;
  c->p_ib[1] = 0.06659;
  c->p_ib[2] = 0.00617;
  c->num_legs = 4;
  c->leg_odom_type = 0;
  c->contact_effort_threshold = 150.0;
  c->rate = 200;
  c->N = 20;
  c->est_type = 0;
  c->window_solve = DEKF_SOLVE_FULL;  // the reference's semantics: every update(T) re-solves the whole window
  c->rho = 0.1;
  c->alpha = 1.6;
  c->delta = 0.00001;
  c->sigma = 0.00001;
  c->verbose = 0;
  c->adaptRho = 1;
  c->polish = 0;
  c->maxQPIter = 4000;
  c->primTol = 1e-6;
  c->dualTol = 1e-6;
  c->realtiveTol = 1e-6;
  c->absTol = 1e-6;
  c->timeLimit = 0.0028;
  for (int i = 0; i < 4; ++i) {
    c->ekf_init_std[i] = 0.001;
    c->ekf_vo_meas_std[i] = 0.0001;
  }
  c->ekf_quaternion_init[0] = 1.0;
  c->ekf_rate = 500;
  // Vector3d p_imu_2_opti(0.016041, 0.089061, 0.0579875), DecentralEst.cpp:181-185
  c->p_imu_2_opti[0] = 0.016041;
  c->p_imu_2_opti[1] = 0.089061;
  c->p_imu_2_opti[2] = 0.0579875;
}

inline int robot_num_legs(int robot) { return robot == DEKF_ROBOT_CASSIE ? 2 : (robot == DEKF_ROBOT_POGOX ? 1 : 4); }
inline int robot_nj(int robot) { return robot == DEKF_ROBOT_CASSIE ? 5 : 3; }

inline Dims make_dims(const dekf_config &c) {
  Dims d;
  d.n = c.n_instances;
  d.ns = (c.n_instances + 127) / 128 * 128;
  d.N = c.N;
  d.NW = c.N + 2;  // window stages T-N .. T plus one slot so that stage T+1 can be assembled while update(T) is in flight
  d.HR = 4 * c.N + 1;
  d.D = c.ekf_hist_depth;
  d.tile0 = 0;
  return d;
}

template <typename T>
inline EkfConst<T> make_ekf_const(const dekf_config &c) {
  EkfConst<T> e;
  e.dt = (T)(1 / static_cast<double>(c.ekf_rate));  // orien_ekf.cpp:25
  for (int i = 0; i < 3; ++i) {
    e.Cg[i] = (T)std::pow(c.ekf_process_std[i], 2);       // :28
    e.Ca[i] = (T)std::pow(c.ekf_gravity_meas_std[i], 2);  // :29
  }
  for (int i = 0; i < 4; ++i) {
    e.Cvo[i] = (T)std::pow(c.ekf_vo_meas_std[i], 2);  // :30
    e.P0[i] = (T)std::pow(c.ekf_init_std[i], 2);      // :27
    e.q0[i] = (T)c.ekf_quaternion_init[i];            // :31-33
  }
  e.g[0] = (T)0;
  e.g[1] = (T)0;
  e.g[2] = (T)9.81;  // :11
  return e;
}

template <typename T>
inline MheConst<T> make_mhe_const(const dekf_config &c) {
  MheConst<T> m;
  const double dt = 1.0 / c.rate;  // DecentralEst.cpp:15
  m.dt = (T)dt;
  m.dt_d = dt;
  m.N = c.N;
  m.est_type = c.est_type;
  m.window_solve = (c.est_type == 0 && !c.v_box_enable && !c.x_box_mask) ? c.window_solve : 0;
  m.thr = c.contact_effort_threshold;
  for (int i = 0; i < 3; ++i) {
    const double Cp = std::pow(c.p_process_std[i], 2), Ca = std::pow(c.accel_input_std[i], 2);
    // Q_dyn^-1 = G C G' with G = [[R dt, R dt^2/2],[0, R dt]], C = blkdiag(C_p, C_accel)
    // (DecentralEst.cpp:409-418) = blk(R,R) [[dt^2 Cp + dt^4/4 Ca, dt^3/2 Ca],[., dt^2 Ca]] blk(R,R)'
    m.d1[i] = (T)(dt * dt * Cp + 0.25 * dt * dt * dt * dt * Ca);
    m.d2[i] = (T)(0.5 * dt * dt * dt * Ca);
    m.d3[i] = (T)(dt * dt * Ca);
    m.cab[i] = (T)(dt * dt * std::pow(c.accel_bias_std[i], 2));  // (Q_accel_bias/dt^2)^-1, :422-424
    m.cvo[i] = (T)std::pow(c.vo_p_std[i], 2);                    // Q_cam^-1 = R diag R', :477
    m.cgy[i] = (T)std::pow(c.gyro_input_std[i], 2);
    m.q_swing[i] = (T)(1 / std::pow(c.foot_swing_std[i], 2));
    m.P0[0 + i] = (T)std::pow(c.p_init_std[i], 2);  // Q_prior^-1, :239-253
    m.P0[3 + i] = (T)std::pow(c.v_init_std[i], 2);
    m.P0[6 + i] = (T)std::pow(c.accel_bias_init_std[i], 2);
    m.p_ib[i] = (T)c.p_ib[i];
    m.lever[i] = (T)c.p_imu_2_opti[i];
  }
  for (int i = 0; i < 3; ++i) {
    m.n1[i] = (i == 0) ? m.d1[0] : m.d1[i] - m.d1[0];
    m.n2[i] = (i == 0) ? m.d2[0] : m.d2[i] - m.d2[0];
    m.n3[i] = (i == 0) ? m.d3[0] : m.d3[i] - m.d3[0];
    m.nvo[i] = (i == 0) ? m.cvo[0] : m.cvo[i] - m.cvo[0];
    const double Ca0 = std::pow(c.accel_input_std[0], 2), Cai = std::pow(c.accel_input_std[i], 2);
    const double Cp0 = dt * dt * std::pow(c.p_process_std[0], 2), Cpi = dt * dt * std::pow(c.p_process_std[i], 2);
    m.ae[i] = (T)((i == 0) ? Ca0 : Cai - Ca0);
    m.pe[i] = (T)((i == 0) ? Cp0 : Cpi - Cp0);
  }
  for (int i = 0; i < 8; ++i) {
    m.cenc_v[i] = (T)std::pow(c.joint_velocity_std[i], 2);
    m.cenc_p[i] = (T)std::pow(c.joint_position_std[i], 2);
  }
  return m;
}

// constants of the state-constrained solve (always double)
inline BoxConst make_box_const(const dekf_config &c) {
  BoxConst b;
  std::memset(&b, 0, sizeof(b));
  b.mask9 = (c.v_box_enable ? 0x38 : 0) | (c.x_box_mask & 0x1ff);
  b.enable = b.mask9 != 0;
  b.general = (b.mask9 & ~0x38) != 0;
  for (int a = 0; a < 9; ++a) {
    const bool gx = ((c.x_box_mask >> a) & 1) != 0, gv = c.v_box_enable && a >= 3 && a < 6;
    b.lo9[a] = gx ? c.x_box_lo[a] : (gv ? c.v_box_lo[a - 3] : -1e300);
    b.hi9[a] = gx ? c.x_box_hi[a] : (gv ? c.v_box_hi[a - 3] : 1e300);
  }
  b.max_iter = c.v_box_max_iter > 0 ? c.v_box_max_iter : 50;
  const double dt = 1.0 / c.rate;
  b.dt = dt;
  for (int i = 0; i < 3; ++i) {
    b.lo[i] = b.lo9[3 + i];  // the velocity box as the team kernel reads it
    b.hi[i] = b.hi9[3 + i];
    b.lever[i] = c.p_imu_2_opti[i];
    const double Cp = std::pow(c.p_process_std[i], 2), Ca = std::pow(c.accel_input_std[i], 2);
    const double d1 = dt * dt * Cp + 0.25 * dt * dt * dt * dt * Ca, d2 = 0.5 * dt * dt * dt * Ca, d3 = dt * dt * Ca;
    const double det = d1 * d3 - d2 * d2;  // (G C G')^-1 per axis, DecentralEst.cpp:409-418
    b.qa[i] = d3 / det;
    b.qb[i] = -d2 / det;
    b.qc[i] = d1 / det;
    b.qab[i] = 1.0 / (dt * dt * std::pow(c.accel_bias_std[i], 2));  // Q_accel_bias / dt^2, :422-424
    b.qvo[i] = 1.0 / std::pow(c.vo_p_std[i], 2);                    // Q_vo, :477
  }
  return b;
}

// Row basis of the general linear rows (dekf_add_state_rows): W = [the count rows of `a`, normalised; the component bounds of the config
// as unit rows; an orthonormal completion (Gram-Schmidt residuals of the unit vectors farthest from the span so far)], V = W^-1 by Gauss-Jordan
// with partial pivoting; lo / hi in row order.  Returns the number of bounded rows m, or -1 (dependent rows, more than 9, lb >= ub).
inline int make_row_basis(const dekf_config &c, int count, const double *a, const double *lb, const double *ub, double *W, double *V,
                          double *lo, double *hi) {
  double Q[81];
  int nq = 0, m = 0;
  auto push = [&](const double *row, double l, double u) -> bool {
    if (m >= 9 || !(l < u)) return false;
    double r[9], na = 0.0, nr = 0.0;
    for (int k = 0; k < 9; ++k) {
      r[k] = row[k];
      na += row[k] * row[k];
    }
    for (int q = 0; q < nq; ++q) {
      double d = 0.0;
      for (int k = 0; k < 9; ++k) d += Q[q * 9 + k] * r[k];
      for (int k = 0; k < 9; ++k) r[k] -= d * Q[q * 9 + k];
    }
    for (int k = 0; k < 9; ++k) nr += r[k] * r[k];
    if (na == 0.0 || !(nr > 1e-16 * na)) return false;
    for (int k = 0; k < 9; ++k) Q[nq * 9 + k] = r[k] / std::sqrt(nr);
    ++nq;
    const double inv = 1.0 / std::sqrt(na);  // rows enter the basis with unit length (conditioning of W): bounds scale along
    for (int k = 0; k < 9; ++k) W[m * 9 + k] = row[k] * inv;
    lo[m] = l * inv;
    hi[m] = u * inv;
    ++m;
    return true;
  };
  for (int i = 0; i < count; ++i)
    if (!push(a + (size_t)i * 9, lb[i], ub[i])) return -1;
  for (int i = 0; i < 9; ++i) {
    const bool gx = ((c.x_box_mask >> i) & 1) != 0, gv = c.v_box_enable && i >= 3 && i < 6;
    if (!gx && !gv) continue;
    double e[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    e[i] = 1.0;
    if (!push(e, gx ? c.x_box_lo[i] : c.v_box_lo[i - 3], gx ? c.x_box_hi[i] : c.v_box_hi[i - 3])) return -1;
  }
  const int rows = m;
  while (nq < 9) {
    int best = -1;
    double bestn = -1.0, br[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 9; ++k) {
      double r[9], nr = 0.0;
      for (int cc = 0; cc < 9; ++cc) r[cc] = cc == k ? 1.0 : 0.0;
      for (int q = 0; q < nq; ++q) {
        const double d = Q[q * 9 + k];
        for (int cc = 0; cc < 9; ++cc) r[cc] -= d * Q[q * 9 + cc];
      }
      for (int cc = 0; cc < 9; ++cc) nr += r[cc] * r[cc];
      if (nr > bestn) {
        bestn = nr;
        best = k;
        for (int cc = 0; cc < 9; ++cc) br[cc] = r[cc];
      }
    }
    (void)best;
    for (int cc = 0; cc < 9; ++cc) Q[nq * 9 + cc] = br[cc] / std::sqrt(bestn);
    for (int cc = 0; cc < 9; ++cc) W[nq * 9 + cc] = Q[nq * 9 + cc];  // orthonormal to everything so far: cond(W) = cond(rows)
    ++nq;
  }
  double A[9][18];
  for (int r = 0; r < 9; ++r)
    for (int cc = 0; cc < 9; ++cc) {
      A[r][cc] = W[r * 9 + cc];
      A[r][9 + cc] = r == cc ? 1.0 : 0.0;
    }
  for (int cc = 0; cc < 9; ++cc) {
    int pv = cc;
    for (int r = cc + 1; r < 9; ++r)
      if (std::fabs(A[r][cc]) > std::fabs(A[pv][cc])) pv = r;
    if (A[pv][cc] == 0.0) return -1;
    if (pv != cc)
      for (int k = 0; k < 18; ++k) std::swap(A[cc][k], A[pv][k]);
    const double d = 1.0 / A[cc][cc];
    for (int k = 0; k < 18; ++k) A[cc][k] *= d;
    for (int r = 0; r < 9; ++r) {
      if (r == cc) continue;
      const double f = A[r][cc];
      if (f != 0.0)
        for (int k = 0; k < 18; ++k) A[r][k] -= f * A[cc][k];
    }
  }
  for (int r = 0; r < 9; ++r)
    for (int cc = 0; cc < 9; ++cc) V[r * 9 + cc] = A[r][9 + cc];
  return rows;
}

// constants of the foot-state model (leg_odom_type 1; always double)
inline FootConst make_foot_const(const dekf_config &c) {
  FootConst f;
  std::memset(&f, 0, sizeof(f));
  const double dt = 1.0 / c.rate;
  f.bc = make_box_const(c);
  f.N = c.N;
  f.est_type = c.est_type;
  for (int i = 0; i < 3; ++i) {
    f.q_slide[i] = 1.0 / (dt * dt * std::pow(c.foot_slide_std[i], 2));  // R Q_foot_slide R' / dt^2, DecentralEst.cpp:438-448
    f.q_swing[i] = 1.0 / (dt * dt * std::pow(c.foot_swing_std[i], 2));
    f.M0[0 + i] = 1.0 / std::pow(c.p_init_std[i], 2);  // :239-253
    f.M0[3 + i] = 1.0 / std::pow(c.v_init_std[i], 2);
    f.M0[6 + i] = 1.0 / std::pow(c.accel_bias_init_std[i], 2);
    f.M0_foot[i] = 1.0 / std::pow(c.foot_init_std[i], 2);  // :313-322
  }
  return f;
}
inline int state_dim(const dekf_config &c) { return 9 + 3 * c.leg_odom_type * robot_num_legs(c.robot); }

// fields per instance of every state array, in units of elements
struct StateSizes {
  size_t ekf_q, ekf_P, ekf_hist, ekf_hist_time, arr_P, arr_x, win, hist_time, hist_quat, wp, wp_time,
      wp_count, p_vo, pend_flag, pend, status;
};
inline StateSizes state_sizes(const Dims &d) {
  StateSizes s;
  const size_t n = (size_t)d.ns;
  s.ekf_q = 4 * n;
  s.ekf_P = 16 * n;
  s.ekf_hist = (size_t)d.D * EKF_HIST_FIELDS * n;
  s.ekf_hist_time = (size_t)d.D * n;
  s.arr_P = 45 * n;
  s.arr_x = 9 * n;
  s.win = (size_t)d.NW * REC_SIZE * n;
  s.hist_time = (size_t)d.HR * n;
  s.hist_quat = (size_t)d.HR * 4 * n;
  s.wp = 12 * n;
  s.wp_time = 4 * n;
  s.wp_count = n;
  s.p_vo = 3 * n;
  s.pend_flag = n;
  s.pend = 5 * n;
  s.status = 2 * n;
  return s;
}

}  // namespace dekf
