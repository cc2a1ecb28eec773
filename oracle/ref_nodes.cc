// Lock-step driver of the reference's OWN node classes, compiled unmodified from /root/reference (never copied):
//   orien_ekf::orien_ekf            src/orien_est/src/orien_ekf.cpp          ("orien_sub": the orientation EKF)
//   robotSub::go1Sub / robotSub     src/go1_example/src/go1Sub.cpp, src/decentral_legged_est/src/EstSub.cpp
//   DecentralizedEstimation, MHEproblem, Bezier   src/decentral_legged_est/src/{DecentralEst,MheSrb}.cpp, Spline/
//   SymFunction::*                  src/go1_example/src/Expressions/*.cc     (FROST kinematics)
// against the stand-in headers under oracle/ref_stub/ (Eigen3, OSQP/osqp-eigen and rclcpp are absent from the image).
//
// TEST INFRASTRUCTURE ONLY.  Output: oracle/_ref/libref_nodes.so (git-ignored).  Purpose: pin the plain-C restatement
// (oracle/*.c) against the reference's real control flow, QP bookkeeping, VO synchronisation, marginalisation formulas
// and EKF replay.  NOT covered: the arithmetic inside Eigen and OSQP themselves (stand-ins: dense eager linear
// algebra; exact KKT solve or the oracle's own OSQP-style ADMM).
//
// One tick = what the two ROS timers do when run in lock-step at the same rate (the benchmark's instance-step):
//   messages of the tick are delivered to the subscriptions (unitree/imu, unitree/joint_state, orb/pos, orb/vo), then
//   orien_ekf::timerCallback (publishes imu/filter -> robotSub::orien_filter_callback), then robotSub::timerCallback
//   (initialize at discrete time 0, update(T) afterwards).
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <stdio.h>
#include <string.h>

#include <Eigen/Sparse>
#include <OsqpEigen/OsqpEigen.h>
#include <rclcpp/rclcpp.hpp>
#include <sensor_msgs/msg/imu.hpp>
#include <sensor_msgs/msg/joint_state.hpp>
#include <geometry_msgs/msg/pose_stamped.hpp>
#include <optitrack_broadcast/msg/mocap.hpp>
#include <custom_msgs/msg/vo_realtive_transform.hpp>

// read-only access to the reference classes' internals from this driver (layout is unaffected by access specifiers)
#define private public
#define protected public
#include "go1Sub.hpp"
#include "orien_ekf.hpp"
#undef private
#undef protected

namespace {
struct Nodes {
  std::shared_ptr<orien_ekf::orien_ekf> ekf;
  std::shared_ptr<robotSub::go1Sub> est;
  sensor_msgs::msg::Imu last_filter;
};

// the reference prints timing / diagnostics on every tick: swallow them while its code runs
struct NullBuf : std::streambuf {
  int overflow(int c) override { return traits_type::not_eof(c); }
  std::streamsize xsputn(const char *, std::streamsize n) override { return n; }
};
struct CoutMute {
  NullBuf nb;
  std::streambuf *old;
  CoutMute() : old(std::cout.rdbuf(&nb)) {}
  ~CoutMute() { std::cout.rdbuf(old); }
};
bool g_mute = true;

void stamp_from_ns(int64_t ns, builtin_interfaces::msg::Time &t) {
  t.sec = (int32_t)(ns / 1000000000LL);
  t.nanosec = (uint32_t)(ns % 1000000000LL);
}
}  // namespace

extern "C" {

void ref_set_verbose(int v) { g_mute = !v; }
void ref_set_osqp_mode(int m) { refstub::osqp_mode() = m; }
int ref_last_kkt_bandwidth() { return refstub::osqp_last_bandwidth(); }
int ref_last_admm_iters() { return refstub::osqp_last_iters(); }

void ref_param_clear() { refstub::param_overrides().clear(); }
void ref_param_set_int(const char *key, long long v) { refstub::ParamValue p; p.kind = 1; p.i = v; refstub::param_overrides()[key] = p; }
void ref_param_set_double(const char *key, double v) { refstub::ParamValue p; p.kind = 2; p.d = v; refstub::param_overrides()[key] = p; }
void ref_param_set_bool(const char *key, int v) { refstub::ParamValue p; p.kind = 3; p.b = v != 0; refstub::param_overrides()[key] = p; }
void ref_param_set_string(const char *key, const char *v) { refstub::ParamValue p; p.kind = 4; p.s = v; refstub::param_overrides()[key] = p; }
void ref_param_set_doubles(const char *key, const double *v, int n) {
  refstub::ParamValue p; p.kind = 5; p.v.assign(v, v + n); refstub::param_overrides()[key] = p;
}

// One set of nodes may be live at a time (the stand-in topic bus is process-global, like one ROS graph).
void *ref_nodes_create(int with_ekf, int with_est) {
  CoutMute *mute = g_mute ? new CoutMute : nullptr;
  refstub::clear_bus();
  refstub::now_ns() = 0;  // both constructors latch time_init_ = now() (orien_ekf.cpp:44, EstSub.cpp:27)
  Nodes *nd = new Nodes;
  if (with_ekf) nd->ekf = std::make_shared<orien_ekf::orien_ekf>("orien_sub");
  if (with_est) {
    nd->est = std::make_shared<robotSub::go1Sub>("est_sub");
    nd->est->imu_msg_num_ = 10;  // robotSub::timerCallback starts estimating after 10 IMU messages (EstSub.cpp:62)
  }
  delete mute;
  return nd;
}

void ref_nodes_destroy(void *h) {
  CoutMute *mute = g_mute ? new CoutMute : nullptr;
  delete (Nodes *)h;  // the topic bus is process-global and re-created by the next ref_nodes_create: one node set at a time
  delete mute;
}

// ---- single events (any schedule of messages and timers; ref_nodes_tick below is the 1:1 lock-step schedule)
void ref_nodes_set_clock(long long t_ns) { refstub::now_ns() = t_ns; }
// /unitree/imu at the current clock -> orien_ekf::imu_callback and go1Sub::imu_callback
void ref_nodes_msg_imu(void *h, const double *gyro, const double *accel) {
  (void)h;
  sensor_msgs::msg::Imu imu;
  stamp_from_ns(refstub::now_ns(), imu.header.stamp);
  imu.angular_velocity.x = gyro[0]; imu.angular_velocity.y = gyro[1]; imu.angular_velocity.z = gyro[2];
  imu.linear_acceleration.x = accel[0]; imu.linear_acceleration.y = accel[1]; imu.linear_acceleration.z = accel[2];
  refstub::deliver<sensor_msgs::msg::Imu>("unitree/imu", imu);
}
// /unitree/joint_state -> go1Sub::lo_callback (positions = joints then foot forces, go1Sub.cpp:68-75)
void ref_nodes_msg_joint(void *h, const double *joint_pos, int n_pos, const double *joint_vel, int n_vel) {
  (void)h;
  sensor_msgs::msg::JointState js;
  stamp_from_ns(refstub::now_ns(), js.header.stamp);
  js.position.assign(joint_pos, joint_pos + n_pos);
  js.velocity.assign(joint_vel, joint_vel + n_vel);
  refstub::deliver<sensor_msgs::msg::JointState>("unitree/joint_state", js);
}
// orb/pos -> orien_ekf::vo_pose_callback, orb/vo -> robotSub::vo_callback; stamps are integer nanoseconds as ROS carries them
void ref_nodes_msg_vo(void *h, const double *vo_quat_wxyz, long long vo_now_ns, long long vo_pre_ns, const double *vo_rel_p) {
  (void)h;
  geometry_msgs::msg::PoseStamped ps;
  stamp_from_ns(vo_now_ns, ps.header.stamp);
  ps.pose.orientation.w = vo_quat_wxyz[0]; ps.pose.orientation.x = vo_quat_wxyz[1];
  ps.pose.orientation.y = vo_quat_wxyz[2]; ps.pose.orientation.z = vo_quat_wxyz[3];
  refstub::deliver<geometry_msgs::msg::PoseStamped>("orb/pos", ps);
  custom_msgs::msg::VoRealtiveTransform vt;
  stamp_from_ns(vo_now_ns, vt.header.stamp);
  stamp_from_ns(vo_pre_ns, vt.header_pre.stamp);
  vt.x_relative = vo_rel_p[0]; vt.y_relative = vo_rel_p[1]; vt.z_relative = vo_rel_p[2];
  refstub::deliver<custom_msgs::msg::VoRealtiveTransform>("orb/vo", vt);
}
// imu/filter published by something else than the EKF node (runs without it)
void ref_nodes_msg_filter(void *h, const double *quat_wxyz) {
  (void)h;
  sensor_msgs::msg::Imu f;
  f.orientation.w = quat_wxyz[0]; f.orientation.x = quat_wxyz[1]; f.orientation.y = quat_wxyz[2]; f.orientation.z = quat_wxyz[3];
  refstub::deliver<sensor_msgs::msg::Imu>("imu/filter", f);
}
void ref_nodes_fire_ekf(void *h) {  // orien_ekf::timerCallback; publishes imu/filter -> robotSub::orien_filter_callback
  Nodes *nd = (Nodes *)h;
  CoutMute *mute = g_mute ? new CoutMute : nullptr;
  if (nd->ekf) nd->ekf->fire_timers();
  delete mute;
}
void ref_nodes_fire_est(void *h) {  // robotSub::timerCallback: initialize at discrete time 0, update(T) afterwards
  Nodes *nd = (Nodes *)h;
  CoutMute *mute = g_mute ? new CoutMute : nullptr;
  if (nd->est) nd->est->fire_timers();
  delete mute;
}

// One lock-step tick.  quat_in (or NULL): orientation published on imu/filter when the EKF node is not instantiated.
void ref_nodes_tick(void *h, long long t_ns, const double *gyro, const double *accel, const double *joint_pos, int n_pos,
                    const double *joint_vel, int n_vel, int vo_new, const double *vo_quat_wxyz, long long vo_now_ns,
                    long long vo_pre_ns, const double *vo_rel_p, const double *quat_in_wxyz) {
  Nodes *nd = (Nodes *)h;
  ref_nodes_set_clock(t_ns);
  ref_nodes_msg_imu(h, gyro, accel);
  if (nd->est) ref_nodes_msg_joint(h, joint_pos, n_pos, joint_vel, n_vel);
  if (vo_new) ref_nodes_msg_vo(h, vo_quat_wxyz, vo_now_ns, vo_pre_ns, vo_rel_p);
  if (nd->ekf) ref_nodes_fire_ekf(h);
  else if (quat_in_wxyz) ref_nodes_msg_filter(h, quat_in_wxyz);
  ref_nodes_fire_est(h);
}

void ref_nodes_get_quat(void *h, double *q4, double *P16) {
  Nodes *nd = (Nodes *)h;
  for (int i = 0; i < 4; ++i) q4[i] = nd->ekf->quaternion_(i);
  if (P16) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) P16[i * 4 + j] = nd->ekf->Cov_q_(i, j);
}
int ref_nodes_state_dim(void *h) { return ((Nodes *)h)->est->mhe.dim_state_; }
// x_MHE_ / x_KF_ (ds), v_*_b_ (3), p_vo_accmulate_ (3), R_sb_ (row-major 9), contact (num_legs)
void ref_nodes_get_est(void *h, double *x, double *v_body, double *p_vo, double *R_sb, double *contact) {
  Nodes *nd = (Nodes *)h;
  DecentralizedEstimation &m = nd->est->mhe;
  bool kf = m.est_type_ == 1;
  const Eigen::Mat &xs = kf ? (const Eigen::Mat &)m.x_KF_ : (const Eigen::Mat &)m.x_MHE_;
  if (x) for (int i = 0; i < xs.size(); ++i) x[i] = xs(i);
  if (v_body) for (int i = 0; i < 3; ++i) v_body[i] = kf ? m.v_KF_b_(i) : m.v_MHE_b_(i);
  if (p_vo) for (int i = 0; i < 3; ++i) p_vo[i] = m.p_vo_accmulate_(i);
  if (R_sb) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R_sb[i * 3 + j] = m.R_sb_(i, j);
  if (contact) for (int i = 0; i < nd->est->robot_store_->contact_.size(); ++i) contact[i] = nd->est->robot_store_->contact_(i);
}
int ref_nodes_get_x_size(void *h) {
  DecentralizedEstimation &m = ((Nodes *)h)->est->mhe;
  return m.est_type_ == 1 ? m.x_KF_.size() : m.x_MHE_.size();
}
void ref_nodes_get_kf_cov(void *h, double *C) {
  DecentralizedEstimation &m = ((Nodes *)h)->est->mhe;
  for (int i = 0; i < m.C_KF_.rows(); ++i) for (int j = 0; j < m.C_KF_.cols(); ++j) C[i * m.C_KF_.cols() + j] = m.C_KF_(i, j);
}
// K_KF_ (DecentralEst.hpp:290, DecentralEst.cpp:858), row-major rows x cols; returns cols (0 before the first correction)
int ref_nodes_get_kf_gain(void *h, double *K) {
  DecentralizedEstimation &m = ((Nodes *)h)->est->mhe;
  for (int i = 0; i < m.K_KF_.rows(); ++i) for (int j = 0; j < m.K_KF_.cols(); ++j) K[i * m.K_KF_.cols() + j] = m.K_KF_(i, j);
  return m.K_KF_.cols();
}
// leg kinematics handed to the estimator by go1Sub::lo_callback: p_imu_2_foot_ (3*legs), J_imu_2_foot_ (3*legs x 3, row-major)
void ref_nodes_get_kin(void *h, double *p, double *J) {
  Nodes *nd = (Nodes *)h;
  const Eigen::Mat &pm = nd->est->robot_store_->p_imu_2_foot_, &Jm = nd->est->robot_store_->J_imu_2_foot_;
  for (int i = 0; i < pm.rows(); ++i) p[i] = pm(i, 0);
  for (int i = 0; i < Jm.rows(); ++i) for (int j = 0; j < Jm.cols(); ++j) J[i * Jm.cols() + j] = Jm(i, j);
}
// arrival cost (MheSrb.hpp:86-87); returns its dimension (0 before the first marginalisation)
int ref_nodes_get_arrival(void *h, double *M, double *nvec) {
  MHEproblem &q = ((Nodes *)h)->est->mhe.mhe_qp_;
  int d = q.M_p.rows();
  if (d == 0) return 0;
  for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) M[i * d + j] = q.M_p(i, j);
  for (int i = 0; i < d; ++i) nvec[i] = q.n_p(i);
  return d;
}
// the QP handed to OSQP at the last initQP (MheSrb.cpp:272-293): dims, then dense row-major copies
void ref_nodes_qp_dims(void *h, int *nVar, int *nCon) {
  MHEproblem &q = ((Nodes *)h)->est->mhe.mhe_qp_;
  *nVar = q.nVar;
  *nCon = q.nConstraints;
}
void ref_nodes_export_qp(void *h, double *H, double *g, double *A, double *l, double *u, double *z) {
  MHEproblem &q = ((Nodes *)h)->est->mhe.mhe_qp_;
  int n = q.nVar, m = q.nConstraints;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) H[(size_t)i * n + j] = q.Hsparse(i, j);
  for (int i = 0; i < n; ++i) g[i] = q.g(i);
  for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = q.Aconstrsparse(i, j);
  for (int i = 0; i < m; ++i) { l[i] = q.lb_all(i); u[i] = q.ub_all(i); }
  if (z) for (int i = 0; i < n && i < q.solution.size(); ++i) z[i] = q.solution(i);
}
// integer VO bookkeeping (DecentralEst.cpp:883-945,987-1009): out[0] = vo_insert_idx_stack_.size(), out[1] = its back (-1 if
// empty), out[2] = vo_insert_discrete_time_stack_.back(), out[3] = vo_curve_.node_count(), out[4] = way-point count,
// out[5] = imu_time_stack_.size()
void ref_nodes_vo_debug(void *h, int *out) {
  DecentralizedEstimation &m = ((Nodes *)h)->est->mhe;
  out[0] = (int)m.vo_insert_idx_stack_.size();
  out[1] = m.vo_insert_idx_stack_.empty() ? -1 : m.vo_insert_idx_stack_.back();
  out[2] = m.vo_insert_discrete_time_stack_.empty() ? -1 : m.vo_insert_discrete_time_stack_.back();
  out[3] = m.vo_curve_.node_count();
  out[4] = (int)m.vo_curve_._way_points.size();
  out[5] = (int)m.imu_time_stack_.size();
}

// Batch runner over SoA streams [step][field][instance] (same layout as orc_run_batch / the device ABI), instances
// [i0, i1) one after the other.  Time stamps are integer nanoseconds.  Optional outputs ([S][k][n], NULL to skip).
void ref_run_stream(int n, int S, int i0, int i1, int with_ekf, const double *gyro, const double *accel,
                    const long long *imu_ns, const double *joint_pos, const double *joint_vel, const double *foot_force,
                    int nq, int nl, const unsigned char *vo_flag, const double *vo_quat, const long long *vo_pre_ns,
                    const long long *vo_now_ns, const double *vo_rel_p, const double *quat_in, double *out_quat,
                    double *out_x, int ds, double *out_v_body, double *out_p_vo, unsigned char *out_contact,
                    int *out_vo_dbg /*[S][6][n]*/, double *out_M_p, double *out_n_p) {
  for (int i = i0; i < i1; ++i) {
    void *h = ref_nodes_create(with_ekf, 1);
    std::vector<double> jp((size_t)(nq + nl)), jv((size_t)nq), xs(64);
    for (int s = 0; s < S; ++s) {
      size_t b1 = (size_t)s * n + i;
      double g[3], a[3], vq[4] = {1, 0, 0, 0}, vp[3] = {0, 0, 0}, qi[4] = {1, 0, 0, 0};
      for (int c = 0; c < 3; ++c) { g[c] = gyro[((size_t)s * 3 + c) * n + i]; a[c] = accel[((size_t)s * 3 + c) * n + i]; }
      for (int c = 0; c < nq; ++c) { jp[(size_t)c] = joint_pos[((size_t)s * nq + c) * n + i]; jv[(size_t)c] = joint_vel[((size_t)s * nq + c) * n + i]; }
      for (int c = 0; c < nl; ++c) jp[(size_t)(nq + c)] = foot_force[((size_t)s * nl + c) * n + i];
      int vn = vo_flag ? vo_flag[b1] : 0;
      if (vn) {
        for (int c = 0; c < 4; ++c) vq[c] = vo_quat[((size_t)s * 4 + c) * n + i];
        for (int c = 0; c < 3; ++c) vp[c] = vo_rel_p[((size_t)s * 3 + c) * n + i];
      }
      if (quat_in) for (int c = 0; c < 4; ++c) qi[c] = quat_in[((size_t)s * 4 + c) * n + i];
      ref_nodes_tick(h, imu_ns[b1], g, a, jp.data(), nq + nl, jv.data(), nq, vn, vq, vn ? vo_now_ns[b1] : 0,
                     vn ? vo_pre_ns[b1] : 0, vp, quat_in ? qi : nullptr);
      Nodes *nd = (Nodes *)h;
      if (out_quat && nd->ekf) {
        double q[4];
        ref_nodes_get_quat(h, q, nullptr);
        for (int c = 0; c < 4; ++c) out_quat[((size_t)s * 4 + c) * n + i] = q[c];
      }
      double vb[3], pv[3], ct[8];
      ref_nodes_get_est(h, (s >= 1 || nd->est->mhe.est_type_ == 1) ? xs.data() : nullptr, vb, pv, nullptr, ct);
      if (out_x && (s >= 1 || nd->est->mhe.est_type_ == 1)) for (int c = 0; c < ds; ++c) out_x[((size_t)s * ds + c) * n + i] = xs[(size_t)c];
      if (out_v_body && s >= 1) for (int c = 0; c < 3; ++c) out_v_body[((size_t)s * 3 + c) * n + i] = vb[c];
      if (out_p_vo) for (int c = 0; c < 3; ++c) out_p_vo[((size_t)s * 3 + c) * n + i] = pv[c];
      if (out_contact) for (int c = 0; c < nl; ++c) out_contact[((size_t)s * nl + c) * n + i] = (unsigned char)(ct[c] != 0.0);
      if (out_vo_dbg) {
        int d[6];
        ref_nodes_vo_debug(h, d);
        for (int c = 0; c < 6; ++c) out_vo_dbg[((size_t)s * 6 + c) * n + i] = d[c];
      }
    }
    if (out_M_p) {
      std::vector<double> M((size_t)ds * ds), nv((size_t)ds);
      if (ref_nodes_get_arrival(h, M.data(), nv.data()) == ds) {
        for (int c = 0; c < ds * ds; ++c) out_M_p[(size_t)c * n + i] = M[(size_t)c];
        for (int c = 0; c < ds; ++c) out_n_p[(size_t)c * n + i] = nv[(size_t)c];
      }
    }
    ref_nodes_destroy(h);
  }
}

}  // extern "C"
