// Drop-in replacement of the reference header
//     /root/reference/src/decentral_legged_est/include/decentral_legged_est/DecentralEst.hpp
// for ONE robot: same include path ("decentral_legged_est/DecentralEst.hpp"), same global names (robot_params, robot_store,
// DecentralizedEstimation), same Eigen member types, same methods -- but initialize()/update(T) step a libdekf_b200.so handle
// (n_instances = 1) on the GPU instead of building and solving the QP with OSQP on the CPU.
//
// Use: put  -I <repo>/include/dekf_b200/dropin  BEFORE the reference's own include directory and link libdekf_b200.so.  The
// reference's node sources -- decentral_legged_est/src/EstSub.cpp:58-91 (timer, initialize/update), go1_example/src/
// go1Sub.cpp:53-126 (callbacks that fill robot_store) -- then compile UNMODIFIED against this file; their other headers
// (EstSub.hpp, EigenUtils.hpp, data_logger.hpp) still come from the reference.  tests/test_dropin_reference_nodes.py does
// exactly that (Eigen stand-in: oracle/ref_stub/mini_eigen.hpp, the image has no Eigen3) and replays the golden streams.
//
// What differs from the reference class, by design:
//   * robot_store_->p_imu_2_foot_, J_imu_2_foot_, contact_ (filled on the host by go1Sub::lo_callback through the FROST
//     functions) are NOT read: the device recomputes kinematics, Jacobians and contact flags from joint_states_position_
//     (angles, foot forces in rows 12..15, go1Sub.cpp:68-75) -- same inputs, same results (tests pin both to FROST).
//   * mhe_qp_ (class MHEproblem, MheSrb.hpp:58-112) is replaced by the getters M_p()/n_p() of the arrival cost; the QP is
//     never materialised.  OSQP settings in robot_params are accepted and ignored.
//   * K_KF_ is filled only with est_type_ == 1 (cfg.kf_export_gain), C_KF_ likewise.
#ifndef MHE_EST_HPP
#define MHE_EST_HPP

#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <Eigen/Sparse>

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../dekf_b200.h"

using namespace Eigen;

// DecentralEst.hpp:18-63 (field for field)
struct robot_params {
  std::vector<double> p_process_std_;
  std::vector<double> accel_input_std_;
  std::vector<double> accel_bias_std_;
  std::vector<double> gyro_input_std_;

  std::vector<double> quaternion_ib_;
  std::vector<double> p_ib_;

  int num_legs_;
  int leg_odom_type_;
  std::vector<double> joint_position_std_;
  std::vector<double> joint_velocity_std_;
  std::vector<double> foot_slide_std_;
  std::vector<double> foot_swing_std_;
  double contact_effort_theshold_;

  std::vector<double> p_init_std_;
  std::vector<double> v_init_std_;
  std::vector<double> foot_init_std_;
  std::vector<double> accel_bias_init_std_;

  std::vector<double> vo_p_std_;

  int rate_;
  int N_;
  int est_type_;

  double rho_;
  double alpha_;
  double delta_;
  double sigma_;
  bool verbose_;
  bool adaptRho_;
  bool polish_;
  int maxQPIter_;
  double realtiveTol_;
  double absTol_;
  double primTol_;
  double dualTol_;
  double timeLimit_;
};

// DecentralEst.hpp:65-94 (field for field)
struct robot_store {
  double imu_time_;
  Vector3d accel_b_;
  Vector3d angular_b_;

  VectorXd joint_states_position_;
  VectorXd joint_states_velocity_;
  VectorXd joint_states_effort_;
  VectorXd contact_;

  MatrixXd p_imu_2_foot_;
  MatrixXd J_imu_2_foot_;

  double vo_time_pre_;
  double vo_time_now_;
  bool vo_new_ = false;
  Vector3d vo_p_body_pre_2_body_;
  Quaterniond vo_quaternion_;

  Quaterniond quaternion_;
  Quaterniond offset_quaternion_;

  Vector3d gt_p_;
  Vector3d gt_v_s_;
};

class DecentralizedEstimation {
 public:
  DecentralizedEstimation() {}
  DecentralizedEstimation(const DecentralizedEstimation &) = delete;
  DecentralizedEstimation &operator=(const DecentralizedEstimation &) = delete;
  ~DecentralizedEstimation() {
    if (h_) dekf_destroy(h_);
  }

  // DecentralEst.hpp:101, DecentralEst.cpp:9-150
  void initialize(std::shared_ptr<robot_store> sub_ptr, std::shared_ptr<robot_params> params_ptr) {
    robot_sub_ptr_ = sub_ptr;
    params_ptr_ = params_ptr;
    if (h_) {
      dekf_destroy(h_);
      h_ = nullptr;
    }
    dekf_config c;
    dekf_config_default_go1(&c);  // device / EKF block / lever arm defaults; every estimator parameter is overwritten below
    const robot_params &p = *params_ptr_;
    put(c.p_process_std, p.p_process_std_, 3, "p_process_std_");
    put(c.accel_input_std, p.accel_input_std_, 3, "accel_input_std_");
    put(c.accel_bias_std, p.accel_bias_std_, 3, "accel_bias_std_");
    put(c.gyro_input_std, p.gyro_input_std_, 3, "gyro_input_std_");
    put(c.quaternion_ib, p.quaternion_ib_, 4, "quaternion_ib_");
    put(c.p_ib, p.p_ib_, 3, "p_ib_");
    c.num_legs = p.num_legs_;
    c.leg_odom_type = p.leg_odom_type_;
    put(c.joint_position_std, p.joint_position_std_, 8, "joint_position_std_");  // the reference uses the first 3
    put(c.joint_velocity_std, p.joint_velocity_std_, 8, "joint_velocity_std_");
    put(c.foot_slide_std, p.foot_slide_std_, 3, "foot_slide_std_");
    put(c.foot_swing_std, p.foot_swing_std_, 3, "foot_swing_std_");
    c.contact_effort_threshold = p.contact_effort_theshold_;
    put(c.p_init_std, p.p_init_std_, 3, "p_init_std_");
    put(c.v_init_std, p.v_init_std_, 3, "v_init_std_");
    put(c.foot_init_std, p.foot_init_std_, 3, "foot_init_std_");
    put(c.accel_bias_init_std, p.accel_bias_init_std_, 3, "accel_bias_init_std_");
    put(c.vo_p_std, p.vo_p_std_, 3, "vo_p_std_");
    c.rate = p.rate_;
    c.N = p.N_;
    c.est_type = p.est_type_;
    c.rho = p.rho_;
    c.alpha = p.alpha_;
    c.delta = p.delta_;
    c.sigma = p.sigma_;
    c.verbose = p.verbose_;
    c.adaptRho = p.adaptRho_;
    c.polish = p.polish_;
    c.maxQPIter = p.maxQPIter_;
    c.realtiveTol = p.realtiveTol_;
    c.absTol = p.absTol_;
    c.primTol = p.primTol_;
    c.dualTol = p.dualTol_;
    c.timeLimit = p.timeLimit_;
    c.n_instances = 1;
    c.kf_export_gain = p.est_type_ == 1 && p.leg_odom_type_ == 0;
    check(dekf_create(&c, &h_), "dekf_create");
    if (!rows_lb_.empty())
      check(dekf_add_state_rows(h_, (int32_t)rows_lb_.size(), rows_a_.data(), rows_lb_.data(), rows_ub_.data()), "dekf_add_state_rows");
    nq_ = dekf_num_joints(h_);
    nl_ = c.num_legs;
    ds_ = dekf_state_dim(h_);
    kf_ = c.est_type == 1;
    x_MHE_ = VectorXd::Zero(ds_);
    x_KF_ = VectorXd::Zero(ds_);
    R_sb_ = Matrix3d::Zero();
    step(0);
  }
  // General inequality rows  lb(i) <= A.row(i) . x_k <= ub(i)  on every window state: MHEproblem::addConstraints(name, lb, ub) with a
  // dependency row on x_k (MheSrb.cpp:58-68, :217-270; the reference never calls it with lb < ub).  A is count x 9 over
  // (p_s, v_s, accel bias).  Call before initialize().
  void addStateRows(const MatrixXd &A, const VectorXd &lb, const VectorXd &ub) {
    if (A.cols() != 9 || A.rows() != lb.size() || lb.size() != ub.size()) throw std::runtime_error("addStateRows: A must be count x 9");
    rows_a_.assign((size_t)A.rows() * 9, 0.0);
    rows_lb_.assign((size_t)A.rows(), 0.0);
    rows_ub_.assign((size_t)A.rows(), 0.0);
    for (int i = 0; i < (int)A.rows(); ++i) {
      for (int k = 0; k < 9; ++k) rows_a_[(size_t)i * 9 + k] = A(i, k);
      rows_lb_[i] = lb(i);
      rows_ub_[i] = ub(i);
    }
  }
  // DecentralEst.hpp:102, DecentralEst.cpp:152-198
  void update(int T) {
    if (!h_) throw std::runtime_error("DecentralizedEstimation::update before initialize");
    step(T);
  }
  // DecentralEst.hpp:103
  void reset() {
    if (h_) check(dekf_reset(h_), "dekf_reset");
  }

  // MHEproblem::M_p / n_p (MheSrb.hpp:86-87): the arrival cost of the current window
  MatrixXd M_p() const {
    std::vector<double> m((size_t)ds_ * ds_);
    check(dekf_get_host(h_, DEKF_GET_ARRIVAL_M, m.data()), "dekf_get_host");
    MatrixXd M = MatrixXd::Zero(ds_, ds_);
    for (int r = 0; r < ds_; ++r)
      for (int c = 0; c < ds_; ++c) M(r, c) = m[(size_t)r * ds_ + c];
    return M;
  }
  VectorXd n_p() const {
    std::vector<double> v((size_t)ds_);
    check(dekf_get_host(h_, DEKF_GET_ARRIVAL_N, v.data()), "dekf_get_host");
    VectorXd n = VectorXd::Zero(ds_);
    for (int r = 0; r < ds_; ++r) n(r) = v[(size_t)r];
    return n;
  }
  int32_t status() const { return status_; }  // DEKF_ST_* bits of the last call (the reference prints and goes on)
  dekf_handle *handle() const { return h_; }

 public:
  // DecentralEst.hpp:278-291
  Matrix3d R_sb_;
  Vector3d p_vo_accmulate_ = Vector3d::Zero();

  VectorXd x_MHE_;
  Vector3d v_MHE_b_ = Vector3d::Zero();

  VectorXd x_KF_;
  MatrixXd C_KF_;
  MatrixXd K_KF_;
  Vector3d v_KF_b_ = Vector3d::Zero();

 private:
  static void put(double *dst, const std::vector<double> &src, size_t n, const char *name) {
    if (src.size() < n && !(n == 8 && src.size() >= 3)) throw std::runtime_error(std::string("robot_params.") + name + " is too short");
    for (size_t i = 0; i < n; ++i) dst[i] = i < src.size() ? src[i] : src.back();
  }
  void check(int rc, const char *what) const {
    if (rc != DEKF_OK) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (h_ ? dekf_last_error(h_) : ""));
  }
  void step(int T) {
    robot_store &st = *robot_sub_ptr_;
    double gyro[3], accel[3], quat[4], vo_p[3], jp[64], jv[64], ff[8];
    for (int c = 0; c < 3; ++c) {
      gyro[c] = st.angular_b_(c);
      accel[c] = st.accel_b_(c);
      vo_p[c] = st.vo_p_body_pre_2_body_(c);
    }
    quat[0] = st.quaternion_.w();
    quat[1] = st.quaternion_.x();
    quat[2] = st.quaternion_.y();
    quat[3] = st.quaternion_.z();
    if ((int)st.joint_states_position_.size() < nq_ + nl_ || (int)st.joint_states_velocity_.size() < nq_)
      throw std::runtime_error("robot_store: joint message shorter than num_legs * joints + num_legs foot forces");
    for (int j = 0; j < nq_; ++j) {
      jp[j] = st.joint_states_position_(j);
      jv[j] = st.joint_states_velocity_(j);
    }
    for (int l = 0; l < nl_; ++l) ff[l] = st.joint_states_position_(nq_ + l);  // go1Sub.cpp:74
    uint8_t vo_flag = st.vo_new_ ? 1 : 0;
    dekf_inputs in;
    std::memset(&in, 0, sizeof(in));
    in.gyro = gyro;
    in.accel = accel;
    in.imu_time = &st.imu_time_;
    in.joint_pos = jp;
    in.joint_vel = jv;
    in.foot_force = ff;
    in.quat = quat;
    if (vo_flag) {
      in.vo_flag = &vo_flag;
      in.vo_time_pre = &st.vo_time_pre_;
      in.vo_time_now = &st.vo_time_now_;
      in.vo_rel_p = vo_p;
    }
    double x[32], vb[3];
    dekf_outputs out;
    std::memset(&out, 0, sizeof(out));
    out.x = x;
    out.v_body = vb;
    out.status = &status_;
    check(dekf_mhe_step_host(h_, T, &in, &out), "dekf_mhe_step_host");
    st.vo_new_ = false;  // DecentralEst.cpp:891 (a message latched at T == 0 stays pending inside the handle)
    if (T >= 1 || kf_) {
      VectorXd &xo = kf_ ? x_KF_ : x_MHE_;
      for (int r = 0; r < ds_; ++r) xo(r) = x[r];
      Vector3d &vo = kf_ ? v_KF_b_ : v_MHE_b_;
      for (int c = 0; c < 3; ++c) vo(c) = vb[c];
    }
    double R[9], pv[3];
    check(dekf_get_host(h_, DEKF_GET_R_SB, R), "dekf_get_host");
    check(dekf_get_host(h_, DEKF_GET_P_VO, pv), "dekf_get_host");
    for (int r = 0; r < 3; ++r) {
      p_vo_accmulate_(r) = pv[r];
      for (int c = 0; c < 3; ++c) R_sb_(r, c) = R[r * 3 + c];
    }
    if (kf_ && ds_ == 9) {
      std::vector<double> C(81), K((size_t)27 * nl_);
      check(dekf_get_host(h_, DEKF_GET_ARRIVAL_COV, C.data()), "dekf_get_host");
      check(dekf_get_host(h_, DEKF_GET_KF_GAIN, K.data()), "dekf_get_host");
      C_KF_ = MatrixXd::Zero(9, 9);
      K_KF_ = MatrixXd::Zero(9, 3 * nl_);
      for (int r = 0; r < 9; ++r) {
        for (int c = 0; c < 9; ++c) C_KF_(r, c) = C[(size_t)r * 9 + c];
        for (int c = 0; c < 3 * nl_; ++c) K_KF_(r, c) = K[(size_t)r * 3 * nl_ + c];
      }
    }
  }

  std::shared_ptr<robot_store> robot_sub_ptr_;
  std::shared_ptr<robot_params> params_ptr_;
  std::vector<double> rows_a_, rows_lb_, rows_ub_;  // addStateRows()
  dekf_handle *h_ = nullptr;
  int nq_ = 12, nl_ = 4, ds_ = 9;
  bool kf_ = false;
  int32_t status_ = 0;
};

#endif  // MHE_EST_HPP
