// C++ facade of the orientation EKF node's arithmetic with the reference's class name (header-only, C++17).
//
//   reference: class orien_ekf::orien_ekf  /root/reference/src/orien_est/include/orien_ekf.hpp:21-95
//              timerCallback               /root/reference/src/orien_est/src/orien_ekf.cpp:77-89
//                = get_measurement (:156, history push + delayed-VO rewind/replay) -> gyro_nonlinear_predict (:108)
//                  -> gyro_nonlinear_correct (:125) [-> vo_nonlinear_correct (:144) inside the replay]
//   here:      dekf::orien_ekf::timerCallback(robot_store&) steps the EKF of every instance once.
//
// The ROS plumbing of the reference class (parameter declaration, subscriptions, publisher, wall timer,
// orien_ekf.cpp:8-75,90-105) stays in the unchanged host-side node; it fills robot_store and calls timerCallback().
#pragma once
#include "DecentralEst.hpp"

namespace dekf {

class orien_ekf {
 public:
  explicit orien_ekf(const robot_params &params) {
    const dekf_config cfg = params.to_config();
    detail::check(dekf_create(&cfg, &h_), nullptr, "dekf_create");
    n_ = cfg.n_instances;
    hist_depth_ = cfg.ekf_hist_depth;
    quaternion_.assign(4 * (size_t)n_, 0.0);
    for (int i = 0; i < n_; ++i) quaternion_[i] = cfg.ekf_quaternion_init[0];
    Cov_q_.assign(16 * (size_t)n_, 0.0);
    status_.assign(n_, 0);
  }
  orien_ekf(const orien_ekf &) = delete;
  orien_ekf &operator=(const orien_ekf &) = delete;
  ~orien_ekf() {
    if (h_) dekf_destroy(h_);
  }

  // orien_ekf.cpp:77-89.  Reads angular_b_, accel_b_, imu_time_ and -- when any vo_new_ flag is set -- vo_quaternion_ and
  // vo_time_now_ (the orb/pos message, orien_ekf.cpp:47-58).  Does NOT clear vo_new_: the MHE consumes the same flag.
  void timerCallback(const robot_store &st) {
    dekf_inputs in;
    std::memset(&in, 0, sizeof(in));
    in.gyro = st.angular_b_.data();
    in.accel = st.accel_b_.data();
    in.imu_time = st.imu_time_.data();
    if (st.any_vo()) {
      in.vo_flag = st.vo_new_.data();
      in.vo_quat = st.vo_quaternion_.data();
      in.vo_time_now = st.vo_time_now_.data();
    }
    dekf_outputs out;
    std::memset(&out, 0, sizeof(out));
    out.quat = quaternion_.data();
    out.status = status_.data();
    detail::check(dekf_ekf_step_host(h_, &in, &out), h_, "dekf_ekf_step_host");
    discrete_time_++;
  }

  // orien_ekf.hpp:59-64
  std::vector<double> quaternion_;  // [4][n] w,x,y,z  (published on imu/filter, orien_ekf.cpp:92-95)
  const std::vector<double> &Cov_q() {
    detail::check(dekf_get_host(h_, DEKF_GET_EKF_COV, Cov_q_.data()), h_, "dekf_get_host");
    return Cov_q_;
  }
  std::vector<int32_t> status_;  // [n] DEKF_ST_EKF_* bits of the last tick
  int discrete_time_ = 0;
  // depth of the device's history ring in ticks (the reference's stacks are unbounded, orien_ekf.cpp:158-163): a VO pose older
  // than this is dropped with DEKF_ST_EKF_HIST_OVERFLOW in status_
  int hist_depth() const { return hist_depth_; }

 private:
  dekf_handle *h_ = nullptr;
  int n_ = 0, hist_depth_ = 0;
  std::vector<double> Cov_q_;  // [16][n]
};

}  // namespace dekf
