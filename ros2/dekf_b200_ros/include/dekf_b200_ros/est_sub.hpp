// ROS 2 shell of the estimator node ("est_sub") on top of libdekf_b200.so.
//
// Replaces, for one robot, the reference's robotSub + go1Sub nodes
// (/root/reference/src/decentral_legged_est/src/EstSub.cpp:7-121, src/go1_example/src/go1Sub.cpp:8-140) with the SAME
// node interface -- parameter names and defaults of paramsWrapper (EstSub.cpp:123-208 = parameters_go1.yaml), topics
// (/unitree/imu, /unitree/joint_state, imu/filter, orb/vo, /mocap/RigidBody), timer period `estimation.interval`, the
// 10-IMU-message start gate, Data_Logger-compatible log files -- while the arithmetic of
// DecentralizedEstimation::initialize/update runs on the GPU through dekf::DecentralizedEstimation
// (include/dekf_b200/DecentralEst.hpp).  What the reference node computes on the host in go1Sub::lo_callback (FROST
// forward kinematics, Jacobians, contact flags) is part of the device step here: the callback only stores the raw message.
//
// Builds against real rclcpp on a robot (ros2/dekf_b200_ros/CMakeLists.txt) and, for the tests of this repository,
// against the stand-in rclcpp of oracle/ref_stub (tests/test_ros_shells.py).
#pragma once
#include <chrono>
#include <cmath>
#include <memory>
#include <string>
#include <vector>

#include <rclcpp/rclcpp.hpp>
#include <sensor_msgs/msg/imu.hpp>
#include <sensor_msgs/msg/joint_state.hpp>
#include <optitrack_broadcast/msg/mocap.hpp>
#include <custom_msgs/msg/vo_realtive_transform.hpp>

#include <dekf_b200/DecentralEst.hpp>
#include "dekf_b200_ros/data_logger.hpp"

namespace dekf_ros {

inline void quaternion_to_euler(double w, double x, double y, double z, double rpy[3]) {
  rpy[0] = std::atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y));
  const double sp = 2.0 * (w * y - z * x);
  rpy[1] = std::fabs(sp) >= 1.0 ? std::copysign(M_PI / 2.0, sp) : std::asin(sp);
  rpy[2] = std::atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z));
}

class EstSub : public rclcpp::Node {
 public:
  explicit EstSub(const std::string &name = "est_sub") : rclcpp::Node(name) {
    robot_store_ = std::make_shared<dekf::robot_store>();
    robot_params_ = std::make_shared<dekf::robot_params>(dekf::robot_params::go1());
    load_parameters();
    robot_store_->resize(1, 3 * robot_params_->num_legs_, robot_params_->num_legs_);
    mhe.prepare(robot_params_);  // the device handle exists before the first timer tick (the reference allocates in initialize())
    using std::placeholders::_1;
    imu_sub_ = create_subscription<sensor_msgs::msg::Imu>("/unitree/imu", 10, std::bind(&EstSub::imu_callback, this, _1));
    joint_sub_ = create_subscription<sensor_msgs::msg::JointState>("/unitree/joint_state", 10, std::bind(&EstSub::lo_callback, this, _1));
    mocap_sub_ = create_subscription<optitrack_broadcast::msg::Mocap>("/mocap/RigidBody", 10, std::bind(&EstSub::mocap_callback, this, _1));
    filter_sub_ = create_subscription<sensor_msgs::msg::Imu>("imu/filter", 10, std::bind(&EstSub::orien_filter_callback, this, _1));
    vo_sub_ = create_subscription<custom_msgs::msg::VoRealtiveTransform>("orb/vo", 10, std::bind(&EstSub::vo_callback, this, _1));
    timer_ = create_wall_timer(std::chrono::milliseconds(interval_ms_), std::bind(&EstSub::timerCallback, this));
    time_init_ = static_cast<double>(rclcpp::Clock().now().nanoseconds()) / 1e9;
  }

  // ---- callbacks (same roles as go1Sub::imu_callback/lo_callback/mocap_callback, robotSub::orien_filter_callback/vo_callback)
  void imu_callback(const sensor_msgs::msg::Imu::SharedPtr msg) {
    dekf::robot_store &st = *robot_store_;
    st.imu_time_[0] = static_cast<double>(rclcpp::Clock().now().nanoseconds()) / 1e9 - time_init_;
    st.accel_b_[0] = msg->linear_acceleration.x;
    st.accel_b_[1] = msg->linear_acceleration.y;
    st.accel_b_[2] = msg->linear_acceleration.z;
    st.angular_b_[0] = msg->angular_velocity.x;
    st.angular_b_[1] = msg->angular_velocity.y;
    st.angular_b_[2] = msg->angular_velocity.z;
    imu_msg_num_++;
  }
  // raw joint message: positions of the 3*num_legs joints followed by the foot forces (go1Sub.cpp:68-75); kinematics and
  // contact detection happen inside the device step
  void lo_callback(const sensor_msgs::msg::JointState::SharedPtr msg) {
    dekf::robot_store &st = *robot_store_;
    const size_t np = st.joint_states_position_.size(), nv = st.joint_states_velocity_.size();
    for (size_t i = 0; i < np && i < msg->position.size(); ++i) st.joint_states_position_[i] = msg->position[i];
    for (size_t i = 0; i < nv && i < msg->velocity.size(); ++i) st.joint_states_velocity_[i] = msg->velocity[i];
  }
  void mocap_callback(const optitrack_broadcast::msg::Mocap::SharedPtr msg) {
    for (int i = 0; i < 3; ++i) { gt_p_raw_[i] = msg->position[i]; gt_v_s_[i] = msg->velocity[i]; }
    const double w = msg->quaternion[0], x = msg->quaternion[1], y = msg->quaternion[2], z = msg->quaternion[3];
    quaternion_to_euler(w, x, y, z, gt_euler_.data());
    const double nrm = std::sqrt(w * w + x * x + y * y + z * z);
    const double qw = w / nrm, qx = x / nrm, qy = y / nrm, qz = z / nrm;
    const double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy),
                         2 * (qx * qy + qw * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qw * qx),
                         2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), 1 - 2 * (qx * qx + qy * qy)};
    for (int i = 0; i < 3; ++i) {
      gt_p_[i] = gt_p_raw_[i] - gt_p_offset_[i];
      gt_v_b_[i] = R[3 * i] * gt_v_s_[0] + R[3 * i + 1] * gt_v_s_[1] + R[3 * i + 2] * gt_v_s_[2];
    }
  }
  void orien_filter_callback(const sensor_msgs::msg::Imu::SharedPtr msg) {
    dekf::robot_store &st = *robot_store_;
    st.quaternion_[0] = msg->orientation.w;
    st.quaternion_[1] = msg->orientation.x;
    st.quaternion_[2] = msg->orientation.y;
    st.quaternion_[3] = msg->orientation.z;
    quaternion_to_euler(st.quaternion_[0], st.quaternion_[1], st.quaternion_[2], st.quaternion_[3], filter_euler_.data());
  }
  void vo_callback(const custom_msgs::msg::VoRealtiveTransform::SharedPtr msg) {
    dekf::robot_store &st = *robot_store_;
    st.vo_new_[0] = 1;
    st.vo_time_pre_[0] = static_cast<double>(msg->header_pre.stamp.sec) + static_cast<double>(msg->header_pre.stamp.nanosec) / 1e9 - time_init_;
    st.vo_time_now_[0] = static_cast<double>(msg->header.stamp.sec) + static_cast<double>(msg->header.stamp.nanosec) / 1e9 - time_init_;
    st.vo_p_body_pre_2_body_[0] = msg->x_relative;
    st.vo_p_body_pre_2_body_[1] = msg->y_relative;
    st.vo_p_body_pre_2_body_[2] = msg->z_relative;
  }

  // ---- the estimator timer (EstSub.cpp:58-91): initialize at discrete time 0, update(T) afterwards, log from N+2 on
  void timerCallback() {
    if (imu_msg_num_ < 10) return;
    if (discrete_time_ == 0) {
      mhe.initialize(robot_store_, robot_params_);
      gt_p_offset_ = gt_p_raw_;
    } else {
      mhe.update(discrete_time_);
    }
    discrete_time_++;
    if (discrete_time_ == robot_params_->N_ + 1) init_logging();
    if (discrete_time_ > robot_params_->N_ + 1 && logger_.is_open()) logger_.spin_logging();
  }

  // ---- public state, named like the reference node's (EstSub.hpp:61-72)
  dekf::DecentralizedEstimation mhe;
  std::shared_ptr<dekf::robot_store> robot_store_;
  std::shared_ptr<dekf::robot_params> robot_params_;
  double time_init_ = 0.0;
  int imu_msg_num_ = 0;
  int discrete_time_ = 0;
  std::vector<double> gt_p_offset_{0, 0, 0}, gt_p_{0, 0, 0}, gt_v_b_{0, 0, 0}, gt_euler_{0, 0, 0}, filter_euler_{0, 0, 0};

 private:
  std::vector<double> dparam(const std::string &name, std::vector<double> def) {
    declare_parameter(name, def);
    return get_parameter(name).as_double_array();
  }
  double fparam(const std::string &name, double def) {
    declare_parameter(name, def);
    return get_parameter(name).as_double();
  }
  int iparam(const std::string &name, int def) {
    declare_parameter(name, def);
    return (int)get_parameter(name).as_int();
  }
  bool bparam(const std::string &name, bool def) {
    declare_parameter(name, def);
    return get_parameter(name).as_bool();
  }
  // parameter names and defaults: EstSub.cpp:125-207
  void load_parameters() {
    dekf::robot_params &p = *robot_params_;
    declare_parameter<std::string>("log_name", "exp");
    log_name_ = get_parameter("log_name").as_string();
    p.p_init_std_ = dparam("prior.p_init_std", {0.001, 0.001, 0.001});
    p.v_init_std_ = dparam("prior.v_init_std", {0.001, 0.001, 0.001});
    p.foot_init_std_ = dparam("prior.foot_init_std", {0.001, 0.001, 0.001});
    p.accel_bias_init_std_ = dparam("prior.accel_bias_init_std", {0.001, 0.001, 0.001});
    p.p_process_std_ = dparam("process.p_process_std", {0.01, 0.01, 0.01});
    p.accel_input_std_ = dparam("process.accel_input_std", {0.01, 0.04, 0.001});
    p.gyro_input_std_ = dparam("process.gyro_input_std", {0.01, 0.01, 0.01});
    p.accel_bias_std_ = dparam("process.accel_bias_process_std", {1., 1., 0.1});
    p.quaternion_ib_ = dparam("leg_odom.quaternion_ib", {1.0, 0.0, 0.0, 0.0});
    p.p_ib_ = dparam("leg_odom.p_ib", {0.0, 0.0, 0.0});
    p.num_legs_ = iparam("leg_odom.num_leg", 4);
    p.leg_odom_type_ = iparam("leg_odom.leg_odom_type", 0);
    p.joint_position_std_ = dparam("leg_odom.joint_position_std", {0.01, 0.01, 0.01});
    p.joint_velocity_std_ = dparam("leg_odom.joint_velocity_std", {0.01, 0.01, 0.01});
    p.foot_slide_std_ = dparam("leg_odom.foot_slide_std", {0.001, 0.001, 0.001});
    p.foot_swing_std_ = dparam("leg_odom.foot_swing_std", {10000.0, 10000.0, 10000.0});
    p.contact_effort_theshold_ = fparam("leg_odom.contact_effort_theshold", 150.0);
    p.vo_p_std_ = dparam("visual_odom.vo_p_std", {0.001, 0.001, 0.001});
    p.rate_ = iparam("estimation.rate", 50);
    interval_ms_ = iparam("estimation.interval", 20);
    p.N_ = iparam("estimation.N", 50);
    p.est_type_ = iparam("estimation.est_type", 0);
    p.rho_ = fparam("osqp.rho", 0.1);
    p.alpha_ = fparam("osqp.alpha", 1.6);
    p.delta_ = fparam("osqp.delta", 0.00001);
    p.sigma_ = fparam("osqp.sigma", 0.00001);
    p.verbose_ = bparam("osqp.verbose", true);
    p.adaptRho_ = bparam("osqp.adaptRho", true);
    p.polish_ = bparam("osqp.polish", true);
    p.maxQPIter_ = iparam("osqp.maxQPIter", 1000);
    p.primTol_ = fparam("osqp.primTol", 0.000001);
    p.dualTol_ = fparam("osqp.dualTol", 0.000001);
    p.realtiveTol_ = fparam("osqp.realtiveTol", 1e-3);
    p.absTol_ = fparam("osqp.absTol", 1e-3);
    p.timeLimit_ = fparam("osqp.timeLimit", 0.005);
    p.n_instances_ = 1;
    p.ekf_rate_ = p.rate_;  // the orientation EKF runs in its own node; this handle only steps the MHE / KF
  }
  // channel list of robotSub::init_logging (EstSub.cpp:93-121)
  void init_logging() {
    logger_.init(log_name_);
    if (!logger_.is_open()) return;
    const bool kf = robot_params_->est_type_ == 1;
    logger_.add_vector(gt_p_.data(), 3, "pose");
    logger_.add_vector(gt_v_b_.data(), 3, "GT_v");
    const std::vector<double> &vb = kf ? mhe.v_KF_b_ : mhe.v_MHE_b_;
    const std::vector<double> &x = kf ? mhe.x_KF_ : mhe.x_MHE_;
    logger_.add_vector(vb.data(), 3, "v_body");
    logger_.add_vector(x.data(), (unsigned)x.size(), "x_MHE");
    logger_.add_vector(mhe.p_vo_accmulate_.data(), 3, "p_vo_accmulate_");
    logger_.add_vector(filter_euler_.data(), 3, "filter_euler_");
    logger_.add_vector(gt_euler_.data(), 3, "gt_euler_");
  }

  rclcpp::Subscription<sensor_msgs::msg::Imu>::SharedPtr imu_sub_, filter_sub_;
  rclcpp::Subscription<sensor_msgs::msg::JointState>::SharedPtr joint_sub_;
  rclcpp::Subscription<optitrack_broadcast::msg::Mocap>::SharedPtr mocap_sub_;
  rclcpp::Subscription<custom_msgs::msg::VoRealtiveTransform>::SharedPtr vo_sub_;
  rclcpp::TimerBase::SharedPtr timer_;
  DataLogger logger_;
  std::string log_name_;
  int interval_ms_ = 20;
  std::vector<double> gt_p_raw_{0, 0, 0}, gt_v_s_{0, 0, 0};
};

}  // namespace dekf_ros
