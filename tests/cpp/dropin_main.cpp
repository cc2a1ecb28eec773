// Drop-in proof driver: the reference's OWN est_sub node -- robotSub::robotSub (decentral_legged_est/src/EstSub.cpp) and
// robotSub::go1Sub (go1_example/src/go1Sub.cpp) with the FROST kinematics it calls -- compiled UNMODIFIED from
// /root/reference against include/dekf_b200/dropin/decentral_legged_est/DecentralEst.hpp (instead of the reference's
// DecentralEst.hpp) and linked to libdekf_b200.so: mhe.initialize()/mhe.update(T) inside the reference's timerCallback now run
// on the GPU.  The orientation comes from dekf_ros::OrienSub (the orien_sub shell over the same library).  rclcpp / Eigen are
// the stand-ins of oracle/ref_stub (neither is in the image).  Built by `make -C oracle dropin` into oracle/_ref/ (git-ignored,
// travels to the GPU box); tests/test_dropin_reference_nodes.py plays the golden streams through it.
//   dropin_nodes <stream.bin> <out.bin>      (formats: tests/cpp/ros_shell_main.cpp)
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <vector>

#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>

#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <Eigen/Sparse>
#include <rclcpp/rclcpp.hpp>
#include <sensor_msgs/msg/imu.hpp>
#include <sensor_msgs/msg/joint_state.hpp>
#include <geometry_msgs/msg/pose_stamped.hpp>
#include <optitrack_broadcast/msg/mocap.hpp>
#include <custom_msgs/msg/vo_realtive_transform.hpp>

// read-only access to the reference node's private `mhe` member from this driver (layout is unaffected by access specifiers)
#define private public
#define protected public
#include "go1Sub.hpp"  // the reference's header; pulls decentral_legged_est/EstSub.hpp -> OUR decentral_legged_est/DecentralEst.hpp
#undef private
#undef protected
#include "dekf_b200_ros/orien_sub.hpp"

static void set_doubles(const std::string &key, const double *v, int n) {
  refstub::ParamValue p; p.kind = 5; p.v.assign(v, v + n); refstub::param_overrides()[key] = p;
}
static void set_int(const std::string &key, long long v) { refstub::ParamValue p; p.kind = 1; p.i = v; refstub::param_overrides()[key] = p; }
static void set_double(const std::string &key, double v) { refstub::ParamValue p; p.kind = 2; p.d = v; refstub::param_overrides()[key] = p; }
static void set_bool(const std::string &key, bool v) { refstub::ParamValue p; p.kind = 3; p.b = v; refstub::param_overrides()[key] = p; }
static void set_string(const std::string &key, const std::string &v) { refstub::ParamValue p; p.kind = 4; p.s = v; refstub::param_overrides()[key] = p; }

// the role of go1_example/config/parameters_go1.yaml in the launch file
static void load_go1_yaml(int N, int est_type, int leg_odom_type, int rate) {
  dekf_config c;
  dekf_config_default_go1(&c);
  const std::string e = "est_sub.";
  set_string(e + "log_name", "dropin");
  set_doubles(e + "prior.p_init_std", c.p_init_std, 3);
  set_doubles(e + "prior.v_init_std", c.v_init_std, 3);
  set_doubles(e + "prior.foot_init_std", c.foot_init_std, 3);
  set_doubles(e + "prior.accel_bias_init_std", c.accel_bias_init_std, 3);
  set_doubles(e + "process.p_process_std", c.p_process_std, 3);
  set_doubles(e + "process.accel_input_std", c.accel_input_std, 3);
  set_doubles(e + "process.gyro_input_std", c.gyro_input_std, 3);
  set_doubles(e + "process.accel_bias_process_std", c.accel_bias_std, 3);
  set_doubles(e + "leg_odom.quaternion_ib", c.quaternion_ib, 4);
  set_doubles(e + "leg_odom.p_ib", c.p_ib, 3);
  set_int(e + "leg_odom.num_leg", c.num_legs);
  set_int(e + "leg_odom.leg_odom_type", leg_odom_type);
  set_doubles(e + "leg_odom.joint_position_std", c.joint_position_std, 3);
  set_doubles(e + "leg_odom.joint_velocity_std", c.joint_velocity_std, 3);
  set_doubles(e + "leg_odom.foot_slide_std", c.foot_slide_std, 3);
  set_doubles(e + "leg_odom.foot_swing_std", c.foot_swing_std, 3);
  set_double(e + "leg_odom.contact_effort_theshold", c.contact_effort_threshold);
  set_doubles(e + "visual_odom.vo_p_std", c.vo_p_std, 3);
  set_int(e + "estimation.rate", rate);
  set_int(e + "estimation.interval", 1000 / rate);
  set_int(e + "estimation.N", N);
  set_int(e + "estimation.est_type", est_type);
  set_bool(e + "osqp.verbose", false);
  const std::string o = "orien_sub.";
  set_doubles(o + "init_std", c.ekf_init_std, 4);
  set_doubles(o + "process_std", c.ekf_process_std, 3);
  set_doubles(o + "gravity_meas_std", c.ekf_gravity_meas_std, 3);
  set_doubles(o + "vo_meas_std", c.ekf_vo_meas_std, 4);
  set_doubles(o + "quaternion_init", c.ekf_quaternion_init, 4);
  set_int(o + "rate", rate);
}

static void stamp(long long ns, builtin_interfaces::msg::Time &t) {
  t.sec = (int32_t)(ns / 1000000000LL);
  t.nanosec = (uint32_t)(ns % 1000000000LL);
}

int main(int argc, char **argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s stream.bin out.bin\n", argv[0]); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  int32_t hdr[8];
  f.read((char *)hdr, sizeof(hdr));
  const int S = hdr[0], nq = hdr[1], nl = hdr[2], N = hdr[3], est_type = hdr[4], leg_odom_type = hdr[5], rate = hdr[6];
  load_go1_yaml(N, est_type, leg_odom_type, rate);
  refstub::now_ns() = 0;
  std::streambuf *cout_buf = std::cout.rdbuf();
  std::ofstream devnull("/dev/null");
  try {
    auto ekf = std::make_shared<dekf_ros::OrienSub>("orien_sub");
    auto est = std::make_shared<robotSub::go1Sub>("est_sub");  // the reference's class
    est->imu_msg_num_ = 10;  // start gate of the reference's timerCallback (EstSub.cpp:62)
    std::ofstream out(argv[2], std::ios::binary);
    const int nd = 3 + 3 + nq + nq + nl + 4 + 3;
    std::vector<double> d((size_t)nd);
    for (int s = 0; s < S; ++s) {
      long long h[4];
      f.read((char *)h, sizeof(h));
      f.read((char *)d.data(), (std::streamsize)(sizeof(double) * nd));
      const double *gyro = d.data(), *accel = gyro + 3, *jp = accel + 3, *jv = jp + nq, *ff = jv + nq, *vq = ff + nl, *vp = vq + 4;
      refstub::now_ns() = h[0];
      sensor_msgs::msg::Imu imu;
      stamp(h[0], imu.header.stamp);
      imu.angular_velocity.x = gyro[0]; imu.angular_velocity.y = gyro[1]; imu.angular_velocity.z = gyro[2];
      imu.linear_acceleration.x = accel[0]; imu.linear_acceleration.y = accel[1]; imu.linear_acceleration.z = accel[2];
      refstub::deliver<sensor_msgs::msg::Imu>("unitree/imu", imu);
      sensor_msgs::msg::JointState js;
      js.position.assign(jp, jp + nq);
      js.position.insert(js.position.end(), ff, ff + nl);
      js.velocity.assign(jv, jv + nq);
      refstub::deliver<sensor_msgs::msg::JointState>("unitree/joint_state", js);
      if (h[1]) {
        geometry_msgs::msg::PoseStamped ps;
        stamp(h[3], ps.header.stamp);
        ps.pose.orientation.w = vq[0]; ps.pose.orientation.x = vq[1]; ps.pose.orientation.y = vq[2]; ps.pose.orientation.z = vq[3];
        refstub::deliver<geometry_msgs::msg::PoseStamped>("orb/pos", ps);
        custom_msgs::msg::VoRealtiveTransform vt;
        stamp(h[3], vt.header.stamp);
        stamp(h[2], vt.header_pre.stamp);
        vt.x_relative = vp[0]; vt.y_relative = vp[1]; vt.z_relative = vp[2];
        refstub::deliver<custom_msgs::msg::VoRealtiveTransform>("orb/vo", vt);
      }
      ekf->fire_timers();  // publishes imu/filter -> robotSub::orien_filter_callback (reference code)
      std::cout.rdbuf(devnull.rdbuf());  // the reference's timerCallback prints its loop rate every tick (EstSub.cpp:90)
      est->fire_timers();  // robotSub::timerCallback (reference code) -> mhe.initialize / mhe.update -> libdekf_b200.so
      std::cout.rdbuf(cout_buf);
      const bool kf = est_type == 1;
      const VectorXd &x = kf ? est->mhe.x_KF_ : est->mhe.x_MHE_;
      const Vector3d &vb = kf ? est->mhe.v_KF_b_ : est->mhe.v_MHE_b_;
      out.write((const char *)ekf->quaternion().data(), 4 * sizeof(double));
      for (int r = 0; r < (int)x.size(); ++r) { double v = x(r); out.write((const char *)&v, sizeof(double)); }
      for (int r = 0; r < 3; ++r) { double v = vb(r); out.write((const char *)&v, sizeof(double)); }
      for (int r = 0; r < 3; ++r) { double v = est->mhe.p_vo_accmulate_(r); out.write((const char *)&v, sizeof(double)); }
      for (int l = 0; l < nl; ++l) { double c = est->robot_store_->contact_(l); out.write((const char *)&c, sizeof(double)); }
    }
  } catch (const std::exception &e) {
    std::cout.rdbuf(cout_buf);
    std::fprintf(stderr, "dropin_nodes: %s\n", e.what());
    return 1;
  }
  return 0;
}
