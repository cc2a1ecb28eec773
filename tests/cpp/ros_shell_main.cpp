// Test driver of the ROS 2 node shells (ros2/dekf_b200_ros): instantiates dekf_ros::OrienSub and dekf_ros::EstSub against
// the stand-in rclcpp of oracle/ref_stub (test infrastructure; a robot builds them against real rclcpp), plays one
// instance of a recorded stream through their subscriptions in lock-step -- exactly how oracle/ref_nodes.cc drives the
// reference's own nodes -- and dumps per tick: quaternion(4) x(ds) v_body(3) p_vo(3) contact(nl) as doubles.
//   ros_shell_main <stream.bin> <out.bin>
// stream.bin: int32 header [S, nq, nl, N, est_type, leg_odom_type, rate, 0]; per tick int64 [imu_ns, vo_flag, vo_pre_ns,
// vo_now_ns] then doubles gyro(3) accel(3) joint_pos(nq) joint_vel(nq) foot_force(nl) vo_quat(4) vo_rel_p(3).
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <vector>

#include "dekf_b200_ros/est_sub.hpp"
#include "dekf_b200_ros/orien_sub.hpp"

static void set_doubles(const std::string &key, const double *v, int n) {
  refstub::ParamValue p; p.kind = 5; p.v.assign(v, v + n); refstub::param_overrides()[key] = p;
}
static void set_int(const std::string &key, long long v) { refstub::ParamValue p; p.kind = 1; p.i = v; refstub::param_overrides()[key] = p; }
static void set_double(const std::string &key, double v) { refstub::ParamValue p; p.kind = 2; p.d = v; refstub::param_overrides()[key] = p; }
static void set_bool(const std::string &key, bool v) { refstub::ParamValue p; p.kind = 3; p.b = v; refstub::param_overrides()[key] = p; }
static void set_string(const std::string &key, const std::string &v) { refstub::ParamValue p; p.kind = 4; p.s = v; refstub::param_overrides()[key] = p; }

// the role of parameters_go1.yaml in the launch file: values = dekf_config_default_go1 (the same numbers)
static void load_go1_yaml(int N, int est_type, int leg_odom_type, int rate) {
  dekf_config c;
  dekf_config_default_go1(&c);
  const std::string e = "est_sub.";
  set_string(e + "log_name", "shell");
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
  try {
    auto ekf = std::make_shared<dekf_ros::OrienSub>("orien_sub");
    auto est = std::make_shared<dekf_ros::EstSub>("est_sub");
    est->imu_msg_num_ = 10;  // start gate of the estimator timer (10 IMU messages)
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
      ekf->fire_timers();  // publishes imu/filter -> EstSub::orien_filter_callback
      est->fire_timers();
      const bool kf = est_type == 1;
      const std::vector<double> &x = kf ? est->mhe.x_KF_ : est->mhe.x_MHE_;
      const std::vector<double> &vb = kf ? est->mhe.v_KF_b_ : est->mhe.v_MHE_b_;
      out.write((const char *)ekf->quaternion().data(), 4 * sizeof(double));
      out.write((const char *)x.data(), (std::streamsize)(x.size() * sizeof(double)));
      out.write((const char *)vb.data(), 3 * sizeof(double));
      out.write((const char *)est->mhe.p_vo_accmulate_.data(), 3 * sizeof(double));
      for (int l = 0; l < nl; ++l) { double c = est->robot_store_->contact_[(size_t)l]; out.write((const char *)&c, sizeof(double)); }
    }
  } catch (const std::exception &e) {
    std::fprintf(stderr, "ros_shell_main: %s\n", e.what());
    return 1;
  }
  return 0;
}
