// ROS 2 shell of the orientation-filter node ("orien_sub") on top of libdekf_b200.so.
//
// Same node interface as the reference's orien_ekf::orien_ekf (/root/reference/src/orien_est/src/orien_ekf.cpp:8-106):
// parameters init_std / process_std / gravity_meas_std / vo_meas_std / quaternion_init / rate (:13-18), subscriptions
// orb/pos and unitree/imu (:35-40), publisher imu/filter (:41), wall timer of period 1/rate (:43).  The arithmetic of
// timerCallback (:77-89: history push, delayed-VO rewind/replay, gyro predict, gravity correct) runs on the GPU through
// dekf::orien_ekf (include/dekf_b200/orien_ekf.hpp).
#pragma once
#include <chrono>
#include <memory>
#include <string>
#include <vector>

#include <rclcpp/rclcpp.hpp>
#include <geometry_msgs/msg/pose_stamped.hpp>
#include <sensor_msgs/msg/imu.hpp>

#include <dekf_b200/orien_ekf.hpp>

namespace dekf_ros {

class OrienSub : public rclcpp::Node {
 public:
  explicit OrienSub(const std::string &name = "orien_sub") : rclcpp::Node(name) {
    dekf::robot_params p = dekf::robot_params::go1();
    declare_parameter("init_std", std::vector<double>{0.001, 0.001, 0.001, 0.001});
    declare_parameter("process_std", std::vector<double>{0.1, 0.1, 0.1});
    declare_parameter("gravity_meas_std", std::vector<double>{4.0, 4.0, 4.0});
    declare_parameter("vo_meas_std", std::vector<double>{0.0001, 0.0001, 0.0001, 0.0001});
    declare_parameter("quaternion_init", std::vector<double>{1.0, 0.0, 0.0, 0.0});
    declare_parameter("rate", 500);
    p.ekf_init_std_ = get_parameter("init_std").as_double_array();
    p.ekf_process_std_ = get_parameter("process_std").as_double_array();
    p.ekf_gravity_meas_std_ = get_parameter("gravity_meas_std").as_double_array();
    p.ekf_vo_meas_std_ = get_parameter("vo_meas_std").as_double_array();
    p.ekf_quaternion_init_ = get_parameter("quaternion_init").as_double_array();
    p.ekf_rate_ = (int)get_parameter("rate").as_int();
    p.n_instances_ = 1;
    dt_ = 1.0 / static_cast<double>(p.ekf_rate_);
    store_.resize(1, 3 * p.num_legs_, p.num_legs_);
    ekf_ = std::make_unique<dekf::orien_ekf>(p);
    using std::placeholders::_1;
    vo_pose_sub_ = create_subscription<geometry_msgs::msg::PoseStamped>("orb/pos", 10, std::bind(&OrienSub::vo_pose_callback, this, _1));
    imu_sub_ = create_subscription<sensor_msgs::msg::Imu>("unitree/imu", 10, std::bind(&OrienSub::imu_callback, this, _1));
    publisher_filter_ = create_publisher<sensor_msgs::msg::Imu>("imu/filter", 10);
    timer_ = create_wall_timer(std::chrono::microseconds(int(dt_ * 1e6)), std::bind(&OrienSub::timerCallback, this));
    time_init_ = static_cast<double>(rclcpp::Clock().now().nanoseconds()) / 1e9;
  }

  void vo_pose_callback(const geometry_msgs::msg::PoseStamped::SharedPtr msg) {
    store_.vo_time_now_[0] = static_cast<double>(msg->header.stamp.sec) + static_cast<double>(msg->header.stamp.nanosec) / 1e9 - time_init_;
    store_.vo_quaternion_[0] = msg->pose.orientation.w;
    store_.vo_quaternion_[1] = msg->pose.orientation.x;
    store_.vo_quaternion_[2] = msg->pose.orientation.y;
    store_.vo_quaternion_[3] = msg->pose.orientation.z;
    store_.vo_new_[0] = 1;
  }
  void imu_callback(const sensor_msgs::msg::Imu::SharedPtr msg) {
    store_.imu_time_[0] = static_cast<double>(rclcpp::Clock().now().nanoseconds()) / 1e9 - time_init_;
    store_.accel_b_[0] = msg->linear_acceleration.x;
    store_.accel_b_[1] = msg->linear_acceleration.y;
    store_.accel_b_[2] = msg->linear_acceleration.z;
    store_.angular_b_[0] = msg->angular_velocity.x;
    store_.angular_b_[1] = msg->angular_velocity.y;
    store_.angular_b_[2] = msg->angular_velocity.z;
    init_imu_ = true;
  }
  // orien_ekf.cpp:77-106: one filter tick once an IMU message has arrived, then publish orientation + raw IMU
  void timerCallback() {
    if (init_imu_) {
      ekf_->timerCallback(store_);
      store_.vo_new_[0] = 0;  // the reference clears its flag inside get_measurement (orien_ekf.cpp:169)
      // the device keeps a ring of ekf_hist_depth ticks where the reference keeps unbounded stacks: a VO pose older than the
      // ring cannot be rewound to and is dropped with this bit set -- make that visible (raise ekf_hist_depth)
      if (ekf_->status_[0] & DEKF_ST_EKF_HIST_OVERFLOW) {
        if (hist_overflows_++ % 500 == 0)
          std::fprintf(stderr, "orien_sub: VO pose older than the EKF history ring (%d ticks); dropped (%ld so far)\n",
                       ekf_->hist_depth(), hist_overflows_);
      }
    }
    sensor_msgs::msg::Imu out;
    out.orientation.w = ekf_->quaternion_[0];
    out.orientation.x = ekf_->quaternion_[1];
    out.orientation.y = ekf_->quaternion_[2];
    out.orientation.z = ekf_->quaternion_[3];
    out.linear_acceleration.x = store_.accel_b_[0];
    out.linear_acceleration.y = store_.accel_b_[1];
    out.linear_acceleration.z = store_.accel_b_[2];
    out.angular_velocity.x = store_.angular_b_[0];
    out.angular_velocity.y = store_.angular_b_[1];
    out.angular_velocity.z = store_.angular_b_[2];
    publisher_filter_->publish(out);
  }

  const std::vector<double> &quaternion() const { return ekf_->quaternion_; }
  dekf::orien_ekf &filter() { return *ekf_; }

 private:
  std::unique_ptr<dekf::orien_ekf> ekf_;
  dekf::robot_store store_;
  rclcpp::Subscription<sensor_msgs::msg::Imu>::SharedPtr imu_sub_;
  rclcpp::Subscription<geometry_msgs::msg::PoseStamped>::SharedPtr vo_pose_sub_;
  rclcpp::Publisher<sensor_msgs::msg::Imu>::SharedPtr publisher_filter_;
  rclcpp::TimerBase::SharedPtr timer_;
  double dt_ = 0.002, time_init_ = 0.0;
  bool init_imu_ = false;
  long hist_overflows_ = 0;
};

}  // namespace dekf_ros
