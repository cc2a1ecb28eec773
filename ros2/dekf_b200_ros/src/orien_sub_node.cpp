// `ros2 run dekf_b200_ros orien_sub`: drop-in for the reference's `orien_est orien_sub` executable (orien_ekf.cpp:360-366).
#include "dekf_b200_ros/orien_sub.hpp"

int main(int argc, char **argv) {
  rclcpp::init(argc, argv);
  rclcpp::spin(std::make_shared<dekf_ros::OrienSub>("orien_sub"));
  rclcpp::shutdown();
  return 0;
}
