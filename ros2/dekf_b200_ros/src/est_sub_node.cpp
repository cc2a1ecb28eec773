// `ros2 run dekf_b200_ros est_sub`: drop-in for the reference's `go1_example est_sub` executable (go1Sub.cpp:142-149).
#include "dekf_b200_ros/est_sub.hpp"

int main(int argc, char **argv) {
  rclcpp::init(argc, argv);
  rclcpp::spin(std::make_shared<dekf_ros::EstSub>("est_sub"));
  rclcpp::shutdown();
  return 0;
}
