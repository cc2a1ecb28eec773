// TEST INFRASTRUCTURE ONLY: stand-in header written for this repository so that the reference's VO wrapper node
// (visual_odometry/orbslam3_ros2/src/stereo-decentralized/stereo-pub-node.cpp) compiles UNMODIFIED into oracle/_ref/vo_pin
// (recipe: oracle/Makefile vo_pin; driver: oracle/vo_pin_main.cc).  ORB-SLAM3, OpenCV, Sophus, cv_bridge and message_filters are
// absent from the image and out of scope; only the surface that source file touches exists here.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include "std_msgs/msg/header.hpp"
namespace sensor_msgs { namespace msg {
struct Image {
  typedef std::shared_ptr<Image> SharedPtr;
  typedef std::shared_ptr<const Image> ConstSharedPtr;
  std_msgs::msg::Header header;
  uint32_t height = 0, width = 0, step = 0;
  std::string encoding;
  uint8_t is_bigendian = 0;
  std::vector<uint8_t> data;
};
} }
