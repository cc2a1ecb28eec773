// Stand-in for the generated header of the reference's src/communication/custom_msgs/msg/VoRealtiveTransform.msg.
// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference's sources).
#pragma once
#include "std_msgs/msg/header.hpp"
namespace custom_msgs { namespace msg {
struct VoRealtiveTransform {
  typedef std::shared_ptr<VoRealtiveTransform> SharedPtr;
  std_msgs::msg::Header header, header_pre;
  double x_relative = 0, y_relative = 0, z_relative = 0;
};
} }
