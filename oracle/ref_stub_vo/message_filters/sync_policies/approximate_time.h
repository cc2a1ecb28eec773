// TEST INFRASTRUCTURE ONLY: stand-in header written for this repository so that the reference's VO wrapper node
// (visual_odometry/orbslam3_ros2/src/stereo-decentralized/stereo-pub-node.cpp) compiles UNMODIFIED into oracle/_ref/vo_pin
// (recipe: oracle/Makefile vo_pin; driver: oracle/vo_pin_main.cc).  ORB-SLAM3, OpenCV, Sophus, cv_bridge and message_filters are
// absent from the image and out of scope; only the surface that source file touches exists here.
#pragma once
namespace message_filters { namespace sync_policies {
template <typename M0, typename M1>
struct ApproximateTime {
  typedef M0 Msg0;
  typedef M1 Msg1;
  explicit ApproximateTime(int queue) : queue_size(queue) {}
  int queue_size;
};
} }
