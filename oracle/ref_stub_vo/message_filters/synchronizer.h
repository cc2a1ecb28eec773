// TEST INFRASTRUCTURE ONLY: stand-in header written for this repository so that the reference's VO wrapper node
// (visual_odometry/orbslam3_ros2/src/stereo-decentralized/stereo-pub-node.cpp) compiles UNMODIFIED into oracle/_ref/vo_pin
// (recipe: oracle/Makefile vo_pin; driver: oracle/vo_pin_main.cc).  ORB-SLAM3, OpenCV, Sophus, cv_bridge and message_filters are
// absent from the image and out of scope; only the surface that source file touches exists here.
//
// The driver plays the role of the approximate-time matcher: it calls refstub_vo::stereo_callbacks().back()(left, right) with
// the pair of image messages the real synchronizer would hand to the node.
#pragma once
#include <functional>
#include <memory>
#include <vector>
#include "message_filters/subscriber.h"
namespace refstub_vo {
template <typename M0, typename M1>
std::vector<std::function<void(std::shared_ptr<M0>, std::shared_ptr<M1>)>> &stereo_callbacks() {
  static std::vector<std::function<void(std::shared_ptr<M0>, std::shared_ptr<M1>)>> v;
  return v;
}
}  // namespace refstub_vo
namespace message_filters {
template <typename Policy>
class Synchronizer {
 public:
  typedef typename Policy::Msg0 M0;
  typedef typename Policy::Msg1 M1;
  Synchronizer(Policy, Subscriber<M0> &, Subscriber<M1> &) {}
  template <typename C, typename T>
  void registerCallback(void (C::*fp)(const std::shared_ptr<M0>, const std::shared_ptr<M1>), T *obj) {
    refstub_vo::stereo_callbacks<M0, M1>().push_back([obj, fp](std::shared_ptr<M0> a, std::shared_ptr<M1> b) { (obj->*fp)(a, b); });
  }
};
}  // namespace message_filters
