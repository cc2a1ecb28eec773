// TEST INFRASTRUCTURE ONLY.  Driver of oracle/_ref/vo_pin: instantiates the reference's own VO wrapper node
// (StereoPubNode, visual_odometry/orbslam3_ros2/src/stereo-decentralized/stereo-pub-node.cpp, compiled UNMODIFIED against the
// stand-in headers of oracle/ref_stub_vo + oracle/ref_stub) and plays a scripted sequence of tracked camera poses through its
// GrabStereo callback; prints every message the node publishes on orb/vo and orb/pos next to what the product's host-side front-end
// (include/dekf_b200/vo_frontend.hpp) returns for the same frames.
//
// stdin:  "R_ic(9, row-major) p_ic(3)" / "n" / n lines "recv_time_s image_stamp_s  w x y z  tx ty tz" -- the pose ORB_SLAM3::System::
//         TrackStereo returns for that frame (float, before the node's own inverse, :139)
// stdout: per frame "in stamp w x y z tx ty tz" (the tracked pose after the node's inverse, widened to double: what a caller of
//         the product front-end passes in), and one line pair per published message pair:
//         "ref  t_pre t_now  rel(3)  q_wxyz(4)  pos(3)  tracked_stamp" then "ours t_pre t_now rel(3) q_wxyz(4) pos(3)"
#include <cstdio>
#include <vector>

#include "stereo-pub-node.hpp"

#include "../include/dekf_b200/vo_frontend.hpp"

namespace {
double stamp_s(const builtin_interfaces::msg::Time &t) { return (double)t.sec + (double)t.nanosec / 1e9; }
}

int main() {
  double Ric[9], pic[3];
  for (double &v : Ric)
    if (std::scanf("%lf", &v) != 1) return 1;
  for (double &v : pic)
    if (std::scanf("%lf", &v) != 1) return 1;
  int n = 0;
  if (std::scanf("%d", &n) != 1) return 1;
  std::vector<double> recv(n), img(n);
  std::vector<Sophus::SE3f> poses;
  for (int k = 0; k < n; ++k) {
    double w, x, y, z, tx, ty, tz;
    if (std::scanf("%lf %lf %lf %lf %lf %lf %lf %lf %lf", &recv[k], &img[k], &w, &x, &y, &z, &tx, &ty, &tz) != 9) return 1;
    poses.emplace_back((float)w, (float)x, (float)y, (float)z, (float)tx, (float)ty, (float)tz);
  }
  refstub_vo::script().poses = poses;
  // node parameters (the launch file's YAML in the reference)
  refstub::ParamValue pr, pp;
  pr.kind = 5;
  pr.v.assign(Ric, Ric + 9);
  pp.kind = 5;
  pp.v.assign(pic, pic + 3);
  refstub::param_overrides()["vo_sub.R_ic"] = pr;
  refstub::param_overrides()["vo_sub.p_ic"] = pp;

  ORB_SLAM3::System slam;
  StereoPubNode *node = new StereoPubNode(&slam, "unused.yaml", "false");  // owned by the handle the node gives its subscribers
  (void)node;

  custom_msgs::msg::VoRealtiveTransform last_vo;
  geometry_msgs::msg::PoseStamped last_pos;
  int got_vo = 0, got_pos = 0;
  rclcpp::Node tap("tap");
  auto s1 = tap.create_subscription<custom_msgs::msg::VoRealtiveTransform>(
      "orb/vo", 3, [&](custom_msgs::msg::VoRealtiveTransform::SharedPtr m) { last_vo = *m; ++got_vo; });
  auto s2 = tap.create_subscription<geometry_msgs::msg::PoseStamped>(
      "orb/pos", 3, [&](geometry_msgs::msg::PoseStamped::SharedPtr m) { last_pos = *m; ++got_pos; });

  dekf::Iso3 Tbc;
  for (int i = 0; i < 9; ++i) Tbc.R[i] = Ric[i];
  for (int i = 0; i < 3; ++i) Tbc.t[i] = pic[i];
  dekf::VoFrontEnd ours(Tbc);

  auto &cbs = refstub_vo::stereo_callbacks<sensor_msgs::msg::Image, sensor_msgs::msg::Image>();
  if (cbs.empty()) return 2;
  for (int k = 0; k < n; ++k) {
    refstub::now_ns() = (int64_t)(recv[k] * 1e9 + 0.5);
    auto left = std::make_shared<sensor_msgs::msg::Image>(), right = std::make_shared<sensor_msgs::msg::Image>();
    left->height = right->height = 480;
    left->width = right->width = 848;
    left->header.stamp.sec = (int32_t)img[k];
    left->header.stamp.nanosec = (uint32_t)((img[k] - (double)(int32_t)img[k]) * 1e9 + 0.5);
    right->header = left->header;
    const int v0 = got_vo, p0 = got_pos;
    cbs.back()(left, right);
    // the product front-end gets the tracked pose as the node sees it after its own inverse (:139), widened to double
    const Sophus::SE3f inv = poses[k].inverse();
    const auto &c = inv.unit_quaternion().coeffs();
    // the stamp as a subscriber reads it from the message header (sec + nanosec / 1e9)
    const builtin_interfaces::msg::Time now_msg = rclcpp::Time(refstub::now_ns());
    const dekf::VoMessage m = ours.push(dekf::from_quat(c.w(), c.x(), c.y(), c.z(), inv.translation().x(), inv.translation().y(),
                                                        inv.translation().z()),
                                        stamp_s(now_msg));
    std::printf("in %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", stamp_s(now_msg), (double)c.w(), (double)c.x(), (double)c.y(),
                (double)c.z(), (double)inv.translation().x(), (double)inv.translation().y(), (double)inv.translation().z());
    const bool published = got_vo > v0 && got_pos > p0;
    if (published != m.valid) return 3;
    if (!published) continue;
    std::printf("ref %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", stamp_s(last_vo.header_pre.stamp),
                stamp_s(last_vo.header.stamp), last_vo.x_relative, last_vo.y_relative, last_vo.z_relative, last_pos.pose.orientation.w,
                last_pos.pose.orientation.x, last_pos.pose.orientation.y, last_pos.pose.orientation.z, last_pos.pose.position.x,
                last_pos.pose.position.y, last_pos.pose.position.z, refstub_vo::script().stamps_seen.back());
    std::printf("ours %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", m.t_pre, m.t_now, m.rel_p[0], m.rel_p[1],
                m.rel_p[2], m.quat[0], m.quat[1], m.quat[2], m.quat[3], m.pos[0], m.pos[1], m.pos[2]);
  }
  return 0;
}
