// Host-side VO front-end: turns a sequence of camera poses into the two messages the estimators consume.
//
// Replaces the pose arithmetic of the reference's ORB-SLAM3 wrapper node (ORB-SLAM3 itself is out of scope; its
// tracked camera pose is the input here), /root/reference/src/visual_odometry/orbslam3_ros2/src/stereo-decentralized/
// stereo-pub-node.cpp:
//   :92-94, :141-150   T_world_to_camera from the tracked pose (after the wrapper's own inverse, :139)
//   :56-66             T_body_to_camera from the node parameters R_ic (row-major 3x3), p_ic
//   :153-156           first frame: T_world_to_body_init = T_world_to_camera * T_body_to_camera^-1, nothing published
//   :161               relative_body_12 = T_body_to_camera * T_world_to_camera_pre^-1 * T_world_to_camera * T_body_to_camera^-1
//                      -> topic orb/vo  (custom_msgs/VoRealtiveTransform: stamps pre/now + xyz, :181-192)
//                      -> dekf_inputs.vo_rel_p, vo_time_pre, vo_time_now
//   :163-165           T_world_to_body = T_world_to_body_init^-1 * T_world_to_camera * T_body_to_camera^-1, its rotation
//                      as a quaternion -> topic orb/pos (:168-179) -> dekf_inputs.vo_quat ([w,x,y,z])
// Plain C++17, no Eigen (not available in this image); Eigen's Quaterniond(Matrix3d) branch structure is kept so that
// the quaternion sign convention matches.  Pinned to the reference node itself: stereo-pub-node.cpp compiles unmodified against
// stand-in headers into oracle/_ref/vo_pin, and every field of the orb/vo and orb/pos messages it publishes for scripted tracked
// poses is reproduced (tests/test_vo_frontend.py, tests/golden/vo_frontend_golden.npz).
#pragma once
#include <array>
#include <cmath>

namespace dekf {

struct Iso3 {
  std::array<double, 9> R{1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major
  std::array<double, 3> t{0, 0, 0};
};

inline Iso3 compose(const Iso3 &a, const Iso3 &b) {  // a * b
  Iso3 c;
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) c.R[r * 3 + k] = a.R[r * 3 + 0] * b.R[0 * 3 + k] + a.R[r * 3 + 1] * b.R[1 * 3 + k] + a.R[r * 3 + 2] * b.R[2 * 3 + k];
    c.t[r] = a.R[r * 3 + 0] * b.t[0] + a.R[r * 3 + 1] * b.t[1] + a.R[r * 3 + 2] * b.t[2] + a.t[r];
  }
  return c;
}
inline Iso3 inverse(const Iso3 &a) {
  Iso3 c;
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) c.R[r * 3 + k] = a.R[k * 3 + r];
  for (int r = 0; r < 3; ++r) c.t[r] = -(c.R[r * 3 + 0] * a.t[0] + c.R[r * 3 + 1] * a.t[1] + c.R[r * 3 + 2] * a.t[2]);
  return c;
}
inline Iso3 from_quat(double w, double x, double y, double z, double tx, double ty, double tz) {
  Iso3 c;
  c.R = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
         2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
         2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
  c.t = {tx, ty, tz};
  return c;
}
// Eigen::Quaterniond(const Matrix3d&) (Shepperd's method as Eigen implements it), [w,x,y,z]
inline std::array<double, 4> to_quat(const std::array<double, 9> &m) {
  std::array<double, 4> q{};
  double tr = m[0] + m[4] + m[8];
  if (tr > 0.0) {
    double t = std::sqrt(tr + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (m[7] - m[5]) * t;
    q[2] = (m[2] - m[6]) * t;
    q[3] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[k * 3 + j] - m[j * 3 + k]) * t;
    q[1 + j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    q[1 + k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
  }
  return q;
}

struct VoMessage {
  bool valid = false;              // false for the first frame (nothing is published, stereo-pub-node.cpp:153-159)
  std::array<double, 3> rel_p{};   // orb/vo  x_relative, y_relative, z_relative
  double t_pre = 0.0, t_now = 0.0; // orb/vo  header_pre.stamp, header.stamp
  std::array<double, 4> quat{};    // orb/pos orientation [w,x,y,z]
  std::array<double, 3> pos{};     // orb/pos position
};

class VoFrontEnd {
 public:
  explicit VoFrontEnd(const Iso3 &T_body_to_camera = Iso3()) : T_bc_(T_body_to_camera), T_bc_inv_(inverse(T_body_to_camera)) {}
  // one tracked frame: camera pose in the VO world frame and the stamp the wrapper puts on it
  VoMessage push(const Iso3 &T_world_to_camera, double stamp) {
    VoMessage m;
    if (init_ == 0) {
      T_wb_init_inv_ = inverse(compose(T_world_to_camera, T_bc_inv_));
    } else {
      const Iso3 rel = compose(compose(compose(T_bc_, inverse(T_wc_pre_)), T_world_to_camera), T_bc_inv_);
      const Iso3 T_wb = compose(compose(T_wb_init_inv_, T_world_to_camera), T_bc_inv_);
      m.valid = true;
      m.rel_p = rel.t;
      m.t_pre = stamp_pre_;
      m.t_now = stamp;
      m.quat = to_quat(T_wb.R);
      m.pos = T_wb.t;
    }
    T_wc_pre_ = T_world_to_camera;
    stamp_pre_ = stamp;
    ++init_;
    return m;
  }

 private:
  Iso3 T_bc_, T_bc_inv_, T_wb_init_inv_, T_wc_pre_;
  double stamp_pre_ = 0.0;
  int init_ = 0;
};

}  // namespace dekf
