// TEST INFRASTRUCTURE ONLY: stand-in header written for this repository so that the reference's VO wrapper node
// (visual_odometry/orbslam3_ros2/src/stereo-decentralized/stereo-pub-node.cpp) compiles UNMODIFIED into oracle/_ref/vo_pin
// (recipe: oracle/Makefile vo_pin; driver: oracle/vo_pin_main.cc).  ORB-SLAM3, OpenCV, Sophus, cv_bridge and message_filters are
// absent from the image and out of scope; only the surface that source file touches exists here.
//
// Sophus::SE3f as far as the node uses it: a unit quaternion + translation in FLOAT, inverse() as Sophus defines it
// (SE3(so3().inverse(), so3().inverse() * (translation() * -1)), the rotation of a point as Eigen's
// QuaternionBase::_transformVector: uv = 2 q.vec x v; v + w uv + q.vec x uv).
#pragma once
namespace Sophus {
struct Vec3f { float v[3] = {0, 0, 0}; float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; } };
struct QuatCoeffsf { float c[4] = {0, 0, 0, 1}; float x() const { return c[0]; } float y() const { return c[1]; } float z() const { return c[2]; } float w() const { return c[3]; } };
struct Quatf { QuatCoeffsf cf; const QuatCoeffsf &coeffs() const { return cf; } };
class SE3f {
 public:
  SE3f() {}
  SE3f(float w, float x, float y, float z, float tx, float ty, float tz) {
    q_.cf.c[0] = x; q_.cf.c[1] = y; q_.cf.c[2] = z; q_.cf.c[3] = w;
    t_.v[0] = tx; t_.v[1] = ty; t_.v[2] = tz;
  }
  const Vec3f &translation() const { return t_; }
  const Quatf &unit_quaternion() const { return q_; }
  SE3f inverse() const {
    const float w = q_.cf.c[3], x = -q_.cf.c[0], y = -q_.cf.c[1], z = -q_.cf.c[2];  // conjugate of the unit quaternion
    const float v[3] = {t_.v[0] * -1.f, t_.v[1] * -1.f, t_.v[2] * -1.f};
    float uv[3] = {y * v[2] - z * v[1], z * v[0] - x * v[2], x * v[1] - y * v[0]};
    for (int k = 0; k < 3; ++k) uv[k] += uv[k];
    const float c[3] = {y * uv[2] - z * uv[1], z * uv[0] - x * uv[2], x * uv[1] - y * uv[0]};
    return SE3f(w, x, y, z, v[0] + w * uv[0] + c[0], v[1] + w * uv[1] + c[1], v[2] + w * uv[2] + c[2]);
  }
 private:
  Quatf q_;
  Vec3f t_;
};
}  // namespace Sophus
