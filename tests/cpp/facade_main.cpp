// Drives the C++ facade (include/dekf_b200/*.hpp) the way the reference's two ROS nodes drive their estimator classes:
// orien_ekf::timerCallback publishes the quaternion (orien_ekf.cpp:77-105), robotSub::timerCallback calls
// DecentralizedEstimation::initialize at T==0 and update(T) afterwards (EstSub.cpp:58-91).
// usage: facade_main <stream.bin> <out.bin>     (binary layout: see tests/test_cpp_facade.py)
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "dekf_b200/DecentralEst.hpp"
#include "dekf_b200/orien_ekf.hpp"

static void rd(std::FILE *f, void *p, size_t bytes) {
  if (std::fread(p, 1, bytes, f) != bytes) {
    std::fprintf(stderr, "short read\n");
    std::exit(2);
  }
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  std::FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int32_t hdr[4];
  rd(f, hdr, sizeof(hdr));
  const int S = hdr[0], n = hdr[1], nq = hdr[2], nl = hdr[3];

  auto params = std::make_shared<dekf::robot_params>(dekf::robot_params::go1());
  params->n_instances_ = n;
  params->ekf_rate_ = params->rate_;  // both timers at 200 Hz (lock-step parity streams)
  auto store = std::make_shared<dekf::robot_store>();
  store->resize(n, nq, nl);

  dekf::orien_ekf ekf(*params);
  dekf::DecentralizedEstimation mhe;

  std::FILE *o = std::fopen(argv[2], "wb");
  if (!o) return 2;
  std::vector<double> foot((size_t)nl * n);
  for (int T = 0; T < S; ++T) {
    // ---- "callbacks": fill robot_store from the recorded messages
    rd(f, store->angular_b_.data(), sizeof(double) * 3 * n);
    rd(f, store->accel_b_.data(), sizeof(double) * 3 * n);
    rd(f, store->imu_time_.data(), sizeof(double) * n);
    rd(f, store->joint_states_position_.data(), sizeof(double) * (size_t)nq * n);
    rd(f, store->joint_states_velocity_.data(), sizeof(double) * (size_t)nq * n);
    rd(f, store->joint_states_position_.data() + (size_t)nq * n, sizeof(double) * (size_t)nl * n);  // foot forces
    rd(f, store->vo_new_.data(), n);
    rd(f, store->vo_quaternion_.data(), sizeof(double) * 4 * n);
    rd(f, store->vo_time_pre_.data(), sizeof(double) * n);
    rd(f, store->vo_time_now_.data(), sizeof(double) * n);
    rd(f, store->vo_p_body_pre_2_body_.data(), sizeof(double) * 3 * n);
    // ---- orien_sub timer, then the imu/filter message lands in est_sub's robot_store
    ekf.timerCallback(*store);
    store->quaternion_ = ekf.quaternion_;
    // ---- est_sub timer (EstSub.cpp:65-73)
    if (T == 0)
      mhe.initialize(store, params);
    else
      mhe.update(T);
    std::fwrite(ekf.quaternion_.data(), sizeof(double), 4 * (size_t)n, o);
    std::fwrite(mhe.x_MHE_.data(), sizeof(double), 9 * (size_t)n, o);
    std::fwrite(mhe.v_MHE_b_.data(), sizeof(double), 3 * (size_t)n, o);
    std::fwrite(mhe.R_sb_.data(), sizeof(double), 9 * (size_t)n, o);
    std::fwrite(mhe.p_vo_accmulate_.data(), sizeof(double), 3 * (size_t)n, o);
    std::fwrite(store->contact_.data(), 1, (size_t)nl * n, o);
  }
  const std::vector<double> M = mhe.mhe_qp_.M_p(), np = mhe.mhe_qp_.n_p();
  std::fwrite(M.data(), sizeof(double), M.size(), o);
  std::fwrite(np.data(), sizeof(double), np.size(), o);
  std::fclose(o);
  std::fclose(f);
  std::printf("facade ok: %d ticks x %d instances\n", S, n);
  return 0;
}
