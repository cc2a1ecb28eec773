// Reads "n" then n lines "stamp w x y z tx ty tz" (camera poses) after one line "R_ic(9) p_ic(3)"; prints one line per
// published message: t_pre t_now rel_p(3) quat(4) pos(3).  Driven by tests/test_vo_frontend.py.
#include <cstdio>

#include "../../include/dekf_b200/vo_frontend.hpp"

int main() {
  dekf::Iso3 Tbc;
  for (int i = 0; i < 9; ++i)
    if (std::scanf("%lf", &Tbc.R[i]) != 1) return 1;
  for (int i = 0; i < 3; ++i)
    if (std::scanf("%lf", &Tbc.t[i]) != 1) return 1;
  int n = 0;
  if (std::scanf("%d", &n) != 1) return 1;
  dekf::VoFrontEnd fe(Tbc);
  for (int k = 0; k < n; ++k) {
    double s, w, x, y, z, tx, ty, tz;
    if (std::scanf("%lf %lf %lf %lf %lf %lf %lf %lf", &s, &w, &x, &y, &z, &tx, &ty, &tz) != 8) return 1;
    const dekf::VoMessage m = fe.push(dekf::from_quat(w, x, y, z, tx, ty, tz), s);
    if (!m.valid) continue;
    std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", m.t_pre, m.t_now, m.rel_p[0], m.rel_p[1],
                m.rel_p[2], m.quat[0], m.quat[1], m.quat[2], m.quat[3], m.pos[0], m.pos[1], m.pos[2]);
  }
  return 0;
}
