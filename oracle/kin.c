/* CPU oracle: leg kinematics.  TEST INFRASTRUCTURE ONLY.
 *
 * Go1: closed form of the reference's FROST-generated expressions
 *   /root/reference/src/go1_example/src/Expressions/{FR,FL,RR,RL}_foot.cc:15-202 (foot position) and
 *   J_{FR,FL,RR,RL}.cc:15-~610 (3x22 Jacobian; the adapter keeps columns 6+4*leg..8+4*leg,
 *   go1Sub.cpp:89-121) evaluated with the six floating-base coordinates var1[0..5] = 0 and the
 *   fixed foot joint var1[9+4*leg] = 0, which is how go1Sub.cpp:68-75 always calls them.
 *   With those zeros the generated expression collapses to the 3R chain
 *     hip roll about x at (sx*0.1881, sy*0.04675, 0); thigh pitch about y at (0, sy*0.08, 0);
 *     calf pitch about y at (0,0,-0.213); foot at (0,0,-0.213)
 *   (sx,sy) = FR(+,-) FL(+,+) RR(-,-) RL(-,+).  Pinned against oracle/_ref (compiled from the
 *   reference sources) in tests/test_oracle_kinematics.py to 1e-15.
 *
 * Cassie / PogoX: BUILDER-DEFINED serial chains (the reference ships no such model, SURVEY.md
 * fact 3); restated identically in decentralized_ekf_mhe_b200/csrc/kinematics.cuh. */
#include "oracle.h"
#include <math.h>
#include <string.h>

int orc_robot_num_legs(int robot) {
  switch (robot) {
    case ORC_ROBOT_CASSIE: return 2;
    case ORC_ROBOT_POGOX: return 1;
    default: return 4;
  }
}
int orc_robot_joints_per_leg(int robot) {
  switch (robot) {
    case ORC_ROBOT_CASSIE: return 5;
    case ORC_ROBOT_POGOX: return 3;
    default: return 3;
  }
}

static void go1_leg(int leg, const double *q, double p[3], double *J) {
  const double HX = 0.1881, HY = 0.04675, TY = 0.08, L = 0.213;
  const double sx = (leg < 2) ? 1.0 : -1.0;
  const double sy = (leg % 2 == 0) ? -1.0 : 1.0;
  double s1 = sin(q[0]), c1 = cos(q[0]);
  double s2 = sin(q[1]), c2 = cos(q[1]);
  double s23 = sin(q[1] + q[2]), c23 = cos(q[1] + q[2]);
  double xl = -L * (s2 + s23);
  double zl = -L * (c2 + c23);
  double yl = sy * TY;
  p[0] = sx * HX + xl;
  p[1] = sy * HY + c1 * yl - s1 * zl;
  p[2] = s1 * yl + c1 * zl;
  /* d/d(hip, thigh, calf) */
  double dzl2 = L * (s2 + s23), dzl3 = L * s23;
  J[0 * 3 + 0] = 0.0;
  J[0 * 3 + 1] = -L * (c2 + c23);
  J[0 * 3 + 2] = -L * c23;
  J[1 * 3 + 0] = -s1 * yl - c1 * zl;
  J[1 * 3 + 1] = -s1 * dzl2;
  J[1 * 3 + 2] = -s1 * dzl3;
  J[2 * 3 + 0] = c1 * yl - s1 * zl;
  J[2 * 3 + 1] = c1 * dzl2;
  J[2 * 3 + 2] = c1 * dzl3;
}

/* Generic serial chain used by the builder-defined models. */
typedef struct {
  int nj;
  int type[8];       /* 0 revolute, 1 prismatic */
  double axis[8][3]; /* unit axis in the joint's parent frame */
  double off[8][3];  /* translation from previous joint frame to this joint */
  double tool[3];    /* foot point in the last joint frame */
} chain_t;

static void rot_axis(const double a[3], double th, double R[9]) {
  double c = cos(th), s = sin(th), v = 1 - c;
  R[0] = c + a[0] * a[0] * v;
  R[1] = a[0] * a[1] * v - a[2] * s;
  R[2] = a[0] * a[2] * v + a[1] * s;
  R[3] = a[1] * a[0] * v + a[2] * s;
  R[4] = c + a[1] * a[1] * v;
  R[5] = a[1] * a[2] * v - a[0] * s;
  R[6] = a[2] * a[0] * v - a[1] * s;
  R[7] = a[2] * a[1] * v + a[0] * s;
  R[8] = c + a[2] * a[2] * v;
}

static void chain_fk(const chain_t *c, const double *q, double p[3], double *J) {
  double Rw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, o[3] = {0, 0, 0};
  double origin[8][3], axw[8][3];
  for (int j = 0; j < c->nj; ++j) {
    for (int r = 0; r < 3; ++r)
      o[r] += Rw[r * 3 + 0] * c->off[j][0] + Rw[r * 3 + 1] * c->off[j][1] + Rw[r * 3 + 2] * c->off[j][2];
    for (int r = 0; r < 3; ++r)
      axw[j][r] = Rw[r * 3 + 0] * c->axis[j][0] + Rw[r * 3 + 1] * c->axis[j][1] + Rw[r * 3 + 2] * c->axis[j][2];
    memcpy(origin[j], o, sizeof(o));
    if (c->type[j] == 0) {
      double Rj[9], Rn[9];
      rot_axis(c->axis[j], q[j], Rj);
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k)
          Rn[r * 3 + k] = Rw[r * 3 + 0] * Rj[0 * 3 + k] + Rw[r * 3 + 1] * Rj[1 * 3 + k] + Rw[r * 3 + 2] * Rj[2 * 3 + k];
      memcpy(Rw, Rn, sizeof(Rn));
    } else {
      for (int r = 0; r < 3; ++r) o[r] += axw[j][r] * q[j];
    }
  }
  for (int r = 0; r < 3; ++r)
    p[r] = o[r] + Rw[r * 3 + 0] * c->tool[0] + Rw[r * 3 + 1] * c->tool[1] + Rw[r * 3 + 2] * c->tool[2];
  for (int j = 0; j < c->nj; ++j) {
    if (c->type[j] == 0) {
      double d[3] = {p[0] - origin[j][0], p[1] - origin[j][1], p[2] - origin[j][2]};
      J[0 * c->nj + j] = axw[j][1] * d[2] - axw[j][2] * d[1];
      J[1 * c->nj + j] = axw[j][2] * d[0] - axw[j][0] * d[2];
      J[2 * c->nj + j] = axw[j][0] * d[1] - axw[j][1] * d[0];
    } else {
      J[0 * c->nj + j] = axw[j][0];
      J[1 * c->nj + j] = axw[j][1];
      J[2 * c->nj + j] = axw[j][2];
    }
  }
}

/* Builder-defined Cassie-like leg: hip roll (x), hip yaw (z), hip pitch (y), knee (y), toe (y). */
static void cassie_chain(int leg, chain_t *c) {
  const double sy = (leg == 0) ? 1.0 : -1.0; /* 0 = left, 1 = right */
  memset(c, 0, sizeof(*c));
  c->nj = 5;
  const double axes[5][3] = {{1, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 0}, {0, 1, 0}};
  const double offs[5][3] = {{0.021, sy * 0.135, 0.0}, {0.0, 0.0, -0.07}, {0.0, 0.0, -0.09},
                             {0.12, 0.0, -0.4896}, {0.06, 0.0, -0.5}};
  for (int j = 0; j < 5; ++j) {
    c->type[j] = 0;
    memcpy(c->axis[j], axes[j], sizeof(double) * 3);
    memcpy(c->off[j], offs[j], sizeof(double) * 3);
  }
  c->tool[0] = 0.02;
  c->tool[1] = 0.0;
  c->tool[2] = -0.05;
}

/* Builder-defined PogoX-like leg: gimbal roll (x), gimbal pitch (y), prismatic spring leg (-z). */
static void pogox_chain(chain_t *c) {
  memset(c, 0, sizeof(*c));
  c->nj = 3;
  c->type[0] = 0;
  c->axis[0][0] = 1;
  c->type[1] = 0;
  c->axis[1][1] = 1;
  c->type[2] = 1;
  c->axis[2][2] = -1;
  c->off[0][2] = -0.05;
  c->off[2][2] = -0.25;
  c->tool[2] = -0.05;
}

void orc_leg_fk(int robot, int leg, const double *q, double p[3], double *J) {
  chain_t c;
  switch (robot) {
    case ORC_ROBOT_CASSIE:
      cassie_chain(leg, &c);
      chain_fk(&c, q, p, J);
      break;
    case ORC_ROBOT_POGOX:
      pogox_chain(&c);
      chain_fk(&c, q, p, J);
      break;
    default:
      go1_leg(leg, q, p, J);
      break;
  }
}
