/* CPU oracle: moving-horizon estimator.  TEST INFRASTRUCTURE ONLY.
 *
 * Literal restatement of
 *   /root/reference/src/decentral_legged_est/src/DecentralEst.cpp  (stage assembly, VO sync, KF alt.)
 *   /root/reference/src/decentral_legged_est/src/MheSrb.cpp        (QP bookkeeping, marginalisation)
 *   /root/reference/src/decentral_legged_est/src/Spline/Bezier_simple.cpp
 *   /root/reference/src/go1_example/src/go1Sub.cpp:64-125          (sensor adapter)
 * The string-keyed registries of MheSrb.cpp are kept as per-stage records (same content, no
 * strings); lb_all/ub_all are kept as the literal row-ordered vectors because the reference updates
 * them by ROW INDEX (Update_Image_bound, MheSrb.cpp:449-459) independently of the registry copy
 * (updateConstraintBound, MheSrb.cpp:233-243).
 * OSQP itself is not in /root/reference (unpinned dependency): solve_mode 0 returns
 * the unique optimum of the assembled QP by slack elimination + banded Cholesky, solve_mode 2 runs
 * the OSQP-style ADMM restated in admm.c.  Parity status: see oracle.h (pinned against the compiled reference sources; OSQP's own iterates unpinned). */
#include "oracle.h"
#include "la.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXDM 12
#define MAXNJ 5
#define MAXLEG 4
#define OSQP_INFTY 1e30 /* OsqpEigen::INFTY, MheSrb.hpp:81 */

/* implemented in admm.c */
int orc_admm_solve_triplets(int n, int m, int nnzP, const int *Pi, const int *Pj, const double *Px,
                            int nnzA, const int *Ai, const int *Aj, const double *Ax, const double *gd,
                            const double *ld, const double *ud, const orc_params *prm, double *z_out,
                            int *iters_out);

typedef struct {
  double imu_time;
  int disc;
  double R[9], accel_s[3], p_foot[MAXDM], J[MAXDM * MAXNJ], contact[MAXLEG], joint_vel[MAXLEG * MAXNJ],
      angular_b[3];
} hist_t;

typedef struct {
  int k;
  double *Hx, *gx;                        /* H block / g segment on x_k */
  double *Q_meas, *meas_lb, *meas_ub;     /* Measurement_k */
  int has_dyn;
  double *A_dyn, *Q_dyn, *dyn_lb, *dyn_ub; /* Dynamic_k */
  double Q_cam[9], vo_lb[3], vo_ub[3];     /* VO_measurement_k */
  int vo_equality;
} stage_t;

/* Bezier_simple.hpp */
typedef struct {
  int nwp;
  double wp[4][3], wp_time[4];
  double t_interval, t_start, u_inc, interpolate_num;
  double node_pre[3];
  int nnodes;
  double *nodes, *distances; /* 3 * nnodes */
  int cap;
} bezier_t;

struct orc_mhe {
  orc_params P;
  int ds, dm, dc, nj, nlegs, N, est_type, leg_odom_type;
  double dt, gravity[3], p_ib[3];
  double C_p[3], C_accel[3], C_accel_bias[3], C_enc_pos[8], C_enc_vel[8], C_gyro[3], C_foot_slide[3],
      C_foot_swing[3];
  double Q_accel_bias[3], Q_foot_slide[3], Q_foot_swing[3], Q_vo_p[3];
  double *A_meas; /* dm x ds */
  /* robot_store (persistent between ticks, like the shared_ptr in the reference) */
  orc_sample store;
  double st_contact[MAXLEG], st_p_foot[MAXDM], st_J[MAXDM * MAXNJ];
  /* current measurement */
  double R_sb[9], accel_s[3], angular_b[3];
  /* stacks */
  hist_t *stack;
  int nstack;
  int *vo_insert_idx, *vo_insert_disc;
  int nvo_insert;
  int vo_flag;
  double p_vo_acc[3];
  bezier_t bz;
  int vo_dbg[10];
  /* QP */
  stage_t *win;
  int nwin, capwin;
  double *lb_all, *ub_all;
  int nCon, nVar;
  int have_arrival;
  double *M_p, *n_p;
  int prior_in_regs; /* Cost_regs.count("Prior_0") */
  double *Q_prior, *x_prior;
  double *solution;
  int admm_iters;
  /* results */
  double *x_MHE;
  double v_MHE_b[3];
  double b_meas[MAXDM], *Q_meas_last;
  /* KF */
  double *x_KF, *C_KF;
  double v_KF_b[3];
};

/* ------------------------------------------------------------------------- params */
void orc_params_go1_defaults(orc_params *p) {
  /* /root/reference/src/go1_example/config/parameters_go1.yaml:5-50 */
  memset(p, 0, sizeof(*p));
  for (int i = 0; i < 3; ++i) {
    p->p_init_std[i] = 0.001;
    p->v_init_std[i] = 0.001;
    p->foot_init_std[i] = 0.001;
    p->accel_bias_init_std[i] = 0.0001;
    p->p_process_std[i] = 0.001;
    p->gyro_input_std[i] = 0.03;
    p->foot_slide_std[i] = 0.003;
    p->foot_swing_std[i] = 10000000.0;
    p->vo_p_std[i] = 0.000015;
  }
  for (int i = 0; i < 8; ++i) {
    p->joint_position_std[i] = 0.04;
    p->joint_velocity_std[i] = 0.22;
  }
  p->accel_input_std[0] = 0.025;
  p->accel_input_std[1] = 0.025;
  p->accel_input_std[2] = 0.02;
  p->accel_bias_std[0] = 0.07;
  p->accel_bias_std[1] = 0.02;
  p->accel_bias_std[2] = 0.03;
  p->quaternion_ib[0] = 1.0;
  p->p_ib[0] = 0.01592;
  p->p_ib[1] = 0.06659;
  p->p_ib[2] = 0.00617;
  p->num_legs = 4;
  p->leg_odom_type = 0;
  p->contact_effort_threshold = 150.0;
  p->rate = 200;
  p->N = 20;
  p->est_type = 0;
  p->rho = 0.1;
  p->alpha = 1.6;
  p->delta = 0.00001;
  p->sigma = 0.00001;
  p->verbose = 0;
  p->adapt_rho = 1;
  p->polish = 0;
  p->max_qp_iter = 4000;
  p->prim_tol = 1e-6;
  p->dual_tol = 1e-6;
  p->relative_tol = 1e-6;
  p->abs_tol = 1e-6;
  p->time_limit = 0.0028;
  p->robot = ORC_ROBOT_GO1;
  p->solve_mode = 0;
  /* DecentralEst.cpp:181: Vector3d p_imu_2_opti(0.016041, 0.089061, 0.0579875) */
  p->p_imu_2_opti[0] = 0.016041;
  p->p_imu_2_opti[1] = 0.089061;
  p->p_imu_2_opti[2] = 0.0579875;
}

/* ------------------------------------------------------------------------- small helpers */
static double *dalloc(int n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

/* EigenUtils.hpp:91-97 */
static void vector3dSkew(double S[9], const double v[3]) {
  S[0] = 0;
  S[1] = -v[2];
  S[2] = v[1];
  S[3] = v[2];
  S[4] = 0;
  S[5] = -v[0];
  S[6] = -v[1];
  S[7] = v[0];
  S[8] = 0;
}

static void cross3(double o[3], const double a[3], const double b[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

static int upper_bound_hist(const hist_t *a, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = lo + (hi - lo) / 2;
    if (!(v < a[mid].imu_time))
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

/* ------------------------------------------------------------------------- Bezier_simple.cpp */
/* :12-27 */
static void bezier_add_way_point(bezier_t *b, const double p[3], double t_end) {
  if (b->nwp < 4) {
    memcpy(b->wp[b->nwp], p, sizeof(double) * 3);
    b->wp_time[b->nwp] = t_end;
    b->nwp++;
  } else {
    /* push_back then erase(begin) when size > 4 */
    for (int i = 0; i < 3; ++i) {
      memcpy(b->wp[i], b->wp[i + 1], sizeof(double) * 3);
      b->wp_time[i] = b->wp_time[i + 1];
    }
    memcpy(b->wp[3], p, sizeof(double) * 3);
    b->wp_time[3] = t_end;
  }
  b->t_interval = b->wp_time[b->nwp - 1] - b->wp_time[0];
}

/* :60-71 */
static void bezier_set_interval(bezier_t *b, double t_start, int interpolate_num, double dt) {
  b->t_start = t_start;
  b->u_inc = dt / (double)b->t_interval;
  b->interpolate_num = interpolate_num;
  b->node_pre[0] = b->node_pre[1] = b->node_pre[2] = 0.0;
}

/* :73-82 */
static void bezier_interpolate(double out[3], double u, const double *P0, const double *P1,
                               const double *P2, const double *P3) {
  for (int c = 0; c < 3; ++c) {
    double point = u * u * u * ((-1) * P0[c] + 3 * P1[c] - 3 * P2[c] + P3[c]);
    point += u * u * (3 * P0[c] - 6 * P1[c] + 3 * P2[c]);
    point += u * ((-3) * P0[c] + 3 * P1[c]);
    point += P0[c];
    out[c] = point;
  }
}

/* :29-58 */
static void bezier_interpolate_waypoint(bezier_t *b) {
  b->nnodes = 0;
  if (b->nwp < 4) return;
  int pt = (b->nwp - 1) - 3;
  double u0 = (double)(b->t_start - b->wp_time[0]) / (double)b->t_interval;
  for (double i = 0; i < b->interpolate_num; i++) {
    double u = u0 + b->u_inc * i;
    double node[3];
    bezier_interpolate(node, u, b->wp[pt], b->wp[pt + 1], b->wp[pt + 2], b->wp[pt + 3]);
    if (b->nnodes == b->cap) {
      b->cap = b->cap ? 2 * b->cap : 64;
      b->nodes = (double *)realloc(b->nodes, sizeof(double) * 3 * (size_t)b->cap);
      b->distances = (double *)realloc(b->distances, sizeof(double) * 3 * (size_t)b->cap);
    }
    for (int c = 0; c < 3; ++c) {
      b->distances[3 * b->nnodes + c] = node[c] - b->node_pre[c];
      b->node_pre[c] = node[c];
      b->nodes[3 * b->nnodes + c] = node[c];
    }
    b->nnodes++;
  }
}

/* ------------------------------------------------------------------------- create / destroy */
orc_mhe *orc_mhe_create(const orc_params *p) {
  orc_mhe *m = (orc_mhe *)calloc(1, sizeof(orc_mhe));
  m->P = *p;
  /* DecentralEst.cpp:14-22 */
  m->est_type = p->est_type;
  m->dt = 1.0 / p->rate;
  m->N = p->N;
  m->nlegs = p->num_legs;
  m->nj = orc_robot_joints_per_leg(p->robot);
  m->leg_odom_type = p->leg_odom_type;
  m->ds = 9 + 3 * m->leg_odom_type * m->nlegs;
  m->dm = 3 * m->nlegs;
  m->dc = 3;
  m->gravity[2] = -9.81; /* :27 */
  memcpy(m->p_ib, p->p_ib, sizeof(double) * 3);
  /* :39-51, :1017-1029 */
  for (int i = 0; i < 3; ++i) {
    m->C_p[i] = pow(p->p_process_std[i], 2);
    m->C_accel[i] = pow(p->accel_input_std[i], 2);
    m->C_accel_bias[i] = pow(p->accel_bias_std[i], 2);
    m->C_gyro[i] = pow(p->gyro_input_std[i], 2);
    m->C_foot_slide[i] = pow(p->foot_slide_std[i], 2);
    m->C_foot_swing[i] = pow(p->foot_swing_std[i], 2);
    m->Q_accel_bias[i] = 1 / pow(p->accel_bias_std[i], 2);
    m->Q_foot_slide[i] = 1 / pow(p->foot_slide_std[i], 2);
    m->Q_foot_swing[i] = 1 / pow(p->foot_swing_std[i], 2);
    m->Q_vo_p[i] = 1 / pow(p->vo_p_std[i], 2);
  }
  for (int i = 0; i < 8; ++i) {
    m->C_enc_pos[i] = pow(p->joint_position_std[i], 2);
    m->C_enc_vel[i] = pow(p->joint_velocity_std[i], 2);
  }
  int ds = m->ds, dm = m->dm;
  /* :86-120 A_meas */
  m->A_meas = dalloc(dm * ds);
  for (int i = 0; i < m->nlegs; ++i) {
    if (m->leg_odom_type == 0) {
      for (int c = 0; c < 3; ++c) m->A_meas[(i * 3 + c) * ds + 3 + c] = 1.0;
    } else {
      for (int c = 0; c < 3; ++c) {
        m->A_meas[(i * 3 + c) * ds + c] = -1.0;
        m->A_meas[(i * 3 + c) * ds + 9 + i * 3 + c] = 1.0;
      }
    }
  }
  m->stack = (hist_t *)calloc((size_t)(4 * m->N + 3), sizeof(hist_t));
  m->vo_insert_idx = (int *)calloc((size_t)(m->N + 3), sizeof(int));
  m->vo_insert_disc = (int *)calloc((size_t)(m->N + 3), sizeof(int));
  m->capwin = m->N + 3;
  m->win = (stage_t *)calloc((size_t)m->capwin, sizeof(stage_t));
  int maxcon = (m->N + 2) * (dm + ds + 3);
  m->lb_all = dalloc(maxcon);
  m->ub_all = dalloc(maxcon);
  m->M_p = dalloc(ds * ds);
  m->n_p = dalloc(ds);
  m->Q_prior = dalloc(ds * ds);
  m->x_prior = dalloc(ds);
  m->solution = dalloc((m->N + 2) * (2 * ds + dm + 3));
  m->x_MHE = dalloc(ds);
  m->Q_meas_last = dalloc(dm * dm);
  m->x_KF = dalloc(ds);
  m->C_KF = dalloc(ds * ds);
  return m;
}

static void stage_free(stage_t *s) {
  free(s->Hx);
  free(s->gx);
  free(s->Q_meas);
  free(s->meas_lb);
  free(s->meas_ub);
  free(s->A_dyn);
  free(s->Q_dyn);
  free(s->dyn_lb);
  free(s->dyn_ub);
  memset(s, 0, sizeof(*s));
}

void orc_mhe_destroy(orc_mhe *m) {
  if (!m) return;
  for (int i = 0; i < m->nwin; ++i) stage_free(&m->win[i]);
  free(m->win);
  free(m->A_meas);
  free(m->stack);
  free(m->vo_insert_idx);
  free(m->vo_insert_disc);
  free(m->lb_all);
  free(m->ub_all);
  free(m->M_p);
  free(m->n_p);
  free(m->Q_prior);
  free(m->x_prior);
  free(m->solution);
  free(m->x_MHE);
  free(m->Q_meas_last);
  free(m->x_KF);
  free(m->C_KF);
  free(m->bz.nodes);
  free(m->bz.distances);
  free(m);
}

/* ------------------------------------------------------------------------- sensor adapter */
/* go1Sub.cpp:64-125 (generalised over robot model; Go1: 4 legs x 3 joints, forces at [12+i]) */
static void adapter_lo_callback(orc_mhe *m) {
  int nl = m->nlegs, nj = m->nj;
  for (int i = 0; i < nl; ++i) {
    /* :74 contact(i) = position(12+i) >= threshold ? 1.0 : 0.0 */
    m->st_contact[i] = (m->store.joint_pos[nl * nj + i] >= m->P.contact_effort_threshold) ? 1.0 : 0.0;
    double p[3], J[3 * MAXNJ];
    orc_leg_fk(m->P.robot, i, &m->store.joint_pos[i * nj], p, J);
    for (int c = 0; c < 3; ++c) m->st_p_foot[i * 3 + c] = p[c] + m->p_ib[c]; /* :86, :97, ... */
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < nj; ++c) m->st_J[(i * 3 + r) * nj + c] = J[r * nj + c]; /* :91 block<3,3>(0, 6+4i) */
  }
}

/* ------------------------------------------------------------------------- GetMeasurement */
/* DecentralEst.cpp:864-985 */
static void get_measurement(orc_mhe *m, int T) {
  int N = m->N;
  orc_quat_to_rot(m->store.quaternion, m->R_sb); /* :867 */
  double imu_time = m->store.imu_time;
  la_mm(m->accel_s, m->R_sb, m->store.accel_b, 3, 3, 1); /* :871 */
  for (int c = 0; c < 3; ++c) m->accel_s[c] += m->gravity[c];
  memcpy(m->angular_b, m->store.angular_b, sizeof(double) * 3);

  if (m->store.vo_new && m->nstack > 0) { /* :884 */
    double vo_p[3] = {m->store.vo_p[0], m->store.vo_p[1], m->store.vo_p[2]};
    double t_pre = m->store.vo_time_pre, t_now = m->store.vo_time_now;
    m->store.vo_new = 0; /* :891 */
    for (int q = 0; q < 10; ++q) m->vo_dbg[q] = -2;
    m->vo_dbg[0] = 0;
    m->vo_dbg[8] = 0;
    m->vo_dbg[9] = m->nstack;
    int ub = upper_bound_hist(m->stack, m->nstack, t_pre); /* :895 */
    if (ub == 0) {
      /* :898-904 drop */
      m->vo_dbg[1] = -1;
    } else {
      int i_pre = ub - 1;                                               /* :907 */
      const double *R_pre = m->stack[i_pre].R;                          /* :909 */
      int i_now = upper_bound_hist(m->stack, m->nstack, t_now) - 1;     /* :911-913 */
      double d[3];
      la_mm(d, R_pre, vo_p, 3, 3, 1);
      for (int c = 0; c < 3; ++c) m->p_vo_acc[c] += d[c];               /* :915 */
      int w0 = (int)(m->nstack - (N < T ? N : T));                      /* :917 */
      int i0 = (w0 > i_pre) ? w0 : i_pre;                               /* :918 */
      double t_start = m->stack[i0].imu_time;                           /* :919 */
      int disc_start = m->stack[i0].disc;                               /* :920 */
      bezier_add_way_point(&m->bz, m->p_vo_acc, t_now);                 /* :923 */
      m->vo_dbg[0] = 1;
      m->vo_dbg[1] = i_pre;
      m->vo_dbg[2] = i_now;
      m->vo_dbg[3] = w0;
      m->vo_dbg[4] = i0;
      if (i_now > w0 && m->bz.nwp >= 4) {                               /* :925 */
        double insert_relative_idx = i0 - w0;                           /* :927 */
        double interpolate_num = i_now - i0 + 1;                        /* :928 */
        bezier_set_interval(&m->bz, t_start, (int)interpolate_num, m->dt); /* :930 */
        bezier_interpolate_waypoint(&m->bz);                            /* :933 */
        m->vo_insert_idx[m->nvo_insert] = (int)insert_relative_idx;     /* :935 */
        m->vo_insert_disc[m->nvo_insert] = disc_start;                  /* :936 */
        m->nvo_insert++;
        m->vo_flag = 1;                                                 /* :938 */
        m->vo_dbg[5] = (int)insert_relative_idx;
        m->vo_dbg[6] = (int)interpolate_num;
        m->vo_dbg[7] = disc_start;
        m->vo_dbg[8] = 1;
      }
    }
  }
  /* :949-957 push */
  hist_t *h = &m->stack[m->nstack++];
  h->imu_time = imu_time;
  h->disc = T;
  memcpy(h->R, m->R_sb, sizeof(double) * 9);
  memcpy(h->accel_s, m->accel_s, sizeof(double) * 3);
  memcpy(h->p_foot, m->st_p_foot, sizeof(double) * MAXDM);
  memcpy(h->J, m->st_J, sizeof(double) * MAXDM * MAXNJ);
  memcpy(h->contact, m->st_contact, sizeof(double) * MAXLEG);
  memcpy(h->joint_vel, m->store.joint_vel, sizeof(double) * MAXLEG * MAXNJ);
  memcpy(h->angular_b, m->angular_b, sizeof(double) * 3);
  /* :963-975 */
  if (m->nstack > 4 * N + 1) {
    memmove(&m->stack[0], &m->stack[1], sizeof(hist_t) * (size_t)(m->nstack - 1));
    m->nstack--;
  }
  /* :977-984 */
  if (m->nvo_insert >= N + 1) {
    memmove(&m->vo_insert_idx[0], &m->vo_insert_idx[1], sizeof(int) * (size_t)(m->nvo_insert - 1));
    memmove(&m->vo_insert_disc[0], &m->vo_insert_disc[1], sizeof(int) * (size_t)(m->nvo_insert - 1));
    m->nvo_insert--;
  }
}

/* ------------------------------------------------------------------------- measurement rows */
/* b_meas / Q_meas (MHE) or C_meas (KF) from stack.back(): DecentralEst.cpp:269-331, :509-570 and
 * :636-690, :803-855 (the four copies are the same computation). out_is_cov: 1 -> covariance. */
static void build_meas(orc_mhe *m, const double R_sb[9], double *b_meas, double *QorC, int out_is_cov) {
  int dm = m->dm, nj = m->nj;
  const hist_t *bk = &m->stack[m->nstack - 1];
  la_zero(QorC, dm * dm);
  double Rt[9];
  la_transpose(Rt, R_sb, 3, 3);
  for (int i = 0; i < m->nlegs; ++i) {
    const double *Ji = &bk->J[(i * 3) * nj]; /* 3 x nj */
    const double *pi = &bk->p_foot[i * 3];
    if (m->leg_odom_type == 0) {
      /* b = -R J dq - R (omega x p)   :275-276, :515-516 */
      double Jdq[3], RJdq[3], wxp[3], Rwxp[3];
      la_mm(Jdq, Ji, &bk->joint_vel[i * nj], 3, nj, 1);
      la_mm(RJdq, R_sb, Jdq, 3, 3, 1);
      cross3(wxp, bk->angular_b, pi);
      la_mm(Rwxp, R_sb, wxp, 3, 3, 1);
      for (int c = 0; c < 3; ++c) b_meas[i * 3 + c] = -RJdq[c] - Rwxp[c];
      double blk[9];
      if (bk->contact[i] == 0.0) { /* :277, :517 */
        la_zero(blk, 9);
        for (int c = 0; c < 3; ++c) blk[c * 3 + c] = out_is_cov ? m->C_foot_swing[c] : m->Q_foot_swing[c];
      } else {
        /* G = [-J, -omega^x J, p^x]; C = blkdiag(C_enc_vel, C_enc_pos, C_gyro)  :283-303 */
        int ng = 2 * nj + 3;
        double G[3 * (2 * MAXNJ + 3)], GC[3 * (2 * MAXNJ + 3)], Cd[2 * MAXNJ + 3];
        double wS[9], pS[9], wJ[3 * MAXNJ];
        vector3dSkew(wS, bk->angular_b);
        vector3dSkew(pS, pi);
        la_mm(wJ, wS, Ji, 3, 3, nj);
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < nj; ++c) {
            G[r * ng + c] = -Ji[r * nj + c];
            G[r * ng + nj + c] = -wJ[r * nj + c];
          }
          for (int c = 0; c < 3; ++c) G[r * ng + 2 * nj + c] = pS[r * 3 + c];
        }
        for (int c = 0; c < nj; ++c) {
          Cd[c] = m->C_enc_vel[c];
          Cd[nj + c] = m->C_enc_pos[c];
        }
        for (int c = 0; c < 3; ++c) Cd[2 * nj + c] = m->C_gyro[c];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < ng; ++c) GC[r * ng + c] = G[r * ng + c] * Cd[c];
        double GCG[9], RGCG[9], Cm[9];
        la_mmt(GCG, GC, G, 3, ng, 3);
        la_mm(RGCG, R_sb, GCG, 3, 3, 3);
        la_mm(Cm, RGCG, Rt, 3, 3, 3); /* R G C G' R' */
        if (out_is_cov)
          la_copy(blk, Cm, 9);
        else
          la_inverse(blk, Cm, 3); /* :305 */
      }
      la_set_block(QorC, dm, i * 3, i * 3, blk, 3, 3);
    } else {
      /* foot position measurement :317-319, :554-561, :677-680 */
      la_mm(&b_meas[i * 3], R_sb, pi, 3, 3, 1);
      double JC[3 * MAXNJ], JCJ[9], inner[9], t[9], blk[9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < nj; ++c) JC[r * nj + c] = Ji[r * nj + c] * m->C_enc_pos[c];
      la_mmt(JCJ, JC, Ji, 3, nj, 3);
      if (out_is_cov)
        la_copy(inner, JCJ, 9);
      else
        la_inverse(inner, JCJ, 3);
      la_mm(t, R_sb, inner, 3, 3, 3);
      la_mm(blk, t, Rt, 3, 3, 3);
      la_set_block(QorC, dm, i * 3, i * 3, blk, 3, 3);
    }
  }
}

/* ------------------------------------------------------------------------- QP bookkeeping */
static stage_t *win_push(orc_mhe *m, int k) {
  stage_t *s = &m->win[m->nwin++];
  memset(s, 0, sizeof(*s));
  s->k = k;
  s->Hx = dalloc(m->ds * m->ds);
  s->gx = dalloc(m->ds);
  s->Q_meas = dalloc(m->dm * m->dm);
  s->meas_lb = dalloc(m->dm);
  s->meas_ub = dalloc(m->dm);
  s->A_dyn = dalloc(m->ds * m->ds);
  s->Q_dyn = dalloc(m->ds * m->ds);
  s->dyn_lb = dalloc(m->ds);
  s->dyn_ub = dalloc(m->ds);
  return s;
}

static void recount(orc_mhe *m) {
  /* nVar = K(ds+dm)+(K-1)(ds+dc); nCon = K dm + (K-1)(ds+dc)  (MheSrb.cpp:35-41,58-68) */
  int K = m->nwin;
  m->nVar = K * (m->ds + m->dm) + (K - 1) * (m->ds + m->dc);
  m->nCon = K * m->dm + (K - 1) * (m->ds + m->dc);
}

/* DecentralEst.cpp:200-351 */
static void initialize_mhe(orc_mhe *m) {
  int ds = m->ds, dm = m->dm;
  get_measurement(m, 0); /* :219 */
  const double *R_sb = m->stack[m->nstack - 1].R; /* :225 */
  la_zero(m->x_prior, ds);                        /* :232-234 */
  la_zero(m->Q_prior, ds * ds);
  for (int c = 0; c < 3; ++c) { /* :239-253 */
    m->Q_prior[(0 + c) * ds + 0 + c] = 1 / pow(m->P.p_init_std[c], 2);
    m->Q_prior[(3 + c) * ds + 3 + c] = 1 / pow(m->P.v_init_std[c], 2);
    m->Q_prior[(6 + c) * ds + 6 + c] = 1 / pow(m->P.accel_bias_init_std[c], 2);
  }
  build_meas(m, R_sb, m->b_meas, m->Q_meas_last, 0); /* :269-331 */
  if (m->leg_odom_type == 1) {
    for (int i = 0; i < m->nlegs; ++i)
      for (int c = 0; c < 3; ++c) {
        m->x_prior[9 + i * 3 + c] = m->b_meas[i * 3 + c];                                   /* :321 */
        m->Q_prior[(9 + 3 * i + c) * ds + 9 + 3 * i + c] = 1 / pow(m->P.foot_init_std[c], 2); /* :322 */
      }
  }
  /* :336-350: x_0, Prior_0, v_0, Measurement_0, updateQP(0) */
  stage_t *s = win_push(m, 0);
  la_copy(s->Hx, m->Q_prior, ds * ds); /* H_00 = I' Q I (MheSrb.cpp:393) */
  double g0[64];
  la_mm(g0, m->Q_prior, m->x_prior, ds, ds, 1);
  for (int i = 0; i < ds; ++i) s->gx[i] = -g0[i]; /* g = -A' Q b (MheSrb.cpp:397) */
  m->prior_in_regs = 1;
  la_copy(s->Q_meas, m->Q_meas_last, dm * dm);
  la_copy(s->meas_lb, m->b_meas, dm);
  la_copy(s->meas_ub, m->b_meas, dm);
  la_copy(m->lb_all, m->b_meas, dm); /* MheSrb.cpp:429-430 */
  la_copy(m->ub_all, m->b_meas, dm);
  recount(m);
}

/* DecentralEst.cpp:353-585 */
static void update_mhe(orc_mhe *m, int T) {
  int ds = m->ds, dm = m->dm;
  double dt = m->dt;
  stage_t *sp = &m->win[m->nwin - 1]; /* stage T-1 */
  const hist_t *bk = &m->stack[m->nstack - 1];
  double R_sb[9], accel_s[3];
  memcpy(R_sb, bk->R, sizeof(R_sb));          /* :374 */
  memcpy(accel_s, bk->accel_s, sizeof(accel_s)); /* :375 */
  /* b_dyn :387-388 */
  double b_dyn[64];
  la_zero(b_dyn, ds);
  for (int c = 0; c < 3; ++c) {
    b_dyn[c] = -dt * dt / 2 * accel_s[c];
    b_dyn[3 + c] = -dt * accel_s[c];
  }
  /* A_dyn :395-398 */
  la_eye(sp->A_dyn, ds);
  for (int r = 0; r < 3; ++r) {
    sp->A_dyn[r * ds + 3 + r] = dt;
    for (int c = 0; c < 3; ++c) {
      sp->A_dyn[r * ds + 6 + c] = -dt * dt / 2 * R_sb[r * 3 + c];
      sp->A_dyn[(3 + r) * ds + 6 + c] = -dt * R_sb[r * 3 + c];
    }
  }
  /* Q_dyn :407-424 */
  la_zero(sp->Q_dyn, ds * ds);
  double G[36], C[36], GC[36], GCG[36], Qpv[36];
  la_zero(G, 36);
  la_zero(C, 36);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      G[r * 6 + c] = R_sb[r * 3 + c] * dt;
      G[r * 6 + 3 + c] = 0.5 * R_sb[r * 3 + c] * dt * dt;
      G[(3 + r) * 6 + 3 + c] = R_sb[r * 3 + c] * dt;
    }
  for (int c = 0; c < 3; ++c) {
    C[c * 6 + c] = m->C_p[c];
    C[(3 + c) * 6 + 3 + c] = m->C_accel[c];
  }
  la_mm(GC, G, C, 6, 6, 6);
  la_mmt(GCG, GC, G, 6, 6, 6);
  la_inverse(Qpv, GCG, 6); /* :418 */
  la_set_block(sp->Q_dyn, ds, 0, 0, Qpv, 6, 6);
  for (int c = 0; c < 3; ++c) sp->Q_dyn[(6 + c) * ds + 6 + c] = 1 / (dt * dt) * m->Q_accel_bias[c]; /* :422-424 */
  if (m->leg_odom_type == 1) { /* :432-452 */
    double Rt[9];
    la_transpose(Rt, R_sb, 3, 3);
    for (int i = 0; i < m->nlegs; ++i) {
      const double *Qd = bk->contact[i] ? m->Q_foot_slide : m->Q_foot_swing;
      double RQ[9], RQR[9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) RQ[r * 3 + c] = 1 / (dt * dt) * R_sb[r * 3 + c] * Qd[c];
      la_mm(RQR, RQ, Rt, 3, 3, 3);
      la_set_block(sp->Q_dyn, ds, 9 + i * 3, 9 + i * 3, RQR, 3, 3);
    }
  }
  la_copy(sp->dyn_lb, b_dyn, ds); /* :461 */
  la_copy(sp->dyn_ub, b_dyn, ds);
  sp->has_dyn = 1;
  /* VO placeholder :474-488 */
  {
    double RQ[9], Rt[9];
    la_transpose(Rt, R_sb, 3, 3);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) RQ[r * 3 + c] = R_sb[r * 3 + c] * m->Q_vo_p[c];
    la_mm(sp->Q_cam, RQ, Rt, 3, 3, 3); /* :477 */
    for (int c = 0; c < 3; ++c) {
      sp->vo_lb[c] = -OSQP_INFTY; /* :481 */
      sp->vo_ub[c] = OSQP_INFTY;
    }
    sp->vo_equality = 0;
  }

  get_measurement(m, T); /* :490 */

  const double *R_new = m->stack[m->nstack - 1].R; /* :496 */
  build_meas(m, R_new, m->b_meas, m->Q_meas_last, 0);
  stage_t *sn = win_push(m, T);
  sp = &m->win[m->nwin - 2];
  la_copy(sn->Q_meas, m->Q_meas_last, dm * dm);
  la_copy(sn->meas_lb, m->b_meas, dm); /* :575 */
  la_copy(sn->meas_ub, m->b_meas, dm);

  /* updateQP(T): append rows [Dyn_{T-1}, VO_{T-1}, Meas_T] (MheSrb.cpp:420-439) */
  int r0 = m->nCon;
  la_copy(&m->lb_all[r0], sp->dyn_lb, ds);
  la_copy(&m->ub_all[r0], sp->dyn_ub, ds);
  la_copy(&m->lb_all[r0 + ds], sp->vo_lb, 3);
  la_copy(&m->ub_all[r0 + ds], sp->vo_ub, 3);
  la_copy(&m->lb_all[r0 + ds + 3], sn->meas_lb, dm);
  la_copy(&m->ub_all[r0 + ds + 3], sn->meas_ub, dm);
  recount(m);
}

/* DecentralEst.cpp:987-1009 + MheSrb.cpp:233-243, :449-459 */
static void update_vo_constraints(orc_mhe *m, int T) {
  (void)T;
  int stride = m->dm + m->ds + m->dc;
  int ins = m->vo_insert_idx[m->nvo_insert - 1];
  int disc0 = m->vo_insert_disc[m->nvo_insert - 1];
  for (int i = 0; i < m->bz.nnodes - 1; ++i) {
    const double *pose = &m->bz.distances[3 * (i + 1)]; /* :995 */
    int idx = (ins + i) * stride + m->dm + m->ds;       /* :997-999 */
    int name_k = disc0 + i;                             /* :1004 */
    /* updateConstraintBound(name, -pose, -pose, true): only if the named constraint exists */
    for (int j = 0; j < m->nwin; ++j)
      if (m->win[j].k == name_k && m->win[j].has_dyn) {
        for (int c = 0; c < 3; ++c) {
          m->win[j].vo_lb[c] = -pose[c];
          m->win[j].vo_ub[c] = -pose[c];
        }
        m->win[j].vo_equality = 1;
      }
    /* Update_Image_bound writes by row index (no existence check in the reference) */
    if (idx >= 0 && idx + 3 <= m->nCon)
      for (int c = 0; c < 3; ++c) {
        m->lb_all[idx + c] = -pose[c];
        m->ub_all[idx + c] = -pose[c];
      }
    else
      fprintf(stderr, "[oracle] Update_Image_bound row %d outside [0,%d): the reference would write out of range\n",
              idx, m->nCon);
  }
}

/* MheSrb.cpp:475-713 */
static void marginalize_qp(orc_mhe *m, int Tm) {
  int ds = m->ds, dm = m->dm, dc = m->dc;
  stage_t *s0 = &m->win[0];
  if (s0->k != Tm || !s0->has_dyn) {
    fprintf(stderr, "[oracle] Error missing dynamic cost at:%d\n", Tm);
    return;
  }
  if (m->prior_in_regs) { /* :517-522 */
    la_copy(m->M_p, m->Q_prior, ds * ds);
    double t[64];
    la_mm(t, m->M_p, m->x_prior, ds, ds, 1);
    for (int i = 0; i < ds; ++i) m->n_p[i] = -t[i];
    m->prior_in_regs = 0;
  }
  double *M_inv = dalloc(ds * ds);
  la_spd_inverse(M_inv, m->M_p, ds); /* :524-525 */
  int na = s0->vo_equality ? ds + dc : ds; /* rows of A_marginalize */
  int nt = na + dm;
  double *Am = dalloc(na * ds), *Qm = dalloc(na * na), *bm = dalloc(na);
  la_set_block(Am, ds, 0, 0, s0->A_dyn, ds, ds);   /* :543 / :607 */
  la_set_block(Qm, na, 0, 0, s0->Q_dyn, ds, ds);   /* :548 / :606 */
  for (int i = 0; i < ds; ++i) bm[i] = -s0->dyn_lb[i]; /* :539 / :608 */
  if (s0->vo_equality) {
    for (int c = 0; c < 3; ++c) Am[(ds + c) * ds + c] = 1.0; /* A_cam_pre :544 */
    la_set_block(Qm, na, ds, ds, s0->Q_cam, 3, 3);           /* :549 */
    for (int c = 0; c < 3; ++c) bm[ds + c] = -s0->vo_lb[c];  /* :535 */
  }
  double *Qm_inv = dalloc(na * na), *R_inv = dalloc(dm * dm);
  la_spd_inverse(Qm_inv, Qm, na);        /* :555-556 / :610-611 */
  la_spd_inverse(R_inv, s0->Q_meas, dm); /* :558-559 / :613-614 */
  const double *Hm = m->A_meas;          /* depVarMap[x] of Measurement :530 */
  double *AMi = dalloc(na * ds), *HMi = dalloc(dm * ds);
  la_mm(AMi, Am, M_inv, na, ds, ds);
  la_mm(HMi, Hm, M_inv, dm, ds, ds);
  double *A11 = dalloc(na * na), *A22 = dalloc(dm * dm), *A12 = dalloc(na * dm);
  la_mmt(A11, AMi, Am, na, ds, na);
  la_mmt(A22, HMi, Hm, dm, ds, dm);
  la_mmt(A12, AMi, Hm, na, ds, dm);
  double *A = dalloc(nt * nt);
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < na; ++j) A[i * nt + j] = -A11[i * na + j] - Qm_inv[i * na + j]; /* :565-567 */
  for (int i = 0; i < dm; ++i)
    for (int j = 0; j < dm; ++j) A[(na + i) * nt + na + j] = -A22[i * dm + j] - R_inv[i * dm + j]; /* :569-570 */
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < dm; ++j) {
      A[i * nt + na + j] = -A12[i * dm + j]; /* :572 */
      A[(na + j) * nt + i] = -A12[i * dm + j]; /* :573 */
    }
  double *B = dalloc(nt * ds); /* :581-583 / :635-636 */
  for (int i = 0; i < ds; ++i) B[i * ds + i] = -1.0;
  if (s0->vo_equality)
    for (int c = 0; c < 3; ++c) B[(ds + c) * ds + c] = -1.0;
  double *A_inv = dalloc(nt * nt);
  la_inverse(A_inv, A, nt); /* :588 / :640 */
  double *u = dalloc(nt), *Min_n = dalloc(ds), *t1 = dalloc(nt);
  la_mm(Min_n, M_inv, m->n_p, ds, ds, 1);
  la_mm(t1, Am, Min_n, na, ds, 1);
  for (int i = 0; i < na; ++i) u[i] = -bm[i] + t1[i]; /* :591 / :643 */
  la_mm(t1, Hm, Min_n, dm, ds, 1);
  for (int i = 0; i < dm; ++i) u[na + i] = s0->meas_lb[i] + t1[i]; /* :592 / :644 */
  /* M_p_next = -C A_inv B ; n_p_next = C A_inv u ; C = B' */
  double *AiB = dalloc(nt * ds), *Aiu = dalloc(nt), *Mn = dalloc(ds * ds), *nn = dalloc(ds);
  la_mm(AiB, A_inv, B, nt, nt, ds);
  la_mm(Aiu, A_inv, u, nt, nt, 1);
  la_mtm(Mn, B, AiB, ds, nt, ds);
  la_mtm(nn, B, Aiu, ds, nt, 1);
  for (int i = 0; i < ds * ds; ++i) m->M_p[i] = -Mn[i];
  for (int i = 0; i < ds; ++i) m->n_p[i] = nn[i];
  m->have_arrival = 1;
  /* H = H[33:,33:] + pad(M_p); g = g[33:] + pad(n_p)  (:505-508, :654-668) */
  stage_t *s1 = &m->win[1];
  for (int i = 0; i < ds * ds; ++i) s1->Hx[i] += m->M_p[i];
  for (int i = 0; i < ds; ++i) s1->gx[i] += m->n_p[i];
  /* rows: drop the first dyn+vo+meas rows (:684-706) */
  int drop = ds + dc + dm;
  memmove(m->lb_all, m->lb_all + drop, sizeof(double) * (size_t)(m->nCon - drop));
  memmove(m->ub_all, m->ub_all + drop, sizeof(double) * (size_t)(m->nCon - drop));
  stage_free(s0);
  memmove(&m->win[0], &m->win[1], sizeof(stage_t) * (size_t)(m->nwin - 1));
  m->nwin--;
  memset(&m->win[m->nwin], 0, sizeof(stage_t));
  recount(m);
  free(M_inv);
  free(Am);
  free(Qm);
  free(bm);
  free(Qm_inv);
  free(R_inv);
  free(AMi);
  free(HMi);
  free(A11);
  free(A22);
  free(A12);
  free(A);
  free(B);
  free(A_inv);
  free(u);
  free(Min_n);
  free(t1);
  free(AiB);
  free(Aiu);
  free(Mn);
  free(nn);
}

/* offsets in the reference ordering */
static int var_x(const orc_mhe *m, int j) { return j * (2 * m->ds + m->dm + m->dc); }
static int var_v(const orc_mhe *m, int j) { return var_x(m, j) + m->ds; }
static int var_w(const orc_mhe *m, int j) { return var_v(m, j) + m->dm; }
static int var_c(const orc_mhe *m, int j) { return var_w(m, j) + m->ds; }
static int row_meas(const orc_mhe *m, int j) { return j * (m->dm + m->ds + m->dc); }
static int row_dyn(const orc_mhe *m, int j) { return row_meas(m, j) + m->dm; }
static int row_vo(const orc_mhe *m, int j) { return row_dyn(m, j) + m->ds; }

void orc_mhe_export_qp(const orc_mhe *m, double *H, double *g, double *A, double *l, double *u) {
  int ds = m->ds, dm = m->dm, nV = m->nVar, nC = m->nCon;
  la_zero(H, nV * nV);
  la_zero(g, nV);
  la_zero(A, nC * nV);
  for (int j = 0; j < m->nwin; ++j) {
    const stage_t *s = &m->win[j];
    la_set_block(H, nV, var_x(m, j), var_x(m, j), s->Hx, ds, ds);
    la_copy(&g[var_x(m, j)], s->gx, ds);
    la_set_block(H, nV, var_v(m, j), var_v(m, j), s->Q_meas, dm, dm);
    /* Measurement_j: A_meas x_j - v_j */
    la_set_block(A, nV, row_meas(m, j), var_x(m, j), m->A_meas, dm, ds);
    for (int i = 0; i < dm; ++i) A[(row_meas(m, j) + i) * nV + var_v(m, j) + i] = -1.0;
    if (j + 1 < m->nwin) {
      la_set_block(H, nV, var_w(m, j), var_w(m, j), s->Q_dyn, ds, ds);
      la_set_block(H, nV, var_c(m, j), var_c(m, j), s->Q_cam, 3, 3);
      la_set_block(A, nV, row_dyn(m, j), var_x(m, j), s->A_dyn, ds, ds);
      for (int i = 0; i < ds; ++i) {
        A[(row_dyn(m, j) + i) * nV + var_w(m, j) + i] = -1.0;
        A[(row_dyn(m, j) + i) * nV + var_x(m, j + 1) + i] = -1.0;
      }
      for (int c = 0; c < 3; ++c) {
        A[(row_vo(m, j) + c) * nV + var_x(m, j) + c] = 1.0;
        A[(row_vo(m, j) + c) * nV + var_x(m, j + 1) + c] = -1.0;
        A[(row_vo(m, j) + c) * nV + var_c(m, j) + c] = -1.0;
      }
    }
  }
  la_copy(l, m->lb_all, nC);
  la_copy(u, m->ub_all, nC);
}

/* Exact optimum of the assembled QP (stand-in for OSQP converged to eps->0): every row carries its
 * own slack with coefficient -I, so v = A_meas x - l etc.; free rows ([-1e30,1e30]) leave their
 * slack at 0.  Normal equations in x are block tridiagonal -> banded Cholesky. */
/* Equality-constrained solve of the window normal equations with the components flagged in act[] (+1: at its upper bound, -1: at its
 * lower bound) fixed: xs = argmin 1/2 x'N0 x - r0'x  s.t.  x[a] = bound(a).  Nm / rhs are scratch. */
static int active_solve(const double *N0, const double *r0, int n, int ds, const int *act, const double *blo, const double *bhi,
                        double *Nm, double *rhs, double *xs) {
  la_copy(Nm, N0, n * n);
  la_copy(rhs, r0, n);
  for (int a = 0; a < n; ++a) {
    if (!act[a]) continue;
    double beta = act[a] > 0 ? bhi[a % ds] : blo[a % ds];
    for (int i = 0; i < n; ++i) rhs[i] -= Nm[i * n + a] * beta;
  }
  for (int a = 0; a < n; ++a) {
    if (!act[a]) continue;
    double beta = act[a] > 0 ? bhi[a % ds] : blo[a % ds];
    for (int i = 0; i < n; ++i) Nm[i * n + a] = Nm[a * n + i] = 0.0;
    Nm[a * n + a] = 1.0;
    rhs[a] = beta;
  }
  int rc = la_chol_solve_banded(Nm, rhs, n, 2 * ds - 1);
  la_copy(xs, rhs, n);
  return rc;
}

/* Textbook PRIMAL active-set method on the same problem (finite and monotone on a strictly convex QP), used when the fast
 * primal-dual iteration has not settled: (1) bounds are only added until the iterate is feasible; (2) while some active bound has
 * a multiplier of the wrong sign, the worst one is dropped and the iterate moves towards the new equality-constrained minimiser,
 * stopping at (and adding) the first bound it would cross.  bmask: bounded components of every state.  Returns iterations used. */
static int primal_active_set(const double *N0, const double *r0, int n, int ds, int K, int bmask, const double *blo,
                             const double *bhi, int *act, double *Nm, double *rhs, double *xs) {
  double *xc = dalloc(n), *xn = dalloc(n);
  int iters = 0;
  for (;;) { /* phase 1: feasibility */
    active_solve(N0, r0, n, ds, act, blo, bhi, Nm, rhs, xs);
    iters++;
    int added = 0;
    for (int j = 0; j < K; ++j)
      for (int c = 0; c < ds && c < 9; ++c) {
        if (!((bmask >> c) & 1)) continue;
        int a = j * ds + c;
        if (act[a]) continue;
        if (xs[a] > bhi[c]) { act[a] = 1; added++; }
        else if (xs[a] < blo[c]) { act[a] = -1; added++; }
      }
    if (!added || iters > 10 * n) break;
  }
  la_copy(xc, xs, n);
  for (int outer = 0; outer < 50 * n; ++outer) { /* phase 2 */
    int drop = -1;
    double worst = 0.0;
    for (int a = 0; a < n; ++a) {
      if (!act[a]) continue;
      double grad = -r0[a], gs = fabs(r0[a]);
      for (int i = 0; i < n; ++i) {
        grad += N0[a * n + i] * xc[i];
        gs += fabs(N0[a * n + i] * xc[i]);
      }
      double mult = act[a] > 0 ? -grad : grad;
      double v = mult / (gs > 0.0 ? gs : 1.0);
      if (v < -1e-10 && (drop < 0 || v < worst)) { worst = v; drop = a; }
    }
    if (drop < 0) break;
    act[drop] = 0;
    for (int inner = 0; inner < 10 * n; ++inner) {
      active_solve(N0, r0, n, ds, act, blo, bhi, Nm, rhs, xn);
      iters++;
      double alpha = 1.0;
      int block = -1, side = 0;
      for (int j = 0; j < K; ++j)
        for (int c = 0; c < ds && c < 9; ++c) {
          if (!((bmask >> c) & 1)) continue;
          int a = j * ds + c;
          if (act[a]) continue;
          double d = xn[a] - xc[a];
          if (d > 0.0 && xn[a] > bhi[c]) {
            double t = (bhi[c] - xc[a]) / d;
            if (t < alpha) { alpha = t < 0.0 ? 0.0 : t; block = a; side = 1; }
          } else if (d < 0.0 && xn[a] < blo[c]) {
            double t = (blo[c] - xc[a]) / d;
            if (t < alpha) { alpha = t < 0.0 ? 0.0 : t; block = a; side = -1; }
          }
        }
      for (int i = 0; i < n; ++i) xc[i] += alpha * (xn[i] - xc[i]);
      if (block < 0) break;
      act[block] = side;
    }
  }
  la_copy(xs, xc, n);
  free(xc);
  free(xn);
  return iters;
}

/* General rows lo <= a_i . x_k <= hi (orc_params.x_row_*), together with the component bounds of v_box / x_box as unit rows:
 * a change of basis y = W x whose first m coordinates ARE the rows turns them into component bounds on y.  W = [rows; unit vectors
 * that complete them to a basis (greedy: the unit vector farthest from the span so far)], V = W^-1 by Gauss-Jordan.  Returns m, or
 * -1 if the rows are linearly dependent / too many. */
static int rows_basis(const orc_params *P, double *W /*81*/, double *V /*81*/, double *lo /*9*/, double *hi /*9*/) {
  int m = 0;
  double Q[81]; /* orthonormal basis of the span so far, row-major */
  int nq = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int cnt = pass == 0 ? P->x_row_count : 9;
    for (int i = 0; i < cnt; ++i) {
      double a[9];
      double l, u;
      if (pass == 0) {
        for (int c = 0; c < 9; ++c) a[c] = P->x_row_a[i * 9 + c];
        l = P->x_row_lo[i];
        u = P->x_row_hi[i];
      } else {
        int gx = (P->x_box_mask >> i) & 1, gv = P->v_box_enable && i >= 3 && i < 6;
        if (!gx && !gv) continue;
        for (int c = 0; c < 9; ++c) a[c] = (c == i) ? 1.0 : 0.0;
        l = gx ? P->x_box_lo[i] : P->v_box_lo[i - 3];
        u = gx ? P->x_box_hi[i] : P->v_box_hi[i - 3];
      }
      if (m >= 9) return -1;
      double r[9], na = 0.0, nr = 0.0;
      for (int c = 0; c < 9; ++c) { r[c] = a[c]; na += a[c] * a[c]; }
      for (int q = 0; q < nq; ++q) {
        double d = 0.0;
        for (int c = 0; c < 9; ++c) d += Q[q * 9 + c] * r[c];
        for (int c = 0; c < 9; ++c) r[c] -= d * Q[q * 9 + c];
      }
      for (int c = 0; c < 9; ++c) nr += r[c] * r[c];
      if (!(nr > 1e-16 * na) || na == 0.0) return -1;
      for (int c = 0; c < 9; ++c) Q[nq * 9 + c] = r[c] / sqrt(nr);
      nq++;
      for (int c = 0; c < 9; ++c) W[m * 9 + c] = a[c] / sqrt(na); /* unit-length rows: bounds scale along */
      lo[m] = l / sqrt(na);
      hi[m] = u / sqrt(na);
      m++;
    }
  }
  int rows = m;
  while (nq < 9) { /* completion */
    int best = -1;
    double bestn = -1.0, br[9];
    for (int k = 0; k < 9; ++k) {
      double r[9], nr = 0.0;
      for (int c = 0; c < 9; ++c) r[c] = (c == k) ? 1.0 : 0.0;
      for (int q = 0; q < nq; ++q) {
        double d = Q[q * 9 + k];
        for (int c = 0; c < 9; ++c) r[c] -= d * Q[q * 9 + c];
      }
      for (int c = 0; c < 9; ++c) nr += r[c] * r[c];
      if (nr > bestn) { bestn = nr; best = k; for (int c = 0; c < 9; ++c) br[c] = r[c]; }
    }
    (void)best;
    for (int c = 0; c < 9; ++c) Q[nq * 9 + c] = br[c] / sqrt(bestn);
    for (int c = 0; c < 9; ++c) W[nq * 9 + c] = Q[nq * 9 + c]; /* orthonormal completion: cond(W) = cond(rows) */
    nq++;
  }
  /* V = W^-1 */
  double A[9][18];
  for (int r = 0; r < 9; ++r)
    for (int c = 0; c < 9; ++c) { A[r][c] = W[r * 9 + c]; A[r][9 + c] = (r == c) ? 1.0 : 0.0; }
  for (int c = 0; c < 9; ++c) {
    int pv = c;
    for (int r = c + 1; r < 9; ++r) if (fabs(A[r][c]) > fabs(A[pv][c])) pv = r;
    if (A[pv][c] == 0.0) return -1;
    if (pv != c) for (int k = 0; k < 18; ++k) { double t = A[c][k]; A[c][k] = A[pv][k]; A[pv][k] = t; }
    double d = 1.0 / A[c][c];
    for (int k = 0; k < 18; ++k) A[c][k] *= d;
    for (int r = 0; r < 9; ++r) {
      if (r == c) continue;
      double f = A[r][c];
      if (f != 0.0) for (int k = 0; k < 18; ++k) A[r][k] -= f * A[c][k];
    }
  }
  for (int r = 0; r < 9; ++r)
    for (int c = 0; c < 9; ++c) V[r * 9 + c] = A[r][9 + c];
  return rows;
}

static int solve_exact(orc_mhe *m) {
  int ds = m->ds, dm = m->dm, K = m->nwin, n = K * ds;
  double *Nm = dalloc(n * n), *rhs = dalloc(n);
  double *t = dalloc(ds * ds > dm * ds ? ds * ds : dm * ds), *t2 = dalloc(ds * ds), *tv = dalloc(2 * ds + dm);
  for (int j = 0; j < K; ++j) {
    const stage_t *s = &m->win[j];
    int o = j * ds;
    for (int a = 0; a < ds; ++a) {
      for (int b = 0; b < ds; ++b) Nm[(o + a) * n + o + b] += s->Hx[a * ds + b];
      rhs[o + a] -= s->gx[a];
    }
    /* measurement */
    const double *bj = &m->lb_all[row_meas(m, j)];
    la_mtm(t, m->A_meas, s->Q_meas, ds, dm, dm); /* A' Q : ds x dm */
    la_mm(t2, t, m->A_meas, ds, dm, ds);
    la_mm(tv, t, bj, ds, dm, 1);
    for (int a = 0; a < ds; ++a) {
      for (int b = 0; b < ds; ++b) Nm[(o + a) * n + o + b] += t2[a * ds + b];
      rhs[o + a] += tv[a];
    }
    if (j + 1 < K) {
      int o1 = o + ds;
      const double *bd = &m->lb_all[row_dyn(m, j)];
      la_mtm(t, s->A_dyn, s->Q_dyn, ds, ds, ds); /* A'Q */
      la_mm(t2, t, s->A_dyn, ds, ds, ds);        /* A'QA */
      la_mm(tv, t, bd, ds, ds, 1);               /* A'Q b */
      double *Qb = tv + ds;
      la_mm(Qb, s->Q_dyn, bd, ds, ds, 1);
      for (int a = 0; a < ds; ++a) {
        for (int b = 0; b < ds; ++b) {
          Nm[(o + a) * n + o + b] += t2[a * ds + b];
          Nm[(o1 + a) * n + o1 + b] += s->Q_dyn[a * ds + b];
          Nm[(o + a) * n + o1 + b] -= t[a * ds + b];
          Nm[(o1 + b) * n + o + a] -= t[a * ds + b];
        }
        rhs[o + a] += tv[a];
        rhs[o1 + a] -= Qb[a];
      }
      const double *lv = &m->lb_all[row_vo(m, j)], *uv = &m->ub_all[row_vo(m, j)];
      int eq = (lv[0] == uv[0]) && (lv[1] == uv[1]) && (lv[2] == uv[2]);
      if (eq) {
        double Qc_c[3];
        la_mm(Qc_c, s->Q_cam, lv, 3, 3, 1);
        for (int a = 0; a < 3; ++a) {
          for (int b = 0; b < 3; ++b) {
            double q = s->Q_cam[a * 3 + b];
            Nm[(o + a) * n + o + b] += q;
            Nm[(o1 + a) * n + o1 + b] += q;
            Nm[(o + a) * n + o1 + b] -= q;
            Nm[(o1 + a) * n + o + b] -= q;
          }
          rhs[o + a] += Qc_c[a];
          rhs[o1 + a] -= Qc_c[a];
        }
      }
    }
  }
  int rc;
  /* general rows: solve in y = W x (block-diagonal congruence of the normal equations), where the rows are component bounds */
  double Wm[81], Vm[81], rlo[9], rhi[9];
  int nrows = 0;
  if (m->P.x_row_count > 0 && ds == 9) {
    nrows = rows_basis(&m->P, Wm, Vm, rlo, rhi);
    if (nrows > 0) {
      double *blk = dalloc(81), *tmpb = dalloc(81);
      for (int j = 0; j < K; ++j) {
        for (int j2 = (j > 0 ? j - 1 : 0); j2 <= (j + 1 < K ? j + 1 : j); ++j2) {
          for (int a = 0; a < 9; ++a)
            for (int b = 0; b < 9; ++b) blk[a * 9 + b] = Nm[(j * 9 + a) * n + j2 * 9 + b];
          for (int a = 0; a < 9; ++a) /* tmp = V' blk */
            for (int b = 0; b < 9; ++b) {
              double v = 0.0;
              for (int k = 0; k < 9; ++k) v += Vm[k * 9 + a] * blk[k * 9 + b];
              tmpb[a * 9 + b] = v;
            }
          for (int a = 0; a < 9; ++a) /* N' = tmp V */
            for (int b = 0; b < 9; ++b) {
              double v = 0.0;
              for (int k = 0; k < 9; ++k) v += tmpb[a * 9 + k] * Vm[k * 9 + b];
              Nm[(j * 9 + a) * n + j2 * 9 + b] = v;
            }
        }
        double rv[9];
        for (int a = 0; a < 9; ++a) {
          double v = 0.0;
          for (int k = 0; k < 9; ++k) v += Vm[k * 9 + a] * rhs[j * 9 + k];
          rv[a] = v;
        }
        for (int a = 0; a < 9; ++a) rhs[j * 9 + a] = rv[a];
      }
      /* symmetrise what round-off left */
      for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b) {
          double v = 0.5 * (Nm[a * n + b] + Nm[b * n + a]);
          Nm[a * n + b] = Nm[b * n + a] = v;
        }
      free(blk);
      free(tmpb);
    }
  }
  if (m->P.v_box_enable || m->P.x_box_mask || nrows > 0) {
    /* Builder extension (never exercised by the reference): rows lo <= x_j[a] <= hi for the bounded components a of every
     * state of the window at solve time -- what MHEproblem::addConstraints(name, lb, ub) + a dependency on x_j with a
     * selector row would add (MheSrb.cpp:58-68).  v_box bounds v_s (a = 3..5), x_box_mask any of the 9 base components.
     * Exact optimum by a primal-dual active-set iteration on the normal equations in x: fix the active components at their
     * bound, solve, read the multipliers off the residual of the free system, update the set; stops when the set repeats. */
    double blo[9], bhi[9];
    int bmask = 0;
    for (int a = 0; a < 9; ++a) {
      int gx = (m->P.x_box_mask >> a) & 1, gv = m->P.v_box_enable && a >= 3 && a < 6;
      blo[a] = gx ? m->P.x_box_lo[a] : (gv ? m->P.v_box_lo[a - 3] : -1e300);
      bhi[a] = gx ? m->P.x_box_hi[a] : (gv ? m->P.v_box_hi[a - 3] : 1e300);
      if (gx || gv) bmask |= 1 << a;
    }
    if (nrows > 0) { /* in y coordinates the first nrows components carry every bound (component bounds became unit rows) */
      bmask = (1 << nrows) - 1;
      for (int a = 0; a < 9; ++a) {
        blo[a] = a < nrows ? rlo[a] : -1e300;
        bhi[a] = a < nrows ? rhi[a] : 1e300;
      }
    }
    const int general = m->P.x_box_mask || nrows > 0;
    double *N0 = dalloc(n * n), *r0 = dalloc(n), *xs = dalloc(n);
    int *act = (int *)calloc((size_t)n, sizeof(int)), *nact = (int *)calloc((size_t)n, sizeof(int));
    la_copy(N0, Nm, n * n);
    la_copy(r0, rhs, n);
    rc = 0;
    int it = 0;
    int settled = 0;
    for (it = 0; it < (general ? 40 : 200); ++it) {
      la_copy(Nm, N0, n * n);
      la_copy(rhs, r0, n);
      for (int a = 0; a < n; ++a) {
        if (!act[a]) continue;
        double beta = act[a] > 0 ? bhi[a % ds] : blo[a % ds];
        for (int i = 0; i < n; ++i) rhs[i] -= Nm[i * n + a] * beta;
      }
      for (int a = 0; a < n; ++a) {
        if (!act[a]) continue;
        double beta = act[a] > 0 ? bhi[a % ds] : blo[a % ds];
        for (int i = 0; i < n; ++i) Nm[i * n + a] = Nm[a * n + i] = 0.0;
        Nm[a * n + a] = 1.0;
        rhs[a] = beta;
      }
      rc = la_chol_solve_banded(Nm, rhs, n, 2 * ds - 1);
      la_copy(xs, rhs, n);
      int changed = 0;
      /* general bounds: after 8 plain primal-dual iterations bounds are only added, and when nothing is violated the single
       * bound with the most negative (scaled) multiplier is dropped -- the plain rule can cycle when bounds of several
       * components interact (the velocity box never did) */
      int safe = general && it >= 8, added = 0, drop_a = -1;
      double drop_val = 0.0;
      for (int j = 0; j < K; ++j)
        for (int c = 0; c < 9; ++c) {
          if (!((bmask >> c) & 1)) continue;
          int a = j * ds + c;
          double grad = -r0[a], gs = fabs(r0[a]);
          for (int i = 0; i < n; ++i) {
            grad += N0[a * n + i] * xs[i];
            gs += fabs(N0[a * n + i] * xs[i]);
          }
          /* a multiplier within round-off of zero keeps its bound (general bounds only); the velocity box keeps the exact
           * sign test it was validated with */
          double gtol = general ? 1e-10 * gs : 0.0;
          int na;
          if (act[a] != 0) {
            double mult = act[a] > 0 ? -grad : grad;
            int keep = mult > -gtol;
            if (!keep && safe) {
              double v = mult / (gs > 0.0 ? gs : 1.0);
              if (drop_a < 0 || v < drop_val) {
                drop_val = v;
                drop_a = a;
              }
              keep = 1;
            }
            na = keep ? act[a] : 0;
          } else {
            na = (xs[a] > bhi[c]) ? 1 : ((xs[a] < blo[c]) ? -1 : 0);
            if (na) added++;
          }
          nact[a] = na;
          if (na != act[a]) changed = 1;
        }
      if (safe && added == 0 && drop_a >= 0) {
        nact[drop_a] = 0;
        changed = 1;
      }
      if (!changed) {
        settled = 1;
        break;
      }
      memcpy(act, nact, sizeof(int) * (size_t)n);
    }
    if (general && !settled) /* the fast iteration cycled: the finite method, warm-started from its last set */
      it += primal_active_set(N0, r0, n, ds, K, bmask, blo, bhi, act, Nm, rhs, xs);
    m->admm_iters = it + 1;
    la_copy(rhs, xs, n);
    if (nrows > 0) /* back to x = V y */
      for (int j = 0; j < K; ++j) {
        double xv[9];
        for (int a = 0; a < 9; ++a) {
          double v = 0.0;
          for (int k = 0; k < 9; ++k) v += Vm[a * 9 + k] * xs[j * 9 + k];
          xv[a] = v;
        }
        for (int a = 0; a < 9; ++a) rhs[j * 9 + a] = xv[a];
      }
    free(N0);
    free(r0);
    free(xs);
    free(act);
    free(nact);
  } else {
    rc = la_chol_solve_banded(Nm, rhs, n, 2 * ds - 1);
  }
  /* full primal in reference ordering */
  la_zero(m->solution, m->nVar);
  for (int j = 0; j < K; ++j) {
    const stage_t *s = &m->win[j];
    const double *xj = &rhs[j * ds];
    la_copy(&m->solution[var_x(m, j)], xj, ds);
    la_mm(tv, m->A_meas, xj, dm, ds, 1);
    for (int i = 0; i < dm; ++i) m->solution[var_v(m, j) + i] = tv[i] - m->lb_all[row_meas(m, j) + i];
    if (j + 1 < K) {
      const double *x1 = &rhs[(j + 1) * ds];
      la_mm(tv, s->A_dyn, xj, ds, ds, 1);
      for (int i = 0; i < ds; ++i) m->solution[var_w(m, j) + i] = tv[i] - x1[i] - m->lb_all[row_dyn(m, j) + i];
      const double *lv = &m->lb_all[row_vo(m, j)], *uv = &m->ub_all[row_vo(m, j)];
      int eq = (lv[0] == uv[0]) && (lv[1] == uv[1]) && (lv[2] == uv[2]);
      for (int c = 0; c < 3; ++c) m->solution[var_c(m, j) + c] = eq ? (xj[c] - x1[c] - lv[c]) : 0.0;
    }
  }
  free(Nm);
  free(rhs);
  free(t);
  free(t2);
  free(tv);
  return rc;
}

/* sparse (triplet) form of orc_mhe_export_qp for the ADMM mode: same entries, nonzeros only */
typedef struct {
  int n, cap;
  int *i, *j;
  double *x;
} trip_t;
static void trip_put(trip_t *t, int i, int j, double v) {
  if (t->n == t->cap) {
    t->cap = t->cap ? 2 * t->cap : 4096;
    t->i = (int *)realloc(t->i, sizeof(int) * (size_t)t->cap);
    t->j = (int *)realloc(t->j, sizeof(int) * (size_t)t->cap);
    t->x = (double *)realloc(t->x, sizeof(double) * (size_t)t->cap);
  }
  t->i[t->n] = i;
  t->j[t->n] = j;
  t->x[t->n] = v;
  t->n++;
}
static void trip_block_upper(trip_t *t, int o, const double *B, int n) {
  for (int a = 0; a < n; ++a)
    for (int b = a; b < n; ++b) trip_put(t, o + a, o + b, B[a * n + b]);
}
static void trip_block(trip_t *t, int r0, int c0, const double *B, int m, int n) {
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < n; ++b) trip_put(t, r0 + a, c0 + b, B[a * n + b]);
}

static void solve_qp(orc_mhe *m) {
  if (m->P.solve_mode == 2) {
    int ds = m->ds, dm = m->dm, nV = m->nVar, nC = m->nCon;
    trip_t P = {0, 0, NULL, NULL, NULL}, A = {0, 0, NULL, NULL, NULL};
    double *g = dalloc(nV);
    for (int j = 0; j < m->nwin; ++j) {
      const stage_t *s = &m->win[j];
      /* Patterns follow what the reference inserts (explicit zeros included, EigenUtils.hpp:36-49):
       * Q_meas / A_meas come from dense matrices; the arrival block M_p is dense; before the first
       * marginalisation the prior holds three dense 3x3 blocks (DecentralEst.cpp:251-253). */
      if (j == 0 && m->have_arrival)
        trip_block_upper(&P, var_x(m, j), s->Hx, ds);
      else if (j == 0)
        for (int b3 = 0; b3 < ds / 3; ++b3) {
          double blk[9];
          la_get_block(blk, s->Hx, ds, 3 * b3, 3 * b3, 3, 3);
          trip_block_upper(&P, var_x(m, j) + 3 * b3, blk, 3);
        }
      la_copy(&g[var_x(m, j)], s->gx, ds);
      trip_block_upper(&P, var_v(m, j), s->Q_meas, dm);
      trip_block(&A, row_meas(m, j), var_x(m, j), m->A_meas, dm, ds);
      for (int i = 0; i < dm; ++i) trip_put(&A, row_meas(m, j) + i, var_v(m, j) + i, -1.0);
      if (j + 1 < m->nwin) {
        { /* Q_dyn: dense 6x6 + 3x3 blocks (DecentralEst.cpp:420-424, :438-448) */
          double b6[36], b3[9];
          la_get_block(b6, s->Q_dyn, ds, 0, 0, 6, 6);
          trip_block_upper(&P, var_w(m, j), b6, 6);
          for (int q3 = 2; q3 < ds / 3; ++q3) {
            la_get_block(b3, s->Q_dyn, ds, 3 * q3, 3 * q3, 3, 3);
            trip_block_upper(&P, var_w(m, j) + 3 * q3, b3, 3);
          }
        }
        trip_block_upper(&P, var_c(m, j), s->Q_cam, 3);
        /* A_dyn: identity + three dense 3x3 blocks (DecentralEst.cpp:395-398) */
        for (int i = 0; i < ds; ++i) trip_put(&A, row_dyn(m, j) + i, var_x(m, j) + i, 1.0);
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) {
            trip_put(&A, row_dyn(m, j) + a, var_x(m, j) + 3 + b, s->A_dyn[a * ds + 3 + b]);
            trip_put(&A, row_dyn(m, j) + a, var_x(m, j) + 6 + b, s->A_dyn[a * ds + 6 + b]);
            trip_put(&A, row_dyn(m, j) + 3 + a, var_x(m, j) + 6 + b, s->A_dyn[(3 + a) * ds + 6 + b]);
          }
        for (int i = 0; i < ds; ++i) {
          trip_put(&A, row_dyn(m, j) + i, var_w(m, j) + i, -1.0);
          trip_put(&A, row_dyn(m, j) + i, var_x(m, j + 1) + i, -1.0);
        }
        for (int c = 0; c < 3; ++c) {
          trip_put(&A, row_vo(m, j) + c, var_x(m, j) + c, 1.0);
          trip_put(&A, row_vo(m, j) + c, var_x(m, j + 1) + c, -1.0);
          trip_put(&A, row_vo(m, j) + c, var_c(m, j) + c, -1.0);
        }
      }
    }
    if (m->P.v_box_enable) {
      /* builder extension: inequality rows lo <= v_s of x_j <= hi appended after the reference's rows */
      int nC2 = nC + 3 * m->nwin;
      double *l2 = dalloc(nC2), *u2 = dalloc(nC2);
      la_copy(l2, m->lb_all, nC);
      la_copy(u2, m->ub_all, nC);
      for (int j = 0; j < m->nwin; ++j)
        for (int c = 0; c < 3; ++c) {
          trip_put(&A, nC + 3 * j + c, var_x(m, j) + 3 + c, 1.0);
          l2[nC + 3 * j + c] = m->P.v_box_lo[c];
          u2[nC + 3 * j + c] = m->P.v_box_hi[c];
        }
      orc_admm_solve_triplets(nV, nC2, P.n, P.i, P.j, P.x, A.n, A.i, A.j, A.x, g, l2, u2, &m->P, m->solution,
                              &m->admm_iters);
      free(l2);
      free(u2);
    } else
      orc_admm_solve_triplets(nV, nC, P.n, P.i, P.j, P.x, A.n, A.i, A.j, A.x, g, m->lb_all, m->ub_all, &m->P,
                              m->solution, &m->admm_iters);
    free(g);
    free(P.i);
    free(P.j);
    free(P.x);
    free(A.i);
    free(A.j);
    free(A.x);
  } else {
    solve_exact(m);
  }
}

/* ------------------------------------------------------------------------- KF alternative */
/* DecentralEst.cpp:592-700 */
static void kf_measurement_update(orc_mhe *m, const double *C_meas) {
  int ds = m->ds, dm = m->dm;
  double *CAt = dalloc(ds * dm), *S = dalloc(dm * dm), *Sinv = dalloc(dm * dm), *K = dalloc(ds * dm);
  la_mmt(CAt, m->C_KF, m->A_meas, ds, ds, dm);
  la_mm(S, m->A_meas, CAt, dm, ds, dm);
  for (int i = 0; i < dm * dm; ++i) S[i] += C_meas[i];
  la_inverse(Sinv, S, dm);
  la_mm(K, CAt, Sinv, ds, dm, dm); /* :697 / :858 */
  double *Ax = dalloc(dm), *innov = dalloc(dm), *dx = dalloc(ds);
  la_mm(Ax, m->A_meas, m->x_KF, dm, ds, 1);
  for (int i = 0; i < dm; ++i) innov[i] = m->b_meas[i] - Ax[i];
  la_mm(dx, K, innov, ds, dm, 1);
  for (int i = 0; i < ds; ++i) m->x_KF[i] += dx[i]; /* :698 / :859 */
  double *KA = dalloc(ds * ds), *IKA = dalloc(ds * ds), *Cn = dalloc(ds * ds);
  la_mm(KA, K, m->A_meas, ds, dm, ds);
  for (int i = 0; i < ds; ++i)
    for (int j = 0; j < ds; ++j) IKA[i * ds + j] = (i == j ? 1.0 : 0.0) - KA[i * ds + j];
  la_mm(Cn, IKA, m->C_KF, ds, ds, ds); /* :699 / :860 */
  la_copy(m->C_KF, Cn, ds * ds);
  free(CAt);
  free(S);
  free(Sinv);
  free(K);
  free(Ax);
  free(innov);
  free(dx);
  free(KA);
  free(IKA);
  free(Cn);
}

static void initialize_kf(orc_mhe *m) {
  int ds = m->ds, dm = m->dm;
  get_measurement(m, 0); /* :594 */
  const double *R_sb = m->stack[m->nstack - 1].R;
  la_zero(m->x_prior, ds);
  double *C_prior = dalloc(ds * ds), *C_meas = dalloc(dm * dm);
  for (int c = 0; c < 3; ++c) { /* :618-624 */
    C_prior[(0 + c) * ds + c] = pow(m->P.p_init_std[c], 2);
    C_prior[(3 + c) * ds + 3 + c] = pow(m->P.v_init_std[c], 2);
    C_prior[(6 + c) * ds + 6 + c] = pow(m->P.accel_bias_init_std[c], 2);
  }
  build_meas(m, R_sb, m->b_meas, C_meas, 1); /* :636-690 */
  if (m->leg_odom_type == 1) {
    for (int i = 0; i < m->nlegs; ++i)
      for (int c = 0; c < 3; ++c) C_prior[(9 + i * 3 + c) * ds + 9 + i * 3 + c] = pow(m->P.foot_init_std[c], 2); /* :681 */
    for (int i = 0; i < dm; ++i) m->x_prior[9 + i] = m->b_meas[i]; /* :683 */
  }
  la_copy(m->x_KF, m->x_prior, ds); /* :693-694 */
  la_copy(m->C_KF, C_prior, ds * ds);
  kf_measurement_update(m, C_meas); /* :697-699 */
  free(C_prior);
  free(C_meas);
}

/* DecentralEst.cpp:702-861 */
static void update_kf(orc_mhe *m) {
  int ds = m->ds, dm = m->dm;
  double dt = m->dt;
  const hist_t *bk = &m->stack[m->nstack - 1];
  double R_sb[9], accel_s[3];
  memcpy(R_sb, bk->R, sizeof(R_sb));
  memcpy(accel_s, bk->accel_s, sizeof(accel_s));
  double *b_dyn = dalloc(ds), *A_dyn = dalloc(ds * ds), *G = dalloc(ds * ds), *Cin = dalloc(ds * ds);
  for (int c = 0; c < 3; ++c) { /* :716-717 */
    b_dyn[c] = -0.5 * dt * dt * accel_s[c];
    b_dyn[3 + c] = -dt * accel_s[c];
  }
  la_eye(A_dyn, ds); /* :727-731 */
  for (int r = 0; r < 3; ++r) {
    A_dyn[r * ds + 3 + r] = dt;
    for (int c = 0; c < 3; ++c) {
      A_dyn[r * ds + 6 + c] = -dt * dt / 2 * R_sb[r * 3 + c];
      A_dyn[(3 + r) * ds + 6 + c] = -dt * R_sb[r * 3 + c];
    }
  }
  for (int r = 0; r < 3; ++r) { /* :742-751 */
    for (int c = 0; c < 3; ++c) {
      G[r * ds + c] = R_sb[r * 3 + c] * dt;
      G[r * ds + 3 + c] = -0.5 * R_sb[r * 3 + c] * dt * dt;
      G[(3 + r) * ds + 3 + c] = -R_sb[r * 3 + c] * dt;
    }
    G[(6 + r) * ds + 6 + r] = dt;
    Cin[r * ds + r] = m->C_p[r];
    Cin[(3 + r) * ds + 3 + r] = m->C_accel[r];
    Cin[(6 + r) * ds + 6 + r] = m->C_accel_bias[r];
  }
  if (m->leg_odom_type == 1) { /* :759-776 */
    for (int i = 0; i < m->nlegs; ++i)
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) G[(9 + i * 3 + r) * ds + 9 + i * 3 + c] = R_sb[r * 3 + c] * dt;
        Cin[(9 + 3 * i + r) * ds + 9 + 3 * i + r] = (bk->contact[i] == 0.0) ? m->C_foot_swing[r] : m->C_foot_slide[r];
      }
  }
  double *x2 = dalloc(ds), *GC = dalloc(ds * ds), *C_dyn = dalloc(ds * ds), *AC = dalloc(ds * ds), *ACA = dalloc(ds * ds);
  la_mm(x2, A_dyn, m->x_KF, ds, ds, 1);
  for (int i = 0; i < ds; ++i) m->x_KF[i] = x2[i] - b_dyn[i]; /* :783 */
  la_mm(GC, G, Cin, ds, ds, ds);
  la_mmt(C_dyn, GC, G, ds, ds, ds); /* :784 */
  la_mm(AC, A_dyn, m->C_KF, ds, ds, ds);
  la_mmt(ACA, AC, A_dyn, ds, ds, ds);
  for (int i = 0; i < ds * ds; ++i) m->C_KF[i] = ACA[i] + C_dyn[i]; /* :785 */
  get_measurement(m, 0); /* :787 (sic: always 0) */
  const double *R_new = m->stack[m->nstack - 1].R;
  double *C_meas = dalloc(dm * dm);
  build_meas(m, R_new, m->b_meas, C_meas, 1);
  kf_measurement_update(m, C_meas);
  free(b_dyn);
  free(A_dyn);
  free(G);
  free(Cin);
  free(x2);
  free(GC);
  free(C_dyn);
  free(AC);
  free(ACA);
  free(C_meas);
}

/* ------------------------------------------------------------------------- public step */
static void body_velocity(const orc_mhe *m, const double *x, double out[3]) {
  /* DecentralEst.cpp:183-185 / :192-194 */
  const double *p_imu_2_opti = m->P.p_imu_2_opti;
  const hist_t *bk = &m->stack[m->nstack - 1];
  double wxp[3], v[3];
  cross3(wxp, bk->angular_b, p_imu_2_opti);
  for (int c = 0; c < 3; ++c) v[c] = x[3 + c] + wxp[c];
  la_mm(out, bk->R, v, 3, 3, 1);
}

void orc_mhe_step(orc_mhe *m, int T, const orc_sample *s) {
  /* callbacks: IMU / joint state / orientation filter overwrite the store every tick; the VO
   * callback (EstSub.cpp:43-55) only when a message arrived, and vo_new_ stays latched until
   * GetMeasurement consumes it (DecentralEst.cpp:891). */
  int pending = m->store.vo_new;
  double pp[3] = {m->store.vo_p[0], m->store.vo_p[1], m->store.vo_p[2]};
  double ptp = m->store.vo_time_pre, ptn = m->store.vo_time_now;
  m->store = *s;
  if (!s->vo_new && pending) {
    m->store.vo_new = 1;
    memcpy(m->store.vo_p, pp, sizeof(pp));
    m->store.vo_time_pre = ptp;
    m->store.vo_time_now = ptn;
  }
  adapter_lo_callback(m);

  if (T == 0) { /* DecentralEst.cpp:131-149 */
    if (m->est_type == 0) {
      initialize_mhe(m);
    } else {
      initialize_kf(m);
      update_kf(m);
    }
    return;
  }
  if (m->est_type == 0) { /* :156-187 */
    update_mhe(m, T);
    if (m->vo_flag) {
      update_vo_constraints(m, T);
      m->vo_flag = 0;
    }
    if (T >= m->N) marginalize_qp(m, T - m->N);
    solve_qp(m);
    /* getsolution(T): x_T (MheSrb.cpp:715-723) */
    la_copy(m->x_MHE, &m->solution[var_x(m, m->nwin - 1)], m->ds);
    body_velocity(m, m->x_MHE, m->v_MHE_b);
  } else {
    update_kf(m);
    body_velocity(m, m->x_KF, m->v_KF_b);
  }
}

/* ------------------------------------------------------------------------- getters */
void orc_mhe_get_x(const orc_mhe *m, double *x) { la_copy(x, m->x_MHE, m->ds); }
void orc_mhe_get_v_body(const orc_mhe *m, double v[3]) { memcpy(v, m->v_MHE_b, sizeof(double) * 3); }
void orc_mhe_get_R_sb(const orc_mhe *m, double R[9]) { memcpy(R, m->R_sb, sizeof(double) * 9); }
void orc_mhe_get_p_vo(const orc_mhe *m, double p[3]) { memcpy(p, m->p_vo_acc, sizeof(double) * 3); }
int orc_mhe_get_arrival(const orc_mhe *m, double *M, double *n) {
  if (!m->have_arrival) return 0;
  la_copy(M, m->M_p, m->ds * m->ds);
  la_copy(n, m->n_p, m->ds);
  return 1;
}
void orc_mhe_get_contact(const orc_mhe *m, double *contact) {
  for (int i = 0; i < m->nlegs; ++i) contact[i] = m->st_contact[i];
}
void orc_mhe_get_meas(const orc_mhe *m, double *b_meas, double *Q_meas) {
  la_copy(b_meas, m->b_meas, m->dm);
  la_copy(Q_meas, m->Q_meas_last, m->dm * m->dm);
}
void orc_mhe_get_kin(const orc_mhe *m, double *p, double *J) {
  la_copy(p, m->st_p_foot, m->dm);
  la_copy(J, m->st_J, m->dm * m->nj);
}
void orc_mhe_get_dims(const orc_mhe *m, int *ds, int *dm, int *dc, int *nVar, int *nCon) {
  *ds = m->ds;
  *dm = m->dm;
  *dc = m->dc;
  *nVar = m->nVar;
  *nCon = m->nCon;
}
void orc_mhe_get_vo_debug(const orc_mhe *m, int out[10]) { memcpy(out, m->vo_dbg, sizeof(int) * 10); }
void orc_mhe_get_solution(const orc_mhe *m, double *z) { la_copy(z, m->solution, m->nVar); }
int orc_mhe_get_admm_iters(const orc_mhe *m) { return m->admm_iters; }
void orc_mhe_get_kf(const orc_mhe *m, double *x, double *C, double v_body[3]) {
  la_copy(x, m->x_KF, m->ds);
  la_copy(C, m->C_KF, m->ds * m->ds);
  memcpy(v_body, m->v_KF_b, sizeof(double) * 3);
}
