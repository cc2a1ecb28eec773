/* CPU oracle for the Decentralized_EKF_MHE estimator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  A plain-C restatement of the reference's algorithm
 * (src/orien_est/src/orien_ekf.cpp, src/decentral_legged_est/src/{DecentralEst,MheSrb}.cpp,
 * src/decentral_legged_est/src/Spline/Bezier_simple.cpp, src/go1_example/src/go1Sub.cpp:64-125),
 * each function citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product path
 * (decentralized_ekf_mhe_b200/ + libdekf_b200.so) never does.
 *
 * PARITY STATUS: the reference ships no tests, golden vectors or fixtures, and Eigen3 / OSQP / osqp-eigen / rclcpp are
 * absent from this image.  The restatement is pinned against THE REFERENCE'S OWN SOURCES instead:
 *   - oracle/_ref/libref_nodes.so = orien_ekf.cpp, go1Sub.cpp, EstSub.cpp, DecentralEst.cpp, MheSrb.cpp,
 *     Bezier_simple.cpp and the FROST expressions, compiled UNMODIFIED where they lie under /root/reference against
 *     stand-in headers (oracle/ref_stub/: eager dense linear algebra for Eigen, exact KKT solve or oracle/admm.c for
 *     OSQP, a synchronous topic bus for rclcpp) and driven through their own ROS callbacks by oracle/ref_nodes.cc;
 *   - tests/golden/go1_refnodes_golden.npz = its outputs on synthetic streams (tests/golden/make_refnodes_golden.py);
 *   - tests/test_refnodes_pin.py: oracle == reference (quaternion 7e-18, x_MHE 5e-11, contact sets / VO index
 *     bookkeeping / accumulated VO translation exact, the QP handed to OSQP equal entry by entry, KF alternative 3e-17,
 *     foot-state model 1e-10).
 * NOT pinned ("parity unpinned" for these): the floating-point rounding inside Eigen's own kernels (PartialPivLU,
 * SimplicialLLT) and OSQP's ADMM iterates -- the stand-ins replace them; BASELINE.json defines parity on the converged
 * optimum (eps 1e-8), which is what both stand-in solvers return.
 * The Go1 kinematics are additionally pinned through oracle/_ref/libfrost_go1.so (tests/test_oracle_kinematics.py,
 * tests/golden/go1_kin_golden.npz).
 */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ EKF (orien_ekf.cpp) */
typedef struct orc_ekf orc_ekf;

/* orien_ekf.cpp:8-33 (constructor parameters -> covariances, dt_ = 1/rate, gravity_=(0,0,9.81)) */
orc_ekf *orc_ekf_create(const double init_std[4], const double process_std[3],
                        const double gravity_meas_std[3], const double vo_meas_std[4],
                        const double quaternion_init[4], int rate);
void orc_ekf_destroy(orc_ekf *e);
/* One timer tick (orien_ekf.cpp:77-89) after the IMU callback stored (gyro, accel, imu_time) and,
 * if vo_new != 0, the VO pose callback stored (vo_quat [w,x,y,z], vo_time) (orien_ekf.cpp:47-75). */
void orc_ekf_tick(orc_ekf *e, const double gyro[3], const double accel[3], double imu_time,
                  int vo_new, const double vo_quat[4], double vo_time);
void orc_ekf_get(const orc_ekf *e, double q[4], double P[16]);
void orc_ekf_set(orc_ekf *e, const double q[4], const double P[16]);
/* Index bookkeeping of the last VO replay (orien_ekf.cpp:175-205): cur = newest history index,
 * idx = synchronised index (-1 = dropped: "not storing enough imu info"), nreplay = loop trips. */
void orc_ekf_last_replay(const orc_ekf *e, int *cur, int *idx, int *nreplay);
/* Stand-alone entry points (orien_ekf.cpp:108-154) for known-answer tests. */
void orc_ekf_predict(const orc_ekf *e, double q_pred[4], const double q[4], const double gyro[3],
                     const double P[16], double P_pred[16]);
void orc_ekf_correct(const orc_ekf *e, double q_corr[4], const double q_pred[4],
                     const double accel[3], const double P_pred[16], double P_corr[16]);
void orc_ekf_vo_correct(const orc_ekf *e, double q_corr[4], const double q_pred[4],
                        const double q_vo[4], const double P_pred[16], double P_corr[16]);
/* Eigen Quaterniond(q).normalized().toRotationMatrix(), q=[w,x,y,z], R row-major 3x3 */
void orc_quat_to_rot(const double q[4], double R[9]);

/* ------------------------------------------------------------------ kinematics / adapter */
/* Robot models.  ORC_ROBOT_GO1 restates the reference (FROST Go1, go1Sub.cpp:64-125); CASSIE and
 * POGOX are builder-defined models (the reference ships no such code). */
enum { ORC_ROBOT_GO1 = 0, ORC_ROBOT_CASSIE = 1, ORC_ROBOT_POGOX = 2 };
int orc_robot_num_legs(int robot);
int orc_robot_joints_per_leg(int robot);
/* Foot position of one leg in the kinematics base frame (no p_ib) and its 3 x nj joint Jacobian
 * (row-major), for joint angles q[nj].  Go1: closed form of FR_foot.cc/J_FR.cc with base dofs = 0. */
void orc_leg_fk(int robot, int leg, const double *q, double p[3], double *J /*3 x nj*/);

/* ------------------------------------------------------------------ MHE (DecentralEst.cpp / MheSrb.cpp) */
typedef struct {
  /* robot_params, DecentralEst.hpp:18-63 */
  double p_process_std[3], accel_input_std[3], accel_bias_std[3], gyro_input_std[3];
  double quaternion_ib[4], p_ib[3];
  int num_legs, leg_odom_type;
  double joint_position_std[8], joint_velocity_std[8]; /* per joint; the reference uses 3 (Go1) */
  double foot_slide_std[3], foot_swing_std[3];
  double contact_effort_threshold;
  double p_init_std[3], v_init_std[3], foot_init_std[3], accel_bias_init_std[3];
  double vo_p_std[3];
  int rate, N, est_type;
  /* osqp params */
  double rho, alpha, delta, sigma;
  int verbose, adapt_rho, polish, max_qp_iter;
  double relative_tol, abs_tol, prim_tol, dual_tol, time_limit;
  /* builder additions */
  int robot;      /* ORC_ROBOT_* */
  int solve_mode; /* 0: exact (slack elimination + banded Cholesky), 2: OSQP-style ADMM */
  /* optional box constraint on v_s of every state (builder extension, never used by the reference):
   * active when v_box_enable != 0; rows lb<=v_s<=ub appended per state in the ADMM mode. */
  int v_box_enable;
  double v_box_lo[3], v_box_hi[3];
  /* lever arm of the body-velocity read-out, DecentralEst.cpp:181-185 (the reference hard-codes the Go1 mocap marker) */
  double p_imu_2_opti[3];
  /* general per-component bounds (builder extension like v_box): bit a of x_box_mask bounds component a (0..8) of every
   * window state, x_box_lo[a] <= x_k[a] <= x_box_hi[a]; combines with v_box (x_box wins on components 3..5) */
  int x_box_mask;
  double x_box_lo[9], x_box_hi[9];
  /* general linear rows (builder extension): x_row_lo[i] <= x_row_a[i] . x_k <= x_row_hi[i] for every window state, i <
   * x_row_count <= 9, rows linearly independent -- MHEproblem::addConstraints(name, lb, ub) with an arbitrary dependency row on
   * x_k (MheSrb.cpp:58-68, :217-270).  Component bounds (v_box / x_box) given at the same time count as unit rows. */
  int x_row_count;
  double x_row_a[81], x_row_lo[9], x_row_hi[9];
} orc_params;

void orc_params_go1_defaults(orc_params *p); /* parameters_go1.yaml */

typedef struct {
  /* robot_store, DecentralEst.hpp:65-94 (only the fields the core reads) + raw joint message */
  double imu_time;
  double accel_b[3], angular_b[3];
  double quaternion[4];   /* [w,x,y,z] from imu/filter (EstSub.cpp:34-40) */
  double joint_pos[40];   /* positions, then foot forces at [num_legs*nj + leg] (go1Sub.cpp:74) */
  double joint_vel[40];
  int vo_new;
  double vo_time_pre, vo_time_now, vo_p[3];
} orc_sample;

typedef struct orc_mhe orc_mhe;
orc_mhe *orc_mhe_create(const orc_params *p);
void orc_mhe_destroy(orc_mhe *m);
/* T==0: DecentralizedEstimation::initialize (DecentralEst.cpp:9-150); T>=1: update(T) (:152-198). */
void orc_mhe_step(orc_mhe *m, int T, const orc_sample *s);

/* results (DecentralEst.hpp:278-291) */
void orc_mhe_get_x(const orc_mhe *m, double *x /*ds*/);
void orc_mhe_get_v_body(const orc_mhe *m, double v[3]);
void orc_mhe_get_R_sb(const orc_mhe *m, double R[9]);
void orc_mhe_get_p_vo(const orc_mhe *m, double p[3]);
int orc_mhe_get_arrival(const orc_mhe *m, double *M /*ds*ds*/, double *n /*ds*/); /* M_p, n_p; returns 0 if none yet */
void orc_mhe_get_contact(const orc_mhe *m, double *contact /*num_legs*/);
void orc_mhe_get_meas(const orc_mhe *m, double *b_meas /*dm*/, double *Q_meas /*dm*dm*/);
void orc_mhe_get_kin(const orc_mhe *m, double *p_imu_2_foot /*dm*/, double *J_imu_2_foot /*dm x nj*/);
void orc_mhe_get_dims(const orc_mhe *m, int *ds, int *dm, int *dc, int *nVar, int *nCon);
/* integer VO bookkeeping of the most recent VO message seen (DecentralEst.cpp:883-945):
 * out[0]=processed(0/1: 0 = dropped or none) out[1]=i_pre out[2]=i_now out[3]=w0 out[4]=i0
 * out[5]=ins out[6]=num out[7]=first discrete time bounded out[8]=flagged(vo_to_be_processed) out[9]=stack size */
void orc_mhe_get_vo_debug(const orc_mhe *m, int out[10]);
/* dense export of the QP OSQP would see (reference ordering, App. B.2 of SURVEY.md) */
void orc_mhe_export_qp(const orc_mhe *m, double *H, double *g, double *A, double *l, double *u);
/* full primal solution in reference ordering (length nVar) */
void orc_mhe_get_solution(const orc_mhe *m, double *z);
int orc_mhe_get_admm_iters(const orc_mhe *m);
/* KF alternative results (est_type 1) */
void orc_mhe_get_kf(const orc_mhe *m, double *x /*ds*/, double *C /*ds*ds*/, double v_body[3]);

#ifdef __cplusplus
}
#endif
#endif
