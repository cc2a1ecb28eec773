"""Configuration half of the drop-in boundary: ``struct robot_params`` (DecentralEst.hpp:18-63), the
``orien_ekf`` constructor parameters (orien_ekf.cpp:13-18) and the batch sizes, as the ctypes mirror
of ``dekf_config`` (include/dekf_b200.h).  Field names follow the reference (trailing underscores
dropped); defaults are go1_example/config/parameters_go1.yaml."""
import ctypes as C

ROBOT_GO1, ROBOT_CASSIE, ROBOT_POGOX = 0, 1, 2
FP64, FP32 = 0, 1
SOLVE_FULL, SOLVE_INCREMENTAL = 0, 1
ROBOT_IDS = {"go1": ROBOT_GO1, "cassie": ROBOT_CASSIE, "pogox": ROBOT_POGOX}

# per-instance status bits (include/dekf_b200.h)
ST_EKF_VO_DROPPED, ST_EKF_VO_NO_REPLAY, ST_EKF_HIST_OVERFLOW = 1, 2, 4
ST_MHE_VO_DROPPED, ST_MHE_VO_BOUNDED, ST_NONFINITE, ST_QP_MAXITER = 8, 16, 32, 64


class DekfConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_instances", C.c_int32), ("device", C.c_int32), ("precision", C.c_int32),
        ("robot", C.c_int32), ("ekf_hist_depth", C.c_int32), ("debug_taps", C.c_int32), ("reserved0", C.c_int32),
        ("p_process_std", C.c_double * 3), ("accel_input_std", C.c_double * 3), ("accel_bias_std", C.c_double * 3),
        ("gyro_input_std", C.c_double * 3), ("quaternion_ib", C.c_double * 4), ("p_ib", C.c_double * 3),
        ("num_legs", C.c_int32), ("leg_odom_type", C.c_int32),
        ("joint_position_std", C.c_double * 8), ("joint_velocity_std", C.c_double * 8),
        ("foot_slide_std", C.c_double * 3), ("foot_swing_std", C.c_double * 3),
        ("contact_effort_threshold", C.c_double),
        ("p_init_std", C.c_double * 3), ("v_init_std", C.c_double * 3), ("foot_init_std", C.c_double * 3),
        ("accel_bias_init_std", C.c_double * 3), ("vo_p_std", C.c_double * 3),
        ("rate", C.c_int32), ("N", C.c_int32), ("est_type", C.c_int32), ("window_solve", C.c_int32),
        ("rho", C.c_double), ("alpha", C.c_double), ("delta", C.c_double), ("sigma", C.c_double),
        ("verbose", C.c_int32), ("adaptRho", C.c_int32), ("polish", C.c_int32), ("maxQPIter", C.c_int32),
        ("realtiveTol", C.c_double), ("absTol", C.c_double), ("primTol", C.c_double), ("dualTol", C.c_double),
        ("timeLimit", C.c_double),
        ("ekf_init_std", C.c_double * 4), ("ekf_process_std", C.c_double * 3),
        ("ekf_gravity_meas_std", C.c_double * 3), ("ekf_vo_meas_std", C.c_double * 4),
        ("ekf_quaternion_init", C.c_double * 4), ("ekf_rate", C.c_int32), ("reserved2", C.c_int32),
        ("v_box_enable", C.c_int32), ("v_box_max_iter", C.c_int32),
        ("v_box_lo", C.c_double * 3), ("v_box_hi", C.c_double * 3),
        ("p_imu_2_opti", C.c_double * 3), ("kf_export_gain", C.c_int32), ("reserved3", C.c_int32),
        ("x_box_mask", C.c_int32), ("reserved4", C.c_int32), ("x_box_lo", C.c_double * 9), ("x_box_hi", C.c_double * 9),
    ]

    def update(self, **over):
        for k, v in over.items():
            cur = getattr(self, k)
            if hasattr(cur, "__len__"):
                for i, x in enumerate(v):
                    cur[i] = x
            else:
                setattr(self, k, v)
        return self
