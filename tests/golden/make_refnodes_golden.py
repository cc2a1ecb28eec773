"""Generates tests/golden/go1_refnodes_golden.npz: synthetic Go1 sensor streams together with the outputs of the
REFERENCE ITSELF on them.

"The reference itself" = oracle/_ref/libref_nodes.so: the reference's own, unmodified node classes and estimator
sources (orien_ekf.cpp, go1Sub.cpp, EstSub.cpp, DecentralEst.cpp, MheSrb.cpp, Bezier_simple.cpp, FROST expressions),
compiled where they lie under /root/reference against the stand-in headers in oracle/ref_stub/ (Eigen3, OSQP/osqp-eigen
and rclcpp are absent from the image) and driven in lock-step through their own ROS callbacks by oracle/ref_nodes.cc.
The QP of every update(T) is solved to its exact optimum (BASELINE.json: "both sides solving to eps 1e-8").  What these
vectors pin: the reference's control flow, contact/VO index logic, QP bookkeeping, marginalisation formulas, EKF
rewind/replay.  What they cannot pin: the rounding inside Eigen's and OSQP's own kernels (stand-ins).

This script needs /root/reference and runs in the build container only; the GPU box reads the committed .npz.

    python tests/golden/make_refnodes_golden.py
"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from decentralized_ekf_mhe_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from oracle import pyref as pr  # noqa: E402

IN_KEYS = ("gyro", "accel", "imu_time", "joint_pos", "joint_vel", "foot_force", "vo_flag", "vo_quat", "vo_time_pre",
           "vo_time_now", "vo_rel_p", "imu_ns", "vo_pre_ns", "vo_now_ns")

# name -> (instances, ticks, stream kwargs, robot_params overrides, EKF rate)
CASES = {
    # BASELINE configs[0]/[1] shape: Go1, MHE, N=20, 200 Hz, ragged VO arrival with 40 ms latency
    "mhe": (8, 200, dict(vo_jitter=True, seed=20240510), dict(), 200),
    # short window: marginalisation from tick 5 on, VO samples older than the window, way-point start-up
    "mhe_n5": (4, 120, dict(vo_jitter=True, seed=7, vo_latency=0.012), dict(N=5), 200),
    # every VO sample is older than the window when it arrives: accumulated but never bounded
    "mhe_n5_late": (2, 80, dict(vo_jitter=True, seed=8, vo_latency=0.055), dict(N=5), 200),
    # KF alternative (est_type_ = 1, DecentralEst.cpp:592-861)
    "kf": (4, 120, dict(vo_jitter=True, seed=11), dict(est_type=1), 200),
    # foot-position states (leg_odom_type_ = 1, ds = 21)
    "foot": (2, 64, dict(vo_jitter=True, seed=13), dict(leg_odom_type=1), 200),
    # KF alternative with foot-position states
    "kf_foot": (2, 80, dict(vo_jitter=True, seed=17), dict(est_type=1, leg_odom_type=1), 200),
}


def main():
    out = {}
    for name, (n, S, skw, pkw, rate) in CASES.items():
        t0 = time.time()
        st = pr.quantize_stream(synth.to_numpy(synth.make_stream(n, S, **skw)))
        prm = po.go1_params(**pkw)
        ep = po.ekf_params(rate=rate)
        ref = pr.run_stream(st, prm, ep, want=("quat", "x", "v_body", "p_vo", "contact", "vo_dbg", "arrival"))
        for k in IN_KEYS:
            out[f"{name}/in_{k}"] = st[k]
        for k, v in ref.items():
            out[f"{name}/out_{k}"] = v
        out[f"{name}/params"] = np.array([prm.N, prm.est_type, prm.leg_odom_type, rate], dtype=np.int32)
        if name in ("mhe", "kf"):
            # what the reference's Data_Logger (data_logger.hpp) writes for instance 0: robotSub::init_logging /
            # spin_logging (EstSub.cpp:77-121) into $HOME/log_exp/<log_name>_{Data,Name.csv}
            home = os.environ.get("HOME")
            with tempfile.TemporaryDirectory() as tmp:
                os.makedirs(os.path.join(tmp, "log_exp"))
                os.environ["HOME"] = tmp
                try:
                    pr.run_stream(st, prm, ep, i0=0, i1=1, want=())
                finally:
                    if home is None:
                        del os.environ["HOME"]
                    else:
                        os.environ["HOME"] = home
                out[f"{name}/log_data"] = np.fromfile(os.path.join(tmp, "log_exp", "refnodes_Data"), dtype=np.float64)
                out[f"{name}/log_names"] = np.frombuffer(open(os.path.join(tmp, "log_exp", "refnodes_Name.csv"), "rb").read(), dtype=np.uint8)
            print(f"  reference log: {out[f'{name}/log_data'].size} doubles, names:",
                  bytes(out[f"{name}/log_names"]).decode().replace("\n", " | "))
        print(f"{name}: {n} instances x {S} ticks in {time.time() - t0:.1f} s", flush=True)
    # the reference's deployment schedule: EKF timer at 500 Hz, estimator timer at 200 Hz (tests/mixed_rate.py)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mixed_rate as mr
    t0 = time.time()
    n, S = 4, 300
    st = pr.quantize_stream(synth.to_numpy(synth.make_stream(n, S, dt=0.002, vo_jitter=True, seed=23)))
    prm, ep = po.go1_params(), po.ekf_params(rate=500)
    runs = [mr.run_reference(pr, st, i, prm, ep) for i in range(n)]
    for k in IN_KEYS:
        out[f"mixed/in_{k}"] = st[k]
    for k in runs[0]:
        out[f"mixed/out_{k}"] = np.stack([r[k] for r in runs], axis=-1)
    out["mixed/params"] = np.array([prm.N, prm.est_type, prm.leg_odom_type, 500], dtype=np.int32)
    print(f"mixed (EKF 500 Hz / MHE 200 Hz): {n} instances x {S} IMU samples, {runs[0]['x'].shape[0]} estimator ticks in {time.time() - t0:.1f} s", flush=True)
    path = os.path.join(os.path.dirname(__file__), "go1_refnodes_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
