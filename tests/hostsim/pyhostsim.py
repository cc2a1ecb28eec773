"""ctypes driver of the CPU math-debug harness (tests/hostsim/hostsim.cpp).  NOT a product path."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _config_struct():
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from decentralized_ekf_mhe_b200.params import DekfConfig
    return DekfConfig


def build():
    so = os.path.join(HERE, "_build", "libhostsim.so")
    srcs = [os.path.join(HERE, "hostsim.cpp")] + [
        os.path.join(ROOT, "decentralized_ekf_mhe_b200", "csrc", f)
        for f in ("estimator_core.cuh", "smallmat.cuh", "kinematics.cuh", "host_setup.hpp", "box_solve.cuh", "footstate.cuh")] + [
        os.path.join(ROOT, "include", "dekf_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                               os.path.join(HERE, "hostsim.cpp")])
    return so


def run(stream, cfg, quat_in=None, rows=None):
    """rows = (a [count][9], lb, ub): general linear rows on every window state (what dekf_add_state_rows sets on a handle)."""
    lib = C.CDLL(build())
    if rows is not None:
        ra, rl, ru = (np.ascontiguousarray(np.asarray(v, dtype=np.float64)) for v in rows)
        lib.hostsim_set_rows(int(rl.size), ra.ctypes.data_as(C.POINTER(C.c_double)), rl.ctypes.data_as(C.POINTER(C.c_double)),
                             ru.ctypes.data_as(C.POINTER(C.c_double)))
    else:
        lib.hostsim_set_rows(0, None, None, None)
    S, _, n = stream["gyro"].shape
    nl = stream["foot_force"].shape[1]
    cfg.n_instances = n
    dp, up, ip = C.POINTER(C.c_double), C.POINTER(C.c_ubyte), C.POINTER(C.c_int32)
    keep = []

    def d(name):
        a = np.ascontiguousarray(stream[name], dtype=np.float64)
        keep.append(a)
        return a.ctypes.data_as(dp)

    flag = np.ascontiguousarray(stream["vo_flag"], dtype=np.uint8)
    xr = 9 + 3 * nl * int(cfg.leg_odom_type)
    res = dict(quat=np.zeros((S, 4, n)), x=np.full((S, xr, n), np.nan), v_body=np.full((S, 3, n), np.nan),
               contact=np.zeros((S, nl, n), np.uint8), vo_dbg=np.zeros((S, 8, n), np.int32),
               ekf_dbg=np.zeros((S, 3, n), np.int32), p_vo=np.zeros((S, 3, n)), status=np.zeros((S, n), np.int32),
               arr_P=np.zeros((45, n)), arr_x=np.zeros((9, n)), qp=np.zeros((S, 2, n), np.int32))
    qi = None
    if quat_in is not None:
        qi = np.ascontiguousarray(quat_in, dtype=np.float64)
    rc = lib.hostsim_run(C.byref(cfg), S, d("gyro"), d("accel"), d("imu_time"), d("joint_pos"), d("joint_vel"),
                         d("foot_force"), flag.ctypes.data_as(up), d("vo_quat"), d("vo_time_pre"), d("vo_time_now"),
                         d("vo_rel_p"), qi.ctypes.data_as(dp) if qi is not None else None,
                         res["quat"].ctypes.data_as(dp), res["x"].ctypes.data_as(dp), res["v_body"].ctypes.data_as(dp),
                         res["contact"].ctypes.data_as(up), res["vo_dbg"].ctypes.data_as(ip),
                         res["ekf_dbg"].ctypes.data_as(ip), res["p_vo"].ctypes.data_as(dp),
                         res["status"].ctypes.data_as(ip), res["arr_P"].ctypes.data_as(dp),
                         res["arr_x"].ctypes.data_as(dp), res["qp"].ctypes.data_as(ip))
    assert rc == 0
    return res
