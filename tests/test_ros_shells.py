"""the host-side ROS 2 node shells re-pointed at the C ABI (ros2/dekf_b200_ros).

dekf_ros::OrienSub / dekf_ros::EstSub keep the node interface of the reference's orien_sub / est_sub (parameter names,
topics, timers, start gate, Data_Logger file format) and call libdekf_b200.so for the arithmetic.  They are compiled here
against the stand-in rclcpp of oracle/ref_stub (test infrastructure; a robot builds them against real rclcpp with
ros2/dekf_b200_ros/CMakeLists.txt) and driven through their subscriptions exactly like oracle/ref_nodes.cc drives the
reference's own nodes.  CPU: they compile and link.  GPU: their outputs and their log files equal those of the reference's
own nodes (tests/golden/go1_refnodes_golden.npz)."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "cpp", "_build", "ros_shell_main")
GOLDEN = os.path.join(HERE, "golden", "go1_refnodes_golden.npz")


def _build_exe():
    from decentralized_ekf_mhe_b200 import build
    so = build.build()
    libdir = os.path.dirname(so)
    src = os.path.join(HERE, "cpp", "ros_shell_main.cpp")
    inc = [os.path.join(ROOT, "ros2", "dekf_b200_ros", "include"), os.path.join(ROOT, "include"), os.path.join(ROOT, "oracle", "ref_stub")]
    deps = [src, so, os.path.join(inc[0], "dekf_b200_ros", "est_sub.hpp"), os.path.join(inc[0], "dekf_b200_ros", "orien_sub.hpp"),
            os.path.join(inc[0], "dekf_b200_ros", "data_logger.hpp"), os.path.join(inc[1], "dekf_b200", "DecentralEst.hpp"),
            os.path.join(inc[2], "rclcpp", "rclcpp.hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Wno-unused-parameter"]
        for i in inc:
            cmd += ["-I", i]
        subprocess.check_call(cmd + [src, "-o", EXE, "-L", libdir, "-ldekf_b200", f"-Wl,-rpath,{libdir}"])
    return EXE


def test_ros_shells_compile_and_link():
    exe = _build_exe()
    out = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("dekf_create", "dekf_mhe_step_host", "dekf_ekf_step_host"):
        assert sym in out
    # the node mains a robot builds are plain rclcpp programs over the same headers
    for f in ("est_sub_node.cpp", "orien_sub_node.cpp"):
        src = os.path.join(ROOT, "ros2", "dekf_b200_ros", "src", f)
        subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "ros2", "dekf_b200_ros", "include"),
                               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "ref_stub"), src])


def _write_stream(path, st, i, N, est_type, leg_odom_type, rate):
    S = st["gyro"].shape[0]
    nq, nl = st["joint_pos"].shape[1], st["foot_force"].shape[1]
    with open(path, "wb") as f:
        f.write(np.array([S, nq, nl, N, est_type, leg_odom_type, rate, 0], np.int32).tobytes())
        for s in range(S):
            f.write(np.array([st["imu_ns"][s, i], st["vo_flag"][s, i], st["vo_pre_ns"][s, i], st["vo_now_ns"][s, i]], np.int64).tobytes())
            f.write(np.concatenate([st[k][s, :, i] for k in ("gyro", "accel", "joint_pos", "joint_vel", "foot_force", "vo_quat", "vo_rel_p")])
                    .astype(np.float64).tobytes())
    return S, nq, nl


@pytest.mark.gpu
@pytest.mark.parametrize("name,inst", [("mhe", 0), ("mhe", 5), ("kf", 0)])
def test_ros_shells_match_the_reference_nodes(tmp_path, name, inst):
    exe = _build_exe()
    g = np.load(GOLDEN)
    st = {k.split("/in_")[1]: g[k] for k in g.files if k.startswith(name + "/in_")}
    ref = {k.split("/out_")[1]: g[k] for k in g.files if k.startswith(name + "/out_")}
    N, est_type, leg_odom_type, rate = (int(v) for v in g[name + "/params"])
    S, nq, nl = _write_stream(str(tmp_path / "stream.bin"), st, inst, N, est_type, leg_odom_type, rate)
    os.makedirs(tmp_path / "log_exp")
    env = dict(os.environ, HOME=str(tmp_path))
    r = subprocess.run([exe, str(tmp_path / "stream.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr + r.stdout
    ds = ref["x"].shape[1]
    out = np.fromfile(tmp_path / "out.bin", dtype=np.float64).reshape(S, 4 + ds + 3 + 3 + nl)
    q, x, vb, pv, ct = out[:, :4], out[:, 4:4 + ds], out[:, 4 + ds:7 + ds], out[:, 7 + ds:10 + ds], out[:, 10 + ds:]
    assert np.abs(q - ref["quat"][:, :, inst]).max() < 1e-9               # north-star: quaternion 1e-9
    assert np.abs(x[1:] - ref["x"][1:, :, inst]).max() < 1e-9             # north-star: velocity 1e-6 m/s
    assert np.abs(vb[1:] - ref["v_body"][1:, :, inst]).max() < 1e-9
    assert np.abs(pv - ref["p_vo"][:, :, inst]).max() < 1e-12
    assert np.array_equal(ct != 0, ref["contact"][:, :, inst] != 0)       # contact sets exact
    if inst == 0:
        # log files against what the reference's Data_Logger wrote for the same instance (same names, same records)
        names = open(tmp_path / "log_exp" / "shell_Name.csv", "rb").read()
        assert names == bytes(g[name + "/log_names"])
        data = np.fromfile(tmp_path / "log_exp" / "shell_Data", dtype=np.float64)
        ref_log = g[name + "/log_data"]
        assert data.shape == ref_log.shape and data.size == (S - N - 1) * 27
        assert np.abs(data - ref_log).max() < 1e-9
