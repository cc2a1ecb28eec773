"""Drop-in proof with the reference's OWN node sources.

oracle/_ref/dropin_nodes = the reference's est_sub node -- decentral_legged_est/src/EstSub.cpp (timerCallback:
mhe.initialize / mhe.update, EstSub.cpp:58-91) and go1_example/src/go1Sub.cpp (callbacks that fill robot_store, FROST
kinematics) -- compiled UNMODIFIED, where they lie under /root/reference, against
include/dekf_b200/dropin/decentral_legged_est/DecentralEst.hpp (our Eigen-typed `DecentralizedEstimation` over
libdekf_b200.so) in place of the reference's DecentralEst.hpp.  Recipe: `make -C oracle dropin` (Eigen / rclcpp: the
stand-ins of oracle/ref_stub, neither library is in the image).  The binary is git-ignored and travels to the GPU box.

CPU: the recipe builds from the reference tree and the binary binds the C ABI.  GPU: golden streams replayed through the
reference's callbacks give the reference's own outputs (tests/golden/go1_refnodes_golden.npz, produced by the reference's
Eigen/OSQP classes): quaternion / x / v_body 1e-9, p_vo 1e-12, contact sets exact."""
import os
import subprocess

import numpy as np
import pytest

from test_ros_shells import _write_stream

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(ROOT, "oracle", "_ref", "dropin_nodes")
GOLDEN = os.path.join(HERE, "golden", "go1_refnodes_golden.npz")
REF = "/root/reference"


def _build():
    from decentralized_ekf_mhe_b200 import build
    build.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dropin"])
    return EXE


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "decentral_legged_est")), reason="reference tree not present")
def test_reference_node_sources_compile_unmodified_against_the_facade():
    exe = _build()
    assert os.path.exists(exe)
    und = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("dekf_create", "dekf_mhe_step_host", "dekf_get_host", "dekf_ekf_step_host"):
        assert sym in und
    # the reference's own classes are in the binary (compiled from its sources), the reference's estimator core is not
    defined = subprocess.run(["nm", "-C", "--defined-only", exe], capture_output=True, text=True).stdout
    assert "robotSub::robotSub::timerCallback()" in defined and "robotSub::go1Sub::lo_callback" in defined
    assert "SymFunction::FR_foot" in defined
    assert "MHEproblem::" not in defined and "DecentralizedEstimation::UpdateMHE" not in defined
    # nothing of the reference is vendored: the recipe compiles the sources where they lie
    mk = open(os.path.join(ROOT, "oracle", "Makefile")).read()
    assert "$(REF_EST)/src/EstSub.cpp" in mk and "$(REF_GO1)/src/go1Sub.cpp" in mk
    assert not os.path.exists(os.path.join(ROOT, "oracle", "EstSub.cpp"))


@pytest.mark.gpu
@pytest.mark.parametrize("name,inst", [("mhe", 0), ("mhe", 5), ("kf", 0), ("mhe_n5", 1)])
def test_reference_nodes_on_the_gpu_path_match_the_reference(tmp_path, name, inst):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/dropin_nodes was not built (needs the reference tree: make -C oracle dropin)")
    g = np.load(GOLDEN)
    st = {k.split("/in_")[1]: g[k] for k in g.files if k.startswith(name + "/in_")}
    ref = {k.split("/out_")[1]: g[k] for k in g.files if k.startswith(name + "/out_")}
    N, est_type, leg_odom_type, rate = (int(v) for v in g[name + "/params"])
    S, nq, nl = _write_stream(str(tmp_path / "stream.bin"), st, inst, N, est_type, leg_odom_type, rate)
    os.makedirs(tmp_path / "log_exp")
    env = dict(os.environ, HOME=str(tmp_path))
    r = subprocess.run([EXE, str(tmp_path / "stream.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr + r.stdout
    ds = ref["x"].shape[1]
    out = np.fromfile(tmp_path / "out.bin", dtype=np.float64).reshape(S, 4 + ds + 3 + 3 + nl)
    q, x, vb, pv, ct = out[:, :4], out[:, 4:4 + ds], out[:, 4 + ds:7 + ds], out[:, 7 + ds:10 + ds], out[:, 10 + ds:]
    assert np.abs(q - ref["quat"][:, :, inst]).max() < 1e-9               # north-star: quaternion 1e-9
    assert np.abs(x[1:] - ref["x"][1:, :, inst]).max() < 1e-9             # north-star: velocity 1e-6 m/s
    assert np.abs(vb[1:] - ref["v_body"][1:, :, inst]).max() < 1e-9
    assert np.abs(pv - ref["p_vo"][:, :, inst]).max() < 1e-12
    assert np.array_equal(ct != 0, ref["contact"][:, :, inst] != 0)       # contact sets (host FROST path of go1Sub) exact
    # the reference's own Data_Logger (data_logger.hpp, compiled unmodified into the node) wrote the reference's log
    if inst == 0 and name in ("mhe", "kf"):
        data = np.fromfile(tmp_path / "log_exp" / "dropin_Data", dtype=np.float64)
        ref_log = g[name + "/log_data"]
        assert data.shape == ref_log.shape and np.abs(data - ref_log).max() < 1e-9
