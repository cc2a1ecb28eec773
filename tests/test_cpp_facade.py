"""The C++ facade (include/dekf_b200/*.hpp: the reference's class names over the C ABI) compiled with g++ and linked
against libdekf_b200.so.  CPU: it compiles and links.  GPU: a recorded stream is played through orien_ekf +
DecentralizedEstimation exactly like the reference's two nodes would, and checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "cpp", "_build", "facade_main")


def _build_exe():
    from decentralized_ekf_mhe_b200 import build
    so = build.build()
    libdir = os.path.dirname(so)
    src = os.path.join(HERE, "cpp", "facade_main.cpp")
    deps = [src, so] + [os.path.join(ROOT, "include", p) for p in ("dekf_b200.h", "dekf_b200/DecentralEst.hpp", "dekf_b200/orien_ekf.hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", libdir, "-ldekf_b200", f"-Wl,-rpath,{libdir}"])
    return EXE


def test_facade_compiles_and_links():
    exe = _build_exe()
    assert os.path.exists(exe)
    out = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("dekf_create", "dekf_mhe_step_host", "dekf_ekf_step_host", "dekf_get_host", "dekf_destroy"):
        assert sym in out


@pytest.mark.gpu
def test_facade_plays_stream_like_the_reference_nodes(oracle, tmp_path):
    from decentralized_ekf_mhe_b200 import synth
    exe = _build_exe()
    n, S = 48, 120
    st = synth.to_numpy(synth.make_stream(n, S, vo_jitter=True))
    nq, nl = st["joint_pos"].shape[1], st["foot_force"].shape[1]
    path, outp = str(tmp_path / "stream.bin"), str(tmp_path / "out.bin")
    with open(path, "wb") as f:
        f.write(np.array([S, n, nq, nl], np.int32).tobytes())
        for s in range(S):
            for k in ("gyro", "accel", "imu_time", "joint_pos", "joint_vel", "foot_force"):
                f.write(np.ascontiguousarray(st[k][s], np.float64).tobytes())
            f.write(np.ascontiguousarray(st["vo_flag"][s], np.uint8).tobytes())
            for k in ("vo_quat", "vo_time_pre", "vo_time_now", "vo_rel_p"):
                f.write(np.ascontiguousarray(st[k][s], np.float64).tobytes())
    r = subprocess.run([exe, path, outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    raw = open(outp, "rb").read()
    per = (4 + 9 + 3 + 9 + 3) * n * 8 + nl * n
    q = np.zeros((S, 4, n)); x = np.zeros((S, 9, n)); vb = np.zeros((S, 3, n)); R = np.zeros((S, 9, n)); pvo = np.zeros((S, 3, n))
    c = np.zeros((S, nl, n), np.uint8)
    for s in range(S):
        blk = raw[s * per:(s + 1) * per]
        d = np.frombuffer(blk[:(28) * n * 8], np.float64)
        q[s] = d[0:4 * n].reshape(4, n)
        x[s] = d[4 * n:13 * n].reshape(9, n)
        vb[s] = d[13 * n:16 * n].reshape(3, n)
        R[s] = d[16 * n:25 * n].reshape(9, n)
        pvo[s] = d[25 * n:28 * n].reshape(3, n)
        c[s] = np.frombuffer(blk[28 * n * 8:], np.uint8).reshape(nl, n)
    tail = np.frombuffer(raw[S * per:], np.float64)
    M = tail[:81 * n].reshape(81, n)
    npv = tail[81 * n:].reshape(9, n)
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(), oracle.ekf_params(rate=200), nthreads=os.cpu_count() or 1,
                                want=("quat", "x", "v_body", "contact", "p_vo", "arrival"))
    assert np.abs(q - ro["quat"]).max() < 1e-9
    assert np.abs(x[1:, 3:6] - ro["x"][1:, 3:6]).max() < 1e-6
    assert np.abs(vb[1:] - ro["v_body"][1:]).max() < 1e-6
    assert np.array_equal(c, ro["contact"])
    assert np.abs(pvo - ro["p_vo"]).max() < 1e-12
    scale = np.abs(ro["M_p"]).max(axis=0)
    assert (np.abs(M - ro["M_p"]).max(axis=0) / scale).max() < 1e-7
    assert np.abs(npv - ro["n_p"]).max() < 1e-7 * max(1.0, np.abs(ro["n_p"]).max())
    # R_sb_ is the rotation of the normalised EKF quaternion (DecentralEst.cpp:867)
    w, xx, y, z = q[-1]
    assert np.abs(R[-1, 0] - (1 - 2 * (y * y + z * z))).max() < 1e-12
