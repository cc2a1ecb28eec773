"""Kernel-math check without a GPU: the per-instance kernel bodies (csrc/estimator_core.cuh, written
__host__ __device__) are compiled for the host by tests/hostsim (a debug harness, not shipped) and
compared with the oracle.  The real parity tests run the CUDA path through the C ABI (-m gpu)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))


def _cfg(**over):
    import pyhostsim as hs
    from decentralized_ekf_mhe_b200.params import DekfConfig
    lib = C.CDLL(hs.build())
    cfg = DekfConfig()
    lib.hostsim_default_go1(C.byref(cfg))
    cfg.ekf_rate = 200
    cfg.update(**over)
    return cfg


def test_kernel_math_matches_oracle_fp64(oracle, go1_stream_small):
    import pyhostsim as hs
    st = go1_stream_small
    r = hs.run(st, _cfg())
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(), oracle.ekf_params(rate=200), nthreads=4)
    assert np.abs(r["quat"] - ro["quat"]).max() < 1e-9           # north-star: quaternions 1e-9
    assert np.abs(r["x"][1:, 3:6] - ro["x"][1:, 3:6]).max() < 1e-6  # north-star: velocity 1e-6 m/s
    assert np.abs(r["x"][1:] - ro["x"][1:]).max() < 1e-8
    assert np.array_equal(r["contact"], ro["contact"])            # contact sets bit-exact
    assert np.array_equal(r["vo_dbg"], ro["vo_dbg"][:, :8])        # VO index logic bit-exact
    assert np.array_equal(r["ekf_dbg"], ro["ekf_dbg"])             # EKF replay index logic bit-exact
    assert np.abs(r["p_vo"] - ro["p_vo"]).max() < 1e-12


def test_kernel_math_matches_oracle_fp32(oracle, go1_stream_small):
    import pyhostsim as hs
    st = go1_stream_small
    r = hs.run(st, _cfg(precision=1))
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(), oracle.ekf_params(rate=200), nthreads=4)
    assert np.abs(r["x"][1:, 3:6] - ro["x"][1:, 3:6]).max() < 1e-4  # north-star: fp32 1e-4 m/s
    assert np.array_equal(r["contact"], ro["contact"])
    assert np.array_equal(r["vo_dbg"], ro["vo_dbg"][:, :8])
    assert np.array_equal(r["ekf_dbg"], ro["ekf_dbg"])


def test_arrival_cost_matches_reference_form(oracle, go1_stream_small):
    import pyhostsim as hs
    st = go1_stream_small
    r = hs.run(st, _cfg())
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(), oracle.ekf_params(rate=200), nthreads=4, want=("arrival",))
    n = st["gyro"].shape[2]
    for i in range(n):
        P = np.zeros((9, 9))
        a = r["arr_P"][:, i]
        sym = lambda v: np.array([[v[0], v[1], v[2]], [v[1], v[3], v[4]], [v[2], v[4], v[5]]])
        P[0:3, 0:3], P[3:6, 3:6], P[6:9, 6:9] = sym(a[0:6]), sym(a[6:12]), sym(a[12:18])
        P[0:3, 3:6], P[0:3, 6:9], P[3:6, 6:9] = a[18:27].reshape(3, 3), a[27:36].reshape(3, 3), a[36:45].reshape(3, 3)
        P = np.triu(P) + np.triu(P, 1).T
        M = np.linalg.inv(P)
        Mo = ro["M_p"][:, i].reshape(9, 9)
        np.testing.assert_allclose(M, Mo, rtol=0, atol=1e-7 * np.abs(Mo).max())
        np.testing.assert_allclose(-M @ r["arr_x"][:, i], ro["n_p"][:, i], rtol=0, atol=1e-7 * max(1, np.abs(ro["n_p"][:, i]).max()))


@pytest.mark.parametrize("robot,rid,nl", [("cassie", 1, 2), ("pogox", 2, 1)])
def test_builder_models_match_generalised_oracle(oracle, robot, rid, nl):
    import pyhostsim as hs
    from decentralized_ekf_mhe_b200 import synth
    st = synth.to_numpy(synth.make_stream(4, 140, robot=robot, vo_jitter=True))
    thr = 150.0 if robot == "cassie" else 100.0
    r = hs.run(st, _cfg(robot=rid, num_legs=nl, contact_effort_threshold=thr, p_ib=(0.0, 0.0, 0.0)))
    prm = oracle.go1_params(robot=rid, num_legs=nl, contact_effort_threshold=thr, p_ib=(0.0, 0.0, 0.0))
    ro, _, _ = oracle.run_batch(st, prm, oracle.ekf_params(rate=200), nthreads=4)
    assert np.abs(r["x"][1:, 3:6] - ro["x"][1:, 3:6]).max() < 1e-6
    assert np.array_equal(r["contact"], ro["contact"])
    assert np.array_equal(r["vo_dbg"], ro["vo_dbg"][:, :8])
    assert r["contact"].any() and not r["contact"].all()


def test_kf_alternative_matches_oracle(oracle, go1_stream_small):
    """est_type 1 (DecentralEst.cpp:592-861): x_KF_, v_KF_b_ and p_vo_accmulate_ against the oracle's KF path."""
    import pyhostsim as hs
    st = go1_stream_small
    r = hs.run(st, _cfg(est_type=1))
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(est_type=1), oracle.ekf_params(rate=200), nthreads=4)
    assert np.abs(r["x"][1:] - ro["x"][1:]).max() < 1e-10
    assert np.abs(r["v_body"][1:] - ro["v_body"][1:]).max() < 1e-10
    assert np.abs(r["p_vo"] - ro["p_vo"]).max() < 1e-12
    assert np.abs(ro["p_vo"]).max() > 0.1  # VO messages were consumed


BOX_LO, BOX_HI = (-0.45, -0.03, -0.015), (0.55, 0.03, 0.015)


def test_state_constrained_solve_matches_oracle(oracle):
    """BASELINE config 4 (PogoX, box on the velocity states that binds in >= 20 % of the steps): the kernel's
    primal-dual active-set solve against the oracle's exact constrained optimum."""
    import pyhostsim as hs
    from decentralized_ekf_mhe_b200 import synth
    st = synth.to_numpy(synth.make_stream(6, 140, robot="pogox", vo_jitter=True))
    kw = dict(robot=2, num_legs=1, contact_effort_threshold=100.0, p_ib=(0.0, 0.0, 0.0))
    r = hs.run(st, _cfg(v_box_enable=1, v_box_lo=BOX_LO, v_box_hi=BOX_HI, **kw))
    prm = oracle.go1_params(v_box_enable=1, v_box_lo=BOX_LO, v_box_hi=BOX_HI, **kw)
    ro, _, _ = oracle.run_batch(st, prm, oracle.ekf_params(rate=200), nthreads=4, want=("x", "v_body"))
    assert np.abs(r["x"][1:] - ro["x"][1:]).max() < 1e-8
    assert np.abs(r["v_body"][1:] - ro["v_body"][1:]).max() < 1e-8
    v = r["x"][1:, 3:6]
    lo, hi = np.array(BOX_LO)[None, :, None], np.array(BOX_HI)[None, :, None]
    assert (v <= hi + 1e-12).all() and (v >= lo - 1e-12).all()          # constraint violation: none
    assert ((v == hi) | (v == lo)).any(axis=1).mean() > 0.2              # the box binds in >= 20 % of the steps
    assert not (r["status"] & 64).any()                                  # active set always converged
    assert r["qp"][1:, 0].mean() < 6                                     # warm start: few factorisations per tick


def test_general_component_bounds_match_oracle(oracle):
    """cfg.x_box_mask (bounds on any state component, here v_x + p_z + the accel bias): box_solve<T, true> (the kernel body of
    k_solve_box, compiled for the host) against the oracle's exact constrained optimum; and the velocity box expressed through
    x_box_* gives what v_box_* gives."""
    import pyhostsim as hs
    from decentralized_ekf_mhe_b200 import synth
    st = synth.to_numpy(synth.make_stream(4, 120, robot="pogox", vo_jitter=True))
    kw = dict(robot=2, num_legs=1, contact_effort_threshold=100.0, p_ib=(0.0, 0.0, 0.0))
    mask = (1 << 3) | (1 << 2) | (7 << 6)
    lo9 = (0, 0, -2e-4, 0.47, 0, 0, -0.004, -0.004, -0.004)
    hi9 = (0, 0, 2e-4, 0.52, 0, 0, 0.004, 0.004, 0.004)
    r = hs.run(st, _cfg(x_box_mask=mask, x_box_lo=lo9, x_box_hi=hi9, **kw))
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(x_box_mask=mask, x_box_lo=lo9, x_box_hi=hi9, **kw), oracle.ekf_params(rate=200),
                                nthreads=4, want=("x", "v_body"))
    assert np.abs(r["x"][1:] - ro["x"][1:]).max() < 1e-8
    x = r["x"][1:]
    for a in (2, 3, 6, 7, 8):
        assert (x[:, a] <= hi9[a] + 1e-12).all() and (x[:, a] >= lo9[a] - 1e-12).all()
    assert (np.abs(x[:, 6:9]) == 0.004).any() and (np.abs(x[:, 2]) == 2e-4).any()   # bias and position rows bind
    assert not (r["status"] & 64).any()
    # the velocity box through the general interface == the velocity box through v_box_*
    rv = hs.run(st, _cfg(v_box_enable=1, v_box_lo=BOX_LO, v_box_hi=BOX_HI, **kw))
    lo_v = (0, 0, 0) + BOX_LO + (0, 0, 0)
    hi_v = (0, 0, 0) + BOX_HI + (0, 0, 0)
    rg = hs.run(st, _cfg(x_box_mask=0x38, x_box_lo=lo_v, x_box_hi=hi_v, **kw))
    assert np.abs(rg["x"][1:] - rv["x"][1:]).max() < 1e-9


def test_general_linear_rows_match_oracle(oracle):
    """General rows  lb <= a . x_k <= ub  (dekf_add_state_rows; three rows mixing velocity / bias / position components, plus a
    component bound from the config): box_solve<T, true> in the row basis y = W x (the kernel body of k_solve_box, compiled for
    the host) against the oracle's exact constrained optimum (KKT-certified in tests/test_oracle_mhe.py)."""
    import pyhostsim as hs
    from decentralized_ekf_mhe_b200 import synth
    st = synth.to_numpy(synth.make_stream(4, 120, robot="pogox", vo_jitter=True))
    kw = dict(robot=2, num_legs=1, contact_effort_threshold=100.0, p_ib=(0.0, 0.0, 0.0))
    A3 = np.zeros((3, 9))
    A3[0, 3], A3[0, 5] = 1.0, 0.5
    A3[1, 6], A3[1, 7] = 1.0, -1.0
    A3[2, 2], A3[2, 8] = 1.0, 0.02
    lo3, hi3 = np.array([0.47, -0.003, -2e-4]), np.array([0.52, 0.003, 2e-4])
    xlo, xhi = [0.0] * 9, [0.0] * 9
    xlo[4], xhi[4] = -0.02, 0.02
    r = hs.run(st, _cfg(x_box_mask=1 << 4, x_box_lo=tuple(xlo), x_box_hi=tuple(xhi), **kw), rows=(A3, lo3, hi3))
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(x_row_count=3, x_row_a=tuple(A3.reshape(-1)) + (0.0,) * 54,
                                                      x_row_lo=tuple(lo3) + (0.0,) * 6, x_row_hi=tuple(hi3) + (0.0,) * 6,
                                                      x_box_mask=1 << 4, x_box_lo=tuple(xlo), x_box_hi=tuple(xhi), **kw),
                                oracle.ekf_params(rate=200), nthreads=4, want=("x", "v_body"))
    assert np.abs(r["x"][1:] - ro["x"][1:]).max() < 1e-8
    rows = np.vstack([A3, np.eye(9)[4:5]])
    lo, hi = np.concatenate([lo3, [-0.02]]), np.concatenate([hi3, [0.02]])
    val = np.einsum("rc,scn->srn", rows, r["x"][1:])
    assert (val <= hi[None, :, None] + 1e-10).all() and (val >= lo[None, :, None] - 1e-10).all()
    bind = ((val >= hi[None, :, None] - 1e-10) | (val <= lo[None, :, None] + 1e-10)).sum(axis=(0, 2))
    assert (bind[:3] > 0).all(), bind
    assert not (r["status"] & 64).any()


def _dup_first(st):
    """Stream with sample 0 delivered twice: the KF alternative runs InitializeKF + UpdateKF on the same sample at
    T == 0 (DecentralEst.cpp:139-141), so x_KF_(T) is the MHE/filter estimate of this stream at T + 1."""
    return {k: np.concatenate([v[:1], v], axis=0) for k, v in st.items()}


def test_foot_state_model_matches_exact_solution(oracle):
    """leg_odom_type 1 (foot-position states, DecentralEst.cpp:101-111, :310-325, :432-452, :550-564).
    The kernel math carries the sweep in information form and matches the EXACT optimum -- the oracle solving the
    whole history in one banded system (N larger than the run: no marginalisation) -- to 1e-8.  The oracle with
    N = 20 restates the reference's marginalizeQP literally (MheSrb.cpp:475-713: dense inverse of the stacked
    observation covariance, which holds the swing-foot variance dt^2 * 1e14 here) and is itself only 2.5e-6 away
    from that optimum; the same bound is asserted for it."""
    import pyhostsim as hs
    from decentralized_ekf_mhe_b200 import synth
    st = synth.to_numpy(synth.make_stream(4, 150, vo=False))
    r = hs.run(st, _cfg(leg_odom_type=1))
    exact, _, _ = oracle.run_batch(st, oracle.go1_params(leg_odom_type=1, N=400), oracle.ekf_params(rate=200), nthreads=4, want=("x", "v_body"))
    ref20, _, _ = oracle.run_batch(st, oracle.go1_params(leg_odom_type=1, N=20), oracle.ekf_params(rate=200), nthreads=4, want=("x",))
    assert r["x"].shape[1] == 21
    assert np.abs(r["x"][1:] - exact["x"][1:]).max() < 1e-8
    assert np.abs(r["v_body"][1:] - exact["v_body"][1:]).max() < 1e-8
    assert np.abs(ref20["x"][1:] - exact["x"][1:]).max() < 1e-5      # noise floor of the reference-form marginalisation
    assert np.abs(r["x"][1:] - ref20["x"][1:]).max() < 1e-5
    # with delayed VO the window matters, so only the literal (N = 20) oracle applies
    st = synth.to_numpy(synth.make_stream(4, 120, vo_jitter=True))
    r = hs.run(st, _cfg(leg_odom_type=1))
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(leg_odom_type=1), oracle.ekf_params(rate=200), nthreads=4, want=("x", "vo_dbg"))
    assert np.abs(r["x"][1:, :9] - ro["x"][1:, :9]).max() < 1e-6 and np.abs(r["x"][1:] - ro["x"][1:]).max() < 1e-5
    assert np.array_equal(r["vo_dbg"], ro["vo_dbg"][:, :8])
    # KF alternative on the same model: exact reference through the duplicated-first-sample stream
    st = synth.to_numpy(synth.make_stream(4, 100, vo=False))
    r = hs.run(st, _cfg(leg_odom_type=1, est_type=1))
    ex, _, _ = oracle.run_batch(_dup_first(st), oracle.go1_params(leg_odom_type=1, N=400), oracle.ekf_params(rate=200), nthreads=4,
                                run_ekf=False, quat_in=_dup_first({"q": r["quat"]})["q"], want=("x",))
    assert np.abs(r["x"][1:] - ex["x"][2:]).max() < 1e-8
    okf, _, _ = oracle.run_batch(st, oracle.go1_params(leg_odom_type=1, est_type=1), oracle.ekf_params(rate=200), nthreads=4, want=("x",))
    assert np.abs(r["x"][1:] - okf["x"][1:]).max() < 1e-5           # the literal covariance-form KF has the same noise floor


@pytest.mark.parametrize("precision", [0, 1])
def test_incremental_solve_is_bit_identical_to_full_resweep(precision):
    """window_solve = DEKF_SOLVE_INCREMENTAL (tier B): restart at the first changed stage from the checkpoint ring.
    Same operations on the same operands => every output and the arrival cost equal the full re-sweep bit for bit."""
    import pyhostsim as hs
    from decentralized_ekf_mhe_b200 import synth
    st = synth.to_numpy(synth.make_stream(8, 260, vo_jitter=True))
    a = hs.run(st, _cfg(window_solve=0, precision=precision))
    b = hs.run(st, _cfg(window_solve=1, precision=precision))
    assert np.array_equal(a["x"][1:], b["x"][1:]) and np.array_equal(a["v_body"][1:], b["v_body"][1:])
    assert np.array_equal(a["arr_P"], b["arr_P"]) and np.array_equal(a["arr_x"], b["arr_x"])
    assert np.array_equal(a["status"], b["status"]) and (a["status"] & 16).any()
