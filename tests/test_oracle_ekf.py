"""Oracle pinning, EKF part: known-answer values of a literal restatement (the reference ships
no tests, so these are hand-derived numbers).  The pin against the reference's OWN compiled sources -- including the EKF
with its VO rewind/replay -- is tests/test_refnodes_pin.py."""
import numpy as np


def _mk(oracle):
    return oracle.Ekf(oracle.ekf_params(rate=500, init_std=(1e-3,) * 4, process_std=(0.1,) * 3,
                                        gravity_meas_std=(4.0,) * 3, vo_meas_std=(1e-4,) * 4))


def test_predict_kat(oracle):
    e = _mk(oracle)
    q, P = e.predict([1.0, 0, 0, 0], [0.1, -0.2, 0.3], 1e-6 * np.eye(4))
    np.testing.assert_allclose(q, [9.999999300000075e-01, 9.999999300000076e-05, -1.999999860000015e-04,
                                   2.999999790000022e-04], rtol=0, atol=1e-15)
    # P[3,3] gets NO process noise because of the W indexing bug (orien_ekf.cpp:285-291)
    np.testing.assert_allclose(np.diag(P), [1.00000014e-06, 1.01000014e-06, 1.01000014e-06, 1.00000014e-06],
                               rtol=1e-9)


def test_correct_and_vo_kat(oracle):
    e = _mk(oracle)
    qp, Pp = e.predict([1.0, 0, 0, 0], [0.1, -0.2, 0.3], 1e-6 * np.eye(4))
    q, P = e.correct(qp, [0.3, -0.2, 9.7], Pp)
    np.testing.assert_allclose(q, [9.999999299504795e-01, 9.974465607449966e-05, -2.003746483096468e-04,
                                   3.000000666421896e-04], rtol=0, atol=1e-15)
    np.testing.assert_allclose(np.diag(P), [9.999755667641671e-07, 1.009975072847103e-06, 1.009975072846351e-06,
                                            1.000000139996560e-06], rtol=1e-12)
    qv = np.array([0.9999, 0.01, -0.005, 0.002])
    qv /= np.linalg.norm(qv)
    q2, P2 = e.vo_correct(q, qv, P)
    np.testing.assert_allclose(q2, [0.999936727054549, 0.009903293689719, -0.004953122394661, 0.001983239792578],
                               rtol=0, atol=2e-15)
    np.testing.assert_allclose(np.diag(P2), [9.900987703771373e-09, 9.901958388334927e-09, 9.901958388334842e-09,
                                             9.900990112733726e-09], rtol=1e-12)


def test_identity_pure_gravity_is_fixed_point(oracle):
    e = _mk(oracle)
    q, P = e.correct([1.0, 0, 0, 0], [0.0, 0.0, 9.81], 1e-6 * np.eye(4))
    np.testing.assert_allclose(q, [1, 0, 0, 0], atol=1e-16)


def test_zero_gyro_predict_adds_buggy_WCW(oracle):
    e = _mk(oracle)
    q0 = np.array([0.9, 0.1, -0.3, 0.2])
    q0 /= np.linalg.norm(q0)
    P0 = np.diag([1e-6, 2e-6, 3e-6, 4e-6])
    q, P = e.predict(q0, [0, 0, 0], P0)
    w, x, y, z = q0
    W = 0.5 * (1 / 500) * np.array([[-x, -y, -z], [w, -z, y], [z, x, w], [-y, 0, 0]])
    np.testing.assert_allclose(P, P0 + W @ (0.01 * np.eye(3)) @ W.T, rtol=1e-12, atol=1e-22)
    np.testing.assert_allclose(q, q0, atol=1e-16)


def test_replay_index_logic(oracle):
    """orien_ekf.cpp:186-205: (cur, idx) -> replayed samples; sample cur-1 is never applied and
    rel <= 1 silently drops the VO measurement."""
    dt = 0.002
    for lag, want_idx, want_n in ((0, 10, 0), (1, 9, 0), (2, 8, 1), (7, 3, 6)):
        e = _mk(oracle)
        rng = np.random.default_rng(1)
        for k in range(10):
            e.tick(rng.normal(0, 0.1, 3), [0, 0, 9.81] + rng.normal(0, 0.05, 3), k * dt)
        q_before, _ = e.get()
        qv = np.array([0.999, 0.02, 0.01, -0.03])
        qv /= np.linalg.norm(qv)
        e.tick([0.0, 0.0, 0.0], [0, 0, 9.81], 10 * dt, vo_quat=qv, vo_time=(10 - lag) * dt + 1e-4)
        cur, idx, n = e.last_replay()
        assert (cur, idx, n) == (10, want_idx, want_n)
        q_after, _ = e.get()
        moved = np.abs(q_after - q_before).max()
        assert (moved > 1e-3) == (want_n > 0)  # VO only applied when at least one sample is replayed


def test_vo_before_history_is_dropped(oracle):
    e = _mk(oracle)
    e.tick([0, 0, 0], [0, 0, 9.81], 1.0, vo_quat=[0, 1, 0, 0], vo_time=0.5)
    assert e.last_replay()[1] == -1
    np.testing.assert_allclose(e.get()[0], [1, 0, 0, 0], atol=1e-12)
