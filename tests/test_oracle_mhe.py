"""Oracle pinning, MHE part.  No reference fixtures exist, so the oracle is
pinned by (1) the QP sizes and delayed-VO index table of a literal reading,
(2) internal consistency: slack-elimination solve == dense KKT of the exported reference-ordered QP,
reference-form marginalisation == information-form Schur complement, OSQP-style ADMM -> same point."""
import numpy as np
import pytest


def _drive(oracle, st, i, steps, prm, on_step=None):
    ekf = oracle.Ekf(oracle.ekf_params(rate=prm.rate))
    m = oracle.Mhe(prm)
    for s in range(steps):
        vo = vq = None
        if st["vo_flag"][s, i]:
            vo = (st["vo_time_pre"][s, i], st["vo_time_now"][s, i], st["vo_rel_p"][s, :, i])
            vq = st["vo_quat"][s, :, i]
        ekf.tick(st["gyro"][s, :, i], st["accel"][s, :, i], st["imu_time"][s, i], vq, st["vo_time_now"][s, i])
        q, _ = ekf.get()
        m.step(s, imu_time=st["imu_time"][s, i], accel=st["accel"][s, :, i], gyro=st["gyro"][s, :, i], quat=q,
               joint_pos=st["joint_pos"][s, :, i], joint_vel=st["joint_vel"][s, :, i],
               foot_force=st["foot_force"][s, :, i], vo=vo)
        if on_step:
            on_step(s, m)
    return m


@pytest.fixture(scope="module")
def lockstep_stream():
    from decentralized_ekf_mhe_b200 import synth
    return synth.to_numpy(synth.make_stream(2, 135))


def test_qp_sizes(oracle, lockstep_stream):
    sizes = {}
    _drive(oracle, lockstep_stream, 0, 25, oracle.go1_params(), lambda s, m: sizes.__setitem__(s, m.dims()[3:]))
    assert sizes[1] == (54, 36)
    assert sizes[19] == (648, 468) and sizes[24] == (648, 468)


def test_vo_index_table(oracle, lockstep_stream):
    """SURVEY.md App. F, N=20, imu_time=T*0.005, frame j at j/30 s, 40 ms latency."""
    rows = {}

    def grab(s, m):
        if lockstep_stream["vo_flag"][s, 0]:
            rows[s] = m.vo_debug()

    _drive(oracle, lockstep_stream, 0, 130, oracle.go1_params(), grab)
    # T: (i_pre, i_now, w0)
    for T, want in {15: (0, 6, 0), 22: (6, 13, 2), 28: (13, 20, 8)}.items():
        assert tuple(rows[T][1:4]) == want and rows[T][8] == 0  # fewer than 4 way-points: no bounds
    # T: (i_pre, i_now, w0, i0, ins, num, first discrete time)
    assert rows[35][1:8] == [20, 26, 15, 20, 5, 7, 20]
    assert rows[122][1:8] == [65, 72, 61, 65, 4, 8, 106] and rows[122][9] == 81  # ring saturated at 4N+1
    assert rows[128][1:8] == [66, 73, 61, 66, 5, 8, 113]


def test_vo_bound_rows(oracle, lockstep_stream):
    """Rows written by Update_Image_bound (MheSrb.cpp:449-459) at T=35: 141,165,...,261."""
    got = {}

    def grab(s, m):
        if s in (34, 35):
            _, _, _, l, u = m.export_qp()
            got[s] = np.where(l == u)[0], l.copy()

    _drive(oracle, lockstep_stream, 0, 36, oracle.go1_params(), grab)
    eq34 = set(got[34][0])
    eq35 = set(got[35][0])
    # at T=35 the window was 15..35 before marginalising stage 15 (24 rows) -> rows shift by -24
    new_rows = sorted(r for r in eq35 if (r + 24) not in eq34 and (r % 24) in (21, 22, 23))
    assert [r for r in new_rows if r % 24 == 21] == [141 - 24 + 24 * i for i in range(6)]


def test_exact_solve_equals_dense_kkt(oracle, go1_stream_small):
    st = go1_stream_small

    def check(s, m):
        if s in (1, 7, 19, 20, 33, 60, 90):
            H, g, A, l, u = m.export_qp()
            eq = l == u
            assert np.all((l[~eq] <= -1e29) & (u[~eq] >= 1e29))  # reference: equality or free rows only
            Ae, be = A[eq], l[eq]
            nV, ne = H.shape[0], Ae.shape[0]
            K = np.block([[H, Ae.T], [Ae, np.zeros((ne, ne))]])
            z = np.linalg.solve(K, np.concatenate([-g, be]))[:nV]
            np.testing.assert_allclose(m.solution(), z, rtol=0, atol=5e-9)
            np.testing.assert_allclose(m.x()[3:6], z[nV - 21 + 3:nV - 21 + 6], rtol=0, atol=1e-9)

    _drive(oracle, st, 3, 91, oracle.go1_params(), check)


def test_marginalisation_equals_information_form(oracle, lockstep_stream):
    """MheSrb.cpp:475-713 (dense (ds+dc+dm)^2 inverse) vs M+ = B'QB - B'QA (M_post + A'QA)^-1 A'QB.
    Lock-step stream: VO bounds arrive 8 ticks late (<< N), so the stage being marginalised at step s
    already carries its final bounds in the QP exported at step s-1."""
    prev = {}
    seen = {"eq": 0, "free": 0}

    def check(s, m):
        if s < 19:
            return
        H, g, A, l, u = m.export_qp()
        if s >= 20:
            Hp, gp, Ap, lp, up = prev["qp"]
            M, n = Hp[:9, :9], gp[:9]
            R, y, Hm = Hp[9:21, 9:21], lp[0:12], Ap[0:12, 0:9]
            Ad, Qd, bd = Ap[12:21, 0:9], Hp[21:30, 21:30], lp[12:21]
            eqvo = bool(np.all(lp[21:24] == up[21:24]))
            seen["eq" if eqvo else "free"] += 1
            At, Bt, Qt, lt = Ad, np.eye(9), Qd, bd
            if eqvo:
                S = np.hstack([np.eye(3), np.zeros((3, 6))])
                At, Bt = np.vstack([Ad, S]), np.vstack([np.eye(9), S])
                Qt = np.block([[Qd, np.zeros((9, 3))], [np.zeros((3, 9)), Hp[30:33, 30:33]]])
                lt = np.concatenate([bd, lp[21:24]])
            Mpost = M + Hm.T @ R @ Hm
            npost = n - Hm.T @ R @ y
            G = Mpost + At.T @ Qt @ At
            Mn = Bt.T @ Qt @ Bt - Bt.T @ Qt @ At @ np.linalg.solve(G, At.T @ Qt @ Bt)
            nn = Bt.T @ Qt @ lt + Bt.T @ Qt @ At @ np.linalg.solve(G, npost - At.T @ Qt @ lt)
            Ma, na = m.arrival()
            scale = np.abs(Mn).max()
            np.testing.assert_allclose(Ma, Mn, rtol=0, atol=1e-10 * scale)
            np.testing.assert_allclose(na, nn, rtol=0, atol=1e-10 * max(1.0, np.abs(nn).max()))
            assert np.abs(Ma - Ma.T).max() <= 1e-12 * scale
            # the arrival cost is what sits on the first state of the new window
            np.testing.assert_array_equal(H[:9, :9], Ma)
            np.testing.assert_array_equal(g[:9], na)
        prev["qp"] = (H, g, A, l, u)

    _drive(oracle, lockstep_stream, 0, 70, oracle.go1_params(), check)
    assert seen["eq"] > 10 and seen["free"] > 10


def test_admm_converges_to_exact(oracle, go1_stream_small):
    st = go1_stream_small
    r0, _, _ = oracle.run_batch(st, oracle.go1_params(solve_mode=0), oracle.ekf_params(rate=200), i1=4, nthreads=4,
                                want=("x",))
    r2, _, _ = oracle.run_batch(st, oracle.go1_params(solve_mode=2, abs_tol=1e-8, relative_tol=1e-8, time_limit=0.0),
                                oracle.ekf_params(rate=200), i1=4, nthreads=4, want=("x", "admm_iters"))
    err = np.abs(r2["x"][1:, 3:6, :4] - r0["x"][1:, 3:6, :4])
    assert r2["admm_iters"][1:, :4].max() < 4000
    # OSQP-style iterates at eps=1e-8 sit within a few 1e-6 m/s of the unique optimum (median ~1e-9)
    assert err.max() < 5e-6 and np.median(err) < 1e-7


def test_kf_alternative_matches_mhe_without_vo(oracle):
    """with no delayed VO the MHE's x_T equals the Kalman filter estimate.
    The reference's KF path runs InitializeKF() AND UpdateKF() on the first tick
    (DecentralEst.cpp:140-141), i.e. it consumes sample 0 twice; feeding the MHE the same duplicated
    sequence aligns the two: x_MHE(T+1 | s0,s0,s1,..) == x_KF(T | s0,s1,..)."""
    from decentralized_ekf_mhe_b200 import synth
    st = synth.to_numpy(synth.make_stream(2, 80, vo=False))
    st_dup = {k: np.concatenate([v[:1], v], axis=0) for k, v in st.items()}
    ep = oracle.ekf_params(rate=200)
    quat = oracle.run_batch(st, oracle.go1_params(), ep, nthreads=2, run_mhe=False, want=("quat",))[0]["quat"]
    quat_dup = np.concatenate([quat[:1], quat], axis=0)
    r_m = oracle.run_batch(st_dup, oracle.go1_params(est_type=0), ep, nthreads=2, run_ekf=False, quat_in=quat_dup,
                           want=("x",))[0]
    r_k = oracle.run_batch(st, oracle.go1_params(est_type=1), ep, nthreads=2, run_ekf=False, quat_in=quat,
                           want=("x",))[0]
    np.testing.assert_allclose(r_m["x"][2:, 3:6], r_k["x"][1:, 3:6], rtol=0, atol=1e-9)


def test_box_constrained_optimum_certificate(oracle):
    """The oracle's constrained solve (builder extension, PogoX config) is pinned by two independent checks:
    (1) KKT certificate on the exported reference-ordered QP: stationarity with multipliers of the right sign on
    the active box rows; (2) the OSQP-style ADMM restatement with the same rows converges to the same point."""
    from decentralized_ekf_mhe_b200 import synth
    lo, hi = np.array((-0.45, -0.03, -0.015)), np.array((0.55, 0.03, 0.015))
    st = synth.to_numpy(synth.make_stream(1, 60, robot="pogox", vo_jitter=True, truth=True))
    kw = dict(robot=2, num_legs=1, contact_effort_threshold=100.0, p_ib=(0.0, 0.0, 0.0), v_box_enable=1,
              v_box_lo=tuple(lo), v_box_hi=tuple(hi))
    m = oracle.Mhe(oracle.go1_params(**kw))
    ma = oracle.Mhe(oracle.go1_params(solve_mode=2, abs_tol=1e-9, relative_tol=1e-9, max_qp_iter=20000, time_limit=0.0, **kw))
    n_active_seen = 0
    worst_admm = 0.0
    for s in range(60):
        q = st["quat_true"][s, :, 0]
        args = dict(imu_time=st["imu_time"][s, 0], accel=st["accel"][s, :, 0], gyro=st["gyro"][s, :, 0], quat=q,
                    joint_pos=st["joint_pos"][s, :, 0], joint_vel=st["joint_vel"][s, :, 0],
                    foot_force=st["foot_force"][s, :, 0],
                    vo=(st["vo_time_pre"][s, 0], st["vo_time_now"][s, 0], st["vo_rel_p"][s, :, 0]) if st["vo_flag"][s, 0] else None)
        m.step(s, **args)
        ma.step(s, **args)
        if s < 1:
            continue
        ds, dm, dc, nV, nC = m.dims()
        H, g, A, l, u = m.export_qp()
        z = m.solution()
        K = (nV + ds + dc) // (2 * ds + dm + dc)
        xi = [j * (2 * ds + dm + dc) for j in range(K)]
        v = np.array([z[o + 3:o + 6] for o in xi])
        assert (v <= hi + 1e-12).all() and (v >= lo - 1e-12).all()
        eq = np.abs(u - l) < 1e-9
        assert np.abs((A @ z - l)[eq]).max() < 1e-9 * max(1.0, np.abs(l[eq]).max())   # primal feasibility of the reference rows
        act = [(j, c, +1 if v[j, c] >= hi[c] else -1) for j in range(K) for c in range(3) if v[j, c] >= hi[c] or v[j, c] <= lo[c]]
        n_active_seen += len(act)
        # stationarity: H z + g + A_eq' y + B_act' mu = 0 for some y and mu with sign(mu) = side of the bound
        B = np.zeros((len(act), nV))
        for r, (j, c, sgn) in enumerate(act):
            B[r, xi[j] + 3 + c] = 1.0
        G = np.vstack([A[eq], B]).T
        rhs = -(H @ z + g)
        scale = np.abs(rhs).max() + 1.0
        sol, *_ = np.linalg.lstsq(G, rhs, rcond=None)
        assert np.abs(G @ sol - rhs).max() < 1e-7 * scale
        mu = sol[eq.sum():]
        for (j, c, sgn), w in zip(act, mu):
            assert sgn * w >= -1e-7 * scale
        if ma.admm_iters() < 20000:
            worst_admm = max(worst_admm, np.abs(ma.x()[3:6] - m.x()[3:6]).max())
    assert n_active_seen > 20
    assert worst_admm < 1e-6


def test_foot_state_oracle_consistency(oracle):
    """leg_odom_type 1 in the oracle: the exact banded solve equals the dense KKT solution of the exported
    reference-ordered QP (nVar/nCon as MheSrb.cpp would count them), and the literal reference-form marginalisation
    stays within its 1e-5 noise floor of the no-marginalisation optimum (see tests/test_hostsim.py)."""
    from decentralized_ekf_mhe_b200 import synth
    st = synth.to_numpy(synth.make_stream(1, 30, vo_jitter=True, truth=True))
    m = oracle.Mhe(oracle.go1_params(leg_odom_type=1))
    for s in range(30):
        m.step(s, imu_time=st["imu_time"][s, 0], accel=st["accel"][s, :, 0], gyro=st["gyro"][s, :, 0], quat=st["quat_true"][s, :, 0],
               joint_pos=st["joint_pos"][s, :, 0], joint_vel=st["joint_vel"][s, :, 0], foot_force=st["foot_force"][s, :, 0],
               vo=(st["vo_time_pre"][s, 0], st["vo_time_now"][s, 0], st["vo_rel_p"][s, :, 0]) if st["vo_flag"][s, 0] else None)
        if s in (1, 5, 19, 20, 29):
            ds, dm, dc, nV, nC = m.dims()
            assert ds == 21 and dm == 12
            K = min(s + 1, 20)
            assert nV == K * (ds + dm) + (K - 1) * (ds + dc) and nC == K * dm + (K - 1) * (ds + dc)
            H, g, A, l, u = m.export_qp()
            eq = np.abs(u - l) < 1e-9
            Ae, be = A[eq], l[eq]
            KKT = np.block([[H, Ae.T], [Ae, np.zeros((Ae.shape[0], Ae.shape[0]))]])
            sol = np.linalg.solve(KKT, np.concatenate([-g, be]))[:nV]
            z = m.solution()
            xT, xk = z[nV - ds - dm:nV - dm], sol[nV - ds - dm:nV - dm]
            assert np.abs(xT - m.x()).max() == 0.0
            assert np.abs(xT - xk).max() < 1e-7, (s, np.abs(xT - xk).max())


def test_general_component_bounds_optimum_certificate(oracle):
    """General rows  lb <= x_k[a] <= ub  on ANY state component (cfg.x_box_mask: here v_x, the z position and the three accel-bias
    components at once -- what MHEproblem::addConstraints(name, lb, ub) with a selector row per component would add,
    MheSrb.cpp:58-68): the oracle's constrained optimum carries a KKT certificate on the exported reference-ordered QP."""
    from decentralized_ekf_mhe_b200 import synth
    mask = (1 << 3) | (1 << 2) | (7 << 6)
    lo9 = np.array([0, 0, -2e-4, 0.47, 0, 0, -0.004, -0.004, -0.004], float)
    hi9 = np.array([0, 0, 2e-4, 0.52, 0, 0, 0.004, 0.004, 0.004], float)
    st = synth.to_numpy(synth.make_stream(1, 60, robot="pogox", vo_jitter=True, truth=True))
    kw = dict(robot=2, num_legs=1, contact_effort_threshold=100.0, p_ib=(0.0, 0.0, 0.0), x_box_mask=mask,
              x_box_lo=tuple(lo9), x_box_hi=tuple(hi9))
    m = oracle.Mhe(oracle.go1_params(**kw))
    comps = [a for a in range(9) if (mask >> a) & 1]
    seen = {a: 0 for a in comps}
    for s in range(60):
        q = st["quat_true"][s, :, 0]
        m.step(s, imu_time=st["imu_time"][s, 0], accel=st["accel"][s, :, 0], gyro=st["gyro"][s, :, 0], quat=q,
               joint_pos=st["joint_pos"][s, :, 0], joint_vel=st["joint_vel"][s, :, 0], foot_force=st["foot_force"][s, :, 0],
               vo=(st["vo_time_pre"][s, 0], st["vo_time_now"][s, 0], st["vo_rel_p"][s, :, 0]) if st["vo_flag"][s, 0] else None)
        if s < 1:
            continue
        ds, dm, dc, nV, nC = m.dims()
        H, g, A, l, u = m.export_qp()
        z = m.solution()
        K = (nV + ds + dc) // (2 * ds + dm + dc)
        xi = [j * (2 * ds + dm + dc) for j in range(K)]
        X = np.array([z[o:o + 9] for o in xi])
        for a in comps:
            assert (X[:, a] <= hi9[a] + 1e-12).all() and (X[:, a] >= lo9[a] - 1e-12).all()
        eq = np.abs(u - l) < 1e-9
        assert np.abs((A @ z - l)[eq]).max() < 1e-9 * max(1.0, np.abs(l[eq]).max())
        act = [(j, a, +1 if X[j, a] >= hi9[a] else -1) for j in range(K) for a in comps if X[j, a] >= hi9[a] or X[j, a] <= lo9[a]]
        for _, a, _ in act:
            seen[a] += 1
        B = np.zeros((len(act), nV))
        for r, (j, a, sgn) in enumerate(act):
            B[r, xi[j] + a] = 1.0
        G = np.vstack([A[eq], B]).T
        rhs = -(H @ z + g)
        scale = np.abs(rhs).max() + 1.0
        sol, *_ = np.linalg.lstsq(G, rhs, rcond=None)
        assert np.abs(G @ sol - rhs).max() < 1e-7 * scale                      # stationarity
        for (j, a, sgn), w in zip(act, sol[eq.sum():]):
            assert sgn * w >= -1e-7 * scale                                     # multiplier signs
    assert seen[2] > 5 and sum(seen[a] for a in (6, 7, 8)) > 20 and seen[3] > 0   # position, bias and velocity rows all bind


def _general_rows():
    """Three non-selector rows on the PogoX state: a velocity combination, a bias difference, position against bias."""
    A = np.zeros((3, 9))
    A[0, 3], A[0, 5] = 1.0, 0.5          # v_x + 0.5 v_z
    A[1, 6], A[1, 7] = 1.0, -1.0         # b_x - b_y
    A[2, 2], A[2, 8] = 1.0, 0.02         # p_z + 0.02 b_z
    lo = np.array([0.47, -0.003, -2e-4])
    hi = np.array([0.52, 0.003, 2e-4])
    return A, lo, hi


def test_general_linear_rows_optimum_certificate(oracle):
    """Arbitrary rows  lb <= a . x_k <= ub  on every window state (orc_params.x_row_*: what MHEproblem::addConstraints(name, lb, ub)
    with a non-selector dependency row on x_k would add, MheSrb.cpp:58-68, :217-270), here three rows that mix velocity, bias and
    position components, TOGETHER with a component bound (x_box on v_y): the oracle solves in the basis y = W x in which the rows are
    coordinates; its optimum carries a KKT certificate on the exported reference-ordered QP with the rows themselves as the
    active constraint normals."""
    from decentralized_ekf_mhe_b200 import synth
    A3, lo3, hi3 = _general_rows()
    rows = np.vstack([A3, np.eye(9)[4:5]])                 # + the component bound on v_y as a unit row, for the certificate
    lo = np.concatenate([lo3, [-0.02]])
    hi = np.concatenate([hi3, [0.02]])
    st = synth.to_numpy(synth.make_stream(1, 60, robot="pogox", vo_jitter=True, truth=True))
    xlo, xhi = [0.0] * 9, [0.0] * 9
    xlo[4], xhi[4] = -0.02, 0.02
    kw = dict(robot=2, num_legs=1, contact_effort_threshold=100.0, p_ib=(0.0, 0.0, 0.0), x_row_count=3,
              x_row_a=tuple(A3.reshape(-1)) + (0.0,) * 54, x_row_lo=tuple(lo3) + (0.0,) * 6, x_row_hi=tuple(hi3) + (0.0,) * 6,
              x_box_mask=1 << 4, x_box_lo=tuple(xlo), x_box_hi=tuple(xhi))
    m = oracle.Mhe(oracle.go1_params(**kw))
    seen = np.zeros(len(rows), int)
    for s in range(60):
        q = st["quat_true"][s, :, 0]
        m.step(s, imu_time=st["imu_time"][s, 0], accel=st["accel"][s, :, 0], gyro=st["gyro"][s, :, 0], quat=q,
               joint_pos=st["joint_pos"][s, :, 0], joint_vel=st["joint_vel"][s, :, 0], foot_force=st["foot_force"][s, :, 0],
               vo=(st["vo_time_pre"][s, 0], st["vo_time_now"][s, 0], st["vo_rel_p"][s, :, 0]) if st["vo_flag"][s, 0] else None)
        if s < 1:
            continue
        ds, dm, dc, nV, nC = m.dims()
        H, g, A, l, u = m.export_qp()
        z = m.solution()
        K = (nV + ds + dc) // (2 * ds + dm + dc)
        xi = [j * (2 * ds + dm + dc) for j in range(K)]
        X = np.array([z[o:o + 9] for o in xi])
        val = X @ rows.T                                                      # [K, rows]
        tol = 1e-9
        assert (val <= hi + tol).all() and (val >= lo - tol).all()            # feasibility of every row at every state
        eq = np.abs(u - l) < 1e-9
        assert np.abs((A @ z - l)[eq]).max() < 1e-9 * max(1.0, np.abs(l[eq]).max())
        act = [(j, r, +1 if val[j, r] >= hi[r] - tol else -1) for j in range(K) for r in range(len(rows))
               if val[j, r] >= hi[r] - tol or val[j, r] <= lo[r] + tol]
        for _, r, _ in act:
            seen[r] += 1
        B = np.zeros((len(act), nV))
        for k, (j, r, sgn) in enumerate(act):
            B[k, xi[j]:xi[j] + 9] = rows[r]
        G = np.vstack([A[eq], B]).T
        rhs = -(H @ z + g)
        scale = np.abs(rhs).max() + 1.0
        sol, *_ = np.linalg.lstsq(G, rhs, rcond=None)
        assert np.abs(G @ sol - rhs).max() < 1e-7 * scale                      # stationarity
        for (j, r, sgn), w in zip(act, sol[eq.sum():]):
            assert sgn * w >= -1e-7 * scale                                     # multiplier signs
    assert (seen[:3] > 0).all(), seen                                           # every general row binds somewhere
