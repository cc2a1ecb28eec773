"""The reference's DEPLOYMENT schedule: orientation EKF at 500 Hz (orien_est/config/parameters.yaml `rate: 500`) and MHE at
200 Hz (parameters_go1.yaml `estimation.rate: 200`, `interval: 5`) as two timers fed by the same sensor topics -- not the
1:1 lock-step of the benchmark.  Shared by tests/golden/make_refnodes_golden.py (reference side), tests/test_refnodes_pin.py
(oracle side) and tests/test_gpu_parity.py (CUDA side) so that all three replay exactly the same event order:

  t = 2 ms * s : /unitree/imu and /unitree/joint_state sample s (and the VO messages that arrive at s), then the EKF timer
  t = 5 ms * T : the estimator timer (after the EKF timer when both fall on the same millisecond): initialize at T = 0,
                 update(T) afterwards, reading the LATEST IMU / joint sample, the latest imu/filter orientation and the VO
                 message latched since its previous tick.
"""
import numpy as np

IMU_PERIOD_NS = 2_000_000
MHE_PERIOD_NS = 5_000_000


def events(S):
    """[(kind, index)] in time order for S IMU samples: ('imu', s) / ('mhe', T)."""
    ev = []
    t_end = (S - 1) * IMU_PERIOD_NS
    T = 0
    for s in range(S):
        t = s * IMU_PERIOD_NS
        while T * MHE_PERIOD_NS < t:
            ev.append(("mhe", T))
            T += 1
        ev.append(("imu", s))
    while T * MHE_PERIOD_NS <= t_end:
        ev.append(("mhe", T))
        T += 1
    return ev, T  # T = number of estimator ticks


def run_reference(pr, st, i, prm, ep):
    """One instance through the compiled reference nodes (oracle/pyref.RefNodes).  Returns quat [S][4], x [T][ds], v_body [T][3],
    p_vo [T][3], contact [T][nl]."""
    S = st["gyro"].shape[0]
    nl = prm.num_legs
    ev, nT = events(S)
    rn = pr.RefNodes(prm, ep)
    ds = 9 + 3 * prm.leg_odom_type * nl
    out = dict(quat=np.zeros((S, 4)), x=np.full((nT, ds), np.nan), v_body=np.full((nT, 3), np.nan), p_vo=np.zeros((nT, 3)),
               contact=np.zeros((nT, nl), np.uint8))
    for kind, k in ev:
        if kind == "imu":
            rn.set_clock(int(st["imu_ns"][k, i]))
            rn.msg_imu(st["gyro"][k, :, i], st["accel"][k, :, i])
            rn.msg_joint(np.concatenate([st["joint_pos"][k, :, i], st["foot_force"][k, :nl, i]]), st["joint_vel"][k, :, i])
            if st["vo_flag"][k, i]:
                rn.msg_vo(st["vo_quat"][k, :, i], st["vo_now_ns"][k, i], st["vo_pre_ns"][k, i], st["vo_rel_p"][k, :, i])
            rn.fire_ekf()
            out["quat"][k] = rn.quat()[0]
        else:
            rn.set_clock(k * MHE_PERIOD_NS)
            rn.fire_est()
            e = rn.est()
            if k >= 1:
                out["x"][k], out["v_body"][k] = e["x"], e["v_body"]
            out["p_vo"][k], out["contact"][k] = e["p_vo"], e["contact"] != 0
    del rn
    return out


def run_oracle(po, st, i, prm, ep):
    """The same schedule through the oracle objects (oracle/pyoracle.Ekf / Mhe)."""
    S = st["gyro"].shape[0]
    nl = prm.num_legs
    ev, nT = events(S)
    e, m = po.Ekf(ep), po.Mhe(prm)
    ds = 9 + 3 * prm.leg_odom_type * nl
    out = dict(quat=np.zeros((S, 4)), x=np.full((nT, ds), np.nan), v_body=np.full((nT, 3), np.nan), p_vo=np.zeros((nT, 3)),
               contact=np.zeros((nT, nl), np.uint8))
    last, vo, q = None, None, np.array([1.0, 0.0, 0.0, 0.0])
    for kind, k in ev:
        if kind == "imu":
            last = k
            vn = bool(st["vo_flag"][k, i])
            if vn:
                vo = (st["vo_time_pre"][k, i], st["vo_time_now"][k, i], st["vo_rel_p"][k, :, i])
            e.tick(st["gyro"][k, :, i], st["accel"][k, :, i], st["imu_time"][k, i], st["vo_quat"][k, :, i] if vn else None,
                   st["vo_time_now"][k, i] if vn else 0.0)
            q = e.get()[0]
            out["quat"][k] = q
        else:
            m.step(k, imu_time=st["imu_time"][last, i], accel=st["accel"][last, :, i], gyro=st["gyro"][last, :, i], quat=q,
                   joint_pos=st["joint_pos"][last, :, i], joint_vel=st["joint_vel"][last, :, i],
                   foot_force=st["foot_force"][last, :, i], vo=vo)
            vo = None
            if k >= 1:
                out["x"][k], out["v_body"][k] = m.x(), m.v_body()
            out["p_vo"][k], out["contact"][k] = m.p_vo(), np.asarray(m.contact()) != 0
    return out


def run_cuda(E, torch, st, prm_kw, ekf_rate=500):
    """All instances at once through the CUDA path with the reference's class API (E = decentralized_ekf_mhe_b200.estimator):
    E.orien_ekf.timerCallback on every IMU sample, E.DecentralizedEstimation.initialize/update on every estimator tick."""
    S, _, n = st["gyro"].shape
    ev, nT = events(S)
    d = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in st.items() if not k.endswith("_ns")}
    prm = E.robot_params("go1", ekf_rate=ekf_rate, **prm_kw)
    ekf = E.orien_ekf(prm, n)
    est = E.DecentralizedEstimation(n)
    ds = 9
    out = dict(quat=np.zeros((S, 4, n)), x=np.full((nT, ds, n), np.nan), v_body=np.full((nT, 3, n), np.nan),
               p_vo=np.zeros((nT, 3, n)), contact=np.zeros((nT, prm.num_legs, n), np.uint8))
    store = E.robot_store()  # the estimator's robot_store: latest samples, VO latched until the estimator timer consumes it
    z3, z1 = torch.zeros(3, n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    store.vo_p_body_pre_2_body_, store.vo_time_pre_, store.vo_time_now_ = z3.clone(), z1.clone(), z1.clone()
    store.vo_quaternion_ = torch.zeros(4, n, dtype=torch.float64, device="cuda")
    latched = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for kind, k in ev:
        if kind == "imu":
            se = E.robot_store.from_stream(d, k)           # the EKF node's view of sample k (its VO flag is consumed at once)
            ekf.timerCallback(se)
            out["quat"][k] = ekf.quaternion_.cpu().numpy()
            store.imu_time_, store.accel_b_, store.angular_b_ = d["imu_time"][k], d["accel"][k], d["gyro"][k]
            store.joint_states_position_, store.joint_states_velocity_, store.foot_force_ = d["joint_pos"][k], d["joint_vel"][k], d["foot_force"][k]
            store.quaternion_ = ekf.quaternion_.clone()
            f = d["vo_flag"][k].bool()
            if bool(f.any()):
                store.vo_p_body_pre_2_body_ = torch.where(f[None, :], d["vo_rel_p"][k], store.vo_p_body_pre_2_body_)
                store.vo_time_pre_ = torch.where(f, d["vo_time_pre"][k], store.vo_time_pre_)
                store.vo_time_now_ = torch.where(f, d["vo_time_now"][k], store.vo_time_now_)
                latched = latched | f.to(torch.uint8)
        else:
            store.vo_new_ = latched.clone() if bool(latched.any()) else None
            latched = torch.zeros_like(latched)
            if k == 0:
                est.initialize(store, prm)
            else:
                est.update(k)
                out["x"][k] = est.x_MHE_.cpu().numpy()
                out["v_body"][k] = est.v_MHE_b_.cpu().numpy()
            out["p_vo"][k] = est.p_vo_accmulate_.cpu().numpy()
            out["contact"][k] = est.contact_.cpu().numpy()
    return out
