"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  The product package ``decentralized_ekf_mhe_b200`` never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
up = C.POINTER(C.c_ubyte)

ROBOT_GO1, ROBOT_CASSIE, ROBOT_POGOX = 0, 1, 2


class Params(C.Structure):
    """orc_params (oracle.h) == robot_params (DecentralEst.hpp:18-63) + builder additions."""
    _fields_ = [
        ("p_process_std", C.c_double * 3), ("accel_input_std", C.c_double * 3),
        ("accel_bias_std", C.c_double * 3), ("gyro_input_std", C.c_double * 3),
        ("quaternion_ib", C.c_double * 4), ("p_ib", C.c_double * 3),
        ("num_legs", C.c_int), ("leg_odom_type", C.c_int),
        ("joint_position_std", C.c_double * 8), ("joint_velocity_std", C.c_double * 8),
        ("foot_slide_std", C.c_double * 3), ("foot_swing_std", C.c_double * 3),
        ("contact_effort_threshold", C.c_double),
        ("p_init_std", C.c_double * 3), ("v_init_std", C.c_double * 3),
        ("foot_init_std", C.c_double * 3), ("accel_bias_init_std", C.c_double * 3),
        ("vo_p_std", C.c_double * 3),
        ("rate", C.c_int), ("N", C.c_int), ("est_type", C.c_int),
        ("rho", C.c_double), ("alpha", C.c_double), ("delta", C.c_double), ("sigma", C.c_double),
        ("verbose", C.c_int), ("adapt_rho", C.c_int), ("polish", C.c_int), ("max_qp_iter", C.c_int),
        ("relative_tol", C.c_double), ("abs_tol", C.c_double), ("prim_tol", C.c_double),
        ("dual_tol", C.c_double), ("time_limit", C.c_double),
        ("robot", C.c_int), ("solve_mode", C.c_int),
        ("v_box_enable", C.c_int), ("v_box_lo", C.c_double * 3), ("v_box_hi", C.c_double * 3),
        ("p_imu_2_opti", C.c_double * 3),
        ("x_box_mask", C.c_int), ("x_box_lo", C.c_double * 9), ("x_box_hi", C.c_double * 9),
        ("x_row_count", C.c_int), ("x_row_a", C.c_double * 81), ("x_row_lo", C.c_double * 9), ("x_row_hi", C.c_double * 9),
    ]


class Sample(C.Structure):
    _fields_ = [
        ("imu_time", C.c_double), ("accel_b", C.c_double * 3), ("angular_b", C.c_double * 3),
        ("quaternion", C.c_double * 4), ("joint_pos", C.c_double * 40), ("joint_vel", C.c_double * 40),
        ("vo_new", C.c_int), ("vo_time_pre", C.c_double), ("vo_time_now", C.c_double),
        ("vo_p", C.c_double * 3),
    ]


class Stream(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("S", C.c_int), ("nq", C.c_int), ("nlegs", C.c_int),
        ("gyro", dp), ("accel", dp), ("imu_time", dp), ("joint_pos", dp), ("joint_vel", dp),
        ("foot_force", dp), ("vo_flag", up), ("vo_quat", dp), ("vo_time_pre", dp),
        ("vo_time_now", dp), ("vo_rel_p", dp), ("quat_in", dp),
    ]


class EkfParams(C.Structure):
    _fields_ = [
        ("init_std", C.c_double * 4), ("process_std", C.c_double * 3),
        ("gravity_meas_std", C.c_double * 3), ("vo_meas_std", C.c_double * 4),
        ("quaternion_init", C.c_double * 4), ("rate", C.c_int),
    ]


class Outputs(C.Structure):
    _fields_ = [
        ("quat", dp), ("x", dp), ("v_body", dp), ("p_vo", dp), ("contact", up),
        ("vo_dbg", ip), ("ekf_dbg", ip), ("M_p", dp), ("n_p", dp), ("admm_iters", ip),
    ]


def build(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in
            ("la.c", "ekf.c", "kin.c", "mhe.c", "admm.c", "runner.c", "la.h", "oracle.h")]
    stale = force or not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    # oracle/_ref: the reference's own sources compiled where they lie (build container only; the GPU box uses the
    # prebuilt files and the committed golden vectors)
    if os.path.isdir("/root/reference/src/go1_example/src/Expressions"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_ekf_create.restype = C.c_void_p
        L.orc_ekf_create.argtypes = [dp, dp, dp, dp, dp, C.c_int]
        L.orc_ekf_destroy.argtypes = [C.c_void_p]
        L.orc_ekf_tick.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_int, dp, C.c_double]
        L.orc_ekf_get.argtypes = [C.c_void_p, dp, dp]
        L.orc_ekf_set.argtypes = [C.c_void_p, dp, dp]
        L.orc_ekf_last_replay.argtypes = [C.c_void_p, ip, ip, ip]
        for f in (L.orc_ekf_predict, L.orc_ekf_correct, L.orc_ekf_vo_correct):
            f.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
        L.orc_quat_to_rot.argtypes = [dp, dp]
        L.orc_leg_fk.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.orc_params_go1_defaults.argtypes = [C.POINTER(Params)]
        L.orc_mhe_create.restype = C.c_void_p
        L.orc_mhe_create.argtypes = [C.POINTER(Params)]
        L.orc_mhe_destroy.argtypes = [C.c_void_p]
        L.orc_mhe_step.argtypes = [C.c_void_p, C.c_int, C.POINTER(Sample)]
        L.orc_mhe_get_x.argtypes = [C.c_void_p, dp]
        L.orc_mhe_get_v_body.argtypes = [C.c_void_p, dp]
        L.orc_mhe_get_R_sb.argtypes = [C.c_void_p, dp]
        L.orc_mhe_get_p_vo.argtypes = [C.c_void_p, dp]
        L.orc_mhe_get_arrival.argtypes = [C.c_void_p, dp, dp]
        L.orc_mhe_get_contact.argtypes = [C.c_void_p, dp]
        L.orc_mhe_get_meas.argtypes = [C.c_void_p, dp, dp]
        L.orc_mhe_get_kin.argtypes = [C.c_void_p, dp, dp]
        L.orc_mhe_get_dims.argtypes = [C.c_void_p, ip, ip, ip, ip, ip]
        L.orc_mhe_get_vo_debug.argtypes = [C.c_void_p, ip]
        L.orc_mhe_export_qp.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
        L.orc_mhe_get_solution.argtypes = [C.c_void_p, dp]
        L.orc_mhe_get_admm_iters.argtypes = [C.c_void_p]
        L.orc_mhe_get_kf.argtypes = [C.c_void_p, dp, dp, dp]
        L.orc_run_batch.restype = C.c_double
        L.orc_run_batch.argtypes = [C.POINTER(Params), C.POINTER(EkfParams), C.POINTER(Stream),
                                    C.POINTER(Outputs), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, dp, dp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(dp)


def go1_params(**over):
    p = Params()
    lib().orc_params_go1_defaults(C.byref(p))
    for k, v in over.items():
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(p, k, v)
    return p


def ekf_params(rate=500, init_std=(1e-3,) * 4, process_std=(0.1,) * 3, gravity_meas_std=(4.0,) * 3,
               vo_meas_std=(1e-4,) * 4, quaternion_init=(1.0, 0.0, 0.0, 0.0)):
    """Defaults of parameters_go1.yaml:68-75 (orien_sub)."""
    e = EkfParams()
    e.init_std[:] = init_std
    e.process_std[:] = process_std
    e.gravity_meas_std[:] = gravity_meas_std
    e.vo_meas_std[:] = vo_meas_std
    e.quaternion_init[:] = quaternion_init
    e.rate = rate
    return e


class Ekf:
    def __init__(self, ep=None):
        self.ep = ep or ekf_params()
        e = self.ep
        self.h = lib().orc_ekf_create(e.init_std, e.process_std, e.gravity_meas_std, e.vo_meas_std,
                                      e.quaternion_init, e.rate)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ekf_destroy(self.h)
            self.h = None

    def tick(self, gyro, accel, t, vo_quat=None, vo_time=0.0):
        g = np.ascontiguousarray(gyro, dtype=np.float64)
        a = np.ascontiguousarray(accel, dtype=np.float64)
        vq = np.ascontiguousarray(vo_quat if vo_quat is not None else [1, 0, 0, 0], dtype=np.float64)
        lib().orc_ekf_tick(self.h, _p(g), _p(a), float(t), int(vo_quat is not None), _p(vq), float(vo_time))

    def get(self):
        q = np.zeros(4)
        P = np.zeros(16)
        lib().orc_ekf_get(self.h, _p(q), _p(P))
        return q, P.reshape(4, 4)

    def set(self, q, P):
        q = np.ascontiguousarray(q, dtype=np.float64)
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(-1)
        lib().orc_ekf_set(self.h, _p(q), _p(P))

    def last_replay(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().orc_ekf_last_replay(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def _op(self, fn, q, v, P):
        q = np.ascontiguousarray(q, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(-1)
        qo = np.zeros(4)
        Po = np.zeros(16)
        fn(self.h, _p(qo), _p(q), _p(v), _p(P), _p(Po))
        return qo, Po.reshape(4, 4)

    def predict(self, q, gyro, P):
        return self._op(lib().orc_ekf_predict, q, gyro, P)

    def correct(self, q, accel, P):
        return self._op(lib().orc_ekf_correct, q, accel, P)

    def vo_correct(self, q, q_vo, P):
        return self._op(lib().orc_ekf_vo_correct, q, q_vo, P)


def leg_fk(robot, leg, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    p = np.zeros(3)
    J = np.zeros(3 * len(q))
    lib().orc_leg_fk(robot, leg, _p(q), _p(p), _p(J))
    return p, J.reshape(3, len(q))


def quat_to_rot(q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    R = np.zeros(9)
    lib().orc_quat_to_rot(_p(q), _p(R))
    return R.reshape(3, 3)


class Mhe:
    def __init__(self, prm=None):
        self.prm = prm or go1_params()
        self.h = lib().orc_mhe_create(C.byref(self.prm))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_mhe_destroy(self.h)
            self.h = None

    def dims(self):
        v = [C.c_int() for _ in range(5)]
        lib().orc_mhe_get_dims(self.h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def step(self, T, *, imu_time, accel, gyro, quat, joint_pos, joint_vel, foot_force, vo=None):
        s = Sample()
        s.imu_time = imu_time
        s.accel_b[:] = list(accel)
        s.angular_b[:] = list(gyro)
        s.quaternion[:] = list(quat)
        nq = len(joint_pos)
        for i in range(nq):
            s.joint_pos[i] = joint_pos[i]
            s.joint_vel[i] = joint_vel[i]
        for i, f in enumerate(foot_force):
            s.joint_pos[nq + i] = f
        if vo is not None:
            s.vo_new = 1
            s.vo_time_pre, s.vo_time_now = vo[0], vo[1]
            s.vo_p[:] = list(vo[2])
        lib().orc_mhe_step(self.h, T, C.byref(s))

    def _vec(self, fn, n):
        a = np.zeros(n)
        fn(self.h, _p(a))
        return a

    def x(self):
        return self._vec(lib().orc_mhe_get_x, self.dims()[0])

    def v_body(self):
        return self._vec(lib().orc_mhe_get_v_body, 3)

    def R_sb(self):
        return self._vec(lib().orc_mhe_get_R_sb, 9).reshape(3, 3)

    def p_vo(self):
        return self._vec(lib().orc_mhe_get_p_vo, 3)

    def arrival(self):
        ds = self.dims()[0]
        M = np.zeros(ds * ds)
        n = np.zeros(ds)
        ok = lib().orc_mhe_get_arrival(self.h, _p(M), _p(n))
        return (M.reshape(ds, ds), n) if ok else None

    def contact(self):
        return self._vec(lib().orc_mhe_get_contact, self.prm.num_legs)

    def meas(self):
        dm = self.dims()[1]
        b = np.zeros(dm)
        Q = np.zeros(dm * dm)
        lib().orc_mhe_get_meas(self.h, _p(b), _p(Q))
        return b, Q.reshape(dm, dm)

    def kin(self):
        """p_imu_2_foot_ (3*legs) and J_imu_2_foot_ (3*legs x joints-per-leg, row-major) of the last sample."""
        dm = self.dims()[1]
        nj = lib().orc_robot_joints_per_leg(self.prm.robot)
        p = np.zeros(dm)
        J = np.zeros(dm * nj)
        lib().orc_mhe_get_kin(self.h, _p(p), _p(J))
        return p, J.reshape(dm, nj)

    def vo_debug(self):
        a = (C.c_int * 10)()
        lib().orc_mhe_get_vo_debug(self.h, a)
        return list(a)

    def export_qp(self):
        ds, dm, dc, nV, nC = self.dims()
        H = np.zeros(nV * nV)
        g = np.zeros(nV)
        A = np.zeros(nC * nV)
        l = np.zeros(nC)
        u = np.zeros(nC)
        lib().orc_mhe_export_qp(self.h, _p(H), _p(g), _p(A), _p(l), _p(u))
        return H.reshape(nV, nV), g, A.reshape(nC, nV), l, u

    def solution(self):
        return self._vec(lib().orc_mhe_get_solution, self.dims()[3])

    def admm_iters(self):
        return lib().orc_mhe_get_admm_iters(self.h)

    def kf(self):
        ds = self.dims()[0]
        x = np.zeros(ds)
        Cm = np.zeros(ds * ds)
        v = np.zeros(3)
        lib().orc_mhe_get_kf(self.h, _p(x), _p(Cm), _p(v))
        return x, Cm.reshape(ds, ds), v


def pin_threads(on=True):
    """Pin worker thread t of run_batch to the t-th CPU of the process's affinity mask (bench.py's CPU baselines)."""
    lib().orc_set_pin_threads(int(bool(on)))


def run_batch(stream, prm=None, ep=None, *, i0=0, i1=None, nthreads=1, run_ekf=True, run_mhe=True,
              quat_in=None, t_steady=0, want=("quat", "x", "v_body", "contact", "vo_dbg", "ekf_dbg", "p_vo")):
    """Run the oracle over a stream dict of numpy arrays laid out [S][k][n] (see synth.py).

    Returns (outputs dict, wall seconds, busy seconds)."""
    prm = prm or go1_params()
    ep = ep or ekf_params(rate=prm.rate)
    S, _, n = stream["gyro"].shape
    i1 = n if i1 is None else i1
    nl = prm.num_legs
    nq = stream["joint_pos"].shape[1]
    ds = 9 + 3 * prm.leg_odom_type * nl
    keep = {}

    def arr(name, dtype=np.float64):
        a = np.ascontiguousarray(stream[name], dtype=dtype)
        keep[name] = a
        return a

    st = Stream()
    st.n, st.S, st.nq, st.nlegs = n, S, nq, nl
    st.gyro = _p(arr("gyro"))
    st.accel = _p(arr("accel"))
    st.imu_time = _p(arr("imu_time"))
    st.joint_pos = _p(arr("joint_pos"))
    st.joint_vel = _p(arr("joint_vel"))
    st.foot_force = _p(arr("foot_force"))
    st.vo_flag = arr("vo_flag", np.uint8).ctypes.data_as(up)
    st.vo_quat = _p(arr("vo_quat"))
    st.vo_time_pre = _p(arr("vo_time_pre"))
    st.vo_time_now = _p(arr("vo_time_now"))
    st.vo_rel_p = _p(arr("vo_rel_p"))
    if quat_in is not None:
        keep["quat_in"] = np.ascontiguousarray(quat_in, dtype=np.float64)
        st.quat_in = _p(keep["quat_in"])
    out = Outputs()
    res = {}
    shapes = {"quat": (S, 4, n), "x": (S, ds, n), "v_body": (S, 3, n), "p_vo": (S, 3, n)}
    for k, shp in shapes.items():
        if k in want:
            res[k] = np.full(shp, np.nan)
            setattr(out, k, _p(res[k]))
    if "contact" in want:
        res["contact"] = np.zeros((S, nl, n), dtype=np.uint8)
        out.contact = res["contact"].ctypes.data_as(up)
    if "vo_dbg" in want:
        res["vo_dbg"] = np.full((S, 10, n), -2, dtype=np.int32)
        out.vo_dbg = res["vo_dbg"].ctypes.data_as(ip)
    if "ekf_dbg" in want:
        res["ekf_dbg"] = np.full((S, 3, n), -2, dtype=np.int32)
        out.ekf_dbg = res["ekf_dbg"].ctypes.data_as(ip)
    if "arrival" in want:
        res["M_p"] = np.full((ds * ds, n), np.nan)
        res["n_p"] = np.full((ds, n), np.nan)
        out.M_p = _p(res["M_p"])
        out.n_p = _p(res["n_p"])
    if "admm_iters" in want:
        res["admm_iters"] = np.zeros((S, n), dtype=np.int32)
        out.admm_iters = res["admm_iters"].ctypes.data_as(ip)
    busy = C.c_double()
    bmax = C.c_double()
    wall = lib().orc_run_batch(C.byref(prm), C.byref(ep), C.byref(st), C.byref(out), i0, i1, nthreads,
                               int(run_ekf), int(run_mhe), t_steady, C.byref(busy), C.byref(bmax))
    res["_busy_max"] = bmax.value
    return res, wall, busy.value


class FrostRef:
    """oracle/_ref/libfrost_go1.so: the reference's own FROST kinematics (compiled from its sources)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libfrost_go1.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)

    def foot(self, leg, var22):
        v = np.ascontiguousarray(var22, dtype=np.float64)
        p = np.zeros(3)
        self.lib.ref_go1_foot(leg, _p(v), _p(p))
        return p

    def jac(self, leg, var22):
        v = np.ascontiguousarray(var22, dtype=np.float64)
        J = np.zeros(66)
        self.lib.ref_go1_jac(leg, _p(v), _p(J))
        return J.reshape(22, 3).T.copy()
