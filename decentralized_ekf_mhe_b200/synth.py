"""Synthetic sensor streams for the batched estimator.

The reference consumes live ROS topics (``/unitree/imu``, ``/unitree/joint_state``, ``orb/pos``,
``orb/vo``; go1Sub.cpp:13-21, EstSub.cpp:17-23, orien_ekf.cpp:35-40); ORB-SLAM3 is out of scope, so
visual odometry is replaced here by synthetic quaternion / relative-translation messages with the
semantics of stereo-pub-node.cpp:161-192 (pose stamped with the current frame time, relative
body translation stamped with previous + current frame times).

Everything is torch (float64) so the same code generates small CPU streams for parity tests and
large on-device streams for the benchmark.  Layout of every array: ``[step][field][instance]``
(instance fastest) -- exactly what the C ABI (include/dekf_b200.h) takes per step.
"""
import math

import torch

# Serial-chain descriptions of the builder-defined models; must match oracle/kin.c and
# csrc/kinematics.cuh.  Go1 is the 3R chain the reference's FROST expressions collapse to when the
# floating-base coordinates are zero (see oracle/kin.c header).
ROBOTS = {
    "go1": dict(
        id=0, num_legs=4, nj=3,
        legs=[dict(types=[0, 0, 0], axes=[(1, 0, 0), (0, 1, 0), (0, 1, 0)],
                   offs=[(sx * 0.1881, sy * 0.04675, 0.0), (0.0, sy * 0.08, 0.0), (0.0, 0.0, -0.213)],
                   tool=(0.0, 0.0, -0.213))
              for sx, sy in ((1, -1), (1, 1), (-1, -1), (-1, 1))],
        nominal=[(0.0, 0.8, -1.5)] * 4, swing_dir=(0.2, 1.0, -0.5),
        gait_period=0.4, duty=0.5, leg_phase=(0.0, 0.5, 0.5, 0.0),
        force_stance=200.0, force_swing=20.0, p_ib=(0.01592, 0.06659, 0.00617)),
    "cassie": dict(
        id=1, num_legs=2, nj=5,
        legs=[dict(types=[0] * 5, axes=[(1, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 0), (0, 1, 0)],
                   offs=[(0.021, sy * 0.135, 0.0), (0.0, 0.0, -0.07), (0.0, 0.0, -0.09),
                         (0.12, 0.0, -0.4896), (0.06, 0.0, -0.5)],
                   tool=(0.02, 0.0, -0.05))
              for sy in (1, -1)],
        nominal=[(0.0, 0.0, 0.5, -1.2, 0.7)] * 2, swing_dir=(0.05, 0.05, 0.6, -0.8, 0.3),
        gait_period=0.7, duty=0.6, leg_phase=(0.0, 0.5),
        force_stance=330.0, force_swing=15.0, p_ib=(0.0, 0.0, 0.0)),
    "pogox": dict(
        id=2, num_legs=1, nj=3,
        legs=[dict(types=[0, 0, 1], axes=[(1, 0, 0), (0, 1, 0), (0, 0, -1)],
                   offs=[(0.0, 0.0, -0.05), (0.0, 0.0, 0.0), (0.0, 0.0, -0.25)],
                   tool=(0.0, 0.0, -0.05))],
        nominal=[(0.0, 0.0, 0.10)], swing_dir=(0.3, 0.3, -0.08),
        gait_period=0.3, duty=0.3, leg_phase=(0.0,),
        force_stance=250.0, force_swing=5.0, p_ib=(0.0, 0.0, 0.0)),
}


def _rot_axis(axis, th):
    """Rodrigues rotation about a constant unit axis; th: [n] -> [n,3,3]."""
    a = torch.tensor(axis, dtype=th.dtype, device=th.device)
    K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]],
                     dtype=th.dtype, device=th.device)
    c = torch.cos(th)[:, None, None]
    s = torch.sin(th)[:, None, None]
    eye = torch.eye(3, dtype=th.dtype, device=th.device)
    return c * eye + s * K + (1 - c) * torch.outer(a, a)


def chain_fk(leg, q):
    """Foot position [n,3] and Jacobian [n,3,nj] of one serial-chain leg for joint values q [n,nj]."""
    n, nj = q.shape
    dt_, dev = q.dtype, q.device
    Rw = torch.eye(3, dtype=dt_, device=dev).expand(n, 3, 3).clone()
    o = torch.zeros(n, 3, dtype=dt_, device=dev)
    origins, axes = [], []
    for j in range(nj):
        off = torch.tensor(leg["offs"][j], dtype=dt_, device=dev)
        ax = torch.tensor(leg["axes"][j], dtype=dt_, device=dev)
        o = o + Rw @ off
        aw = Rw @ ax
        origins.append(o.clone())
        axes.append(aw)
        if leg["types"][j] == 0:
            Rw = Rw @ _rot_axis(leg["axes"][j], q[:, j])
        else:
            o = o + aw * q[:, j:j + 1]
    tool = torch.tensor(leg["tool"], dtype=dt_, device=dev)
    p = o + Rw @ tool
    cols = []
    for j in range(nj):
        if leg["types"][j] == 0:
            cols.append(torch.linalg.cross(axes[j], p - origins[j]))
        else:
            cols.append(axes[j])
    return p, torch.stack(cols, dim=2)


def euler_to_rot(r, p, y):
    cr, sr, cp, sp, cy, sy = torch.cos(r), torch.sin(r), torch.cos(p), torch.sin(p), torch.cos(y), torch.sin(y)
    R = torch.stack([
        torch.stack([cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr], -1),
        torch.stack([sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr], -1),
        torch.stack([-sp, cp * sr, cp * cr], -1)], -2)
    return R


def euler_to_quat(r, p, y):
    cr, sr, cp, sp, cy, sy = (torch.cos(r / 2), torch.sin(r / 2), torch.cos(p / 2), torch.sin(p / 2),
                              torch.cos(y / 2), torch.sin(y / 2))
    return torch.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                        cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], 0)  # [4,n] w,x,y,z


class Trajectory:
    """Per-instance analytic base motion (SURVEY.md 8d "Go1 trot")."""

    def __init__(self, n, seed, device, amp_jitter=False, speed=0.5):
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        u = torch.rand(10, n, generator=g, dtype=torch.float64)
        self.phi_v = (2 * math.pi * u[0:3]).to(device)
        self.phi_rpy = ((u[3:6] - 0.5) * 0.04).to(device)
        self.phi_gait = u[6].to(device)
        self.phi_vo = u[7].to(device)
        self.amp = (0.5 + u[8] if amp_jitter else torch.ones(n, dtype=torch.float64)).to(device)
        self.speed = speed
        self.n = n
        self.device = device

    def _t(self, t):
        if not torch.is_tensor(t):
            t = torch.full((self.n,), float(t), dtype=torch.float64, device=self.device)
        return t

    def vel(self, t):
        t = self._t(t)
        a = self.amp
        return torch.stack([self.speed + a * 0.1 * torch.sin(math.pi * t + self.phi_v[0]),
                            a * 0.05 * torch.sin(2 * math.pi * t + self.phi_v[1]),
                            a * 0.02 * torch.sin(8 * math.pi * t + self.phi_v[2])], 0)

    def acc(self, t):
        t = self._t(t)
        a = self.amp
        return torch.stack([a * 0.1 * math.pi * torch.cos(math.pi * t + self.phi_v[0]),
                            a * 0.05 * 2 * math.pi * torch.cos(2 * math.pi * t + self.phi_v[1]),
                            a * 0.02 * 8 * math.pi * torch.cos(8 * math.pi * t + self.phi_v[2])], 0)

    def pos(self, t):
        t = self._t(t)
        a = self.amp
        return torch.stack([
            self.speed * t - a * 0.1 / math.pi * (torch.cos(math.pi * t + self.phi_v[0]) - torch.cos(self.phi_v[0])),
            -a * 0.05 / (2 * math.pi) * (torch.cos(2 * math.pi * t + self.phi_v[1]) - torch.cos(self.phi_v[1])),
            -a * 0.02 / (8 * math.pi) * (torch.cos(8 * math.pi * t + self.phi_v[2]) - torch.cos(self.phi_v[2]))], 0)

    def rpy(self, t):
        t = self._t(t)
        return (0.03 * torch.sin(4 * math.pi * t) + self.phi_rpy[0],
                0.03 * torch.cos(4 * math.pi * t) + self.phi_rpy[1],
                0.2 * t + self.phi_rpy[2])

    def rpy_rate(self, t):
        t = self._t(t)
        return (0.03 * 4 * math.pi * torch.cos(4 * math.pi * t),
                -0.03 * 4 * math.pi * torch.sin(4 * math.pi * t),
                torch.full_like(t, 0.2))


def make_stream(n, S, *, robot="go1", seed=20240510, dt=0.005, device="cpu", s0=0, vo=True,
                vo_rate=30.0, vo_latency=0.040, vo_jitter=False, amp_jitter=False, truth=False,
                device_rng=False):
    """Generate steps ``s0 .. s0+S-1`` of the synthetic stream for ``n`` instances.

    vo_jitter=False: every instance sees VO frames at t_j = j / vo_rate (lock-step arrival, the
    benchmark stream).  vo_jitter=True: per-instance capture phase -> ragged arrival ticks.
    """
    spec = ROBOTS[robot]
    nl, nj = spec["num_legs"], spec["nj"]
    f64 = torch.float64
    tr = Trajectory(n, seed, device, amp_jitter=amp_jitter)
    out = {
        "gyro": torch.empty(S, 3, n, dtype=f64, device=device),
        "accel": torch.empty(S, 3, n, dtype=f64, device=device),
        "imu_time": torch.empty(S, n, dtype=f64, device=device),
        "joint_pos": torch.empty(S, nl * nj, n, dtype=f64, device=device),
        "joint_vel": torch.empty(S, nl * nj, n, dtype=f64, device=device),
        "foot_force": torch.empty(S, nl, n, dtype=f64, device=device),
        "vo_flag": torch.zeros(S, n, dtype=torch.uint8, device=device),
        "vo_quat": torch.zeros(S, 4, n, dtype=f64, device=device),
        "vo_time_pre": torch.zeros(S, n, dtype=f64, device=device),
        "vo_time_now": torch.zeros(S, n, dtype=f64, device=device),
        "vo_rel_p": torch.zeros(S, 3, n, dtype=f64, device=device),
    }
    if truth:
        out["quat_true"] = torch.empty(S, 4, n, dtype=f64, device=device)
        out["v_true"] = torch.empty(S, 3, n, dtype=f64, device=device)
        out["p_true"] = torch.empty(S, 3, n, dtype=f64, device=device)
    bias = torch.tensor([0.05, -0.03, 0.02], dtype=f64, device=device)[:, None]
    acc_std = torch.tensor([0.025, 0.025, 0.02], dtype=f64, device=device)[:, None]
    grav = torch.tensor([0.0, 0.0, 9.81], dtype=f64, device=device)[:, None]
    p_ib = torch.tensor(spec["p_ib"], dtype=f64, device=device)
    swing_dir = torch.tensor(spec["swing_dir"], dtype=f64, device=device)
    vo_phase = tr.phi_vo if vo_jitter else torch.zeros(n, dtype=f64, device=device)
    for si in range(S):
        k = s0 + si
        t = k * dt
        if device_rng:  # large benchmark streams: draw on the device (not reproducible across devices)
            g = torch.Generator(device=device).manual_seed(int(seed) * 1000003 + k)
            noise = torch.randn(16 + 2 * nl * nj + nl, n, generator=g, dtype=f64, device=device)
        else:
            g = torch.Generator(device="cpu").manual_seed(int(seed) * 1000003 + k)
            noise = torch.randn(16 + 2 * nl * nj + nl, n, generator=g, dtype=f64).to(device)
        r, p, y = tr.rpy(t)
        dr, dp_, dy = tr.rpy_rate(t)
        R = euler_to_rot(r, p, y)  # [n,3,3]
        w = torch.stack([dr - dy * torch.sin(p),
                         dp_ * torch.cos(r) + dy * torch.sin(r) * torch.cos(p),
                         -dp_ * torch.sin(r) + dy * torch.cos(r) * torch.cos(p)], 0)  # [3,n]
        v_s = tr.vel(t)
        a_s = tr.acc(t)
        a_b = torch.einsum("nji,jn->in", R, a_s + grav) + bias + acc_std * noise[0:3]
        out["gyro"][si] = w
        out["accel"][si] = a_b
        out["imu_time"][si] = k * dt
        v_b = torch.einsum("nji,jn->ni", R, v_s)  # R' v_s : [n,3]
        # legs
        for leg in range(nl):
            psi = torch.remainder(t / spec["gait_period"] + tr.phi_gait + spec["leg_phase"][leg], 1.0)
            stance = psi < spec["duty"]
            sw = torch.clamp((psi - spec["duty"]) / (1.0 - spec["duty"]), 0.0, 1.0)
            nominal = torch.tensor(spec["nominal"][leg], dtype=f64, device=device)
            q = nominal[None, :] + (0.2 * torch.sin(math.pi * sw) * (~stance))[:, None] * swing_dir[None, :]
            pf, J = chain_fk(spec["legs"][leg], q)
            pf = pf + p_ib
            rhs = -(v_b + torch.linalg.cross(w.T, pf))  # J dq = -(R'v + w x p)
            if nj == 3:
                dq_st = torch.linalg.solve(J, rhs[:, :, None])[:, :, 0]
            else:
                dq_st = (torch.linalg.pinv(J) @ rhs[:, :, None])[:, :, 0]
            nz = noise[16 + leg * nj:16 + (leg + 1) * nj].T
            dq = torch.where(stance[:, None], dq_st + 0.22 * nz, 1.0 * nz)
            out["joint_pos"][si, leg * nj:(leg + 1) * nj] = q.T
            out["joint_vel"][si, leg * nj:(leg + 1) * nj] = dq.T
            fz = torch.where(stance, torch.full_like(psi, spec["force_stance"]),
                             torch.full_like(psi, spec["force_swing"]))
            out["foot_force"][si, leg] = fz + 5.0 * noise[16 + 2 * nl * nj + leg]
        if truth:
            out["quat_true"][si] = euler_to_quat(r, p, y)
            out["v_true"][si] = v_s
            out["p_true"][si] = tr.pos(t)
        if vo:
            # frame j captured at t_j = (j + phase)/rate, message visible at the first tick with
            # k*dt >= t_j + latency
            jf = (t - vo_latency) * vo_rate - vo_phase + 1e-9
            j_now = torch.floor(jf)
            jf_prev = ((k - 1) * dt - vo_latency) * vo_rate - vo_phase + 1e-9
            j_prev = torch.floor(jf_prev)
            new = (j_now > j_prev) & (j_now >= 1)
            if bool(new.any()):
                t_now = (j_now + vo_phase) / vo_rate
                t_pre = (j_now - 1 + vo_phase) / vo_rate
                rn, pn, yn = tr.rpy(t_now)
                rp, pp, yp = tr.rpy(t_pre)
                R_pre = euler_to_rot(rp, pp, yp)
                dpw = tr.pos(t_now) - tr.pos(t_pre)
                rel = torch.einsum("nji,jn->in", R_pre, dpw) + 1.5e-5 * noise[3:6]
                qv = euler_to_quat(rn, pn, yn) + 1e-4 * noise[6:10]
                qv = qv / torch.linalg.norm(qv, dim=0, keepdim=True)
                m = new
                out["vo_flag"][si] = m.to(torch.uint8)
                out["vo_quat"][si] = torch.where(m[None, :], qv, torch.zeros_like(qv))
                out["vo_time_pre"][si] = torch.where(m, t_pre, torch.zeros_like(t_pre))
                out["vo_time_now"][si] = torch.where(m, t_now, torch.zeros_like(t_now))
                out["vo_rel_p"][si] = torch.where(m[None, :], rel, torch.zeros_like(rel))
    return out


def to_numpy(stream):
    return {k: v.detach().cpu().numpy() for k, v in stream.items()}
