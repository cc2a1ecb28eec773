"""Foot-position-state model (leg_odom_type 1): one warp per instance (k_foot_team) against one thread per instance
(k_solve_foot, DEKF_FOOT_SERIAL=1) -- outputs and per-tick device time.  GPU box:  python tools/foot_probe.py [n] [ticks]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from decentralized_ekf_mhe_b200 import build, estimator as E, synth  # noqa: E402

build.build()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
S = int(sys.argv[2]) if len(sys.argv) > 2 else 60
st = {k: v.contiguous() for k, v in synth.make_stream(n, S, vo_jitter=True, device="cuda").items()}


def run(serial, est_type):
    os.environ["DEKF_FOOT_SERIAL"] = "1" if serial else "0"
    est = E.BatchedEstimator(E.robot_params("go1", ekf_rate=200, leg_odom_type=1, est_type=est_type), n)
    xs, sts = [], []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(S + 1)]
    ev[0].record()
    for s in range(S):
        est.step(s, E.robot_store.from_stream(st, s))
        ev[s + 1].record()
        xs.append(est.x_MHE_.clone())
        sts.append(est.status_.clone())
    torch.cuda.synchronize()
    ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(S)]
    est.close()
    return torch.stack(xs).cpu().numpy(), torch.stack(sts).cpu().numpy(), ms


run(False, 0) if S > 12 else None  # warm-up
for est_type in (0, 1):
    xa, sa, ma = run(False, est_type)
    xb, sb, mb = run(True, est_type)
    steady = slice(25, S)
    d = np.abs(xa[1:] - xb[1:])
    print(json.dumps(dict(est_type=est_type, instances=n, ticks=S, max_abs_dx_base=float(np.nanmax(d[:, :9])), max_abs_dx_all=float(np.nanmax(d)),
                          max_abs_dv=float(np.nanmax(d[:, 3:6])), status_equal=bool(np.array_equal(sa, sb)), nonfinite=int((sa & 32).any(axis=0).sum()),
                          team_ms_median=float(np.median(ma[steady])), serial_ms_median=float(np.median(mb[steady])),
                          team_instance_steps_per_s=n / (np.median(ma[steady]) * 1e-3), serial_instance_steps_per_s=n / (np.median(mb[steady]) * 1e-3))), flush=True)
