"""PogoX state-constrained solve (BASELINE config 4, 16,384 instances): team kernel (9 lanes per instance) against the
one-thread-per-instance kernel (DEKF_BOX_SERIAL=1) -- outputs, active-set statistics and per-tick device time.
GPU box:  python tools/box_probe.py [n] [ticks]"""
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
lo, hi = (-0.45, -0.03, -0.015), (0.55, 0.03, 0.015)
st = {k: v.contiguous() for k, v in synth.make_stream(n, S, robot="pogox", vo_jitter=True, device="cuda").items()}


def run(serial, precision="fp64"):
    os.environ["DEKF_BOX_SERIAL"] = "1" if serial else "0"
    prm = E.robot_params("pogox", ekf_rate=200, v_box_enable=1, v_box_lo=lo, v_box_hi=hi)
    est = E.BatchedEstimator(prm, n, precision=precision)
    xs, its, nas, sts = [], [], [], []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(S + 1)]
    ev[0].record()
    for s in range(S):
        est.step(s, E.robot_store.from_stream(st, s))
        ev[s + 1].record()
        xs.append(est.x_MHE_.clone())
        it, na = est.qp_info()
        its.append(it.clone())
        nas.append(na.clone())
        sts.append(est.status_.clone())
    torch.cuda.synchronize()
    ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(S)]
    est.close()
    return torch.stack(xs).cpu().numpy(), torch.stack(its).cpu().numpy(), torch.stack(nas).cpu().numpy(), torch.stack(sts).cpu().numpy(), ms


_S = S
S = 12
run(False, "fp64")  # warm-up: module load, allocator, clocks
S = _S
for precision in ("fp64", "fp32"):
    xa, ia, na_, sa, ma = run(False, precision)
    xb, ib, nb, sb, mb = run(True, precision)
    steady = slice(25, S)
    print(json.dumps(dict(
        precision=precision, instances=n, ticks=S,
        max_abs_dx=float(np.nanmax(np.abs(xa[1:] - xb[1:]))), iters_equal=bool(np.array_equal(ia[1:], ib[1:])),
        nactive_equal=bool(np.array_equal(na_[1:], nb[1:])), status_equal=bool(np.array_equal(sa, sb)),
        iters_mean=float(ia[steady].mean()), nactive_mean=float(na_[steady].mean()), maxiter_flags=int((sa & 64).any(axis=0).sum()),
        nonfinite_flags=int((sa & 32).any(axis=0).sum()),
        team_ms_per_tick=float(np.mean(ma[steady])), serial_ms_per_tick=float(np.mean(mb[steady])),
        team_ms_median=float(np.median(ma[steady])), team_ms_p10=float(np.percentile(ma[steady], 10)), team_ms_p90=float(np.percentile(ma[steady], 90)),
        team_instance_steps_per_s=n / (np.mean(ma[steady]) * 1e-3), serial_instance_steps_per_s=n / (np.mean(mb[steady]) * 1e-3))), flush=True)
