"""Prints, per golden case, the largest difference between the CUDA path (through the C ABI) and the outputs of the
reference's own compiled sources (tests/golden/go1_refnodes_golden.npz).  GPU box:  python tools/ref_parity_report.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from decentralized_ekf_mhe_b200 import build, estimator as E  # noqa: E402

build.build()
g = np.load(os.path.join(ROOT, "tests", "golden", "go1_refnodes_golden.npz"))
names = sorted({k.split("/")[0] for k in g.files})
rows = []
for name in names:
    st = {k.split("/in_")[1]: g[k] for k in g.files if k.startswith(name + "/in_") and not k.endswith("_ns")}
    ref = {k.split("/out_")[1]: g[k] for k in g.files if k.startswith(name + "/out_")}
    N, est_type, leg_odom_type, rate = (int(v) for v in g[name + "/params"])
    S, _, n = st["gyro"].shape
    for precision in ("fp64", "fp32"):
        if precision == "fp32" and leg_odom_type == 1:
            continue
        for ws in (0, 1):
            if ws == 1 and (est_type == 1 or leg_odom_type == 1):
                continue
            est = E.BatchedEstimator(E.robot_params("go1", ekf_rate=rate, N=N, est_type=est_type, leg_odom_type=leg_odom_type,
                                                    window_solve=ws), n, precision=precision)
            d = {k: torch.as_tensor(v).cuda().contiguous() for k, v in st.items()}
            dq = dx = dv = dp = 0.0
            contact_ok = True
            for s in range(S):
                est.step(s, E.robot_store.from_stream(d, s))
                dq = max(dq, float(np.abs(est.quaternion_.cpu().numpy() - ref["quat"][s]).max()))
                dp = max(dp, float(np.abs(est.p_vo_accmulate_.cpu().numpy() - ref["p_vo"][s]).max()))
                contact_ok &= bool(np.array_equal(est.contact_.cpu().numpy(), ref["contact"][s]))
                if s >= 1:
                    x = est.x_MHE_.cpu().numpy()
                    dx = max(dx, float(np.abs(x - ref["x"][s]).max()))
                    dv = max(dv, float(np.abs(x[3:6] - ref["x"][s, 3:6]).max()))
            est.close()
            rows.append(dict(case=name, N=N, est_type=est_type, leg_odom_type=leg_odom_type, instances=n, ticks=S,
                             precision=precision, window_solve="incremental" if ws else "full", max_dq=dq, max_dx=dx,
                             max_dv_mps=dv, max_dp_vo=dp, contact_sets_exact=contact_ok))
            print(json.dumps(rows[-1]), flush=True)
