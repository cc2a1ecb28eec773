import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from decentralized_ekf_mhe_b200 import estimator as E, synth
n, S = 512, 60
st_t = synth.make_stream(n, S, robot="pogox", vo_jitter=True, device="cuda")
mask = (1 << 3) | (1 << 2) | (7 << 6)
lo9 = (0, 0, -2e-4, 0.47, 0, 0, -0.004, -0.004, -0.004)
hi9 = (0, 0, 2e-4, 0.52, 0, 0, 0.004, 0.004, 0.004)
for mi in (50, 200):
    est = E.BatchedEstimator(E.robot_params("pogox", ekf_rate=200, x_box_mask=mask, x_box_lo=lo9, x_box_hi=hi9, v_box_max_iter=mi), n)
    d = {k: v.contiguous() for k, v in st_t.items()}
    for s in range(S):
        est.step(s, E.robot_store.from_stream(d, s))
        if s in (1, 2, 5, 10, 30, 59):
            it, na = est.qp_info()
            stt = est.status_
            print(mi, s, 'iters mean', float(it.double().mean()), 'max', int(it.max()), 'nact mean', float(na.double().mean()), 'maxiter flags', int((stt & 64).ne(0).sum()), 'nonfinite', int((stt & 32).ne(0).sum()))
    est.close()
