"""Small run of the foot-state kernels (team and serial, MHE and KF) for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
for serial in ("0", "1"):
    os.environ["DEKF_FOOT_SERIAL"] = serial
    for est_type in (0, 1):
        n, S = 70, 16
        st = {k: v.contiguous() for k, v in synth.make_stream(n, S, vo_jitter=True, device="cuda").items()}
        est = estimator.BatchedEstimator(estimator.robot_params("go1", ekf_rate=200, N=6, leg_odom_type=1, est_type=est_type), n)
        for s in range(S):
            est.step(s, estimator.robot_store.from_stream(st, s))
        torch.cuda.synchronize()
        assert torch.isfinite(est.x_MHE_).all()
        est.mhe_qp_.M_p
        est.close()
        print("ok foot", "serial" if serial == "1" else "team", "kf" if est_type else "mhe", flush=True)
