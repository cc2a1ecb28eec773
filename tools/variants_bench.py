"""Throughput of the estimator variants beside the headline path (device-resident streams, dekf_run, CUDA events):
KF alternative, leg_odom_type 1, state-constrained PogoX, Cassie, N = 100.  One JSON line per variant."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
dev = torch.device("cuda", 0)
CASES = [
    ("go1_full_resweep_fp64", "go1", 65536, 20, "fp64", dict(window_solve=0)),
    ("go1_incremental_fp64", "go1", 65536, 20, "fp64", dict(window_solve=1)),
    ("go1_full_resweep_fp32", "go1", 65536, 20, "fp32", dict(window_solve=0)),
    ("go1_incremental_fp32", "go1", 65536, 20, "fp32", dict(window_solve=1)),
    ("go1_kf_alternative_fp64", "go1", 65536, 20, "fp64", dict(est_type=1)),
    ("go1_foot_states_fp64 (leg_odom_type 1)", "go1", 16384, 20, "fp64", dict(leg_odom_type=1)),
    ("cassie_full_resweep_fp64", "cassie", 65536, 20, "fp64", dict(window_solve=0)),
    ("cassie_full_resweep_fp32", "cassie", 65536, 20, "fp32", dict(window_solve=0)),
    ("pogox_box_constrained_fp64", "pogox", 16384, 20, "fp64", dict(v_box_enable=1, v_box_lo=(-0.45, -0.03, -0.015), v_box_hi=(0.55, 0.03, 0.015))),
    ("go1_N100_full_resweep_fp64", "go1", 65536, 100, "fp64", dict(window_solve=0)),
    ("go1_N100_incremental_fp64", "go1", 65536, 100, "fp64", dict(window_solve=1)),
]
only = sys.argv[1:]
for name, robot, n, N, prec, over in CASES:
    if only and not any(o in name for o in only):
        continue
    fill = N + 6
    K = 40 if N == 20 else 20
    if "foot" in name or "box" in name:
        K = 12
    S = fill + K
    st = synth.make_stream(n, S, robot=robot, device=dev, device_rng=True)
    vo = [bool(st["vo_flag"][s].any()) for s in range(S)]
    sub = {k: v for k, v in st.items() if torch.is_tensor(v) and v.shape[0] == S}
    est = estimator.BatchedEstimator(estimator.robot_params(robot, ekf_rate=200, N=N, **over), n, precision=prec)
    est.run(0, fill, sub, vo[:fill])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    est.run(fill, K, {k: v[fill:] for k, v in sub.items()}, vo[fill:])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    extra = {}
    if "box" in name:
        it, na = est.qp_info()
        extra = {"factorisations_mean": float(it.double().mean()), "active_bounds_mean": float(na.double().mean())}
    print(json.dumps({"variant": name, "instances": n, "N": N, "precision": prec, "ms_per_tick": ms,
                      "instance_steps_per_s": n / (ms * 1e-3), "device_bytes": est.device_bytes(), **extra}), flush=True)
    est.close()
    del st, sub
    torch.cuda.empty_cache()
