"""Small run of every kernel path for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
dev = torch.device("cuda", 0)
def run(robot, n, S, prec="fp64", **over):
    st = synth.make_stream(n, S, robot=robot, vo_jitter=True, device=dev)
    vo = [bool(st["vo_flag"][s].any()) for s in range(S)]
    sub = {k: v.contiguous() for k, v in st.items() if torch.is_tensor(v) and v.shape[0] == S}
    est = estimator.BatchedEstimator(estimator.robot_params(robot, ekf_rate=200, N=8, **over), n, precision=prec)
    est.run(0, S - 6, sub, vo)
    for s in range(S - 6, S):
        est.step(s, estimator.robot_store.from_stream(sub, s))
    torch.cuda.synchronize()
    assert torch.isfinite(est.x_MHE_).all()
    est.close()
    print("ok", robot, n, prec, over, flush=True)
os.environ["DEKF_FUSED_MAX_N"] = "0"   # split kernels + dekf_run pipeline even at this size
run("go1", 300, 40, window_solve=0)
run("go1", 300, 40, window_solve=1)
run("go1", 300, 40, "fp32", window_solve=1)
run("go1", 200, 30, est_type=1)
run("go1", 160, 30, leg_odom_type=1)
run("pogox", 200, 40, v_box_enable=1, v_box_lo=(-0.45, -0.03, -0.015), v_box_hi=(0.55, 0.03, 0.015))
os.environ["DEKF_BOX_SERIAL"] = "1"    # the one-thread-per-instance form of the constrained solve
run("pogox", 100, 30, v_box_enable=1, v_box_lo=(-0.45, -0.03, -0.015), v_box_hi=(0.55, 0.03, 0.015))
os.environ["DEKF_BOX_SERIAL"] = "0"
run("cassie", 200, 30, window_solve=0)
# host paths: chunked pipeline with double and with float32 sensor streams
def run_host(n, S, f32):
    st = synth.make_stream(n, S, vo_jitter=True)
    vo = [bool(st["vo_flag"][s].any()) for s in range(S)]
    keys = estimator.BatchedEstimator._IN_KEYS
    host = {k: (st[k].float() if (f32 and k in estimator.BatchedEstimator.F32_KEYS) else st[k]).contiguous().pin_memory() for k in keys}
    out = {"quat": torch.empty(S, 4, n, dtype=torch.float64).pin_memory(), "x": torch.empty(S, 9, n, dtype=torch.float64).pin_memory(),
           "v_body": torch.empty(S, 3, n, dtype=torch.float64).pin_memory()}
    est = estimator.BatchedEstimator(estimator.robot_params("go1", ekf_rate=200, N=8), n)
    (est.run_host_f32 if f32 else est.run_host)(0, S, host, vo, out=out, out_per_step=True)
    assert torch.isfinite(out["x"][1:]).all()
    est.close()
    print("ok host path", n, "f32 sensor streams" if f32 else "double streams", flush=True)
os.environ["DEKF_HOST_CHUNK"] = "5"
run_host(300, 33, False)
run_host(300, 33, True)
os.environ["DEKF_FUSED_MAX_N"] = "4096"
run("go1", 100, 30, window_solve=1)
