"""End-to-end probe of dekf_run_host_f32io (float32 sensor streams in, float32 results out, pinned host buffers), the bench's
headline `e2e` contract, for tuning the chunk pipeline (DEKF_HOST_CHUNK, DEKF_HOST_RAMP).  usage: e2e_f32_probe.py [K=100]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n, N, FILL, Kw = 65536, 20, 34, 8
S = FILL + Kw + 3 * (K + Kw)
dev = torch.device("cuda", 0)
stream = synth.make_stream(n, S, device=dev, device_rng=True)
vo = [bool(stream["vo_flag"][s].any()) for s in range(S)]
keys = ["gyro", "accel", "imu_time", "joint_pos", "joint_vel", "foot_force", "vo_quat", "vo_time_pre", "vo_time_now", "vo_rel_p"]
F32 = estimator.BatchedEstimator.F32_KEYS
rows = {k: (stream[k][0].numel() // n) for k in keys}
est = estimator.BatchedEstimator(estimator.robot_params("go1", ekf_rate=200, N=N), n)
est.run(0, FILL, {k: v[:FILL] for k, v in stream.items() if torch.is_tensor(v) and v.shape[0] == S}, vo[:FILL])


def host_slice(a, b):
    h = {k: stream[k][a:b].reshape(b - a, rows[k], n).cpu() for k in keys}
    for k in F32:
        h[k] = h[k].float()
    h = {k: v.pin_memory() for k, v in h.items()}
    h["vo_flag"] = stream["vo_flag"][a:b].cpu().pin_memory()
    return h


def host_out(k):
    return {"quat": torch.empty(k, 4, n, dtype=torch.float32).pin_memory(), "x": torch.empty(k, 9, n, dtype=torch.float32).pin_memory(),
            "v_body": torch.empty(k, 3, n, dtype=torch.float32).pin_memory(), "contact": torch.empty(k, 4, n, dtype=torch.uint8).pin_memory(),
            "status": torch.empty(k, n, dtype=torch.int32).pin_memory()}


T = FILL
est.run_host_f32(T, Kw, host_slice(T, T + Kw), vo[T:T + Kw], out=host_out(Kw), out_per_step=True)
T += Kw
res = []
WARM = os.environ.get("PROBE_WARM", "0") == "1"
for rep in range(3):
    hst, hout = host_slice(T + (Kw if WARM else 0), T + (Kw if WARM else 0) + K), host_out(K)
    if WARM:  # a short untimed call right before the timed one (the GPU / the link idled while the host buffers were filled)
        est.run_host_f32(T, Kw, host_slice(T, T + Kw), vo[T:T + Kw], out=host_out(Kw), out_per_step=True)
        T += Kw
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    est.run_host_f32(T, K, hst, vo[T:T + K], out=hout, out_per_step=True)
    chk = float(hout["x"][:, 3, 0].double().sum())
    dt = time.perf_counter() - t0
    res.append(dt)
    T += K
knobs = {k: v for k, v in os.environ.items() if k.startswith("DEKF_")}
print(f"e2e_f32io K={K} knobs={knobs}: " + "  ".join(f"{1e6*d/K:.1f} us/tick ({n*K/d:.3e})" for d in res) + f"  chk {chk:.6f}")
