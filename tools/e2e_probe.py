"""Probe of the host-buffer path (dekf_run_host): per-call wall time for several chunk sizes, cold and warm."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
n, N, Ke = 65536, 20, 48
S = 24 + 6 * Ke
dev = torch.device("cuda", 0)
stream = synth.make_stream(n, S, device=dev, device_rng=True)
vo = [bool(stream["vo_flag"][s].any()) for s in range(S)]
keys = ["gyro", "accel", "imu_time", "joint_pos", "joint_vel", "foot_force", "vo_quat", "vo_time_pre", "vo_time_now", "vo_rel_p"]
rows = {k: (stream[k][0].numel() // n) for k in keys}
for B in [int(x) for x in (sys.argv[1:] or ["8"])]:
    os.environ["DEKF_HOST_CHUNK"] = str(B)
    est = estimator.BatchedEstimator(estimator.robot_params("go1", ekf_rate=200, N=N), n)
    dsub = {k: v[:24] for k, v in stream.items() if torch.is_tensor(v) and v.shape[0] == S}
    est.run(0, 24, dsub, vo[:24])
    torch.cuda.synchronize()
    T = 24
    hout = {"quat": torch.empty(Ke, 4, n, dtype=torch.float64).pin_memory(), "x": torch.empty(Ke, 9, n, dtype=torch.float64).pin_memory(),
            "v_body": torch.empty(Ke, 3, n, dtype=torch.float64).pin_memory(), "contact": torch.empty(Ke, 4, n, dtype=torch.uint8).pin_memory(),
            "status": torch.empty(Ke, n, dtype=torch.int32).pin_memory()}
    for rep in range(5):
        hst = {k: stream[k][T:T + Ke].reshape(Ke, rows[k], n).cpu().pin_memory() for k in keys}
        hst["vo_flag"] = stream["vo_flag"][T:T + Ke].cpu().pin_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        est.run_host(T, Ke, hst, vo[T:T + Ke], out=hout, out_per_step=True)
        dt = time.perf_counter() - t0
        print(f"chunk {B} rep {rep}: {dt*1e3/Ke:.3f} ms/tick  {n*Ke/dt:.3e} inst-steps/s", flush=True)
        T += Ke
    est.close()
