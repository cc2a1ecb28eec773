"""ms per tick of dekf_run at the benchmark size (65,536 Go1 instances, N=20, per-tick outputs), for tuning the tick pipeline.
Environment knobs (DEKF_PRIO, DEKF_NO_SPLIT, ...) are read by dekf_create.  usage: run_probe.py [window_solve] [K] [n] [ragged]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
ws = int(sys.argv[1]) if len(sys.argv) > 1 else 0
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
n = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
ragged = len(sys.argv) > 4 and sys.argv[4] == "ragged"  # per-instance camera phase: every tick carries VO messages
N, FILL = 20, 34
S = FILL + K
dev = torch.device("cuda", 0)
st = synth.make_stream(n, S, device=dev, device_rng=True, vo_jitter=ragged)
vo = [bool(st["vo_flag"][s].any()) for s in range(S)]
cut = {k: v for k, v in st.items() if torch.is_tensor(v) and v.shape[0] == S}
best = None
for rep in range(3):
    est = estimator.BatchedEstimator(estimator.robot_params("go1", ekf_rate=200, N=N, window_solve=ws), n)
    est.run(0, FILL, {k: v[:FILL] for k, v in cut.items()}, vo[:FILL])
    outs = {"quat": torch.empty(K, 4, n, dtype=torch.float64, device=dev), "x": torch.empty(K, 9, n, dtype=torch.float64, device=dev),
            "v_body": torch.empty(K, 3, n, dtype=torch.float64, device=dev), "contact": torch.empty(K, 4, n, dtype=torch.uint8, device=dev),
            "status": torch.empty(K, n, dtype=torch.int32, device=dev)}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    import time
    e0.record()
    th = time.perf_counter()
    est.run(FILL, K, {k: v[FILL:] for k, v in cut.items()}, vo[FILL:], out=outs, out_per_step=True)
    th = time.perf_counter() - th  # host time to QUEUE the K ticks (the call returns before the GPU is done)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    best = ms if best is None else min(best, ms)
    chk = float(outs["x"][:, 3].sum().item())
    est.close()
    del outs
knobs = {k: v for k, v in os.environ.items() if k.startswith("DEKF_")}
print(f"host enqueue {1e6 * th / K:.1f} us/tick; ", end="")
print(f"run_probe ws={ws} n={n} K={K} {'ragged ' if ragged else ''}knobs={knobs}: {best*1e3:.1f} us/tick  {n/best/1e3:.4g} instance-steps/s  checksum {chk:.6f}")
