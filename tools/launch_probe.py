"""Is dekf_run launch-bound?  Host enqueue time of K ticks (call returns without sync) vs device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
n, N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 20, 200
S = 30 + K
dev = torch.device("cuda", 0)
stream = synth.make_stream(n, S, device=dev, device_rng=True)
vo = [bool(stream["vo_flag"][s].any()) for s in range(S)]
for ws in (1, 0):
    est = estimator.BatchedEstimator(estimator.robot_params("go1", ekf_rate=200, N=N, window_solve=ws), n)
    sub = {k: v for k, v in stream.items() if torch.is_tensor(v) and v.shape[0] == S}
    est.run(0, 30, sub, vo[:30])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    est.run(30, K, {k: v[30:] for k, v in sub.items()}, vo[30:])
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"window_solve={ws} n={n}: host enqueue {1e6*(t1-t0)/K:.1f} us/tick, device {1e3*e0.elapsed_time(e1)/K:.1f} us/tick, wall {1e6*(t2-t0)/K:.1f} us/tick")
    est.close()
