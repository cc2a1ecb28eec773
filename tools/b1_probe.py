"""Batch-1 tick loop (k_fused, one launch per tick) for ncu captures and quick latency numbers.
    python tools/b1_probe.py [window_solve=0] [ticks=120] [n=1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
ws = int(sys.argv[1]) if len(sys.argv) > 1 else 0
K = int(sys.argv[2]) if len(sys.argv) > 2 else 120
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
N = 20
S = 40 + K
dev = torch.device("cuda", 0)
st = synth.make_stream(n, S, seed=777, device=dev, device_rng=True)
vo = [bool(st["vo_flag"][s].any()) for s in range(S)]
est = estimator.BatchedEstimator(estimator.robot_params("go1", ekf_rate=200, N=N, window_solve=ws), n)
sub = {k: v for k, v in st.items() if torch.is_tensor(v) and v.shape[0] == S}
est.run(0, 40, sub, vo[:40])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
est.run(40, K, {k: v[40:] for k, v in sub.items()}, vo[40:])
e1.record()
torch.cuda.synchronize()
print(f"batch-{n} window_solve={ws} roles_max={os.environ.get('DEKF_ROLES_MAX_N', 'default')}: {1e3 * e0.elapsed_time(e1) / K:.1f} us device time per tick over {K} ticks")
