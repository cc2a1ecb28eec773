"""Per-tick kernel times (event pairs around every launch, one dekf_step per tick) split by VO / non-VO ticks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
ws = int(sys.argv[1]) if len(sys.argv) > 1 else 1
prec = sys.argv[2] if len(sys.argv) > 2 else "fp64"
ragged = len(sys.argv) > 3 and sys.argv[3] == "ragged"  # per-instance camera phase: every tick carries VO messages
n, N, K = 65536, 20, 80
S = 30 + K
dev = torch.device("cuda", 0)
stream = synth.make_stream(n, S, device=dev, device_rng=True, vo_jitter=ragged)
vo = [bool(stream["vo_flag"][s].any()) for s in range(S)]
est = estimator.BatchedEstimator(estimator.robot_params("go1", ekf_rate=200, N=N, window_solve=ws), n, precision=prec)
sub = {k: v for k, v in stream.items() if torch.is_tensor(v) and v.shape[0] == S}
est.run(0, 30, sub, vo[:30])
torch.cuda.synchronize()
acc = {v: {"ekf": 0, "assemble": 0, "solve": 0, "resweep": 0, "n": 0} for v in (True, False)}
est.profile(True)
for s in range(30, S):
    est.step(s, estimator.robot_store.from_stream(sub, s, with_vo=vo[s]))
    ms, cnt = est.profile_read()
    a = acc[vo[s]]
    for k in ("ekf", "assemble", "solve", "resweep"):
        a[k] += ms[k]
    a["n"] += 1
for v in (False, True):
    a = acc[v]
    if a["n"] == 0:
        continue
    print(f"window_solve={ws} {prec} {'ragged ' if ragged else ''}vo_tick={v}: ticks {a['n']}  ekf {1e3*a['ekf']/a['n']:.1f} us  assemble {1e3*a['assemble']/a['n']:.1f} us  solve {1e3*(a['solve']+a['resweep'])/a['n']:.1f} us")
