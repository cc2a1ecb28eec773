"""Small runs of the paths added in the second half of round 2 for compute-sanitizer (memcheck / synccheck / racecheck): the role
form of the fused tick (named barriers), the VO compaction launches, the proxy-fenced TMA ring, general linear rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from decentralized_ekf_mhe_b200 import build, estimator, synth
build.build()
dev = torch.device("cuda", 0)


def run(robot, n, S, rows=False, **over):
    st = synth.make_stream(n, S, robot=robot, vo_jitter=True, device=dev)
    vo = [bool(st["vo_flag"][s].any()) for s in range(S)]
    sub = {k: v.contiguous() for k, v in st.items() if torch.is_tensor(v) and v.shape[0] == S}
    est = estimator.BatchedEstimator(estimator.robot_params(robot, ekf_rate=200, N=8, **over), n)
    if rows:
        a = np.zeros((2, 9))
        a[0, 3], a[0, 5], a[1, 6], a[1, 7] = 1.0, 0.5, 1.0, -1.0
        est.add_state_rows(a, [0.47, -0.003], [0.52, 0.003])
    est.run(0, S - 6, sub, vo)
    for s in range(S - 6, S):
        est.step(s, estimator.robot_store.from_stream(sub, s))
    torch.cuda.synchronize()
    assert torch.isfinite(est.x_MHE_).all()
    est.close()
    print("ok", robot, n, rows, over, flush=True)


run("go1", 70, 30, window_solve=0)      # k_fused_roles, 3 CTAs, last one partly filled
run("cassie", 33, 24, window_solve=1)   # role kernel, two legs, incremental
os.environ["DEKF_FUSED_MAX_N"] = "0"    # split kernels + dekf_run pipeline + VO compaction at this size
run("go1", 300, 36, window_solve=0)
run("go1", 300, 36, window_solve=1)
run("pogox", 96, 30, rows=True)         # general linear rows (k_solve_box in the row basis)
