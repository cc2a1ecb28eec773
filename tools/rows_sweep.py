"""Randomised sweep of the general linear rows (dekf_add_state_rows) on the CPU: the kernel body of k_solve_box compiled for the host
(tests/hostsim) against the oracle, 1-4 dense random rows with bounds at the 25 % / 75 % quantiles of a.x along the unconstrained
run (so that they bind in a large share of all instance-stages).  Prints max |x - oracle| per state component, the worst row violation,
iteration-cap flags.  No GPU needed.   python tools/rows_sweep.py [trials=6] [seed=3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostsim")):
    sys.path.insert(0, p)
import numpy as np
import pyhostsim as hs
import test_hostsim as T
from decentralized_ekf_mhe_b200 import synth
from oracle import pyoracle as oracle

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
n, S = 24, 100
st = synth.to_numpy(synth.make_stream(n, S, robot="pogox", vo_jitter=True))
kw = dict(robot=2, num_legs=1, contact_effort_threshold=100.0, p_ib=(0.0, 0.0, 0.0))
X = hs.run(st, T._cfg(**kw))["x"][20:]
worst = 0.0
for trial in range(trials):
    m = int(rng.integers(1, 5))
    A = rng.normal(size=(m, 9)) * np.array([1, 1, 1, 1, 1, 1, 3, 3, 3])
    vals = np.einsum("rc,scn->srn", A, X)
    lo, hi = np.quantile(vals, 0.25, axis=(0, 2)), np.quantile(vals, 0.75, axis=(0, 2))
    r = hs.run(st, T._cfg(**kw), rows=(A, lo, hi))
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(x_row_count=m, x_row_a=tuple(A.reshape(-1)) + (0.0,) * (81 - 9 * m),
                                                      x_row_lo=tuple(lo) + (0.0,) * (9 - m), x_row_hi=tuple(hi) + (0.0,) * (9 - m), **kw),
                                oracle.ekf_params(rate=200), nthreads=os.cpu_count() or 1, want=("x",))
    dd = np.abs(r["x"][1:] - ro["x"][1:])
    v = np.einsum("rc,scn->srn", A, r["x"][1:])
    viol = max((v - hi[None, :, None]).max(), (lo[None, :, None] - v).max())
    bind = ((v >= hi[None, :, None] - 1e-10) | (v <= lo[None, :, None] + 1e-10)).mean()
    print(f"trial {trial}: rows {m}  max|x - oracle| {dd.max():.2e} (p {dd[:, 0:3].max():.1e}, v {dd[:, 3:6].max():.1e}, bias {dd[:, 6:9].max():.1e})  "
          f"violation {viol:.1e}  binding share {bind:.2f}  cap flags {int(((r['status'] & 64) != 0).sum())}  max factorisations {r['qp'][1:, 0, :].max()}")
    worst = max(worst, dd.max())
print("worst", worst)
