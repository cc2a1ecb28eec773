"""Generates tests/golden/go1_stream_golden.npz: a small synthetic Go1 stream (4 instances x 130
steps, ragged VO arrival) together with the ORACLE's outputs on it (quaternions, x_MHE, v_MHE_b,
contact flags, VO / EKF index bookkeeping, final arrival cost).

These vectors are produced by oracle/ (the CPU restatement): they pin "CUDA path == oracle" including the integer
debug taps of the VO / EKF-replay index logic and guard the oracle against regressions.  The vectors produced by the
REFERENCE'S OWN sources are tests/golden/go1_refnodes_golden.npz (make_refnodes_golden.py).

    python tests/golden/make_stream_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from decentralized_ekf_mhe_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

st = synth.to_numpy(synth.make_stream(4, 130, vo_jitter=True, seed=20240510))
res, _, _ = po.run_batch(st, po.go1_params(), po.ekf_params(rate=200), nthreads=4,
                         want=("quat", "x", "v_body", "contact", "vo_dbg", "ekf_dbg", "p_vo", "arrival"))
out = {"in_" + k: v for k, v in st.items()}
out.update({"out_" + k: v for k, v in res.items()})
np.savez_compressed(os.path.join(os.path.dirname(__file__), "go1_stream_golden.npz"), **out)
print("wrote go1_stream_golden.npz", {k: v.shape for k, v in out.items() if k.startswith("out_")})
