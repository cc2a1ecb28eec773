"""Generates tests/golden/go1_kin_golden.npz from the REFERENCE's own FROST kinematics
(oracle/_ref/libfrost_go1.so, compiled from /root/reference/src/go1_example/src/Expressions/*.cc by
`make -C oracle ref`).  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_go1_kin_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

ref = po.FrostRef()
rng = np.random.default_rng(20240510)
K = 64
q = rng.uniform(-1.2, 1.2, size=(K, 4, 3))
q[0] = [[0.1, 0.8, -1.5]] * 4  # SURVEY.md App. E probe point
p = np.zeros((K, 4, 3))
J = np.zeros((K, 4, 3, 3))
Jfull = np.zeros((K, 4, 3, 22))
for k in range(K):
    var = np.zeros(22)
    for leg in range(4):
        var[6 + 4 * leg:9 + 4 * leg] = q[k, leg]  # go1Sub.cpp:72-73 packing
    for leg in range(4):
        p[k, leg] = ref.foot(leg, var)
        Jfull[k, leg] = ref.jac(leg, var)
        J[k, leg] = Jfull[k, leg][:, 6 + 4 * leg:9 + 4 * leg]  # go1Sub.cpp:91
np.savez_compressed(os.path.join(os.path.dirname(__file__), "go1_kin_golden.npz"), q=q, p=p, J=J)
print("wrote go1_kin_golden.npz", p[0, 0], J[0, 0])
