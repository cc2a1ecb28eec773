"""Oracle pinning, kinematics part: orc_leg_fk(Go1) against the reference's own FROST-generated
code -- both live (oracle/_ref, compiled from /root/reference when present) and through the
committed golden vectors generated from it (tests/golden/make_go1_kin_golden.py)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def test_go1_fk_matches_frost_golden(oracle):
    g = np.load(os.path.join(HERE, "golden", "go1_kin_golden.npz"))
    for k in range(g["q"].shape[0]):
        for leg in range(4):
            p, J = oracle.leg_fk(oracle.ROBOT_GO1, leg, g["q"][k, leg])
            np.testing.assert_allclose(p, g["p"][k, leg], rtol=0, atol=1e-15)
            np.testing.assert_allclose(J, g["J"][k, leg], rtol=0, atol=1e-15)


def test_go1_probe_point(oracle):
    # FR_foot(hip .1, thigh .8, calf -1.5)
    p, J = oracle.leg_fk(oracle.ROBOT_GO1, 0, [0.1, 0.8, -1.5])
    np.testing.assert_allclose(p, [0.172521520020, -0.095271200774, -0.317741335432], atol=1e-12)
    np.testing.assert_allclose(J, [[0, -0.311309915, -0.162911386], [0.317741335, -0.001555253, 0.013698978],
                                   [-0.048521201, 0.015500652, -0.136532847]], atol=1e-9)


@pytest.mark.skipif(not (os.path.exists(os.path.join(HERE, "..", "oracle", "_ref", "libfrost_go1.so"))
                         or os.path.isdir("/root/reference/src/go1_example")),
                    reason="oracle/_ref not built and /root/reference absent")
def test_go1_fk_matches_frost_live(oracle):
    ref = oracle.FrostRef()
    rng = np.random.default_rng(7)
    for _ in range(50):
        var = np.zeros(22)
        qs = rng.uniform(-1.5, 1.5, (4, 3))
        for leg in range(4):
            var[6 + 4 * leg:9 + 4 * leg] = qs[leg]
        for leg in range(4):
            p, J = oracle.leg_fk(oracle.ROBOT_GO1, leg, qs[leg])
            np.testing.assert_allclose(p, ref.foot(leg, var), atol=1e-15)
            np.testing.assert_allclose(J, ref.jac(leg, var)[:, 6 + 4 * leg:9 + 4 * leg], atol=1e-15)


@pytest.mark.parametrize("robot,rid", [("go1", 0), ("cassie", 1), ("pogox", 2)])
def test_synth_chain_matches_oracle_models(oracle, robot, rid):
    from decentralized_ekf_mhe_b200 import synth
    spec = synth.ROBOTS[robot]
    rng = np.random.default_rng(3)
    for leg in range(spec["num_legs"]):
        q = rng.uniform(-0.8, 0.8, (5, spec["nj"]))
        p, J = synth.chain_fk(spec["legs"][leg], torch.tensor(q))
        for k in range(5):
            po, Jo = oracle.leg_fk(rid, leg, q[k])
            np.testing.assert_allclose(p[k].numpy(), po, atol=1e-14)
            np.testing.assert_allclose(J[k].numpy(), Jo, atol=1e-14)
            # Jacobian is the derivative of the position
            eps = 1e-6
            for j in range(spec["nj"]):
                dq = np.zeros(spec["nj"])
                dq[j] = eps
                pp, _ = oracle.leg_fk(rid, leg, q[k] + dq)
                pm, _ = oracle.leg_fk(rid, leg, q[k] - dq)
                np.testing.assert_allclose((pp - pm) / (2 * eps), Jo[:, j], atol=1e-8)
