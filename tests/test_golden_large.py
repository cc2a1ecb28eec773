"""The enlarged reference goldens (tests/golden/go1_refnodes_golden_large.npz, make_refnodes_golden_large.py):

  mhe64   64 instances x 400 ticks of the headline configuration through the reference's own compiled sources, exact QP optimum
  admm16  16 instances x 200 ticks, the reference's solveQP() driven by the OSQP-style ADMM at eps_abs = eps_rel = 1e-8
          (cold setup every tick, no time limit) -- the other side of BASELINE.json's "both sides solving to eps 1e-8"
  exact16 the same stream with the exact optimum (the ADMM iterates stop up to 2.7e-5 m/s short of it)

Inputs are regenerated from (seed, kwargs) and their checksum is compared with the stored one.  CPU: the oracle restatement
and the kernel math compiled for the host against both.  GPU: the CUDA path through the C ABI against both (quaternion 1e-9,
velocity 1e-6 m/s -- asserted tighter where the exact optimum allows --, contact sets exact)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(HERE, "hostsim"))
LARGE = os.path.join(HERE, "golden", "go1_refnodes_golden_large.npz")


def _case(name):
    import make_refnodes_golden_large as mk
    g = np.load(LARGE)
    st = mk.make_inputs(name)
    assert mk.input_checksum(st) == bytes(g[name + "/sha256"]).decode(), "the synthetic stream generator drifted from the golden inputs"
    ref = {k.split("/out_")[1]: g[k] for k in g.files if k.startswith(name + "/out_")}
    return st, ref


def test_oracle_matches_large_reference_golden(oracle):
    st, ref = _case("mhe64")
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(), oracle.ekf_params(rate=200), nthreads=os.cpu_count() or 1,
                                want=("quat", "x", "v_body", "contact", "p_vo"))
    assert np.abs(ro["quat"] - ref["quat"]).max() < 1e-12
    assert np.abs(ro["x"][1:] - ref["x"][1:]).max() < 1e-9
    assert np.abs(ro["v_body"][1:] - ref["v_body"][1:]).max() < 1e-9
    assert np.array_equal(ro["contact"], ref["contact"])
    assert np.abs(ro["p_vo"] - ref["p_vo"]).max() < 1e-12


def test_kernel_math_matches_large_reference_golden():
    """The sweep of csrc/estimator_core.cuh (version-4/5 stages), compiled for the host, against the reference's outputs."""
    import ctypes as C
    import pyhostsim as hs
    st, ref = _case("mhe64")
    sub = {k: np.ascontiguousarray(v[..., :16]) for k, v in st.items()}
    from decentralized_ekf_mhe_b200.params import DekfConfig
    cfg = DekfConfig()
    C.CDLL(hs.build()).hostsim_default_go1(C.byref(cfg))
    cfg.ekf_rate = 200
    r = hs.run(sub, cfg)
    assert np.abs(r["quat"] - ref["quat"][..., :16]).max() < 1e-9
    assert np.abs(r["x"][1:] - ref["x"][1:, :, :16]).max() < 1e-9
    assert np.array_equal(r["contact"], ref["contact"][..., :16])


def _admm_bars(x, ref_admm, ref_exact):
    """The reference solving by ADMM to eps 1e-8 stops up to 2.7e-5 m/s short of the optimum on this stream (its own run with
    the exact solve, `exact16`, says so; 99 % of its velocity entries are within 1e-6).  What can be asserted of a direct
    solver: it reproduces the exact-optimum reference to round-off, hence it is never farther from the ADMM reference than
    the ADMM reference is from its own optimum, and it meets the 1e-6 m/s bar wherever the ADMM itself converged that far."""
    d_exact = np.abs(x[1:, 3:6] - ref_exact[1:, 3:6])
    d_admm = np.abs(x[1:, 3:6] - ref_admm[1:, 3:6])
    gap = np.abs(ref_admm[1:, 3:6] - ref_exact[1:, 3:6])   # the ADMM's own distance from the optimum
    assert d_exact.max() < 1e-9
    assert (d_admm <= gap + 1e-9).all()
    assert 2e-5 < gap.max() < 5e-5 and (gap < 1e-6).mean() > 0.98    # documents the figure quoted in DESIGN.md
    assert d_admm[gap < 5e-7].max() < 1e-6


def test_oracle_against_the_reference_solving_by_admm_to_1e8(oracle):
    st, ref = _case("admm16")
    _, ref_exact = _case("exact16")
    ro, _, _ = oracle.run_batch(st, oracle.go1_params(), oracle.ekf_params(rate=200), nthreads=os.cpu_count() or 1, want=("x", "v_body"))
    _admm_bars(ro["x"], ref["x"], ref_exact["x"])


@pytest.mark.gpu
@pytest.mark.parametrize("window_solve", [0, 1], ids=["full-resweep", "incremental"])
def test_cuda_path_matches_large_reference_golden(window_solve):
    import torch
    from decentralized_ekf_mhe_b200 import estimator as E
    st, ref = _case("mhe64")
    S, _, n = st["gyro"].shape
    dev = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in st.items() if not k.endswith("_ns")}
    vo = [bool(st["vo_flag"][s].any()) for s in range(S)]
    est = E.BatchedEstimator(E.robot_params("go1", ekf_rate=200, window_solve=window_solve), n)
    o = {"quat": torch.empty(S, 4, n, dtype=torch.float64, device="cuda"), "x": torch.empty(S, 9, n, dtype=torch.float64, device="cuda"),
         "v_body": torch.empty(S, 3, n, dtype=torch.float64, device="cuda"), "contact": torch.empty(S, 4, n, dtype=torch.uint8, device="cuda"),
         "status": torch.empty(S, n, dtype=torch.int32, device="cuda")}
    est.run(0, S, dev, vo, out=o, out_per_step=True)
    torch.cuda.synchronize()
    pv = est.p_vo_accmulate_.cpu().numpy()
    est.close()
    assert np.abs(o["quat"].cpu().numpy() - ref["quat"]).max() < 1e-9
    dx = np.abs(o["x"].cpu().numpy()[1:] - ref["x"][1:])
    assert dx[:, 3:6].max() < 1e-6 and dx.max() < 1e-9
    assert np.abs(o["v_body"].cpu().numpy()[1:] - ref["v_body"][1:]).max() < 1e-9
    assert np.array_equal(o["contact"].cpu().numpy(), ref["contact"])
    assert np.abs(pv - ref["p_vo"][-1]).max() < 1e-12


@pytest.mark.gpu
def test_cuda_path_against_the_reference_solving_by_admm_to_1e8():
    """BASELINE.json's parity definition taken literally: the reference side solves its QP with (OSQP-style) ADMM at
    eps_abs = eps_rel = 1e-8, our side with the direct sweep, 16 instances x 200 ticks (bars: see _admm_bars)."""
    import torch
    from decentralized_ekf_mhe_b200 import estimator as E
    st, ref = _case("admm16")
    _, ref_exact = _case("exact16")
    S, _, n = st["gyro"].shape
    dev = {k: torch.as_tensor(np.ascontiguousarray(v)).cuda() for k, v in st.items() if not k.endswith("_ns")}
    est = E.BatchedEstimator(E.robot_params("go1", ekf_rate=200), n)
    xs = np.full((S, 9, n), np.nan)
    for s in range(S):
        est.step(s, E.robot_store.from_stream(dev, s))
        xs[s] = est.x_MHE_.cpu().numpy()
    est.close()
    _admm_bars(xs, ref["x"], ref_exact["x"])
