import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def go1_stream_small():
    """8 instances x 160 steps of the Go1 trot stream with ragged VO arrival."""
    from decentralized_ekf_mhe_b200 import synth
    return synth.to_numpy(synth.make_stream(8, 160, vo_jitter=True, truth=True))
