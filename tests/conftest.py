import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "gpu_unverified: GPU tests of code that has not run on a B200 yet (FDB_RUN_UNVERIFIED=1 on a GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) libfdb200.so and the oracle libraries once per session."""
    import __graft_entry__ as g
    g.build_cuda()
    from oracle import fdoracle
    fdoracle.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    from featuredetection_b200.detector import Context
    return Context(0)


@pytest.fixture(scope="session")
def face_models(built):
    from featuredetection_b200 import synthetic as syn
    return syn.landmark_models("FaceFrontal")


@pytest.fixture(scope="session")
def face_models_noexit(built):
    from featuredetection_b200 import synthetic as syn
    return syn.landmark_models("FaceFrontal", profile="no-exit")
