import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# fcc_qp_test.py:78-83 -- the reference's settings for the walking log
LOG_OPTS = dict(max_iter=100, rho=5e-5, eps_fcone=1e-6, eps_bound=1e-6)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def walking_log():
    from fcc_qp_b200.logdata import load_walking_log
    return load_walking_log()
