import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle():
    return entry.load_oracle()


def make_case(pkg, h, w, F, occ, seed, n_dynamic=6, static_per_dynamic=3, r_min=3.0):
    dq, sq = pkg.synth.annular_qmaps(h, w, n_dynamic=n_dynamic, static_per_dynamic=static_per_dynamic, r_min=r_min)
    off, idx, val = pkg.synth.sparse_frames(h * w, F, occ, seed=seed)
    return dq, sq, off, idx, val
