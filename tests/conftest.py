import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import build, load_package
    p = load_package()
    if not os.path.exists(p.api.library_path()):
        build()
    return p


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def armtest():
    return dict(np.load(os.path.join(GOLDEN, "armtest.npz")))


SYNTH_NAMES = ["arm_sigma", "acts5", "net3", "mlp64", "odd_tanh_out", "pendulum64"]


def load_synth(name):
    d = dict(np.load(os.path.join(GOLDEN, f"synth_{name}.npz")))
    d["layers"] = [int(x) for x in d["layers"]]
    d["acfunc"] = str(d["acfunc"])
    return d


def rel_err(got, ref):
    """(max|d|/max|ref|, rel-L2): the norm-relative metrics of SURVEY.md section 8a."""
    got, ref = np.asarray(got), np.asarray(ref)
    return float(np.abs(got - ref).max() / np.abs(ref).max()), float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
