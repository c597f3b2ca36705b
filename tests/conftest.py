import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CODES = os.path.join(ROOT, "qldpc_b200", "data", "codes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_code_file(stem, layout="F"):
    d = np.load(os.path.join(CODES, stem + ".npz"))
    H = d["Hx"]
    if layout == "C":
        H = np.ascontiguousarray(H)
    return H, d


@pytest.fixture(scope="session")
def bp_golden():
    d = np.load(os.path.join(GOLDEN, "bp_golden.npz"))
    return d, json.loads(str(d["meta"]))


@pytest.fixture(scope="session")
def osd_golden():
    d = np.load(os.path.join(GOLDEN, "osd_golden.npz"))
    return d, json.loads(str(d["meta"]))


@pytest.fixture(scope="session")
def spacetime_golden():
    return np.load(os.path.join(GOLDEN, "spacetime_golden.npz"))


@pytest.fixture(scope="session")
def reference_stats():
    with open(os.path.join(GOLDEN, "reference_stats.json")) as f:
        return json.load(f)
