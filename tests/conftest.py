"""pytest configuration: marker registration and shared paths.

`-m "not gpu"` runs here (no GPU): oracle vs golden fixtures, host logic, C-ABI symbol checks.
`-m gpu` runs on a B200: the parity tests proper, all through the C-ABI of libdhj.so.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    with np.errstate(all="ignore"):
        e = np.abs(got - want) / np.abs(want)
    both_nan = np.isnan(got) & np.isnan(want)
    return np.where(both_nan | (got == want), 0.0, e)
